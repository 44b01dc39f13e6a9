#!/usr/bin/env python
"""Adds the interactive-session goldens to tests/golden/golden.json from the reference compiled
under oracle/_ref:
    make -C oracle ref && python tests/golden/make_golden_sessions.py

A session = batches of key presses separated by '|'.  ct_ref --session queues each batch through the reference's own
AddEvent and runs one HandleUpdates + workers-to-completion per batch (one main-loop tick, cobbletrace.cpp:88-118);
after every tick it reports the camera HandleKeyboard/HandleUpdates arrived at and the FNV-1a hash of the bitmap.
(The harness appends an 'm' to every batch so that the reference re-dispatches even when a batch is empty --
'm' changes nothing but counts as a change, raythread.cpp:424-429; the session strings below therefore never rely on
"no key -> no new frame", which is tested on the host side alone.)
Output: the "sessions" section of golden.json (cameras + frame hashes only, no frames).
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ct_oracle_py as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SCENES = os.path.join(O.REF_DIR, "scenes")
# (case, scene (golden.json name), scene file, W, H, depth, threads, session)
CASES = [
    ("cube_walk", "scene_file_cube", "scene_file_cube.json", 96, 96, 10, 2, "yp|wd|m|c|oooooooooooooooooooooooooooooo|rrr|sa|i"),
    ("cube_spin", "scene_file_cube", "scene_file_cube.json", 64, 64, 4, 1, "yyyyyyyy|yyyyyyyy|yyyyyyyyyyyyyyyy|pppp|rrrrrrrrrrrr"),
    ("import_tour", "scene_import", "scene_import.json", 80, 60, 10, 2, "dddd|wwww|yx|q p|ooooo"),
    ("bunny_turn", "scene_import_bunny", "scene_import_bunny.json", 64, 64, 10, 2, "y|aaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaa|p"),
]


def main():
    if not O.ref_available():
        sys.exit("oracle/_ref/ct_ref missing: run `make -C oracle ref` first (needs /root/reference)")
    gpath = os.path.join(GOLD, "golden.json")
    gold = json.load(open(gpath))
    gold["sessions"] = {}
    for case, scene, jf, W, H, depth, threads, session in CASES:
        args = ["--scene", jf, "--chdir", SCENES, "--width", str(W), "--height", str(H), "--depth", str(depth), "--threads", str(threads),
                "--session", session]
        out = subprocess.run([os.path.join(O.REF_DIR, "ct_ref")] + args, check=True, capture_output=True, text=True).stdout
        ticks = json.loads(out.strip().splitlines()[-1])["session"]
        assert len(ticks) == session.count("|") + 2
        gold["sessions"][case] = {"scene": scene, "width": W, "height": H, "depth": depth, "session": session,
                                  "ticks": [{"pos": [float(v) for v in t["pos"]], "rot": [float(v) for v in t["rot"]], "fnv1a": t["fnv"]} for t in ticks]}
        print(case, [t["fnv"] for t in ticks], flush=True)
    with open(gpath, "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
