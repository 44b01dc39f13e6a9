#!/usr/bin/env python
"""Adds the sampling-mode goldens to tests/golden/ from the reference compiled under oracle/_ref:
    make -C oracle ref && python tests/golden/make_golden_subsampling.py

settings.subsampling: the UNMODIFIED reference.  settings.supersampling: the reference with its two rand() calls
routed through a counter-based generator (ct_ref --supersampling-hash, see oracle/ref_driver.cpp CT_RAND) -- with
libc rand() on several threads the reference's own output is not reproducible.

One worker thread per frame: with several, the reference's partitions race on the rows between them
(raythread.cpp:526 writes row y-1 of the first row of a partition, which belongs to the partition below).
Outputs frames_sub_<case>.npz (frame only) and the "frames_subsampling" section of golden.json.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cobbletrace_b200.sceneio import frame_fnv1a  # noqa: E402
from oracle import ct_oracle_py as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SCENES = os.path.join(O.REF_DIR, "scenes")
# (case, scene (golden.json name), scene file, W, H, depth, force_reflection)
CASES = [
    ("cube_160", "scene_file_cube", "scene_file_cube.json", 160, 160, 10, None),
    ("cube_tall_90x151", "scene_file_cube", "scene_file_cube.json", 90, 151, 3, None),
    ("bunny_refl_d2_160", "scene_import_bunny", "scene_import_bunny.json", 160, 160, 2, 0.5),
    ("import_wide_200x121", "scene_import", "scene_import.json", 200, 121, 10, None),
]


SS_CASES = [
    ("cube_96", "scene_file_cube", "scene_file_cube.json", 96, 96, 10, None, 3),
    ("bunny_refl_d2_80x60", "scene_import_bunny", "scene_import_bunny.json", 80, 60, 2, 0.5, 2),
    ("import_wide_100x61", "scene_import", "scene_import.json", 100, 61, 10, None, 1),
]


# both settings at once (the reference blends the 16 samples of every traced pixel, then averages the rows between)
BOTH_CASES = [
    ("cube_64x61", "scene_file_cube", "scene_file_cube.json", 64, 61, 4, None),
    ("bunny_refl_d2_72x48", "scene_import_bunny", "scene_import_bunny.json", 72, 48, 2, 0.5),
    ("import_tall_40x66", "scene_import", "scene_import.json", 40, 66, 10, None),
]


def main():
    if not O.ref_available():
        sys.exit("oracle/_ref/ct_ref missing: run `make -C oracle ref` first (needs /root/reference)")
    gpath = os.path.join(GOLD, "golden.json")
    gold = json.load(open(gpath))
    gold["frames_subsampling"] = {}
    tmp = tempfile.mkdtemp(prefix="ctgoldsub")
    for case, scene, jf, W, H, depth, refl in CASES:
        fr = os.path.join(tmp, case + ".frame")
        args = ["--scene", jf, "--chdir", SCENES, "--width", str(W), "--height", str(H), "--depth", str(depth), "--threads", "1",
                "--subsampling", "--frame", fr]
        if refl is not None:
            args += ["--force-reflection", repr(refl)]
        subprocess.run([os.path.join(O.REF_DIR, "ct_ref")] + args, check=True, capture_output=True)
        frame = np.fromfile(fr, np.uint32).reshape(H, W)
        np.savez_compressed(os.path.join(GOLD, f"frames_sub_{case}.npz"), frame=frame)
        gold["frames_subsampling"][case] = {"scene": scene, "width": W, "height": H, "depth": depth, "force_reflection": refl,
                                            "fnv1a": frame_fnv1a(frame)}
        print(case, gold["frames_subsampling"][case]["fnv1a"], flush=True)
    gold["frames_supersampling"] = {}
    for case, scene, jf, W, H, depth, refl, threads in SS_CASES:
        fr = os.path.join(tmp, case + ".ssframe")
        args = ["--scene", jf, "--chdir", SCENES, "--width", str(W), "--height", str(H), "--depth", str(depth), "--threads", str(threads),
                "--supersampling-hash", "--frame", fr]
        if refl is not None:
            args += ["--force-reflection", repr(refl)]
        subprocess.run([os.path.join(O.REF_DIR, "ct_ref")] + args, check=True, capture_output=True)
        frame = np.fromfile(fr, np.uint32).reshape(H, W)
        np.savez_compressed(os.path.join(GOLD, f"frames_ss_{case}.npz"), frame=frame)
        gold["frames_supersampling"][case] = {"scene": scene, "width": W, "height": H, "depth": depth, "force_reflection": refl,
                                              "fnv1a": frame_fnv1a(frame)}
        print("supersampling", case, gold["frames_supersampling"][case]["fnv1a"], flush=True)
    gold["frames_both_sampling"] = {}
    for case, scene, jf, W, H, depth, refl in BOTH_CASES:
        fr = os.path.join(tmp, case + ".bothframe")
        args = ["--scene", jf, "--chdir", SCENES, "--width", str(W), "--height", str(H), "--depth", str(depth), "--threads", "1",
                "--subsampling", "--supersampling-hash", "--frame", fr]
        if refl is not None:
            args += ["--force-reflection", repr(refl)]
        subprocess.run([os.path.join(O.REF_DIR, "ct_ref")] + args, check=True, capture_output=True)
        frame = np.fromfile(fr, np.uint32).reshape(H, W)
        np.savez_compressed(os.path.join(GOLD, f"frames_both_{case}.npz"), frame=frame)
        gold["frames_both_sampling"][case] = {"scene": scene, "width": W, "height": H, "depth": depth, "force_reflection": refl,
                                              "fnv1a": frame_fnv1a(frame)}
        print("both", case, gold["frames_both_sampling"][case]["fnv1a"], flush=True)
    with open(gpath, "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
