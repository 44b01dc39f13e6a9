#!/usr/bin/env python
"""Random scene files through the compiled reference (oracle/_ref):
    make -C oracle ref && python tests/golden/make_golden_fuzz.py

Every case is a generated scene JSON (triangle soups with mirrors and all three specular kinds, 1-6 lights of every
type, spheres that the tracer ignores, numbers written in the notations the reference's GetNumber handles) rendered by
the UNMODIFIED reference at a small, sometimes odd, sometimes non-square size and a random recursion depth, with a few
key presses for the camera.  Stored per case in tests/golden/fuzz_<k>.npz: the JSON text (so that the host parser and
BVH builder can be pinned on it), the reference's flattened scene + BVH (its --dump-scene), frame and primary hit records.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cobbletrace_b200.sceneio import frame_fnv1a, load_ctscene  # noqa: E402
from oracle import ct_oracle_py as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
HIT_DT = np.dtype([("found", "<u4"), ("index", "<u4"), ("t", "<f4")])
N_CASES = 20


def num(rng, v):
    """One of the notations GetNumber (fileBuffer.cpp:160-207) parses: fixed, few digits, exponent."""
    k = rng.integers(0, 4)
    if k == 0:
        return f"{v:.4f}"
    if k == 1:
        return f"{v:.2f}"
    if k == 2:                       # the reference's exponent takes digits and '-' only: no '+'
        return f"{v:.3e}".replace("e+0", "e").replace("e+", "e").replace("e-0", "e-")
    return f"{v:.6f}"


def vec(rng, v):
    return "[" + ", ".join(num(rng, x) for x in v) + "]"


def make_scene(rng):
    n_tri = int(rng.choice([1, 2, 3, 7, 40, 150, 400]))
    spread, size = rng.choice([0.5, 1.5, 3.0]), rng.choice([0.3, 1.0, 2.5])
    objs = []
    for i in range(n_tri):
        c = rng.normal(size=3) * spread + np.array([0.0, 0.0, 6.0])
        a, b = rng.normal(size=3) * size, rng.normal(size=3) * size
        p1 = c - (a + b) / 3
        refl = float(rng.choice([0.0, 0.0, 0.3, 0.75, 1.0]))
        objs.append('{"type": "triangle", "p1": %s, "p2": %s, "p3": %s, "color": [%d, %d, %d], "specular": %d, "reflection": %s}'
                    % (vec(rng, p1), vec(rng, p1 + a), vec(rng, p1 + b), *rng.integers(0, 256, 3), int(rng.choice([-1, 0, 10, 500])), num(rng, refl)))
        if rng.random() < 0.05:
            objs.append('{"type": "sphere", "center": %s, "radius": %s, "color": [1, 2, 3], "specular": 10, "reflection": 0.5}' % (vec(rng, c), num(rng, 0.5)))
    lights = []
    for _ in range(int(rng.integers(1, 7))):          # the reference cannot parse an empty "lights" list
        kind = rng.choice(["ambient", "point", "directional"])
        if kind == "ambient":
            lights.append('{"type": "ambient", "intensity": %s}' % num(rng, rng.uniform(0.05, 0.4)))
        elif kind == "point":
            lights.append('{"type": "point", "intensity": %s, "position": %s}' % (num(rng, rng.uniform(0.1, 0.8)), vec(rng, rng.normal(size=3) * 6)))
        else:
            lights.append('{"type": "directional", "intensity": %s, "direction": %s}' % (num(rng, rng.uniform(0.1, 0.6)), vec(rng, rng.normal(size=3) * 3)))
    cam = rng.normal(size=3) * 0.5 + np.array([0.0, 0.0, -3.0])
    return ('{"objects":[\n  %s\n],\n "lights":[ %s ],\n "camera":{ "position": %s },\n'
            ' "settings":{ "numberOfThreads": 2, "subsampling": false, "wireframe": false, "supersampling": false }\n}\n'
            % (",\n  ".join(objs), ", ".join(lights), vec(rng, cam)))


def main():
    if not O.ref_available():
        sys.exit("oracle/_ref/ct_ref missing: run `make -C oracle ref` first (needs /root/reference)")
    tmp = tempfile.mkdtemp(prefix="ctfuzz")
    meta = {}
    for k in range(N_CASES):
        rng = np.random.default_rng(1000 + k)
        text = make_scene(rng)
        with open(os.path.join(tmp, "s.json"), "w") as f:
            f.write(text)
        W, H = int(rng.integers(20, 70)), int(rng.integers(20, 70))
        depth = int(rng.integers(0, 5))
        keys = "".join(rng.choice(list("yprwsadio"), size=int(rng.integers(0, 6))))
        fr, hi, du = (os.path.join(tmp, n) for n in ("f.bin", "h.bin", "d.ctscene"))
        args = ["--scene", "s.json", "--chdir", tmp, "--width", str(W), "--height", str(H), "--depth", str(depth),
                "--threads", "1" if H % 2 else "2", "--frame", fr, "--hits", hi, "--dump-scene", du]
        if keys:
            args += ["--keys", keys]
        subprocess.run([os.path.join(O.REF_DIR, "ct_ref")] + args, check=True, capture_output=True)
        fs = load_ctscene(du)
        frame = np.fromfile(fr, np.uint32).reshape(H, W)
        hits = np.fromfile(hi, HIT_DT).reshape(H, W)
        np.savez_compressed(os.path.join(GOLD, f"fuzz_{k:02d}.npz"), json=np.frombuffer(text.encode(), np.uint8), depth=depth,
                            keys=np.frombuffer(keys.encode(), np.uint8), frame=frame, found=hits["found"], index=hits["index"], t=hits["t"],
                            tri=fs.tri, mat_color=fs.mat_color, mat_specular=fs.mat_specular, mat_reflection=fs.mat_reflection,
                            light_type=fs.light_type, light_intensity=fs.light_intensity, light_pos=fs.light_pos, light_dir=fs.light_dir,
                            cam_pos=fs.cam_pos, cam_rot=fs.cam_rot, node_min=fs.node_min, node_max=fs.node_max, node_left=fs.node_left,
                            node_first=fs.node_first, node_count=fs.node_count, tri_index=fs.tri_index)
        meta[f"{k:02d}"] = {"n_tri": fs.n_tri, "n_lights": fs.n_lights, "width": W, "height": H, "depth": depth, "keys": keys, "fnv1a": frame_fnv1a(frame)}
        print(k, meta[f"{k:02d}"], flush=True)
    gpath = os.path.join(GOLD, "golden.json")
    gold = json.load(open(gpath))
    gold["fuzz"] = meta
    with open(gpath, "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
