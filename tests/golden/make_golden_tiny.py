#!/usr/bin/env python
"""Adds frames of degenerate bitmap sizes to tests/golden/golden.json ("tiny_frames"), from the reference compiled
under oracle/_ref:  make -C oracle ref && python tests/golden/make_golden_tiny.py

Sizes the reference handles in surprising ways: H = 1 traces nothing; odd H writes row 0 after all; W < H makes
PutPixel (draw2d.h:8-20) wrap columns outside [0, W) into the neighbouring rows.  One worker thread; the frames are
small enough to be stored inline.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ct_oracle_py as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SCENES = os.path.join(O.REF_DIR, "scenes")
SIZES = [(2, 2), (4, 2), (3, 3), (1, 5), (6, 1), (2, 7), (5, 4), (7, 2), (8, 12), (33, 9)]
CASES = [("scene_file_cube", "scene_file_cube.json", 10), ("scene_file_cube", "scene_file_cube.json", 1), ("scene_import", "scene_import.json", 10)]


def main():
    if not O.ref_available():
        sys.exit("oracle/_ref/ct_ref missing: run `make -C oracle ref` first (needs /root/reference)")
    gpath = os.path.join(GOLD, "golden.json")
    gold = json.load(open(gpath))
    gold["tiny_frames"] = []
    tmp = tempfile.mkdtemp(prefix="ctgoldtiny")
    for scene, jf, depth in CASES:
        for W, H in SIZES:
            fr = os.path.join(tmp, "f.bin")
            subprocess.run([os.path.join(O.REF_DIR, "ct_ref"), "--scene", jf, "--chdir", SCENES, "--width", str(W), "--height", str(H),
                            "--depth", str(depth), "--threads", "1", "--frame", fr], check=True, capture_output=True)
            frame = np.fromfile(fr, np.uint32).reshape(H, W)
            gold["tiny_frames"].append({"scene": scene, "width": W, "height": H, "depth": depth, "frame": frame.tolist()})
    print(len(gold["tiny_frames"]), "tiny frames")
    with open(gpath, "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
