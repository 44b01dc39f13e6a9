"""One process, two GPUs (ct_host_boss over devices (0, 1)): the shared-frame mode and the row-tile mode against the
one-GPU frame, several frames each; reports where a frame differs instead of stopping at the first mismatch."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cobbletrace_b200 as ct
from cobbletrace_b200 import api, host
GOLD = os.path.join(ROOT, "tests", "golden")
meta = json.load(open(os.path.join(GOLD, "golden.json")))
fs = ct.load_ctscene(os.path.join(GOLD, meta["scenes"]["scene_import_bunny"]["file"])).with_reflection(0.5)
W, H, depth = 640, 480, 2
hs = host.HostScene.from_flat(fs.without_bvh())
one = host.Boss(hs, W, H, devices=(0,), max_depth=depth)
ref, _ = one.render(np.zeros((H, W), np.uint32))
one.close()
bad = 0
for tile_rows in (0, 32, 7, 0, 5, 7, 3):
    b = host.Boss(hs, W, H, devices=(0, 1), max_depth=depth, tile_rows=tile_rows)
    try:
        for frame in range(3):
            got, st = b.render(np.zeros((H, W), np.uint32))
            diff = got != ref
            rows = np.flatnonzero(diff.any(1))
            print(f"tile_rows={tile_rows} frame {frame}: {int(diff.sum())} pixels differ" + (f", rows {rows[:6].tolist()}..{rows[-1]}, zero there: {int((got[diff] == 0).sum())}" if diff.any() else ""),
                  "tiles", len(b.tiles()), flush=True)
            bad += int(diff.any())
    finally:
        b.close()
print("OK" if not bad else f"{bad} frames differ")
