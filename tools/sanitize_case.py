"""Small frames through every kernel (incl. both k_overflow passes via tiny budgets, subsampling, shared mode) for compute-sanitizer."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cobbletrace_b200 as ct
from cobbletrace_b200 import api
from conftest import load_golden_scene, load_frames, case_scene, GOLD
gold = json.load(open(os.path.join(GOLD, "golden.json")))
loader = lambda n: load_golden_scene(n, gold)
for budget, wb in ((0, 0), (16, 0), (16, 4)):
    api.set_option("traversal_budget", budget); api.set_option("overflow_warp_budget", wb)
    for case in ("bunny_refl_d2_160", "cube_160", "pc_big_96"):
        fs, meta = case_scene(case, gold, loader)
        r = api.GpuRenderer(0).upload(fs, meta["width"], meta["height"], max_depth=meta["depth"], flags=api.CT_FLAG_KEEP_HITS | api.CT_FLAG_COUNT_TESTS)
        r.render_tile(counters=True)
        ok = np.array_equal(r.readback(), load_frames(case)["frame"])
        r.share_attach(r.share_export()); r.share_reset(); r.render_shared()
        ok2 = np.array_equal(r.readback(), load_frames(case)["frame"])
        print(case, budget, wb, ok, ok2, r.overflow_stats(), flush=True)
        r.shutdown()
api.set_option("traversal_budget", 0); api.set_option("overflow_warp_budget", 0)
m = gold["frames_subsampling"]["cube_160"]
fs = loader(m["scene"])
r = api.GpuRenderer(0).upload(fs, m["width"], m["height"], max_depth=m["depth"], flags=api.CT_FLAG_SUBSAMPLING)
r.render_tile()
got = np.zeros((m["height"], m["width"]), np.uint32); r.readback(got)
print("subsampling", np.array_equal(got, np.load(os.path.join(GOLD, "frames_sub_cube_160.npz"))["frame"]))
r.shutdown()
