"""Per-kernel times of ONE rank's share of a shared frame, emulated on one GPU (option "emulate_ranks")."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [sys.argv[0], "cube640"]
import importlib.util
from cobbletrace_b200 import api
spec = importlib.util.spec_from_file_location("ps", os.path.join(os.path.dirname(os.path.abspath(__file__)), "perf_stages.py"))
ps = importlib.util.module_from_spec(spec); spec.loader.exec_module(ps)
mk, W, H, depth = ps.CASES["dragon4k"]
fs = mk()
shifts = [int(v) for v in os.environ.get("CT_SHARED_SHIFTS", "0").split(",")]
budgets = [int(v) for v in os.environ.get("CT_BUDGETS", "0").split(",")]
ranks = [int(v) for v in os.environ.get("CT_RANKS", "1,2,4,8").split(",")]
runs = [int(v) for v in os.environ.get("CT_RUN_SHIFTS", "0").split(",")]
if os.environ.get("CT_PRIMARY_SPLIT"):
    api.set_option("primary_split", int(os.environ["CT_PRIMARY_SPLIT"]))
if os.environ.get("CT_PRIMARY_BUDGET"):
    api.set_option("primary_budget", int(os.environ["CT_PRIMARY_BUDGET"]))
for R, shift, budget, run in [(R, sh, b, rn) for R in ranks for sh in shifts for b in budgets for rn in runs]:
    api.set_option("emulate_ranks", R)
    api.set_option("shared_run_shift", run)
    api.set_option("shared_chunk_shift", shift)
    api.set_option("traversal_budget", budget)
    for flags, what in ((api.CT_FLAG_STAGE_TIMING, "serialised"), (0, "concurrent")):
        r = api.GpuRenderer(0).upload(fs, W, H, max_depth=depth, flags=flags)
        best = 1e9
        for i in range(6):
            r.render_tile(); r.sync()
            if i >= 2: best = min(best, r.last_tile_ms())
        line = f"ranks={R} shift={shift} budget={budget} run_shift={run} {what}: {best:.3f} ms"
        if flags:
            line += "  " + " ".join(f"{nm}[{d}]={ms:.3f}" for nm, d, ms in r.last_tile_stages())
        print(line, flush=True)
        r.shutdown()
api.set_option("emulate_ranks", 0)
api.set_option("shared_run_shift", 0)
api.set_option("shared_chunk_shift", 0)
api.set_option("traversal_budget", 0)
