"""Frame time against the traversal budget (rays whose walk exceeds it are parked for the breadth-first k_overflow)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cobbletrace_b200 import api
import importlib.util
spec = importlib.util.spec_from_file_location("ps", os.path.join(os.path.dirname(os.path.abspath(__file__)), "perf_stages.py"))
sys.argv = [sys.argv[0], "cube640"]          # keep perf_stages' own loop tiny
ps = importlib.util.module_from_spec(spec); spec.loader.exec_module(ps)
for case in ("dragon4k", "pcbig1080", "bunny1080"):
    mk, W, H, depth = ps.CASES[case]
    fs = mk()
    for budget in (256, 320, 384, 448, 512, 768):
        api.set_option("traversal_budget", budget)
        r = api.GpuRenderer(0).upload(fs, W, H, max_depth=depth, flags=api.CT_FLAG_STAGE_TIMING)
        best = None
        for i in range(5):
            r.render_tile(); r.sync()
            ms = r.last_tile_ms()
            if i >= 2 and (best is None or ms < best[0]):
                best = (ms, r.last_tile_stages())
        per = {}
        for nm, d, ms in best[1]:
            per[nm] = per.get(nm, 0.0) + ms
        print(f"{case} budget={budget}: {best[0]:.3f} ms  " + " ".join(f"{k}={v:.3f}" for k, v in per.items()) + f" parked={r.overflow_stats()[0] // 5}", flush=True)
        r.shutdown()
api.set_option("traversal_budget", 0)
