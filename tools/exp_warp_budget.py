import os, sys
sys.path.insert(0, "/root/repo")
sys.argv = [sys.argv[0], "cube640"]
import importlib.util
from cobbletrace_b200 import api
spec = importlib.util.spec_from_file_location("ps", "/root/repo/tools/perf_stages.py")
ps = importlib.util.module_from_spec(spec); spec.loader.exec_module(ps)
mk, W, H, depth = ps.CASES["dragon4k"]
fs = mk()
for b, wb in [(b, wb) for b in (256, 384, 512) for wb in (1024, 2048, 4096)]:
    api.set_option("traversal_budget", b)
    api.set_option("overflow_warp_budget", wb)
    r = api.GpuRenderer(0).upload(fs, W, H, max_depth=depth, flags=api.CT_FLAG_STAGE_TIMING)
    for i in range(3):
        r.render_tile(); r.sync()
    print("budget", b, "warp_budget", wb, r.last_tile_ms(), [(n, d, round(ms, 3)) for n, d, ms in r.last_tile_stages() if "shadow" in n], flush=True)
    r.shutdown()
