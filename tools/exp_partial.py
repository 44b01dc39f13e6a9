"""How do the stage times shrink when a GPU renders only 1/2, 1/4, 1/8 of the dragon4k frame (middle rows)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [sys.argv[0], "cube640"]
import importlib.util
from cobbletrace_b200 import api
spec = importlib.util.spec_from_file_location("ps", os.path.join(os.path.dirname(os.path.abspath(__file__)), "perf_stages.py"))
ps = importlib.util.module_from_spec(spec); spec.loader.exec_module(ps)
mk, W, H, depth = ps.CASES["dragon4k"]
fs = mk()
r = api.GpuRenderer(0).upload(fs, W, H, max_depth=depth, flags=api.CT_FLAG_STAGE_TIMING)
for frac in (1, 2, 4, 8):
    rows = H // frac
    y0 = -rows // 2
    for i in range(4):
        r.render_tile(y0, y0 + rows); r.sync()
    per = {}
    for nm, d, ms in r.last_tile_stages():
        per[nm] = per.get(nm, 0.0) + ms
    c = r.render_tile(y0, y0 + rows, counters=True)
    rays = c["rays_primary"] + c["rays_shadow"] + c["rays_reflection"]
    print(f"1/{frac} of the rows: {r.last_tile_ms():.3f} ms rays={rays} " + " ".join(f"{k}={v:.3f}" for k, v in per.items()), flush=True)
r.shutdown()
r = api.GpuRenderer(0).upload(fs, W, H, max_depth=depth, flags=api.CT_FLAG_STAGE_TIMING)
for (y0, y1, what) in ((H // 2 - 40, H // 2 - 8, "32 background rows"), (-16, 16, "32 rows through the centre"), (-64, 64, "128 centre rows")):
    for i in range(4):
        r.render_tile(y0, y1); r.sync()
    print(f"{what}: {r.last_tile_ms():.3f} ms " + " ".join(f"{nm}[{d}]={ms:.3f}" for nm, d, ms in r.last_tile_stages()), flush=True)
r.shutdown()
print("-- concurrent mode (no stage timing): whole-tile device time", flush=True)
r = api.GpuRenderer(0).upload(fs, W, H, max_depth=depth)
for frac in (1, 2, 4, 8):
    rows = H // frac
    y0 = -rows // 2
    best = 1e9
    for i in range(6):
        r.render_tile(y0, y0 + rows); r.sync()
        if i >= 2: best = min(best, r.last_tile_ms())
    print(f"concurrent 1/{frac} of the rows: {best:.3f} ms", flush=True)
for (y0, y1, what) in ((H // 2 - 40, H // 2 - 8, "32 background rows"), (-16, 16, "32 rows through the centre"), (-64, 64, "128 centre rows")):
    best = 1e9
    for i in range(6):
        r.render_tile(y0, y1); r.sync()
        if i >= 2: best = min(best, r.last_tile_ms())
    print(f"concurrent {what}: {best:.3f} ms", flush=True)
r.shutdown()
