"""A/B of the L2 access-policy window: stage times of a bench workload with option l2_persist = 0 / 1.

    python tools/exp_l2.py [workload ...]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from cobbletrace_b200 import api, host  # noqa: E402

for wl in (sys.argv[1:] or ["dragon4k"]):
    desc, kind, W, H, depth, refl = bench.WORKLOADS[wl][:6]
    hs, _ = bench.load_host_scene(kind, refl)
    fs = hs.to_flat(with_bvh=True)
    for persist in (0, 1, 0, 1):
        api.set_option("l2_persist", persist)
        r = api.GpuRenderer(0).upload(fs, W, H, max_depth=depth, flags=api.CT_FLAG_STAGE_TIMING)
        best = None
        for i in range(6):
            r.render_tile(); r.sync()
            ms = r.last_tile_ms()
            if i >= 2 and (best is None or ms < best[0]):
                best = (ms, r.last_tile_stages())
        per = {}
        for nm, d, ms in best[1]:
            per[nm] = per.get(nm, 0.0) + ms
        r.shutdown()
        r = api.GpuRenderer(0).upload(fs, W, H, max_depth=depth)
        t = []
        for i in range(6):
            r.render_tile(); r.sync(); t.append(r.last_tile_ms())
        r.shutdown()
        print(f"{wl} l2_persist={persist}: serialised {best[0]:.3f} ms  concurrent {min(t[2:]):.3f} ms  " + " ".join(f"{k}={v:.3f}" for k, v in per.items()), flush=True)
api.set_option("l2_persist", 1)
