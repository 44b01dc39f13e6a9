"""Render a few frames of a bench workload with nothing else around them (for ncu captures).

    python tools/prof_frame.py [--workload dragon4k] [--frames 2]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from cobbletrace_b200 import api, host  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="dragon4k")
ap.add_argument("--frames", type=int, default=2)
ap.add_argument("--flags", type=int, default=0)
a = ap.parse_args()
desc, kind, W, H, depth, refl = bench.WORKLOADS[a.workload]
scene, n = bench.ensure_scene(kind)
hs = host.HostScene.load(scene, base_dir=bench.scene_cache_dir())
if refl is not None:
    hs.set_reflection(refl)
fs = hs.to_flat(with_bvh=True)
r = api.GpuRenderer(0).upload(fs, W, H, max_depth=depth, flags=a.flags)
for i in range(a.frames):
    r.render_tile()
    r.sync()
    print("frame", i, "ms", r.last_tile_ms(), flush=True)
print("launches", r.kernel_launches())
r.shutdown()
