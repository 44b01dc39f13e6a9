"""Per-kernel device time of one frame for the BASELINE.json configs (CUDA events between launches).

    python tools/perf_stages.py [dragon4k bunny1080 pcbig1080 cube640 import640 ...]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import cobbletrace_b200 as ct  # noqa: E402
from cobbletrace_b200 import api, host  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
meta = json.load(open(os.path.join(GOLD, "golden.json")))


def golden_scene(name):
    fs = ct.load_ctscene(os.path.join(GOLD, meta["scenes"][name]["file"]))
    if not fs.has_bvh():
        fs = host.HostScene.from_flat(fs).to_flat(with_bvh=True)
    return fs


_dragon = None


def dragon():
    global _dragon
    if _dragon is None:
        scene, n = bench.ensure_scene("dragon")
        hs = host.HostScene.load(scene, base_dir=bench.scene_cache_dir())
        hs.set_reflection(0.5)
        _dragon = hs.to_flat(with_bvh=True)
    return _dragon


CASES = {
    "dragon4k": (dragon, 3840, 2160, 2),
    "dragon8k": (dragon, 7680, 4320, 2),
    "dragon1080": (dragon, 1920, 1080, 2),
    "bunny1080": (lambda: golden_scene("scene_import_bunny"), 1920, 1080, 10),
    "bunny4k_refl": (lambda: golden_scene("scene_import_bunny").with_reflection(0.5), 3840, 2160, 2),
    "pcbig1080": (lambda: golden_scene("pc_big"), 1920, 1080, 10),
    "cube640": (lambda: golden_scene("scene_file_cube"), 640, 640, 10),
    "import640": (lambda: golden_scene("scene_import"), 640, 640, 10),
}

if os.environ.get("CT_PRIMARY_SPLIT"):
    api.set_option("primary_split", int(os.environ["CT_PRIMARY_SPLIT"]))
if os.environ.get("CT_PRIMARY_BUDGET"):
    api.set_option("primary_budget", int(os.environ["CT_PRIMARY_BUDGET"]))      # experiments: -1 = never park a primary walk
names = sys.argv[1:] or ["dragon4k", "bunny1080", "pcbig1080", "cube640", "import640"]
for name in names:
    mk, W, H, depth = CASES[name]
    fs = mk()
    r = api.GpuRenderer(0).upload(fs, W, H, max_depth=depth, flags=api.CT_FLAG_STAGE_TIMING)
    best = None
    for i in range(5):
        c = r.render_tile(counters=(i == 0))
        if i == 0:
            ctr = c
        r.sync()
        ms = r.last_tile_ms()
        if i >= 2 and (best is None or ms < best[0]):
            best = (ms, r.last_tile_stages())
    rays = ctr["rays_primary"] + ctr["rays_shadow"] + ctr["rays_reflection"]
    per = {}
    for nm, d, ms in best[1]:
        per[nm] = per.get(nm, 0.0) + ms
    # one counted frame for the filter statistics
    r.shutdown()
    r = api.GpuRenderer(0).upload(fs, W, H, max_depth=depth, flags=api.CT_FLAG_COUNT_TESTS)
    cc = r.render_tile(counters=True)
    bx, tx = r.filter_stats()
    print(f"   tests: box={cc['box_tests']} (fp64 {bx / max(cc['box_tests'], 1) * 100:.2f}%)  tri={cc['tri_tests']} (fp64 {tx / max(cc['tri_tests'], 1) * 100:.2f}%)")
    print(f"{name}: {best[0]:.3f} ms/frame  {rays / best[0] / 1e3:.1f} Mrays/s  rays={rays}  " +
          " ".join(f"{k}={v:.3f}" for k, v in per.items()) + f"  overflow={r.overflow_stats()}", flush=True)
    r.shutdown()
