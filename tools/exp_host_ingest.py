"""Scene parse time against the thread count (CT_HOST_THREADS); the geometry digest must not change."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cobbletrace_b200 import host
scene, n = bench.ensure_scene("dragon")
dig = {}
for thr in ("1", "2", "4", "8", "16", "32"):
    os.environ["CT_HOST_THREADS"] = thr
    best = 1e9
    for _ in range(2):
        t0 = time.time(); hs = host.HostScene.load(scene, base_dir=bench.scene_cache_dir()); best = min(best, time.time() - t0)
    dig[thr] = hs.to_flat(with_bvh=False).geometry_digest()
    print(thr, "threads: parse", round(best * 1e3, 1), "ms", dig[thr][:16], flush=True)
assert len(set(dig.values())) == 1
print("identical digests")
