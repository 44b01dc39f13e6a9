"""Host BVH build time against the thread count (CT_HOST_THREADS); the digest must not change."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cobbletrace_b200 import host
scene, n = bench.ensure_scene("dragon")
res = {}
for thr in ("1", "2", "4", "8", "16", "32"):
    os.environ["CT_HOST_THREADS"] = thr
    hs = host.HostScene.load(scene, base_dir=bench.scene_cache_dir())
    best = 1e9
    for _ in range(2):
        hs2 = host.HostScene.load(scene, base_dir=bench.scene_cache_dir())
        t0 = time.time(); nn = hs2.build_bvh(); best = min(best, time.time() - t0)
    res[thr] = hs2.to_flat(with_bvh=True).bvh_digest()
    print(thr, "threads:", nn, "nodes", round(best * 1e3, 1), "ms", res[thr][:16], flush=True)
assert len(set(res.values())) == 1
print("identical digests")
