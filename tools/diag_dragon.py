import sys, time, os, numpy as np
sys.path.insert(0, ".")
import cobbletrace_b200 as ct
from cobbletrace_b200 import host, procedural, api
d = "/tmp/ct_diag"; scene, n = procedural.write_dragon_standin(d)
hs = host.HostScene.load(scene, base_dir=d); hs.set_reflection(0.5); fs = hs.to_flat(with_bvh=True)
for (W, H, depth) in [(960, 540, 0), (960, 540, 1), (960, 540, 2), (1920, 1080, 2)]:
    r = ct.GpuRenderer(0).upload(fs, W, H, max_depth=depth, flags=api.CT_FLAG_STAGE_TIMING | api.CT_FLAG_COUNT_TESTS)
    c = r.render_tile(counters=True)
    c = r.render_tile(counters=True)
    print(W, H, "depth", depth, c)
    for name, dd, ms in r.last_tile_stages(): print(f"    {name}[{dd}] {ms:.3f} ms")
    r.shutdown()
import torch
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
for (W, H, depth, do_flush) in [(3840, 2160, 2, False), (3840, 2160, 2, True), (2880, 1620, 2, False)]:
    r = ct.GpuRenderer(0).upload(fs, W, H, max_depth=depth, flags=api.CT_FLAG_STAGE_TIMING)
    for i in range(3):
        if do_flush: flush.fill_(1); torch.cuda.synchronize()
        r.render_tile(); r.sync()
    print(W, H, "depth", depth, "flush", do_flush, "total ms", r.last_tile_ms())
    for name, dd, ms in r.last_tile_stages(): print(f"    {name}[{dd}] {ms:.3f} ms")
    r.shutdown()
