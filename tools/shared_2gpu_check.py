"""2+ ranks (torchrun): render the bunny with ct_gpu_render_shared and compare rank 0's frame with a one-GPU render."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import cobbletrace_b200 as ct
from cobbletrace_b200 import api, host, multi
rank, local, world = multi.init_distributed()
torch.cuda.set_device(local)
GOLD = os.path.join(ROOT, "tests", "golden")
meta = json.load(open(os.path.join(GOLD, "golden.json")))
fs = ct.load_ctscene(os.path.join(GOLD, meta["scenes"]["scene_import_bunny"]["file"]))
if not fs.has_bvh():
    fs = host.HostScene.from_flat(fs).to_flat(with_bvh=True)
fs = fs.with_reflection(0.5)
W, H, depth = 1920, 1080, 2
r = api.GpuRenderer(local).upload(fs, W, H, max_depth=depth)
ref = None
if rank == 0:
    r.render_tile(); ref = r.readback().copy()
    r.upload(fs, W, H, max_depth=depth)        # fresh framebuffer
sf = multi.SharedFrame(r)
for frame in range(3):
    sf.begin(); c = sf.render(counters=True); sf.end()
    t = torch.tensor([c["rays_primary"], c["rays_shadow"], c["rays_reflection"]], device=f"cuda:{local}", dtype=torch.int64)
    per = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(per, t)
    if rank == 0:
        got = r.readback()
        print("frame", frame, "identical to one-GPU frame:", bool(np.array_equal(got, ref)), "primary rays per rank:", [int(p[0]) for p in per], flush=True)
        assert np.array_equal(got, ref)
sf.close()
r.shutdown()
dist.barrier(); dist.destroy_process_group()
