"""N ranks (torchrun): per-kernel times of every rank's share of a REAL shared frame (dragon4k, CT_FLAG_STAGE_TIMING:
the stages are serialised on one stream), to compare with the emulated share of tools/exp_emulate_ranks.py."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib.util
import torch, torch.distributed as dist
from cobbletrace_b200 import api, multi
sys.argv = [sys.argv[0], "cube640"]
spec = importlib.util.spec_from_file_location("ps", os.path.join(ROOT, "tools", "perf_stages.py"))
rank, local, world = multi.init_distributed()
torch.cuda.set_device(local)
if rank != 0:                                   # perf_stages prints on import; keep that to rank 0
    sys.stdout = open(os.devnull, "w")
ps = importlib.util.module_from_spec(spec); spec.loader.exec_module(ps)
sys.stdout = sys.__stdout__
mk, W, H, depth = ps.CASES["dragon4k"]
if rank == 0:
    fs = mk()
dist.barrier()
fs = mk()
stream = torch.cuda.Stream(device=local)
eighths = [int(v) for v in os.environ.get("CT_EIGHTHS", "7").split(",")]      # dealt share of the chunks, in eighths
modes = ((api.CT_FLAG_STAGE_TIMING, "serialised"), (0, "concurrent")) if len(eighths) == 1 else ((0, "concurrent"),)
for e8, (flags, what) in [(e, m) for e in eighths for m in modes]:
    api.set_option("shared_static_eighths", e8)
    r = api.GpuRenderer(local).upload(fs, W, H, max_depth=depth, flags=flags)
    r.set_stream(stream.cuda_stream)
    sf = multi.SharedFrame(r, stream=stream)
    best, stages = 1e9, None
    for i in range(8):
        sf.begin(); sf.render(); sf.end()
        ms = r.last_tile_ms()
        if i >= 3 and ms < best:
            best, stages = ms, (r.last_tile_stages() if flags else None)
    agg = {}
    for nm, d, ms in (stages or []):
        agg[nm] = agg.get(nm, 0.0) + ms
    line = f"rank {rank}/{world} dealt {e8}/8 {what}: {best:.3f} ms  " + " ".join(f"{k}={v:.3f}" for k, v in agg.items())
    out = [None] * world
    dist.all_gather_object(out, line)
    if rank == 0:
        print("\n".join(out), flush=True)
    sf.close(); r.shutdown()
dist.barrier(); dist.destroy_process_group()
