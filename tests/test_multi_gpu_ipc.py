"""The N > 1 path bench.py runs, proven on its pixels: world_size processes (one renderer each, gloo control plane) share ONE
frame -- CUDA-IPC attach to rank 0's framebuffer and chunk cursor, dealt + stolen 32-pixel chunks, peer stores of finished
pixels (multi.SharedFrame / ct_gpu_render_shared) -- and rank 0's read-back bitmap must be the reference's frame.
On a box with fewer GPUs than ranks every process uses device 0 (the IPC mapping, the cursor atomics and the stores take
the same code path; only the wire differs), so the test also runs on a 1-GPU box.
Match: raythread.cpp:574-588 (every row handed out exactly once), :657-661 (all partitions finished before the frame is read)."""
import json
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import dataclasses
    import torch
    import torch.distributed as dist
    import cobbletrace_b200 as ct
    from cobbletrace_b200 import api, host, multi
    multi.init_distributed("gloo")
    dev = rank if torch.cuda.device_count() >= world else 0
    gold = json.load(open(os.path.join(GOLD, "golden.json")))

    def scene(name):
        fs = ct.load_ctscene(os.path.join(GOLD, gold["scenes"][name]["file"]))
        return fs if fs.has_bvh() else host.HostScene.from_flat(fs).to_flat(with_bvh=True)

    results = []
    # (a) a golden frame of the compiled reference: bunny, mirrors forced on, depth 2, 160 x 160
    meta = gold["frames"]["bunny_refl_d2_160"]
    fs = dataclasses.replace(scene(meta["scene"]), cam_pos=np.array(meta["cam_pos"]), cam_rot=np.array(meta["cam_rot"])).with_reflection(meta["force_reflection"])
    want_a = np.load(os.path.join(GOLD, "frames_bunny_refl_d2_160.npz"))["frame"]
    cases = [("bunny_refl_d2_160", fs, meta["width"], meta["height"], meta["depth"], lambda f: np.array_equal(f, want_a))]
    # (b) BASELINE configs[0] at its full size: scene_file_cube.json 640 x 640, depth 10 (one mirror: reflection chains), SURVEY 8c hash
    for name in ("scene_file_cube", "scene_import_bunny"):
        h = gold["scenes"][name]["frame640_fnv1a"]
        cases.append((name + "_640", scene(name), 640, 640, 10, lambda f, h=h: ct.frame_fnv1a(f) == h))
    for label, fs, W, H, depth, ok in cases:
        r = api.GpuRenderer(dev).upload(fs, W, H, max_depth=depth)
        sf = multi.SharedFrame(r, root=0, stream=None)            # no stream token: the rendezvous is a gloo barrier
        for frame in range(3):
            sf.begin()
            c = sf.render(counters=True)
            sf.end()
            mine = torch.tensor([c["rays_primary"]], dtype=torch.int64)
            per = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(per, mine)
            if rank == 0:
                got = r.readback()
                results.append({"case": label, "frame": frame, "ok": bool(ok(got)), "primary_per_rank": [int(p[0]) for p in per],
                                "zero_px": int((got == 0).sum())})
        sf.close()
        r.shutdown()
        dist.barrier()
    if rank == 0:
        json.dump(results, open(out_path, "w"))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_shared_frame_across_processes_reproduces_the_reference(tmp_path, world):
    out = str(tmp_path / "shared.json")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    res = json.load(open(out))
    assert len(res) == 9
    for r in res:
        assert r["ok"], r                                          # rank 0's bitmap is the reference's frame, every frame
        assert all(n > 0 for n in r["primary_per_rank"]), r        # every rank traced a share of it
        W = 160 if r["case"].startswith("bunny_refl") else 640
        assert sum(r["primary_per_rank"]) == W * (W - 1), r        # every traced pixel exactly once (row 0 is never traced)
        assert r["zero_px"] == W, r
