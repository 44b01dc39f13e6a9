"""N > 1 host logic on CPU: world_size-2 gloo processes steal tiles from the shared counter and gather their
framebuffer rows to rank 0 (the same code path bench.py uses with NCCL on GPUs)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, H, W, tile_rows, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from cobbletrace_b200 import host, multi
    multi.init_distributed("gloo")
    name = multi.shared_counter_name()
    ctr = host.TileCounter(name)
    half = H // 2
    y_lo, y_hi = -half, -half + H
    n_tiles = (y_hi - y_lo + tile_rows - 1) // tile_rows
    fb = torch.zeros((H, W), dtype=torch.int32)
    import time
    per_frame = []
    for frame in range(3):
        if rank == 0:
            ctr.reset()
        dist.barrier()
        mine = []
        while True:
            t = ctr.next()
            if t >= n_tiles:
                break
            y0 = y_lo + t * tile_rows; y1 = min(y_hi, y0 + tile_rows)
            mine.append((y0, y1))
            for (r0, r1) in multi.tiles_to_rows([(y0, y1)], H):     # "render": pixel = f(row, frame), tagged with the rank
                rows = torch.arange(r0, r1, dtype=torch.int32)[:, None]
                fb[r0:r1] = rows * 1000 + frame * 7 + torch.arange(W, dtype=torch.int32)[None, :]
            time.sleep(0.003 if rank == 1 else 0.001)               # uneven ranks: stealing rebalances by itself
        owners = multi.exchange_tiles(mine, n_tiles)
        nbytes = multi.gather_rows_to_root(fb, owners, root=0)
        dist.barrier()
        if rank == 0:
            flat = sorted(t for o in owners for t in o)
            assert flat == [(y_lo + i * tile_rows, min(y_hi, y_lo + (i + 1) * tile_rows)) for i in range(n_tiles)], "every tile exactly once"
            rows = torch.arange(H, dtype=torch.int32)[:, None]
            want = rows * 1000 + frame * 7 + torch.arange(W, dtype=torch.int32)[None, :]
            first = 1 - H % 2                                         # even H: row 0 (y = H/2) is never rendered
            assert torch.equal(fb[first:], want[first:]), "gathered frame incomplete"
            assert first == 0 or int(fb[0].abs().sum()) == 0
            assert nbytes == sum((r1 - r0) * W * 4 for r0, r1 in multi.tiles_to_rows(owners[1], H))
            per_frame.append([len(o) for o in owners])
            if frame == 2:
                np.save(out_path, np.array(per_frame))
    ctr.close(unlink=(rank == 0))
    dist.destroy_process_group()


@pytest.mark.parametrize("H,W,tile_rows", [(64, 48, 4), (101, 30, 7)])
def test_two_ranks_steal_and_gather(tmp_path, H, W, tile_rows):
    out = str(tmp_path / "owners.npy")
    mp.spawn(_worker, args=(2, _free_port(), H, W, tile_rows, out), nprocs=2, join=True)
    counts = np.load(out)                                         # [frame, rank] tiles rendered
    assert (counts.sum(1) == (H + tile_rows - 1) // tile_rows).all()
    assert (counts.sum(0) > 0).all(), counts                      # both ranks took part
    assert counts[:, 0].sum() > counts[:, 1].sum(), counts        # ... and the faster rank stole more


def test_tile_counter_local_and_shared():
    from cobbletrace_b200 import host
    a = host.TileCounter()
    assert [a.next() for _ in range(4)] == [0, 1, 2, 3]
    a.reset(); assert a.next() == 0
    a.close()
    name = f"ct_test_{os.getpid()}"
    b, c = host.TileCounter(name), host.TileCounter(name)           # two handles on one shm counter
    b.reset()
    got = [b.next(), c.next(), b.next(), c.next()]
    assert sorted(got) == [0, 1, 2, 3]
    c.close(); b.close(unlink=True)


def test_tiles_to_rows_mapping():
    from cobbletrace_b200 import multi
    # H = 8: y in [-4,4) -> rows 8..1 ; tile [-4,-2) covers y=-4 (row 8, dropped) and y=-3 (row 7)
    assert multi.tiles_to_rows([(-4, -2), (2, 4)], 8) == [(7, 8), (1, 3)]
    assert multi.tiles_to_rows([(-2, 3)], 5) == [(0, 5)]            # odd H reaches row 0
