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


class _FakeRenderer:
    """Stands in for GpuRenderer in the SharedFrame protocol test: the 'root framebuffer' and the chunk cursor live in
    a file-backed numpy memmap that every rank opens -- the CPU analogue of the IPC-mapped GPU-0 memory."""

    def __init__(self, rank, path, n_chunks):
        self.rank, self.path, self.n_chunks = rank, path, n_chunks
        self.calls = []
        self.mem = None
        self.part = None
        self.frames = 0              # shared renders so far: frame f uses cursor f & 1, the root zeroes the other one meanwhile

    def share_export(self):
        m = np.lib.format.open_memmap(self.path, mode="w+", dtype=np.int64, shape=(2 + self.n_chunks,))
        m[:2] = 0; m[2:] = -1
        self.mem = np.load(self.path, mmap_mode="r+")
        self.calls.append("export")
        return self.path.encode()

    def share_attach(self, handle):
        self.calls.append("attach" if handle is not None else "detach")
        if handle is not None:
            self.mem = np.load(handle.decode(), mmap_mode="r+")

    def share_partition(self, index, count):
        self.calls.append("partition" if count > 1 else "unpartition")
        self.part = (index, count) if count > 1 else None

    def share_reset(self):
        self.calls.append("reset")
        self.mem[:2] = 0
        self.mem.flush()

    def render_shared(self, counters=False):
        import fcntl, time
        self.calls.append("render")
        took = 0
        mem = np.load(self.path, mmap_mode="r+")
        slot = self.frames & 1
        if self.rank == 0:                                          # the root prepares the cursor of the NEXT frame (ct_gpu_render_shared)
            mem[slot ^ 1] = 0; mem.flush()
        self.frames += 1

        def trace(idx):
            assert mem[2 + idx] == -1, "chunk rendered twice"
            mem[2 + idx] = self.rank; mem.flush()                   # "peer store" of the finished chunk into the root's frame
            time.sleep(0.002 if self.rank == 1 else 0.0005)
            return 1
        # the numbering of next_chunk (ct_kernels.cuh): groups of 8R chunks, 7R dealt round-robin, R stolen from the cursor
        R, E = (self.part[1], 7) if self.part else (1, 0)
        G = 8 * R
        n_groups = (self.n_chunks + G - 1) // G
        if self.part:
            for g in range(n_groups):
                for j in range(E):
                    idx = g * G + j * R + self.part[0]
                    if idx < self.n_chunks:
                        took += trace(idx)
        per = (8 - E) * R
        with open(self.path + ".lock", "a+") as lock:
            while True:
                fcntl.flock(lock, fcntl.LOCK_EX)                    # "atomicAdd" on the shared cursor
                cur = np.load(self.path, mmap_mode="r+")
                c = int(cur[slot]); cur[slot] = c + 1; cur.flush()
                fcntl.flock(lock, fcntl.LOCK_UN)
                g = c // per
                if g >= n_groups:
                    break
                idx = g * G + E * R + (c - g * per)
                if idx < self.n_chunks:
                    took += trace(idx)
        return {"rays_primary": took}

    def sync(self):
        self.calls.append("sync")


def _shared_worker(rank, world, port, path, n_chunks, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from cobbletrace_b200 import multi
    multi.init_distributed("gloo")
    r = _FakeRenderer(rank, path, n_chunks)
    sf = multi.SharedFrame(r, root=0)
    shares = []
    for frame in range(3):
        sf.begin()
        took = sf.render(counters=True)["rays_primary"]
        sf.end()
        t = torch.tensor([took])
        dist.all_reduce(t)
        assert int(t) == n_chunks, "every chunk exactly once"
        if rank == 0:
            owners = np.load(path, mmap_mode="r")[2:]
            assert set(np.unique(owners).tolist()) <= {0, 1} and (owners >= 0).all(), "frame incomplete on the root"
            shares.append([(owners == 0).sum(), (owners == 1).sum()])
        dist.barrier()
        if rank == 0:
            np.load(path, mmap_mode="r+")[2:] = -1
    sf.close()
    want = (["export", "partition"] + ["render", "sync"] * 3 + ["unpartition"] if rank == 0
            else ["attach", "partition"] + ["render", "sync"] * 3 + ["unpartition", "detach"])
    assert r.calls == want, r.calls
    if rank == 0:
        np.save(out_path, np.array(shares))
    dist.destroy_process_group()


def test_shared_frame_protocol_two_ranks(tmp_path):
    """multi.SharedFrame on 2 gloo ranks: the root exports, the other attaches, every frame is barrier -> render -> sync
    -> barrier (no cursor reset: frames alternate between two cursors), all chunks are taken exactly once (most dealt round-robin, the rest stolen from the shared
    cursor: ct_gpu_share_partition) and land in the root's frame."""
    out = str(tmp_path / "shares.npy")
    mp.spawn(_shared_worker, args=(2, _free_port(), str(tmp_path / "frame.npy"), 60, out), nprocs=2, join=True)
    shares = np.load(out)
    assert (shares.sum(1) == 60).all() and (shares > 0).all(), shares
    assert shares[:, 0].sum() > shares[:, 1].sum(), shares          # the faster rank stole more
