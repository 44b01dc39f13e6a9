"""One-process-per-GPU plumbing (torch.distributed).

The path shards by pixels and has exactly one exchange step -- collecting 4-byte pixels on GPU 0.
Two ways to run a frame on N GPUs, scene replicated on each:

  SharedFrame (the fast path, used by bench.py): every GPU renders the same tile with ct_gpu_render_shared; its
      primary-ray warps trace 32-pixel chunks -- 7/8 dealt round-robin, 1/8 stolen from ONE cursor in GPU 0's memory
      (atomics over NVLink; ct_gpu_share_partition) -- and its shading kernels store finished pixels straight into
      GPU 0's framebuffer (peer stores through CUDA IPC).  No data-path collective and no host in the loop;
      torch.distributed ships the IPC handle, and per frame a one-word all-reduce on the render stream is the
      rendezvous and a barrier marks the end.
  row tiles (ct_host_boss_* / exchange_tiles / gather_rows_to_root): tiles stolen from a host-side counter and
      gathered with point-to-point sends (NCCL on GPUs, gloo on CPU for the tests) -- the portable variant.
"""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def env_rank() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_distributed(backend: str | None = None):
    rank, local_rank, world = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def shared_counter_name() -> str:
    """One POSIX shm counter per job (all ranks of a torchrun share MASTER_PORT)."""
    return f"ct_tiles_{os.environ.get('MASTER_PORT', '0')}_{os.getuid()}"


class _CudaArray:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can alias it (no copy)."""

    def __init__(self, ptr: int, shape, typestr="<i4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 3, "strides": None}


def framebuffer_tensor(ptr: int, width: int, height: int, device: int) -> torch.Tensor:
    """int32 [H, W] tensor aliasing the renderer's device framebuffer (ct_gpu_framebuffer)."""
    return torch.as_tensor(_CudaArray(ptr, (height, width)), device=torch.device("cuda", device))


def tiles_to_rows(tiles: Sequence[Tuple[int, int]], height: int) -> List[Tuple[int, int]]:
    """canvas-y tile [y0,y1) -> framebuffer rows [r0,r1): row = H/2 - y (raythread.cpp:182)."""
    half = height // 2
    return [(max(0, half - (y1 - 1)), min(height, half - y0 + 1)) for y0, y1 in tiles]


def exchange_tiles(my_tiles: Sequence[Tuple[int, int]], max_tiles: int, device=None) -> List[List[Tuple[int, int]]]:
    """all_gather of who rendered which tile (tiny)."""
    world = dist.get_world_size()
    buf = torch.full((max_tiles + 1, 2), -1, dtype=torch.int32)
    buf[0, 0] = len(my_tiles)
    for i, (a, b) in enumerate(my_tiles):
        buf[1 + i, 0], buf[1 + i, 1] = a, b
    if device is not None:
        buf = buf.to(device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    res = []
    for t in out:
        t = t.cpu()
        n = int(t[0, 0])
        res.append([(int(t[1 + i, 0]), int(t[1 + i, 1])) for i in range(n)])
    return res


def gather_rows_to_root(fb: torch.Tensor, tiles_by_rank: Sequence[Sequence[Tuple[int, int]]], root: int = 0, renderer=None) -> int:
    """Every rank sends the framebuffer rows of the tiles it rendered to `root`, which receives them in place.
    `renderer` (the root's GpuRenderer, when `fb` aliases its framebuffer) is told which rows arrived, so that its
    readback covers them (ct_gpu_mark_rows).  Returns the number of bytes that crossed into root."""
    rank = dist.get_rank()
    H = fb.shape[0]
    ops, nbytes = [], 0
    for r, tiles in enumerate(tiles_by_rank):
        if r == root:
            continue
        for r0, r1 in tiles_to_rows(tiles, H):
            if r1 <= r0:
                continue
            view = fb[r0:r1]
            if rank == r:
                ops.append(dist.P2POp(dist.isend, view, root))
            elif rank == root:
                ops.append(dist.P2POp(dist.irecv, view, r))
                nbytes += view.numel() * view.element_size()
                if renderer is not None:
                    renderer.mark_rows(r0, r1)
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return nbytes


class SharedFrame:
    """N ranks, one GPU each, one frame: see the module docstring.  `renderer` is this rank's GpuRenderer with the
    scene already uploaded; rank `root`'s framebuffer receives the whole frame."""

    def __init__(self, renderer, root: int = 0, stream=None, partition: bool = True):
        """`stream`: the torch.cuda.Stream the renderer launches on (ct_gpu_set_stream).  With it, the per-frame
        rendezvous is an all-reduce of one word ON THAT STREAM: the ranks' kernels start within microseconds of each
        other instead of a host-side barrier's exit skew (which the rank that starts first pays for in full: it
        keeps stealing until the last rank is done).  Without it (CPU tests, gloo): dist.barrier()."""
        self.r, self.root = renderer, root
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self._stream = stream
        self._token = torch.zeros(1, dtype=torch.int32, device=stream.device) if (stream is not None and self.world > 1) else None
        box = [renderer.share_export() if self.rank == root else None]
        if self.world > 1:
            dist.broadcast_object_list(box, src=root)
            if self.rank != root:
                renderer.share_attach(box[0])
        self.handle = box[0]
        if partition and self.world > 1:
            # every rank renders every frame: most chunks are dealt round-robin, the rest stolen (ct_gpu_share_partition)
            renderer.share_partition(self.rank, self.world)

    def begin(self):
        """Rendezvous before the frame: the ranks' kernels should start together (a rank that starts early steals more than
        its share).  The cursor needs no reset: shared frames alternate between two cursors and the root zeroes the idle one
        on its stream (ct_gpu_render_shared)."""
        if self.world > 1:
            if self._token is not None:
                with torch.cuda.stream(self._stream):
                    dist.all_reduce(self._token)      # stream-ordered: after the root's reset, before anybody's kernels
            else:
                dist.barrier()

    def render(self, counters: bool = False):
        """This rank's share of the frame (asynchronous on the renderer's stream)."""
        return self.r.render_shared(counters=counters)

    def end(self):
        """All ranks' pixels are in the root's framebuffer after this."""
        self.r.sync()
        if self.world > 1:
            dist.barrier()

    def close(self):
        if self.world > 1:
            self.r.share_partition(0, 0)
        if self.rank != self.root and self.world > 1:
            self.r.share_attach(None)
