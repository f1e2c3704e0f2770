"""Batch sharding across the B200s of one host.

Transforms are independent, so multi-GPU is contiguous row partitioning with NO data-path
collective: rank r of G owns rows [begin, end) and runs its own plan on its own device with
device-resident buffers.  Two ways to drive it:

* one process per GPU (bench.py under torchrun): each rank calls `partition()` with its RANK /
  WORLD_SIZE and creates a plan for its rows; the only cross-rank traffic is the timing reduction
  (`max_over_ranks`).
* one process, several devices: `ShardedSplitFFT` keeps one plan + stream per device and launches
  them back to back (launches are asynchronous), then synchronises all.
"""
from __future__ import annotations

import numpy as np

from . import _cabi as C
from .contexts import Plan


def partition(batch: int, world: int, rank: int | None = None):
    """Contiguous, balanced row ranges: the first (batch % world) ranks get one extra row.
    Returns [(begin, end)] for all ranks, or the single pair for `rank`."""
    if world < 1 or batch < 0:
        raise ValueError("bad partition request")
    base, extra = divmod(batch, world)
    bounds, b = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        bounds.append((b, b + n))
        b += n
    return bounds if rank is None else bounds[rank]


def max_over_ranks(value: float, backend_group=None) -> float:
    """Timing reduction used by bench.py: the job time is the slowest rank's time."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = "cuda" if dist.get_backend(backend_group) == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=backend_group)
    return float(t.item())


class ShardedSplitFFT:
    """c2c f32 split-format transform of `batch` rows sharded over `devices` from one process."""

    def __init__(self, size: int, batch: int, devices):
        self.size, self.batch, self.devices = size, batch, list(devices)
        self.bounds = partition(batch, len(self.devices))
        self.plans = [Plan(C.C2C, C.F32, C.SPLIT, size, e - b, d) if e > b else None
                      for (b, e), d in zip(self.bounds, self.devices)]

    def buffers(self, rank):
        p = self.plans[rank]
        return p.host(0).reshape(-1, self.size), p.host(1).reshape(-1, self.size)

    def scatter(self, re: np.ndarray, im: np.ndarray):
        for r, (b, e) in enumerate(self.bounds):
            if self.plans[r] is not None:
                hr, hi = self.buffers(r)
                hr[:], hi[:] = re[b:e], im[b:e]

    def gather(self):
        re = np.empty((self.batch, self.size), np.float32)
        im = np.empty_like(re)
        for r, (b, e) in enumerate(self.bounds):
            if self.plans[r] is not None:
                hr, hi = self.buffers(r)
                re[b:e], im[b:e] = hr, hi
        return re, im

    def run(self, inverse=False):
        d = C.INVERSE if inverse else C.FORWARD
        for p in self.plans:            # asynchronous: all devices work concurrently
            if p is not None:
                p.exec(d, C.STAGE_H2D | C.STAGE_D2H)
        for p in self.plans:
            if p is not None:
                p.sync()

    def dispose(self):
        for p in self.plans:
            if p is not None:
                p.destroy()
