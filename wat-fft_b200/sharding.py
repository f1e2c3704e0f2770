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


def rank_device(local_rank: int, world: int, visible: int, order: str = "spread") -> int:
    """GPU ordinal for a rank of a one-node job.  "spread": when the node shows more GPUs than there are ranks (and a
    multiple of them), rank i takes GPU i * visible / world, so the ranks sit on different host bridges of the board --
    on this pool GPUs 0-3 share one bridge and get no more host bandwidth together than one GPU alone
    (profiles/r02_pcie.md).  "seq" (and every other case): rank i -> GPU i."""
    if not 0 <= local_rank < world:
        raise ValueError("local_rank outside the world")
    if order == "spread" and world > 1 and visible > world and visible % world == 0:
        return local_rank * (visible // world)
    return local_rank


def grains(batch: int, row_bytes: int, grain_bytes: int = 32 << 20):
    """Row ranges [(first row, rows)] of the pulled-grain schedule (ShardedSplitFFT schedule="dynamic"): grains of
    `grain_bytes` per plane, whole kernel tiles (multiples of 256 rows) once a grain holds 512 rows or more, the last
    grain whatever is left.  Returns (rows per full grain, list)."""
    if batch < 0 or row_bytes < 1 or grain_bytes < 1:
        raise ValueError("bad grain request")
    g = max(1, grain_bytes // row_bytes)
    if g >= 512:
        g &= ~255
    return g, [(r0, min(g, batch - r0)) for r0 in range(0, batch, g)]


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
    """c2c f32 split-format transform of `batch` rows sharded over `devices` from one process, host buffers in and out.

    schedule="static":  contiguous partition, one plan (with its own pinned host buffers) per device, launched back to back.
    schedule="dynamic": ONE pinned host arena per plane for the whole batch; the rows are cut into grains and
        `workers_per_device` host threads per device pull grains from a shared counter, each running
        H2D -> kernel -> D2H for its grain on its own plan (wfb_exec_host).  The host fabric of a multi-GPU box does not
        share its bandwidth evenly between GPUs (profiles/r02_pcie.md: the slowest of 8 GPUs gets 8.0 GB/s while the mean
        is 9.7), so a static partition finishes with its slowest device; pulled grains finish together, at the SUM of the
        devices' rates.  `last_counts` holds the grains each device processed in the latest run()."""

    def __init__(self, size: int, batch: int, devices, schedule: str = "static", grain_bytes: int = 32 << 20,
                 workers_per_device: int = 2):
        if schedule not in ("static", "dynamic"):
            raise ValueError("schedule must be 'static' or 'dynamic'")
        self.size, self.batch, self.devices, self.schedule = size, batch, list(devices), schedule
        self.bounds = partition(batch, len(self.devices))
        self.last_counts = [0] * len(self.devices)
        if schedule == "static":
            self.plans = [Plan(C.C2C, C.F32, C.SPLIT, size, e - b, d) if e > b else None
                          for (b, e), d in zip(self.bounds, self.devices)]
            return
        from .contexts import HostMemory
        row = 4 * size
        self._mem = [HostMemory(max(1, batch) * row), HostMemory(max(1, batch) * row)]
        self.real = self._mem[0].view(np.float32, 0, batch * size).reshape(batch, size)
        self.imag = self._mem[1].view(np.float32, 0, batch * size).reshape(batch, size)
        self.grain, self.grains = grains(batch, row, grain_bytes)
        g = self.grain
        flags = C.PLAN_NO_HOST_BUFFERS
        self.workers = []                              # (device slot, plan for full grains)
        for slot, d in enumerate(self.devices):
            for _ in range(max(1, workers_per_device)):
                self.workers.append((slot, Plan(C.C2C, C.F32, C.SPLIT, size, min(g, max(1, batch)), d, flags)))
        self.plans = [w[1] for w in self.workers]
        for pl in self.plans:
            pl.set_option(C.OPT_MAPPED_MAX_BYTES, 0)   # grains are large: always the copy path

    def buffers(self, rank):
        if self.schedule == "dynamic":
            b, e = self.bounds[rank]
            return self.real[b:e], self.imag[b:e]
        p = self.plans[rank]
        return p.host(0).reshape(-1, self.size), p.host(1).reshape(-1, self.size)

    def scatter(self, re: np.ndarray, im: np.ndarray):
        if self.schedule == "dynamic":
            self.real[:], self.imag[:] = re, im
            return
        for r, (b, e) in enumerate(self.bounds):
            if self.plans[r] is not None:
                hr, hi = self.buffers(r)
                hr[:], hi[:] = re[b:e], im[b:e]

    def gather(self):
        if self.schedule == "dynamic":
            return self.real.copy(), self.imag.copy()
        re = np.empty((self.batch, self.size), np.float32)
        im = np.empty_like(re)
        for r, (b, e) in enumerate(self.bounds):
            if self.plans[r] is not None:
                hr, hi = self.buffers(r)
                re[b:e], im[b:e] = hr, hi
        return re, im

    def run(self, inverse=False):
        d = C.INVERSE if inverse else C.FORWARD
        if self.schedule == "dynamic":
            return self._run_dynamic(d)
        for p in self.plans:            # asynchronous: all devices work concurrently
            if p is not None:
                p.exec(d, C.STAGE_H2D | C.STAGE_D2H)
        for p in self.plans:
            if p is not None:
                p.sync()

    def _run_dynamic(self, direction):
        import threading
        lock = threading.Lock()
        state = {"next": 0}
        counts = [0] * len(self.devices)
        errors = []
        row = 4 * self.size
        base = (self._mem[0].ptr, self._mem[1].ptr)

        def work(slot, plan):
            tail = None
            try:
                while True:
                    with lock:                          # claim the next grain (and count it for this device)
                        i = state["next"]
                        if i >= len(self.grains) or errors:
                            break
                        state["next"] = i + 1
                        counts[slot] += 1
                    r0, rows = self.grains[i]
                    pl = plan
                    if rows != plan.batch:             # the batch's last, shorter grain: a plan of its own size
                        tail = pl = Plan(C.C2C, C.F32, C.SPLIT, self.size, rows, plan.device, C.PLAN_NO_HOST_BUFFERS)
                        pl.set_option(C.OPT_MAPPED_MAX_BYTES, 0)
                    ptrs = (base[0] + r0 * row, base[1] + r0 * row)
                    pl.exec_host(direction, ptrs, ptrs, C.SYNC)      # in place, like the contexts; the GIL is released inside
            except Exception as ex:                     # surfaced by run(): a worker must not die silently
                errors.append(ex)
            finally:
                if tail is not None:
                    tail.destroy()

        threads = [threading.Thread(target=work, args=w) for w in self.workers]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        self.last_counts = counts
        if errors:
            raise errors[0]

    def dispose(self):
        for p in self.plans:
            if p is not None:
                p.destroy()
        self.plans = []
        if self.schedule == "dynamic":
            self.real = self.imag = None
            for m in self._mem:
                m.free()
            self._mem = []
