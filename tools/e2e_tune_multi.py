#!/usr/bin/env python3
"""e2e staging knobs under concurrency: every rank (torchrun, one per GPU) runs createFFTf32Split(n, batch).forward()/inverse()
on pinned host buffers at the same time, for each (chunk MiB, streams) setting; rank 0 prints per-GPU and total GB/s per
direction next to the pinned-copy ceiling measured the same way (wfb_pcie_probe on all ranks at once, before and after).

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/e2e_tune_multi.py [n=1024]"""
import ctypes, json, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist
import watfft_b200 as wf

C = wf._cabi
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", os.devnull)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
lib = C.lib()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()


def reduce(v, op):
    if world == 1:
        return float(v)
    t = torch.tensor([v], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=op)
    return float(t.item())


def probe(tag):
    g = (ctypes.c_double * 4)()
    barrier()
    C.check(lib.wfb_pcie_probe(local, 256 << 20, 6, g))
    sums = [reduce(x, dist.ReduceOp.SUM) if world > 1 else x for x in g]
    if rank == 0:
        print(json.dumps({"probe": tag, "ranks": world, "h2d_alone_sum": round(sums[0], 1), "d2h_alone_sum": round(sums[1], 1),
                          "h2d_duplex_sum": round(sums[2], 1), "d2h_duplex_sum": round(sums[3], 1)}), flush=True)


probe("before")
batch = (1 << 30) // (8 * n)
ctx = wf.createFFTf32Split(n, batch=batch, device=local)
ctx.getRealBuffer()[:] = 0.5
ctx.getImagBuffer()[:] = 0.25
ctx.forward(); ctx.inverse()
for chunk_mb, streams in ((32, 3), (16, 3), (64, 3), (128, 3), (256, 2), (64, 2), (64, 4), (32, 6), (8, 6), (512, 1)):
    ctx.plan.set_option(C.OPT_STAGE_CHUNK_BYTES, chunk_mb << 20)
    ctx.plan.set_option(C.OPT_STAGE_STREAMS, streams)
    ctx.forward()
    barrier()
    t0 = time.perf_counter()
    for _ in range(2):
        ctx.forward(); ctx.inverse()
    dt = (time.perf_counter() - t0) / 4
    dt = reduce(dt, dist.ReduceOp.MAX) if world > 1 else dt
    if rank == 0:
        per = (1 << 30) / dt / 1e9
        print(json.dumps({"chunk_MiB": chunk_mb, "streams": streams, "n": n, "ms_per_exec": round(dt * 1e3, 2),
                          "GBs_per_dir_per_gpu": round(per, 2), "GBs_per_dir_total": round(per * world, 1)}), flush=True)
ctx.dispose()
probe("after")
if world > 1:
    dist.destroy_process_group()
