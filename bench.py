#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 batched-FFT hot path.

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): batched complex
f32 split-format fft_split + ifft_split at N = 16, 32, ..., 4096, batch sized to 1 GiB of input per
launch per GPU (B = 2^30 / 8N).  One STEP = forward then inverse at each of the 9 sizes = 18
kernel launches, each reading 1 GiB and writing 1 GiB (all far larger than the 126 MB L2, so no
flush is needed between iterations).  Inputs are synthetic (uniform [-1,1)), resident in HBM when
the timed region starts.  N > 1 GPUs: the batch is sharded, every rank runs the same per-GPU
workload on its own device with no data-path collective (weak scaling).

Printed JSON (one line, rank 0): see the contract in the task statement.  `value` is whole-job
transforms/s; `e2e` is the same workload through the public context API with pinned HOST buffers
(H2D + kernel + D2H inside the timed region); `roofline` is for the kernel with the largest share
of the step; `cpu_baseline` is the reference's own modules (transpiled, oracle/_ref) on all host
cores over a bounded sample of the same workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SIZES = [16, 32, 64, 128, 256, 512, 1024, 2048, 4096]
GIB = 1 << 30
METRIC = "batched FFT transforms/s (c2c f32 split fft+ifft, N=16-4096, 1 GiB/launch/GPU)"
UNIT = "transforms/s"


def workload_config(n_gpus):
    return {
        "workload": "configs[1]: batched complex f32 split fft_split+ifft_split, N=16..4096 (9 sizes), "
                    "batch = 2^30/(8N) per GPU (1 GiB input per launch)",
        "sizes": SIZES,
        "input_bytes_per_launch_per_gpu": GIB,
        "launches_per_step": 2 * len(SIZES),
        "cache": "inputs (1 GiB per launch) larger than L2 (126 MB); no flush needed",
        "parallelism": f"batch-sharded x{n_gpus}, no data-path collective",
    }


def peaks():
    try:
        p = json.load(open(ROOT / "MEASURED_PEAKS.json"))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi during the timed region)
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        rows = [l.strip().split(", ") for l in open(self.tmp.name) if l.strip()]
        os.unlink(self.tmp.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if sm:
            busy = [s for s in sm if s > 0.5 * max(mx)] or sm
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# --------------------------------------------------------------------------------------------
# the reference arm / CPU baseline
# --------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, sample_div=16, min_seconds=0.25):
    """Times the reference's CPU implementation of the SAME workload mix on all host cores.
    Each step processes 1/sample_div of every batch (forward + inverse at each N), rows copied into
    a private module memory per thread before every transform, like bench() in
    benchmarks/lib/wat-contexts.js:125-129.  Returns (transforms_per_s, cores, kind, sample, s_per_step)."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import numpy as np
    import oracle as om
    om.build()
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(1234)
    use_ref = om.WatRef.available()
    ref = om.WatRef() if use_ref else None
    port = None if use_ref else om.Oracle()
    rows_per_n = {n: max(cores, (GIB // (8 * n)) // sample_div) for n in SIZES}
    data = {}
    for n in SIZES:
        b = rows_per_n[n]
        data[n] = (rng.uniform(-1, 1, (b, n)).astype(np.float32), rng.uniform(-1, 1, (b, n)).astype(np.float32))

    def one_step():
        t = 0.0
        for n in SIZES:
            re, im = data[n]
            for fn in ("fft_split", "ifft_split"):
                if use_ref:
                    t += ref.run_batch("fft_split_native_f32", "precompute_twiddles_split", fn, n,
                                       re, 0, im, om.SPLIT_IMAG_OFFSET, threads=cores)
                else:
                    t0 = time.perf_counter()
                    r2, i2 = re.copy(), im.copy()
                    port.lib.wfo_fft_split_f32_batch(n, re.shape[0], r2.ctypes.data_as(om._c_f32p),
                                                     i2.ctypes.data_as(om._c_f32p), int(fn == "ifft_split"))
                    t += time.perf_counter() - t0
        return t

    for _ in range(max(1, min(warmup, 2))):
        one_step()
    times = [one_step() for _ in range(max(1, steps))]
    per_step = statistics.median(times)
    transforms = 2 * sum(rows_per_n.values())
    kind = "reference" if use_ref else "port"
    sample = (f"1/{sample_div} of each batch (N=16..4096, fwd+inv; {transforms} transforms, "
              f"{2 * len(SIZES) * (GIB // sample_div) >> 20} MiB in per step), memcpy-in per transform, "
              + ("transpiled reference WAT modules (oracle/_ref/libwatref.so), one module memory per pthread"
                 if use_ref else "single-thread C port (oracle/watfft_oracle.c)"))
    return transforms / per_step, (cores if use_ref else 1), kind, sample, per_step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    value, cores, kind, sample, per_step = cpu_reference_run(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import watfft_b200 as wf
    C = wf._cabi
    lib = C.lib()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout; keep stdout to the single JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", os.devnull)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    n_gpus = world
    rc = lib.wfb_require_b200(local)
    if rc != 0:
        raise SystemExit("bench.py: " + lib.wfb_strerror(rc).decode() + " -- " + lib.wfb_last_cuda_error().decode())
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    assert sptr != 0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # device-resident buffers: 2 planes in, 2 planes out (1 GiB each pair), reused for every N
    nfloat = GIB // 8
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    a_re = torch.rand(nfloat, device=dev, generator=g) * 2 - 1
    a_im = torch.rand(nfloat, device=dev, generator=g) * 2 - 1
    b_re, b_im = torch.empty_like(a_re), torch.empty_like(a_im)
    ref_re = a_re[: 1 << 20].clone()
    ref_im = a_im[: 1 << 20].clone()
    flags = C.PLAN_NO_HOST_BUFFERS | C.PLAN_NO_DEVICE_BUFFERS
    plans = {n: wf.Plan(C.C2C, C.F32, C.SPLIT, n, GIB // (8 * n), local, flags) for n in SIZES}
    A = (a_re.data_ptr(), a_im.data_ptr())
    B = (b_re.data_ptr(), b_im.data_ptr())

    def step(events=None):
        for i, n in enumerate(SIZES):
            if events is not None:
                events[2 * i].record(stream)
            plans[n].exec_device(C.FORWARD, A, B, sptr)       # A -> B
            if events is not None:
                events[2 * i + 1].record(stream)
            plans[n].exec_device(C.INVERSE, B, A, sptr)       # B -> A (round trip restores the data)
        if events is not None:
            events[2 * len(SIZES)].record(stream)

    sampler = ClockSampler(local) if rank == 0 else None     # nvidia-smi needs ~0.3 s to start sampling
    for _ in range(max(3, args.warmup)):
        step()
    t_w = time.perf_counter()
    extra_warm = 0
    while time.perf_counter() - t_w < 0.5:                    # keep the GPU under load until the sampler is live
        step()
        torch.cuda.synchronize()
        extra_warm += 1
    barrier()
    launches0 = lib.wfb_kernel_launch_count()
    K = args.steps
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(2 * len(SIZES) + 1)] for _ in range(K)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin.record(stream)
    for k in range(K):
        step(evs[k])
    t_end.record(stream)
    barrier()
    launches = lib.wfb_kernel_launch_count() - launches0
    assert launches == K * 2 * len(SIZES), launches
    elapsed_ms = t_begin.elapsed_time(t_end)
    clocks = sampler.stop() if sampler else None
    # the round trips must have restored the input: proof the timed launches did the work.  f32 rounding
    # accumulates over the ~10^3 fft->ifft round trips of a run, so the bound scales with their number ...
    trips = len(SIZES) * (K + max(3, args.warmup) + extra_warm)
    drift = float((a_re[: 1 << 20] - ref_re).abs().max())
    assert drift < max(1e-3, 4e-6 * trips), f"round-trip drift {drift} after {trips} round trips"
    # ... and one fresh step on pristine data must come back within the reference's own round-trip tolerance
    a_re[: 1 << 20] = ref_re
    a_im[: 1 << 20] = ref_im
    step()
    torch.cuda.synchronize()
    fresh = max(float((a_re[: 1 << 20] - ref_re).abs().max()), float((a_im[: 1 << 20] - ref_im).abs().max()))
    assert fresh < 1e-4, f"fresh-step round-trip error {fresh}"
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())

    transforms_per_step_gpu = 2 * sum(GIB // (8 * n) for n in SIZES)
    ms_per_step = elapsed_ms / K
    value = n_gpus * transforms_per_step_gpu / (ms_per_step * 1e-3)
    bytes_per_launch = 2 * GIB
    agg_gbs = n_gpus * 2 * len(SIZES) * bytes_per_launch / (ms_per_step * 1e-3) / 1e9

    # per-kernel launch durations (rank-local), averaged over the K timed steps
    per_kernel = []
    for i, n in enumerate(SIZES):
        for d, nm in ((0, "fwd"), (1, "inv")):
            ms = statistics.fmean(evs[k][2 * i + d].elapsed_time(evs[k][2 * i + d + 1]) for k in range(K))
            per_kernel.append({"kernel": f"k_c2c<f32,N={n},split,{nm}>", "variant": plans[n].current_variant(d), "n": n, "dir": nm, "ms": ms,
                               "GBs": bytes_per_launch / ms / 1e6, "Mtransforms_s": (GIB // (8 * n)) / ms / 1e3})
    peak, peak_src = peaks()
    dom = max(per_kernel, key=lambda r: r["ms"])
    traffic = None
    try:
        prof = json.load(open(ROOT / "profiles" / "traffic.json"))
        traffic = prof.get(dom["kernel"])
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom["kernel"], "variant": dom["variant"], "achieved": dom["GBs"], "peak": peak, "unit": "GB/s",
                "frac": dom["GBs"] / peak, "traffic": traffic, "peak_source": peak_src,
                "peak_note": "the measured peak is torch's copy_ rate; a TMA-pipelined copy (tools/microbench/copy_pipe.cu) reaches 6.85 TB/s on the same board, so fractions slightly above 1.0 are possible",
                "algorithmic_bytes_per_launch": bytes_per_launch,
                "share_of_step": dom["ms"] / ms_per_step,
                "aggregate_GBs_per_gpu": agg_gbs / n_gpus, "aggregate_frac": agg_gbs / n_gpus / peak}
    for r in per_kernel:
        r["frac"] = r["GBs"] / peak
        r["ms"] = round(r["ms"], 4); r["GBs"] = round(r["GBs"], 1); r["Mtransforms_s"] = round(r["Mtransforms_s"], 2)
        r["frac"] = round(r["frac"], 3)

    for p in plans.values():
        p.destroy()
    del a_re, a_im, b_re, b_im
    torch.cuda.empty_cache()

    # ---- e2e: same workload through the public context API with pinned HOST buffers -------------
    e2e_ms = 0.0
    h2d = d2h = 0
    e2e_iters = max(1, min(K, 2))
    for n in SIZES:
        ctx = wf.createFFTf32Split(n, batch=GIB // (8 * n), device=local)
        re, im = ctx.getRealBuffer(), ctx.getImagBuffer()
        re[:] = 0.25
        im[:] = -0.5
        ctx.forward(); ctx.inverse()                        # warm-up (also pages in the pinned buffers)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_iters):
            ctx.forward()                                   # H2D 1 GiB, kernel, D2H 1 GiB, sync
            ctx.inverse()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / e2e_iters
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e_ms += dt * 1e3
        h2d += 2 * GIB
        d2h += 2 * GIB
        assert abs(float(re[0]) - 0.25) < 1e-4
        ctx.dispose()
    e2e_value = n_gpus * transforms_per_step_gpu / (e2e_ms * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": K, "warmup": max(3, args.warmup),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(n_gpus),
        "achieved_GBs": agg_gbs,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * n_gpus, "d2h_bytes_per_step": d2h * n_gpus,
                "ms_per_step": e2e_ms, "api": "createFFTf32Split(n, batch).forward()/inverse() on pinned host buffers"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "per_kernel": per_kernel,
    }
    if n_gpus == 1 and not args.no_cpu_baseline:
        v, cores, kind, sample, _ = cpu_reference_run(steps=3, warmup=1)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def quiet_stdout():
    """stdout carries exactly ONE JSON line: libraries that chat on fd 1 (NCCL prints its version banner there on
    the first collective, whatever NCCL_DEBUG_FILE says) are sent to stderr for the rest of the run."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not (args.impl == "ours" and args.gpus > 1 and world == 1):      # (the torchrun re-launch prints through its ranks)
        quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29541", __file__, "--gpus", str(args.gpus),
               "--steps", str(args.steps), "--warmup", str(args.warmup)]
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
