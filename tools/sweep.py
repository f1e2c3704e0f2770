#!/usr/bin/env python3
"""Device-resident throughput sweep over transforms / sizes / kernel variants (development tool).

python tools/sweep.py [--gib 1.0] [--kinds c2c_split,c2c_il,r2c,c2r,c2c_f64,r2c_f64] [--sizes 16,...]
Writes one JSON line per (kind, n, variant) to stdout and to gpurun_out/sweep.jsonl."""
import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import watfft_b200 as wf  # noqa: E402

C = wf._cabi
PEAK = 6547.5
try:
    PEAK = json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass


SUSTAIN = 0.0


def time_plan(plan, direction, d_in, d_out, iters=10, warm=3):
    s = torch.cuda.current_stream().cuda_stream
    if SUSTAIN > 0:
        # steady state under the power cap: run for SUSTAIN seconds, time the second half
        import time as _t
        t0 = _t.perf_counter()
        n = 0
        while _t.perf_counter() - t0 < SUSTAIN / 2:
            for _ in range(20):
                plan.exec_device(direction, d_in, d_out, s)
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = _t.perf_counter()
        while _t.perf_counter() - t0 < SUSTAIN / 2:
            for _ in range(20):
                plan.exec_device(direction, d_in, d_out, s)
            n += 20
            torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        return ms, ms
    assert s != 0, "time on a non-default stream: stream 0 means 'the plan's own stream' to the C ABI"
    for _ in range(warm):
        plan.exec_device(direction, d_in, d_out, s)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        plan.exec_device(direction, d_in, d_out, s)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=1.0)
    ap.add_argument("--kinds", default="c2c_split,c2c_il,r2c,c2r,c2c_f64,r2c_f64")
    ap.add_argument("--sizes", default="16,32,64,128,256,512,1024,2048,4096")
    ap.add_argument("--inverse", action="store_true")
    ap.add_argument("--sustain", type=float, default=0.0, help="seconds per variant of back-to-back launches (power-capped steady state)")
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "sweep.jsonl"))
    a = ap.parse_args()
    global SUSTAIN
    SUSTAIN = a.sustain
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    sizes = [int(x) for x in a.sizes.split(",")]
    total = int(a.gib * (1 << 30))
    dev = torch.device("cuda:0")
    outf = open(a.out, "a")
    torch.cuda.set_stream(torch.cuda.Stream())
    for kind in a.kinds.split(","):
        for n in sizes:
            if kind == "stft":
                # fused spectrogram: 64 Mi samples, hop = n/4, hann; bytes = samples read once + dB output written once
                if n < 64:
                    continue
                ns = (1 << 26)
                sp = wf.Spectrogram(ns, n, n // 4, "hann", 1, flags=C.PLAN_NO_HOST_BUFFERS | C.PLAN_NO_DEVICE_BUFFERS)
                x = torch.rand(ns, device=dev) * 2 - 1
                out = torch.empty(sp.numFrames * sp.numBins, device=dev)
                s_ = torch.cuda.current_stream().cuda_stream
                for _ in range(3):
                    sp.run_device(x.data_ptr(), out.data_ptr(), s_)
                torch.cuda.synchronize()
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
                ev[0].record()
                for i in range(10):
                    sp.run_device(x.data_ptr(), out.data_ptr(), s_)
                    ev[i + 1].record()
                torch.cuda.synchronize()
                ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(10))
                med = ts[5]
                nbytes = sp.algorithmic_bytes()
                rec = {"kind": kind, "n": n, "batch": sp.numFrames, "variant": "k_stft", "ms": round(med, 4), "ms_best": round(ts[0], 4),
                       "Mtransforms_s": round(sp.numFrames / med / 1e3, 2), "GBs": round(nbytes / med / 1e6, 1),
                       "frac": round(nbytes / med / 1e6 / PEAK, 3), "inverse": False}
                line = json.dumps(rec)
                print(line, flush=True)
                outf.write(line + "\n")
                sp.dispose()
                del x, out
                torch.cuda.empty_cache()
                continue
            f64 = kind.endswith("f64")
            e = 8 if f64 else 4
            if kind.startswith("c2c"):
                batch = total // (2 * e * n)
                layout = C.SPLIT if kind == "c2c_split" else C.INTERLEAVED
                if n < 4:
                    continue
                plan = wf.Plan(C.C2C, C.F64 if f64 else C.F32, layout, n, batch, 0, C.PLAN_NO_HOST_BUFFERS | C.PLAN_NO_DEVICE_BUFFERS)
                dt = torch.float64 if f64 else torch.float32
                if layout == C.SPLIT:
                    a0 = torch.rand(batch * n, device=dev, dtype=dt) * 2 - 1
                    a1 = torch.rand(batch * n, device=dev, dtype=dt) * 2 - 1
                    b0, b1 = torch.empty_like(a0), torch.empty_like(a1)
                    d_in, d_out = (a0.data_ptr(), a1.data_ptr()), (b0.data_ptr(), b1.data_ptr())
                else:
                    a0 = torch.rand(batch * 2 * n, device=dev, dtype=dt) * 2 - 1
                    b0 = torch.empty_like(a0)
                    d_in, d_out = (a0.data_ptr(), None), (b0.data_ptr(), None)
                direction = C.INVERSE if a.inverse else C.FORWARD
            else:
                lo = 8 if f64 else 32
                if n < lo:
                    continue
                batch = total // (e * n)
                plan = wf.Plan(C.R2C, C.F64 if f64 else C.F32, C.INTERLEAVED, n, batch, 0, C.PLAN_NO_HOST_BUFFERS | C.PLAN_NO_DEVICE_BUFFERS)
                dt = torch.float64 if f64 else torch.float32
                t = torch.rand(batch * n, device=dev, dtype=dt) * 2 - 1
                sp = torch.rand(batch * (n + 2), device=dev, dtype=dt) * 2 - 1
                if kind.startswith("c2r"):
                    d_in, d_out, direction = (sp.data_ptr(), None), (t.data_ptr(), None), C.INVERSE
                else:
                    d_in, d_out, direction = (t.data_ptr(), None), (sp.data_ptr(), None), C.FORWARD
            nbytes = plan.algorithmic_bytes()
            for vi, vname in enumerate(plan.variants()):
                plan.set_variant(vi)
                med, best = time_plan(plan, direction, d_in, d_out)
                rec = {"kind": kind, "n": n, "batch": batch, "variant": vname, "ms": round(med, 4), "ms_best": round(best, 4),
                       "Mtransforms_s": round(batch / med / 1e3, 2), "GBs": round(nbytes / med / 1e6, 1),
                       "frac": round(nbytes / med / 1e6 / PEAK, 3), "inverse": bool(a.inverse)}
                line = json.dumps(rec)
                print(line, flush=True)
                outf.write(line + "\n")
            plan.destroy()
            del plan
            torch.cuda.empty_cache()
    outf.close()


if __name__ == "__main__":
    main()
