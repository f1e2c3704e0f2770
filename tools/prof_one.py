#!/usr/bin/env python3
"""Launch a few kernels of chosen transforms/sizes (for ncu).  python tools/prof_one.py c2c_split:4096 r2c:1024 ... [--gib 1]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import watfft_b200 as wf  # noqa: E402

C = wf._cabi
gib = 1.0
specs = []
args = sys.argv[1:]
while args:
    a = args.pop(0)
    if a == "--gib":
        gib = float(args.pop(0))
    else:
        specs.append(a)
total = int(gib * (1 << 30))
dev = torch.device("cuda:0")
for spec in specs:
    parts = spec.split(":")
    kind, n = parts[0], int(parts[1])
    variant = int(parts[2]) if len(parts) > 2 else 0
    if kind == "stft":                                  # stft:N  (WFB_STFT_PIPE_MIN_N selects the pipelined kernel)
        ns = 1 << 26
        sp = wf.Spectrogram(ns, n, n // 4, "hann", 1, flags=C.PLAN_NO_HOST_BUFFERS | C.PLAN_NO_DEVICE_BUFFERS)
        x = torch.rand(ns, device=dev) * 2 - 1
        out = torch.empty(sp.numFrames * sp.numBins, device=dev)
        for _ in range(3):
            sp.run_device(x.data_ptr(), out.data_ptr(), 0)
        torch.cuda.synchronize()
        print("ran", spec, "frames", sp.numFrames, flush=True)
        sp.dispose()
        continue
    f64 = kind.endswith("f64")
    e = 8 if f64 else 4
    dt = torch.float64 if f64 else torch.float32
    flags = C.PLAN_NO_HOST_BUFFERS | C.PLAN_NO_DEVICE_BUFFERS
    if kind.startswith("c2c"):
        batch = total // (2 * e * n)
        layout = C.SPLIT if kind == "c2c_split" else C.INTERLEAVED
        plan = wf.Plan(C.C2C, C.F64 if f64 else C.F32, layout, n, batch, 0, flags)
        a0 = torch.rand(batch * n * (1 if layout == C.SPLIT else 2), device=dev, dtype=dt)
        a1 = torch.rand(batch * n, device=dev, dtype=dt) if layout == C.SPLIT else None
        b0 = torch.empty_like(a0)
        b1 = torch.empty_like(a1) if a1 is not None else None
        d_in = (a0.data_ptr(), a1.data_ptr() if a1 is not None else None)
        d_out = (b0.data_ptr(), b1.data_ptr() if b1 is not None else None)
        direction = C.INVERSE if kind.endswith("_inv") else C.FORWARD
    else:
        batch = total // (e * n)
        plan = wf.Plan(C.R2C, C.F64 if f64 else C.F32, 0, n, batch, 0, flags)
        t = torch.rand(batch * n, device=dev, dtype=dt)
        sp = torch.rand(batch * (n + 2), device=dev, dtype=dt)
        if kind.startswith("c2r"):
            d_in, d_out, direction = (sp.data_ptr(), None), (t.data_ptr(), None), C.INVERSE
        else:
            d_in, d_out, direction = (t.data_ptr(), None), (sp.data_ptr(), None), C.FORWARD
    plan.set_variant(variant)
    for _ in range(3):
        plan.exec_device(direction, d_in, d_out)
    plan.sync()
    print("ran", spec, plan.variants()[variant], "batch", batch, flush=True)
    plan.destroy()
