#!/usr/bin/env python3
"""Parity of EVERY compiled kernel variant (not just the default) against the CPU oracle."""
import math
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
import numpy as np  # noqa: E402
import watfft_b200 as wf  # noqa: E402
from oracle import Oracle  # noqa: E402

C = wf._cabi
O = Oracle()
rng = np.random.default_rng(0)
bad = 0
VERBOSE = bool(os.environ.get("WFB_CHECK_VERBOSE"))   # print every worst error as a fraction of its bound


def rel(a, b, x):
    return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))) / np.linalg.norm(np.asarray(x, np.float64)))


sizes = [int(a) for a in sys.argv[1:]] or [4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192]
for n in sizes:
    b = 37 if n <= 1024 else 7          # odd: exercises the packed-lane tail
    re = rng.uniform(-1, 1, (b, n)).astype(np.float32)
    im = rng.uniform(-1, 1, (b, n)).astype(np.float32)
    il = np.empty((b, 2 * n), np.float32); il[:, 0::2] = re; il[:, 1::2] = im
    row = []
    for layout in (C.SPLIT, C.INTERLEAVED):
        plan = wf.Plan(C.C2C, C.F32, layout, n, b)
        for vi, vn in enumerate(plan.variants()):
            plan.set_variant(vi)
            for inv in (False, True):
                if layout == C.SPLIT:
                    plan.host(0)[:] = re.ravel(); plan.host(1)[:] = im.ravel()
                else:
                    plan.host(0)[:] = il.ravel()
                plan.exec(C.INVERSE if inv else C.FORWARD)
                worst = 0.0
                for r in range(b):
                    if layout == C.SPLIT:
                        o = np.r_[O.fft_split_f32(re[r], im[r], inv)]
                        g = np.r_[plan.host(0).reshape(b, n)[r], plan.host(1).reshape(b, n)[r]]
                        worst = max(worst, rel(g, o, np.r_[re[r], im[r]]))
                    else:
                        worst = max(worst, rel(plan.host(0).reshape(b, 2 * n)[r], O.fft_interleaved_f32(il[r], inv), il[r]))
                ok = worst <= 2e-6 * math.log2(n)
                bad += (not ok)
                row.append(f"{vn}{'/il' if layout else ''}{'/inv' if inv else ''}:{('ok %.3f' % (worst / (2e-6 * math.log2(n))) if VERBOSE else 'ok') if ok else 'FAIL %.2e' % worst}")
        plan.destroy()
    print(n, " ".join(row), flush=True)
# real transforms and f64: every variant against the oracle
for n in sizes:
    b = 37 if n <= 1024 else 7
    row = []
    if n >= 32:
        x = rng.uniform(-1, 1, (b, n)).astype(np.float32)
        spec = np.stack([O.rfft_split_f32(x[r]) for r in range(b)])
        plan = wf.Plan(C.R2C, C.F32, 0, n, b)
        for vi, vn in enumerate(plan.variants()):
            plan.set_variant(vi)
            plan.host(C.BUF_TIME)[:] = x.ravel(); plan.exec(C.FORWARD)
            g = plan.host(C.BUF_SPECTRUM).reshape(b, n + 2)
            w1 = max(rel(g[r], spec[r], x[r]) for r in range(b))
            plan.host(C.BUF_SPECTRUM)[:] = spec.ravel(); plan.exec(C.INVERSE)
            t = plan.host(C.BUF_TIME).reshape(b, n)
            w2 = max(rel(t[r], O.irfft_split_f32(spec[r]), spec[r]) for r in range(b))
            ok = max(w1, w2) <= 2e-6 * math.log2(n); bad += (not ok)
            row.append(f"r2c:{vn}:{('ok %.3f %.3f' % (w1 / (2e-6 * math.log2(n)), w2 / (2e-6 * math.log2(n))) if VERBOSE else 'ok') if ok else 'FAIL %.2e %.2e' % (w1, w2)}")
        plan.destroy()
    d = rng.uniform(-1, 1, (b, 2 * n))
    if n <= 8192:
        plan = wf.Plan(C.C2C, C.F64, C.INTERLEAVED, n, b)
        for vi, vn in enumerate(plan.variants()):
            plan.set_variant(vi)
            w = 0.0
            for inv in (False, True):
                plan.host(0)[:] = d.ravel(); plan.exec(C.INVERSE if inv else C.FORWARD)
                g = plan.host(0).reshape(b, 2 * n)
                w = max(w, max(rel(g[r], O.fft_f64(d[r], inv), d[r]) for r in range(b)))
            ok = w <= 1e-14 * math.log2(n); bad += (not ok)
            row.append(f"f64:{vn}:{'ok' if ok else 'FAIL %.2e' % w}")
        plan.destroy()
    if n >= 8:
        x = rng.uniform(-1, 1, (b, n))
        plan = wf.Plan(C.R2C, C.F64, 0, n, b)
        for vi, vn in enumerate(plan.variants()):
            plan.set_variant(vi)
            plan.host(C.BUF_TIME)[:] = x.ravel(); plan.exec(C.FORWARD)
            g = plan.host(C.BUF_SPECTRUM).reshape(b, n + 2).copy()
            w = max(rel(g[r], O.rfft_f64(x[r]), x[r]) for r in range(b))
            plan.exec(C.INVERSE)
            rt = float(np.max(np.abs(plan.host(C.BUF_TIME).reshape(b, n) - x)))
            ok = w <= 1e-14 * math.log2(n) and rt < 1e-9; bad += (not ok)
            row.append(f"r2c64:{vn}:{'ok' if ok else 'FAIL %.2e rt %.2e' % (w, rt)}")
        plan.destroy()
    print(n, " ".join(row), flush=True)
print("FAILURES:", bad)
sys.exit(1 if bad else 0)
