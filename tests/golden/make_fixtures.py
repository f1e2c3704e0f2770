#!/usr/bin/env python3
"""Generates tests/golden/watref_vectors.npz from the reference's OWN modules (transpiled:
oracle/_ref/libwatref.so).  Run in the authoring container (needs /root/reference at build time):

    python tests/golden/make_fixtures.py

Inputs are the reference's seeded LCG signals (tools/accuracy_report.js:46-55, seeds 12345+n,
54321+n, 98765+n); outputs are what the reference modules produce for them.  The committed file
is what pins parity on the GPU box, where /root/reference does not exist."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "oracle"))
import oracle as om  # noqa: E402

SIZES = [4, 8, 16, 32, 64, 128, 256, 1024, 4096]


def main():
    om.build()
    w = om.WatRef()
    out = {}
    for n in SIZES:
        re = om.lcg_signal(n, 12345 + n).astype(np.float32)
        im = om.lcg_signal(n, 54321 + n).astype(np.float32)
        xr = om.lcg_signal(n, 98765 + n)
        out[f"in_re_{n}"], out[f"in_im_{n}"], out[f"in_real_{n}"] = re, im, xr
        for inv, tag in ((False, "fwd"), (True, "inv")):
            a, b = w.fft_split_f32(re, im, inv)
            out[f"split_{tag}_re_{n}"], out[f"split_{tag}_im_{n}"] = a, b
            x = np.empty(2 * n, np.float32)
            x[0::2], x[1::2] = re, im
            out[f"dual_{tag}_{n}"] = w.fft_interleaved_f32(x, inv)
            d = np.empty(2 * n)
            d[0::2], d[1::2] = om.lcg_signal(n, 12345 + n), om.lcg_signal(n, 54321 + n)
            out[f"f64_{tag}_{n}"] = w.fft_f64(d, inv)
        if n >= 32:
            spec = w.rfft_split_f32(xr.astype(np.float32))
            out[f"rfft32_{n}"] = spec
            out[f"irfft32_{n}"] = w.irfft_split_f32(spec)
        if n >= 8:
            out[f"rfft64_{n}"] = w.rfft_f64(xr)
        if 8 <= n <= 256:
            # the module behind the public createRFFTf32 context (fft_real_f32_dual): pins N = 8, 16
            spec = w.rfft_f32_dual(xr.astype(np.float32))
            out[f"rfft32dual_{n}"] = spec
            out[f"irfft32dual_{n}"] = w.irfft_f32_dual(spec)
    path = Path(__file__).resolve().parent / "watref_vectors.npz"
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({path.stat().st_size} bytes, {len(out)} arrays)")


if __name__ == "__main__":
    main()
