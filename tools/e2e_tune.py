#!/usr/bin/env python3
"""Times ctx.forward() (H2D + kernel + D2H) for one size under the current WFB_STAGE_* env settings."""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import watfft_b200 as wf
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
batch = (1 << 30) // (8 * n)
ctx = wf.createFFTf32Split(n, batch=batch)
ctx.getRealBuffer()[:] = 0.5; ctx.getImagBuffer()[:] = 0.25
ctx.forward(); ctx.inverse()
t0 = time.perf_counter()
for _ in range(3):
    ctx.forward(); ctx.inverse()
dt = (time.perf_counter() - t0) / 6
print(f"chunk={os.environ.get('WFB_STAGE_CHUNK_MB','16')}MB streams={os.environ.get('WFB_STAGE_STREAMS','3')} n={n}: {dt*1e3:.2f} ms/exec  {2*(1<<30)/dt/1e9:.1f} GB/s (H2D+D2H)")
ctx.dispose()
