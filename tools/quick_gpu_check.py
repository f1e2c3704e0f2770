import sys, math, time
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import numpy as np
import watfft_b200 as wf
from oracle import Oracle
O = Oracle()
def rel(a,b,x): return float(np.max(np.abs(np.asarray(a,np.float64)-np.asarray(b,np.float64)))/np.linalg.norm(np.asarray(x, np.float64)))
rng = np.random.default_rng(0)
print("n  split_f/i  il_f/i  r2c/c2r  f64_f/i r2c64  (ratio to bound)")
for n in [4,8,16,32,64,128,256,512,1024,2048,4096,8192]:
    b = 7
    row = []
    re = rng.uniform(-1,1,(b,n)).astype(np.float32); im = rng.uniform(-1,1,(b,n)).astype(np.float32)
    for inv in (False, True):
        try:
            c = wf.createFFTf32Split(n, batch=b); c.getRealBuffer()[:] = re.ravel(); c.getImagBuffer()[:] = im.ravel()
            c.inverse() if inv else c.forward()
            g = np.c_[c.getRealBuffer().reshape(b,n), c.getImagBuffer().reshape(b,n)]
            e = max(rel(g[r], np.r_[O.fft_split_f32(re[r],im[r],inv)], np.r_[re[r],im[r]]) for r in range(b))
            row.append(e/(2e-6*math.log2(n))); c.dispose()
        except Exception as ex: row.append(str(ex)[:40])
    x = rng.uniform(-1,1,(b,2*n)).astype(np.float32)
    for inv in (False, True):
        try:
            c = wf.createFFTf32(n, batch=b); c.getInputBuffer()[:] = x.ravel(); c.inverse() if inv else c.forward()
            g = c.getOutputBuffer().reshape(b,2*n)
            e = max(rel(g[r], O.fft_interleaved_f32(x[r],inv), x[r]) for r in range(b)); row.append(e/(2e-6*math.log2(n))); c.dispose()
        except Exception as ex: row.append(str(ex)[:40])
    if n >= 32:
        try:
            xr = rng.uniform(-1,1,(b,n)).astype(np.float32)
            c = wf.createRFFTf32(n, batch=b); c.getInputBuffer()[:] = xr.ravel(); c.forward()
            g = c.getOutputBuffer().reshape(b,n+2).copy()
            e = max(rel(g[r], O.rfft_split_f32(xr[r]), xr[r]) for r in range(b)); row.append(e/(2e-6*math.log2(n)))
            c.inverse(); t = c.getInputBuffer().reshape(b,n)
            e = max(rel(t[r], O.irfft_split_f32(g[r]), g[r]) for r in range(b)); row.append(e/(2e-6*math.log2(n))); c.dispose()
        except Exception as ex: row.append(str(ex)[:40])
    else: row += [float('nan')]*2
    d = rng.uniform(-1,1,(b,2*n))
    for inv in (False, True):
        try:
            c = wf.createFFT(n, batch=b); c.getInputBuffer()[:] = d.ravel(); c.inverse() if inv else c.forward()
            g = c.getOutputBuffer().reshape(b,2*n)
            e = max(rel(g[r], O.fft_f64(d[r],inv), d[r]) for r in range(b)); row.append(e/(1e-14*math.log2(n))); c.dispose()
        except Exception as ex: row.append(str(ex)[:40])
    if n >= 8:
        try:
            xr = rng.uniform(-1,1,(b,n))
            c = wf.createRFFT(n, batch=b); c.getInputBuffer()[:] = xr.ravel(); c.forward()
            g = c.getOutputBuffer().reshape(b,n+2)
            e = max(rel(g[r], O.rfft_f64(xr[r]), xr[r]) for r in range(b)); row.append(e/(1e-14*math.log2(n))); c.dispose()
        except Exception as ex: row.append(str(ex)[:40])
    print(n, " ".join(("%.3f"%v if isinstance(v,float) else v) for v in row), flush=True)
