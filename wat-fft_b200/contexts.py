"""Context objects mirroring the reference's JS surface (index.js:69-178, index.d.ts:42-153).

    ctx = createFFTf32(1024, batch=4096)
    ctx.getInputBuffer()[:] = ...        # numpy view over PINNED host memory (stable for the
    ctx.forward()                        #   life of the context, like memory.buffer views)
    out = ctx.getOutputBuffer()

Semantics kept from the reference at batch = 1: complex contexts are in place (input and output
views are the same bytes, index.js:78-83); real contexts read `size` reals and write
(size/2+1)*2 interleaved values over the same bytes (index.js:136-141); forward()/inverse() are
synchronous and return nothing.  New: `batch` (rows are contiguous per transform), `dispose()`,
and creation raises WatFFTError when no B200 is present (no CPU fallback).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _cabi as C


def _view(ptr, nbytes, dtype):
    if not ptr:
        return None
    buf = (ctypes.c_uint8 * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype)


class Plan:
    """Thin owner of a wfb_plan*."""

    def __init__(self, kind, precision, layout, n, batch=1, device=0, flags=0):
        self._lib = C.lib()
        err = ctypes.c_int(0)
        self._p = self._lib.wfb_plan_create_ex(kind, precision, layout, int(n), int(batch), int(device), int(flags),
                                               ctypes.byref(err))
        if not self._p:
            raise C.WatFFTError(err.value)
        self.kind, self.precision, self.layout, self.n, self.batch, self.device = kind, precision, layout, n, batch, device
        self.dtype = np.float64 if precision == C.F64 else np.float32

    # -- buffers
    def host(self, which):
        return _view(self._lib.wfb_host_buffer(self._p, which), self._lib.wfb_host_bytes(self._p, which), self.dtype)

    def device_ptr(self, which):
        return self._lib.wfb_device_buffer(self._p, which)

    def nbytes(self, which):
        return self._lib.wfb_host_bytes(self._p, which)

    # -- execution
    def exec(self, direction, flags=C.EXEC_DEFAULT):
        C.check(self._lib.wfb_exec(self._p, direction, flags))

    def exec_device(self, direction, d_in, d_out, stream=None):
        a = (ctypes.c_void_p * 2)(*[ctypes.c_void_p(x) if x else None for x in d_in])
        b = (ctypes.c_void_p * 2)(*[ctypes.c_void_p(x) if x else None for x in d_out])
        C.check(self._lib.wfb_exec_device(self._p, direction, a, b, ctypes.c_void_p(stream) if stream else None))

    def sync(self):
        C.check(self._lib.wfb_sync(self._p))

    @property
    def stream(self):
        return self._lib.wfb_plan_stream(self._p)

    # -- introspection
    def variants(self):
        return [self._lib.wfb_plan_variant_name(self._p, i).decode() for i in range(self._lib.wfb_plan_variant_count(self._p))]

    def current_variant(self, direction=0):
        """name of the kernel variant exec(direction) launches"""
        return self.variants()[self._lib.wfb_plan_current_variant(self._p, direction)]

    def set_variant(self, i):
        C.check(self._lib.wfb_plan_set_variant(self._p, i))

    def algorithmic_bytes(self):
        return self._lib.wfb_plan_algorithmic_bytes(self._p)

    def set_option(self, option, value):
        C.check(self._lib.wfb_plan_set_option(self._p, option, int(value)))

    def get_option(self, option):
        return self._lib.wfb_plan_get_option(self._p, option)

    def last_path(self):
        """C.PATH_*: how the latest exec() moved its data (staged copies, chunked pipeline, or zero-copy)."""
        return self._lib.wfb_plan_last_path(self._p)

    def destroy(self):
        if getattr(self, "_p", None):
            self._lib.wfb_plan_destroy(self._p)
            self._p = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class _ComplexContext:
    """createFFT / createFFTf32 shape: {size, getInputBuffer, getOutputBuffer, forward, inverse}."""

    def __init__(self, size, precision, batch, device):
        self.size, self.batch = size, batch
        self.plan = Plan(C.C2C, precision, C.INTERLEAVED, size, batch, device)
        self._buf = self.plan.host(0)           # batch * 2 * size values

    def getInputBuffer(self):
        return self._buf

    def getOutputBuffer(self):                  # same bytes: in-place contract (index.js:78-83)
        return self._buf

    def forward(self):
        self.plan.exec(C.FORWARD)

    def inverse(self):
        self.plan.exec(C.INVERSE)

    def dispose(self):
        self._buf = None
        self.plan.destroy()


class _SplitContext:
    """Split-format f32 context (the reference exposes this only as raw exports)."""

    def __init__(self, size, batch, device):
        self.size, self.batch = size, batch
        self.plan = Plan(C.C2C, C.F32, C.SPLIT, size, batch, device)
        self._re, self._im = self.plan.host(0), self.plan.host(1)

    def getRealBuffer(self):
        return self._re

    def getImagBuffer(self):
        return self._im

    def getInputBuffer(self):
        return self._re, self._im

    def getOutputBuffer(self):
        return self._re, self._im

    def forward(self):
        self.plan.exec(C.FORWARD)

    def inverse(self):
        self.plan.exec(C.INVERSE)

    def dispose(self):
        self._re = self._im = None
        self.plan.destroy()


class _RealContext:
    """createRFFT / createRFFTf32 shape: input = size reals, output = (size/2+1)*2 values.

    forward() reads getInputBuffer() and writes getOutputBuffer(); inverse() reads
    getOutputBuffer() (the spectrum) and writes getInputBuffer().  At batch = 1 both views start
    at the same address, exactly like the reference's views over memory offset 0."""

    def __init__(self, size, precision, batch, device):
        self.size, self.batch = size, batch
        self.plan = Plan(C.R2C, precision, C.INTERLEAVED, size, batch, device)
        spec = self.plan.host(C.BUF_SPECTRUM)
        if batch == 1:
            self._time = spec[:size]
        else:
            self._time = self.plan.host(C.BUF_TIME)
        self._spec = spec

    def getInputBuffer(self):
        return self._time

    def getOutputBuffer(self):
        return self._spec

    def forward(self):
        self.plan.exec(C.FORWARD)

    def inverse(self):
        self.plan.exec(C.INVERSE)

    def dispose(self):
        self._time = self._spec = None
        self.plan.destroy()


# ---- factories: same names as index.js (async there, synchronous here) -------------------
def createFFT(size, batch=1, device=0):
    """f64 interleaved complex FFT (index.js:69-91, backed by fft_combined)."""
    return _ComplexContext(size, C.F64, batch, device)


def createFFTf32(size, batch=1, device=0):
    """f32 interleaved complex FFT (index.js:98-120, backed by fft_stockham_f32_dual)."""
    return _ComplexContext(size, C.F32, batch, device)


def createRFFT(size, batch=1, device=0):
    """f64 real FFT (index.js:127-149).  inverse() is an extension: the reference's export is missing."""
    return _RealContext(size, C.F64, batch, device)


def createRFFTf32(size, batch=1, device=0):
    """f32 real FFT with the rfft_split contract (index.js:156-178).  N >= 8: sizes 8 and 16 follow the
    public createRFFTf32 context (fft_real_f32_dual, benchmarks/shared/wat-surfaces.mjs:133-141)."""
    return _RealContext(size, C.F32, batch, device)


def createFFTf32Split(size, batch=1, device=0):
    """f32 split-format complex FFT: the flagship fft_split/ifft_split path."""
    return _SplitContext(size, batch, device)


createRFFTf32Split = createRFFTf32


# ---- exports-shaped facade ------------------------------------------------------------------
class SplitExportsFacade:
    """Looks like `instance.exports` of a reference module so the reference's own suites (which
    poke raw exports) can be pointed at the GPU: a host `memory` laid out like
    the WAT memory map, precompute_* that (re)targets the instance at a size, and transform calls
    that stage memory -> GPU -> memory.  One facade per module name:

      fft_split_native_f32 : REAL_OFFSET, IMAG_OFFSET, precompute_twiddles_split,
                             precompute_rfft_twiddles_split, fft_split, ifft_split, rfft_split, irfft_split
      fft_stockham_f32_dual: precompute_twiddles, fft, ifft
      fft_combined         : precompute_twiddles, fft, ifft
      fft_real_combined    : precompute_rfft_twiddles, rfft   (+ irfft extension)
    """

    PAGES = {"fft_split_native_f32": 8, "fft_stockham_f32_dual": 4, "fft_combined": 6, "fft_real_combined": 8}
    REAL_OFFSET = 0
    IMAG_OFFSET = 32768

    def __init__(self, module="fft_split_native_f32", device=0):
        if module not in self.PAGES:
            raise ValueError(module)
        self.module, self.device = module, device
        self.memory = np.zeros(self.PAGES[module] * 65536, np.uint8)
        self._plans = {}

    def _plan(self, kind, precision, layout, n):
        key = (kind, precision, layout, n)
        if key not in self._plans:
            self._plans[key] = Plan(kind, precision, layout, n, 1, self.device)
        return self._plans[key]

    def _mem(self, dtype, offset, count):
        item = np.dtype(dtype).itemsize
        return self.memory[offset: offset + count * item].view(dtype)

    # split module
    def precompute_twiddles_split(self, n):
        self._plan(C.C2C, C.F32, C.SPLIT, n)

    def precompute_rfft_twiddles_split(self, n):
        self._plan(C.R2C, C.F32, C.INTERLEAVED, n)

    def _c2c_split(self, n, direction):
        p = self._plan(C.C2C, C.F32, C.SPLIT, n)
        p.host(0)[:] = self._mem(np.float32, self.REAL_OFFSET, n)
        p.host(1)[:] = self._mem(np.float32, self.IMAG_OFFSET, n)
        p.exec(direction)
        self._mem(np.float32, self.REAL_OFFSET, n)[:] = p.host(0)
        self._mem(np.float32, self.IMAG_OFFSET, n)[:] = p.host(1)

    def fft_split(self, n):
        self._c2c_split(n, C.FORWARD)

    def ifft_split(self, n):
        self._c2c_split(n, C.INVERSE)

    def _real(self, n, precision, direction):
        p = self._plan(C.R2C, precision, C.INTERLEAVED, n)
        dt = p.dtype
        spec = p.host(C.BUF_SPECTRUM)
        if direction == C.FORWARD:
            spec[:n] = self._mem(dt, 0, n)
            p.exec(C.FORWARD)
            self._mem(dt, 0, n + 2)[:] = spec
        else:
            spec[:] = self._mem(dt, 0, n + 2)
            p.exec(C.INVERSE)
            self._mem(dt, 0, n)[:] = spec[:n]

    def rfft_split(self, n):
        self._real(n, C.F32, C.FORWARD)

    def irfft_split(self, n):
        self._real(n, C.F32, C.INVERSE)

    # interleaved modules
    def precompute_twiddles(self, n):
        prec = C.F64 if self.module == "fft_combined" else C.F32
        self._plan(C.C2C, prec, C.INTERLEAVED, n)

    def _c2c_il(self, n, direction):
        prec = C.F64 if self.module == "fft_combined" else C.F32
        p = self._plan(C.C2C, prec, C.INTERLEAVED, n)
        p.host(0)[:] = self._mem(p.dtype, 0, 2 * n)
        p.exec(direction)
        self._mem(p.dtype, 0, 2 * n)[:] = p.host(0)

    def fft(self, n):
        self._c2c_il(n, C.FORWARD)

    def ifft(self, n):
        self._c2c_il(n, C.INVERSE)

    # f64 real module
    def precompute_rfft_twiddles(self, n):
        self._plan(C.R2C, C.F64, C.INTERLEAVED, n)

    def rfft(self, n):
        self._real(n, C.F64, C.FORWARD)

    def irfft(self, n):          # extension (the reference module has no such export)
        self._real(n, C.F64, C.INVERSE)

    def dispose(self):
        for p in self._plans.values():
            p.destroy()
        self._plans.clear()
