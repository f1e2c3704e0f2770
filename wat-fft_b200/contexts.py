"""Context objects mirroring the reference's JS surface (index.js:69-178, index.d.ts:42-153).

    ctx = createFFTf32(1024, batch=4096)
    ctx.getInputBuffer()[:] = ...        # numpy view over PINNED host memory (stable for the
    ctx.forward()                        #   life of the context, like memory.buffer views)
    out = ctx.getOutputBuffer()

Semantics kept from the reference at batch = 1: complex contexts are in place (input and output
views are the same bytes, index.js:78-83); real contexts read `size` reals and write
(size/2+1)*2 interleaved values over the same bytes (index.js:136-141); forward()/inverse() are
synchronous and return nothing; every context carries `exports`, the raw module-shaped instance it
runs on (index.js:72-75), and the four create*Instance() factories return such instances
(index.js:28-58).  At batch = 1 a context is built exactly like the reference builds it: instance ->
precompute(size) -> typed views over `exports.memory` -> forward() = exports.fft(size).
New: `batch` (rows are contiguous per transform), `dispose()`, and creation raises WatFFTError when
no B200 is present (no CPU fallback).  This file and js/index.js are the same design twice.
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

    def exec_host(self, direction, h_in, h_out, flags=C.SYNC):
        """run on plane pointers inside a HostMemory (wfb_host_alloc): the exports facade's path"""
        a = (ctypes.c_void_p * 2)(*[ctypes.c_void_p(x) if x else None for x in h_in])
        b = (ctypes.c_void_p * 2)(*[ctypes.c_void_p(x) if x else None for x in h_out])
        C.check(self._lib.wfb_exec_host(self._p, direction, a, b, flags))

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


class HostMemory:
    """`exports.memory`: one module-sized block of pinned, device-mapped host memory (wfb_host_alloc).  `.buffer` is the
    byte view (the JS side hands out the same bytes as an ArrayBuffer)."""

    def __init__(self, nbytes):
        self._lib = C.lib()
        self.ptr = self._lib.wfb_host_alloc(nbytes)
        if not self.ptr:
            rc = self._lib.wfb_require_b200(0)
            raise C.WatFFTError(rc if rc else C.ERR_ALLOC)
        self.nbytes = nbytes
        self.buffer = _view(self.ptr, nbytes, np.uint8)

    def view(self, dtype, offset, count):
        item = np.dtype(dtype).itemsize
        return self.buffer[offset: offset + count * item].view(dtype)

    def free(self):
        if self.ptr:
            self.buffer = None
            self._lib.wfb_host_free(self.ptr)
            self.ptr = None


class ModuleExports:
    """Looks like `instance.exports` of a reference module: ONE `memory` laid out like the WAT
    memory map, `precompute_*` that (re)targets the instance at a size -- calling it with a new n silently re-targets,
    tests/boundary.test.js:304-332 -- and SYNCHRONOUS transform calls that work in place on that memory.  The memory is
    pinned and device-mapped, so a call is one kernel reading and writing those very bytes (no copies), returning when
    the results are in `memory`: a test may read them on the next line, as the reference suites do
    (tests/fft_split_native.test.js:78-114).

      fft_split_native_f32 : REAL_OFFSET, IMAG_OFFSET, precompute_twiddles_split, precompute_rfft_twiddles_split,
                             fft_split, ifft_split, rfft_split, irfft_split
      fft_stockham_f32_dual: precompute_twiddles, fft, ifft
      fft_combined         : precompute_twiddles, fft, ifft
      fft_real_combined    : precompute_rfft_twiddles, rfft   (+ irfft: extension, the reference export is missing)
      fft_real_f32_dual    : precompute_rfft_twiddles, rfft, irfft
    """

    PAGES = {"fft_split_native_f32": 8, "fft_stockham_f32_dual": 4, "fft_combined": 6, "fft_real_combined": 8,
             "fft_real_f32_dual": 6}
    REAL_OFFSET = 0
    IMAG_OFFSET = 32768

    def __init__(self, module="fft_split_native_f32", device=0):
        if module not in self.PAGES:
            raise ValueError(module)
        lib = C.lib()
        rc = lib.wfb_require_b200(device)
        if rc:
            raise C.WatFFTError(rc)
        self.module, self.device = module, device
        self.memory = HostMemory(self.PAGES[module] * 65536)
        self._plans = {}
        self._f64 = module in ("fft_combined", "fft_real_combined")

    # -- plan cache: (kind, layout, n) -> Plan without buffers of its own (it runs on `memory`)
    def _plan(self, kind, layout, n):
        key = (kind, layout, n)
        p = self._plans.get(key)
        if p is None:
            p = self._plans[key] = Plan(kind, C.F64 if self._f64 else C.F32, layout, n, 1, self.device, C.PLAN_NO_HOST_BUFFERS)
        return p

    def _mem(self, dtype, offset, count):
        return self.memory.view(dtype, offset, count)

    def _run(self, plan, direction, off0, off1=None):
        base = self.memory.ptr
        planes = (base + off0, base + off1 if off1 is not None else None)
        plan.exec_host(direction, planes, planes)          # in place, synchronous
        return plan

    # split module
    def precompute_twiddles_split(self, n):
        self._plan(C.C2C, C.SPLIT, n)

    def precompute_rfft_twiddles_split(self, n):
        self._plan(C.R2C, C.INTERLEAVED, n)

    def fft_split(self, n):
        self._run(self._plan(C.C2C, C.SPLIT, n), C.FORWARD, self.REAL_OFFSET, self.IMAG_OFFSET)

    def ifft_split(self, n):
        self._run(self._plan(C.C2C, C.SPLIT, n), C.INVERSE, self.REAL_OFFSET, self.IMAG_OFFSET)

    def rfft_split(self, n):
        self._run(self._plan(C.R2C, C.INTERLEAVED, n), C.FORWARD, 0)

    def irfft_split(self, n):
        self._run(self._plan(C.R2C, C.INTERLEAVED, n), C.INVERSE, 0)

    # interleaved complex modules
    def precompute_twiddles(self, n):
        self._plan(C.C2C, C.INTERLEAVED, n)

    def fft(self, n):
        self._run(self._plan(C.C2C, C.INTERLEAVED, n), C.FORWARD, 0)

    def ifft(self, n):
        self._run(self._plan(C.C2C, C.INTERLEAVED, n), C.INVERSE, 0)

    # real modules (f64: fft_real_combined, f32: fft_real_f32_dual)
    def precompute_rfft_twiddles(self, n):
        self._plan(C.R2C, C.INTERLEAVED, n)

    def rfft(self, n):
        self._run(self._plan(C.R2C, C.INTERLEAVED, n), C.FORWARD, 0)

    def irfft(self, n):          # f64: extension (the reference module has no such export, index.js:145-147)
        self._run(self._plan(C.R2C, C.INTERLEAVED, n), C.INVERSE, 0)

    def dispose(self):
        for p in self._plans.values():
            p.destroy()
        self._plans.clear()
        if self.memory is not None:
            self.memory.free()
            self.memory = None


SplitExportsFacade = ModuleExports          # (round-1 name)


# ---- low-level instance factories (index.js:28-58; async there, synchronous here) -----------
def createFFTInstance(device=0):
    """raw instance for complex FFT (f64): memory, precompute_twiddles, fft, ifft (index.js:28-31)"""
    return ModuleExports("fft_combined", device)


def createFFTf32Instance(device=0):
    """raw instance for complex FFT (f32, interleaved) (index.js:37-40)"""
    return ModuleExports("fft_stockham_f32_dual", device)


def createRFFTInstance(device=0):
    """raw instance for real FFT (f64) (index.js:46-49)"""
    return ModuleExports("fft_real_combined", device)


def createRFFTf32Instance(device=0):
    """raw instance for real FFT (f32) (index.js:55-58)"""
    return ModuleExports("fft_real_f32_dual", device)


def createFFTf32SplitInstance(device=0):
    """raw instance of the split-format module (the reference exposes it only to its tests and benchmarks)"""
    return ModuleExports("fft_split_native_f32", device)


class _Context:
    """Shared shape: {size, batch, exports, plan, forward, inverse, dispose}.
    batch = 1: `exports` is the instance the context runs on (the reference's construction, index.js:69-91).
    batch > 1: the context owns a plan with batch-sized pinned buffers; `exports` is a module instance of its own,
    created on first use (its `memory` is one transform wide, so it cannot be the batch buffers)."""

    _module = None

    def _init(self, size, batch, device):
        self.size, self.batch, self._device = size, batch, device
        self._exports = ModuleExports(self._module, device) if batch == 1 else None
        self._disposed = False

    @property
    def exports(self):
        self._check()
        if self._exports is None:
            self._exports = ModuleExports(self._module, self._device)
        return self._exports

    def _check(self):
        if self._disposed:
            raise RuntimeError("watfft_b200: context used after dispose()")

    def dispose(self):                      # idempotent
        if self._disposed:
            return
        self._disposed = True
        self._drop_views()
        if self._exports is not None:
            self._exports.dispose()
        if self.batch > 1 and self.plan is not None:
            self.plan.destroy()
        self.plan = None


class _ComplexContext(_Context):
    """createFFT / createFFTf32 shape: {size, exports, getInputBuffer, getOutputBuffer, forward, inverse}."""

    def __init__(self, size, precision, batch, device):
        self._module = "fft_combined" if precision == C.F64 else "fft_stockham_f32_dual"
        dtype = np.float64 if precision == C.F64 else np.float32
        lo, hi = ctypes.c_int(), ctypes.c_int()
        C.check(C.lib().wfb_size_range(C.C2C, precision, C.INTERLEAVED, lo, hi))
        if size < lo.value or size > hi.value or size & (size - 1):
            raise C.WatFFTError(C.ERR_BAD_SIZE)
        self._init(size, batch, device)
        if batch == 1:
            ex = self._exports
            ex.precompute_twiddles(size)
            self.plan = ex._plan(C.C2C, C.INTERLEAVED, size)
            self._buf = ex.memory.view(dtype, 0, 2 * size)
            self._fwd, self._inv = (lambda: ex.fft(size)), (lambda: ex.ifft(size))
        else:
            self.plan = Plan(C.C2C, precision, C.INTERLEAVED, size, batch, device)
            self._buf = self.plan.host(0)           # batch * 2 * size values
            self._fwd, self._inv = (lambda: self.plan.exec(C.FORWARD)), (lambda: self.plan.exec(C.INVERSE))

    def getInputBuffer(self):
        self._check()
        return self._buf

    def getOutputBuffer(self):                  # same bytes: in-place contract (index.js:78-83)
        self._check()
        return self._buf

    def forward(self):
        self._check()
        self._fwd()

    def inverse(self):
        self._check()
        self._inv()

    def _drop_views(self):
        self._buf = None


class _SplitContext(_Context):
    """Split-format f32 context (the reference exposes this only as raw exports)."""

    _module = "fft_split_native_f32"

    def __init__(self, size, batch, device):
        if size < 4 or size > 8192 or size & (size - 1):
            raise C.WatFFTError(C.ERR_BAD_SIZE)
        self._init(size, batch, device)
        if batch == 1:
            ex = self._exports
            ex.precompute_twiddles_split(size)
            self.plan = ex._plan(C.C2C, C.SPLIT, size)
            self._re = ex.memory.view(np.float32, ex.REAL_OFFSET, size)
            self._im = ex.memory.view(np.float32, ex.IMAG_OFFSET, size)
            self._fwd, self._inv = (lambda: ex.fft_split(size)), (lambda: ex.ifft_split(size))
        else:
            self.plan = Plan(C.C2C, C.F32, C.SPLIT, size, batch, device)
            self._re, self._im = self.plan.host(0), self.plan.host(1)
            self._fwd, self._inv = (lambda: self.plan.exec(C.FORWARD)), (lambda: self.plan.exec(C.INVERSE))

    def getRealBuffer(self):
        self._check()
        return self._re

    def getImagBuffer(self):
        self._check()
        return self._im

    def getInputBuffer(self):
        self._check()
        return self._re, self._im

    def getOutputBuffer(self):
        self._check()
        return self._re, self._im

    def forward(self):
        self._check()
        self._fwd()

    def inverse(self):
        self._check()
        self._inv()

    def _drop_views(self):
        self._re = self._im = None


class _RealContext(_Context):
    """createRFFT / createRFFTf32 shape: input = size reals, output = (size/2+1)*2 values.

    forward() reads getInputBuffer() and writes getOutputBuffer(); inverse() reads
    getOutputBuffer() (the spectrum) and writes getInputBuffer().  At batch = 1 both views start
    at the same address, exactly like the reference's views over memory offset 0."""

    def __init__(self, size, precision, batch, device):
        self._module = "fft_real_combined" if precision == C.F64 else "fft_real_f32_dual"
        dtype = np.float64 if precision == C.F64 else np.float32
        if size < 8 or size > 16384 or size & (size - 1):
            raise C.WatFFTError(C.ERR_BAD_SIZE)
        self._init(size, batch, device)
        if batch == 1:
            ex = self._exports
            ex.precompute_rfft_twiddles(size)
            self.plan = ex._plan(C.R2C, C.INTERLEAVED, size)
            self._time = ex.memory.view(dtype, 0, size)
            self._spec = ex.memory.view(dtype, 0, size + 2)
            self._fwd, self._inv = (lambda: ex.rfft(size)), (lambda: ex.irfft(size))
        else:
            self.plan = Plan(C.R2C, precision, C.INTERLEAVED, size, batch, device)
            self._time, self._spec = self.plan.host(C.BUF_TIME), self.plan.host(C.BUF_SPECTRUM)
            self._fwd, self._inv = (lambda: self.plan.exec(C.FORWARD)), (lambda: self.plan.exec(C.INVERSE))

    def getInputBuffer(self):
        self._check()
        return self._time

    def getOutputBuffer(self):
        self._check()
        return self._spec

    def forward(self):
        self._check()
        self._fwd()

    def inverse(self):
        self._check()
        self._inv()

    def _drop_views(self):
        self._time = self._spec = None


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
