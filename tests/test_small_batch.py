"""The small-batch path of wfb_exec: payloads up to WFB_OPT_MAPPED_MAX_BYTES skip the copies --
one direct-load kernel reads and writes the pinned, device-mapped host buffers (the reference's own call shape is
batch = 1, index.js:84-89).  Parity of that path for every transform, equality of its results with the copy path's
contract, the option plumbing, and the two argument-robustness fixes of ADVICE r1."""
import ctypes

import numpy as np
import pytest

import oracle as om
from conftest import f32_bound, f64_bound, rel_err


def test_option_symbols_exist(wf):
    lib = wf._cabi.lib()
    for name in ("wfb_plan_set_option", "wfb_plan_get_option", "wfb_plan_last_path", "wfb_pcie_probe"):
        assert getattr(lib, name)
    # argument checks are host logic: no device needed
    assert lib.wfb_plan_set_option(None, 0, 1) == wf._cabi.ERR_BAD_ARG
    assert lib.wfb_plan_get_option(None, 0) == -1
    assert lib.wfb_pcie_probe(0, 16, 1, (ctypes.c_double * 4)()) == wf._cabi.ERR_BAD_ARG


@pytest.mark.gpu
@pytest.mark.parametrize("n", [4, 16, 64, 256, 1024, 4096, 8192])
@pytest.mark.parametrize("batch", [1, 3])
def test_zero_copy_c2c_split(wf, oracle, n, batch):
    C = wf._cabi
    rng = np.random.default_rng(n + batch)
    re = rng.uniform(-1, 1, (batch, n)).astype(np.float32)
    im = rng.uniform(-1, 1, (batch, n)).astype(np.float32)
    ctx = wf.createFFTf32Split(n, batch=batch)
    ctx.plan.set_option(C.OPT_MAPPED_MAX_BYTES, 1 << 22)
    for inverse in (False, True):
        ctx.getRealBuffer()[:] = re.ravel()
        ctx.getImagBuffer()[:] = im.ravel()
        ctx.inverse() if inverse else ctx.forward()
        assert ctx.plan.last_path() == C.PATH_MAPPED
        gr, gi = ctx.getRealBuffer().reshape(batch, n), ctx.getImagBuffer().reshape(batch, n)
        for r in range(batch):
            o = np.r_[oracle.fft_split_f32(re[r], im[r], inverse)]
            assert rel_err(np.r_[gr[r], gi[r]], o, np.r_[re[r], im[r]]) <= f32_bound(n), (n, r, inverse)
    # the copy path on the same plan gives a result within the same bound (a different kernel, same contract)
    ctx.plan.set_option(C.OPT_MAPPED_MAX_BYTES, 0)
    ctx.getRealBuffer()[:] = re.ravel()
    ctx.getImagBuffer()[:] = im.ravel()
    ctx.forward()
    assert ctx.plan.last_path() == C.PATH_STAGED
    o = np.r_[oracle.fft_split_f32(re[0], im[0])]
    assert rel_err(np.r_[ctx.getRealBuffer()[:n], ctx.getImagBuffer()[:n]], o, np.r_[re[0], im[0]]) <= f32_bound(n)
    ctx.dispose()


@pytest.mark.gpu
@pytest.mark.parametrize("n", [8, 32, 128, 1024, 4096, 16384])
@pytest.mark.parametrize("batch", [1, 2])
def test_zero_copy_real_f32(wf, oracle, n, batch):
    """batch = 1: input and output views are the SAME bytes (index.js:136-141): the kernel runs in place on the host buffer."""
    C = wf._cabi
    rng = np.random.default_rng(7 * n + batch)
    x = rng.uniform(-1, 1, (batch, n)).astype(np.float32)
    ctx = wf.createRFFTf32(n, batch=batch)
    ctx.plan.set_option(C.OPT_MAPPED_MAX_BYTES, 1 << 22)
    ctx.getInputBuffer()[:] = x.ravel()
    ctx.forward()
    assert ctx.plan.last_path() == C.PATH_MAPPED
    g = ctx.getOutputBuffer().reshape(batch, n + 2).copy()
    for r in range(batch):
        assert rel_err(g[r], oracle.rfft_split_f32(x[r]), x[r]) <= f32_bound(n), (n, r)
        assert g[r, 1] == 0.0 and g[r, n + 1] == 0.0
    ctx.inverse()
    assert ctx.plan.last_path() == C.PATH_MAPPED
    t = ctx.getInputBuffer().reshape(batch, n)
    assert np.max(np.abs(t - x)) < 1e-5 * np.sqrt(n)
    ctx.dispose()


@pytest.mark.gpu
@pytest.mark.parametrize("n", [4, 16, 256, 1024, 4096])
def test_zero_copy_interleaved_and_f64(wf, oracle, n):
    C = wf._cabi
    rng = np.random.default_rng(n)
    z = rng.uniform(-1, 1, 2 * n)
    c = wf.createFFTf32(n)
    c.getInputBuffer()[:] = z.astype(np.float32)
    c.forward()
    assert c.plan.last_path() == C.PATH_MAPPED
    assert rel_err(c.getOutputBuffer(), oracle.fft_interleaved_f32(z.astype(np.float32)), z) <= f32_bound(n)
    c.dispose()
    c = wf.createFFT(n)
    c.getInputBuffer()[:] = z
    c.forward()
    assert c.plan.last_path() == C.PATH_MAPPED
    assert rel_err(c.getOutputBuffer(), oracle.fft_f64(z), z) <= f64_bound(n)
    c.inverse()
    # (the reference's f64 twiddles are a Taylor series good to ~5e-11 -- BASELINE.md section 1 -- so fft -> ifft is not the
    #  identity to machine precision; sizes 4 and 16 use exact constants)
    assert np.max(np.abs(c.getOutputBuffer() - z)) < (1e-13 if n <= 16 else 1e-9)
    c.dispose()
    if n >= 8:
        x = rng.uniform(-1, 1, n)
        c = wf.createRFFT(n)
        c.getInputBuffer()[:] = x
        c.forward()
        assert c.plan.last_path() == C.PATH_MAPPED
        assert rel_err(c.getOutputBuffer(), oracle.rfft_f64(x), x) <= f64_bound(n)
        c.dispose()


@pytest.mark.gpu
def test_default_threshold_and_set_variant(wf):
    C = wf._cabi
    small = wf.createFFTf32Split(1024)                      # 16 KiB payload: zero-copy by default
    small.forward()
    assert small.plan.last_path() == C.PATH_MAPPED
    assert small.plan.get_option(C.OPT_MAPPED_MAX_BYTES) == 256 << 10
    small.plan.set_variant(0)                               # an explicit kernel choice switches the detour off
    small.forward()
    assert small.plan.last_path() == C.PATH_STAGED
    small.dispose()
    big = wf.createFFTf32Split(1024, batch=4096)            # 64 MiB payload: copies
    big.forward()
    assert big.plan.last_path() in (C.PATH_STAGED, C.PATH_PIPELINED)
    big.plan.set_option(C.OPT_STAGE_CHUNK_BYTES, 1 << 20)
    big.plan.set_option(C.OPT_STAGE_STREAMS, 2)
    big.getRealBuffer()[:] = 1.0
    big.getImagBuffer()[:] = 0.0
    big.forward()
    assert big.plan.last_path() == C.PATH_PIPELINED
    assert np.all(big.getRealBuffer().reshape(4096, 1024)[:, 0] == 1024.0)
    with pytest.raises(wf.WatFFTError):
        big.plan.set_option(C.OPT_STAGE_STREAMS, 99)
    big.dispose()


@pytest.mark.gpu
def test_device_buffers_on_demand(wf, oracle):
    """Small plans allocate their device buffers only when someone asks for them; the device-pointer entry keeps working."""
    C = wf._cabi
    n = 256
    p = wf.Plan(C.C2C, C.F32, C.SPLIT, n, 1)
    d0, d1 = p.device_ptr(0), p.device_ptr(1)
    assert d0 and d1
    rng = np.random.default_rng(0)
    re, im = rng.uniform(-1, 1, n).astype(np.float32), rng.uniform(-1, 1, n).astype(np.float32)
    p.host(0)[:] = re
    p.host(1)[:] = im
    p.exec(C.FORWARD, C.STAGE_H2D | C.SYNC)                 # H2D only: result stays on the device (copy path)
    assert p.last_path() == C.PATH_STAGED
    p.exec_device(C.INVERSE, (d0, d1), (d0, d1))
    p.sync()
    p.host(0)[:] = 0
    p.exec(C.INVERSE, C.SYNC)                               # (no staging flags: runs on the device buffers again)
    p.destroy()


@pytest.mark.gpu
def test_pcie_probe(wf):
    lib = wf._cabi.lib()
    g = (ctypes.c_double * 4)()
    assert lib.wfb_pcie_probe(0, 64 << 20, 3, g) == 0
    assert all(1.0 < x < 200.0 for x in g), list(g)


@pytest.mark.gpu
def test_pcie_probe_phases(wf):
    """The phase-wise probe (what bench.py drives with a barrier across the ranks in front of every phase)."""
    lib = wf._cabi.lib()
    h = ctypes.c_void_p()
    assert lib.wfb_pcie_probe_open(0, 32 << 20, ctypes.byref(h)) == 0
    sec = (ctypes.c_double * 2)()
    for directions in (1, 2, 3):
        assert lib.wfb_pcie_probe_run(h, directions, 3, sec) == 0
        for bit, t in ((1, sec[0]), (2, sec[1])):
            if directions & bit:
                assert 1.0 < 3 * (32 << 20) / t / 1e9 < 200.0, (directions, list(sec))
            else:
                assert t == 0.0
    assert lib.wfb_pcie_probe_run(h, 0, 3, sec) != 0
    lib.wfb_pcie_probe_close(h)


@pytest.mark.gpu
@pytest.mark.parametrize("batch", [2000, 2001, 1731])
@pytest.mark.parametrize("ramp", [1, 0])
def test_staging_pipeline_chunk_schedule(wf, oracle, batch, ramp):
    """Every row of a pipelined wfb_exec is transformed exactly once whatever the chunk schedule: ramped ends (short first
    and last chunks), the even-remainder chunk in the middle, an odd batch's last row; c2c split and the (n+2)-wide r2c rows."""
    C = wf._cabi
    n = 64
    rng = np.random.default_rng(batch)
    ctx = wf.createFFTf32Split(n, batch=batch)
    ctx.plan.set_option(C.OPT_MAPPED_MAX_BYTES, 0)
    ctx.plan.set_option(C.OPT_STAGE_CHUNK_BYTES, 32 << 10)      # 128 rows per chunk: ramp 16, 32, 64
    ctx.plan.set_option(C.OPT_STAGE_RAMP, ramp)
    assert ctx.plan.get_option(C.OPT_STAGE_RAMP) == ramp
    re = rng.uniform(-1, 1, (batch, n)).astype(np.float32)
    im = rng.uniform(-1, 1, (batch, n)).astype(np.float32)
    ctx.getRealBuffer()[:] = re.ravel()
    ctx.getImagBuffer()[:] = im.ravel()
    ctx.forward()
    assert ctx.plan.last_path() == C.PATH_PIPELINED
    got = ctx.getRealBuffer().reshape(batch, n) + 1j * ctx.getImagBuffer().reshape(batch, n)
    want = np.fft.fft(re.astype(np.float64) + 1j * im.astype(np.float64), axis=1)
    assert np.max(np.abs(got - want)) / np.sqrt(2 * n) <= f32_bound(n)
    o_re, o_im = oracle.fft_split_f32(re[-1], im[-1])
    assert rel_err(np.r_[got[-1].real, got[-1].imag], np.r_[o_re, o_im], np.r_[re[-1], im[-1]]) <= f32_bound(n)
    ctx.inverse()
    assert np.max(np.abs(ctx.getRealBuffer().reshape(batch, n) - re)) < 1e-5
    ctx.dispose()
    r = wf.createRFFTf32(n, batch=batch)
    r.plan.set_option(C.OPT_MAPPED_MAX_BYTES, 0)
    r.plan.set_option(C.OPT_STAGE_CHUNK_BYTES, 32 << 10)
    r.plan.set_option(C.OPT_STAGE_RAMP, ramp)
    x = rng.uniform(-1, 1, (batch, n)).astype(np.float32)
    r.getInputBuffer()[:] = x.ravel()
    r.forward()
    assert r.plan.last_path() == C.PATH_PIPELINED
    spec = r.getOutputBuffer().reshape(batch, n + 2)
    want = np.fft.rfft(x.astype(np.float64), axis=1)
    assert np.max(np.abs((spec[:, 0::2] + 1j * spec[:, 1::2]) - want)) / np.sqrt(n) <= f32_bound(n)
    r.inverse()
    assert np.max(np.abs(r.getInputBuffer().reshape(batch, n) - x)) < 1e-5
    r.dispose()


@pytest.mark.gpu
def test_unused_second_plane_pointer_is_ignored(wf, oracle):
    """d_in[1]/d_out[1] are documented as unused for R2C and interleaved C2C; an odd or stale value there must
    neither demote the launch to the unaligned-pointer fallback nor fail."""
    import torch
    C = wf._cabi
    n, b = 1024, 64
    x = torch.rand(b * n, device="cuda")
    spec = torch.empty(b * (n + 2), device="cuda")
    p = wf.Plan(C.R2C, C.F32, 0, n, b, 0, C.PLAN_NO_HOST_BUFFERS | C.PLAN_NO_DEVICE_BUFFERS)
    torch.cuda.synchronize()      # torch filled the inputs on ITS stream; the plan's stream is non-blocking
    p.exec_device(C.FORWARD, (x.data_ptr(), 0x3), (spec.data_ptr(), 0x7))
    torch.cuda.synchronize()
    g = spec.cpu().numpy().reshape(b, n + 2)
    xs = x.cpu().numpy().reshape(b, n)
    assert rel_err(g[5], oracle.rfft_split_f32(xs[5]), xs[5]) <= f32_bound(n)
    p.destroy()


@pytest.mark.gpu
def test_stft_unaligned_sample_pointer(wf, oracle):
    """a caller's float pointer that is only 4-byte aligned (an odd-sample offset into a larger device buffer)
    must take the scalar-load path instead of faulting on a 64-bit load."""
    import torch
    n_fft, hop, ns = 256, 64, 4096
    buf = torch.rand(ns + 1, device="cuda") * 2 - 1
    sp = wf.Spectrogram(ns, n_fft, hop, "hann", mode="complex", flags=wf._cabi.PLAN_NO_HOST_BUFFERS | wf._cabi.PLAN_NO_DEVICE_BUFFERS)
    torch.cuda.synchronize()      # torch filled the inputs on ITS stream; the plan's stream is non-blocking
    out = torch.empty(sp.numFrames * sp.numBins * 2, device="cuda")
    sp.run_device(buf.data_ptr() + 4, out.data_ptr())
    torch.cuda.synchronize()
    x = buf.cpu().numpy()[1:]
    ref = om.spectrogram_reference(x, n_fft, hop, "hann", rfft=oracle.rfft_split_f32, complex_out=True)
    g = out.cpu().numpy().reshape(ref.shape)
    assert np.max(np.abs(g - ref)) <= f32_bound(n_fft) * np.linalg.norm(x[:n_fft]) * 4
    sp.dispose()
