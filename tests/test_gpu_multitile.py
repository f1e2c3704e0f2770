"""Persistent-kernel pipelines under load: batches large enough that every resident CTA loops over several tiles
(double-buffered bulk loads, bulk stores draining out of the stage that is refilled next, dynamic tile claims), with a
row count that leaves a ragged last tile.  Rows sampled all over the batch are checked against the CPU oracle, and the
forward -> inverse round trip over the WHOLE batch catches any row that was dropped, duplicated or written twice.

The small-batch variant tests (test_gpu_variants.py) give each CTA at most one tile, so they cannot see refill /
drain hazards; the full-size property tests cover only a few sizes."""
import numpy as np
import pytest

from conftest import f32_bound, f64_bound, rel_err

pytestmark = pytest.mark.gpu

REPEATS = 6          # identical launches must give identical bits: a refill/drain race would not

TARGET_BYTES = 96 << 20        # of input per case: >= 3 tiles of 16-32 KB for each of the <= 1776 resident CTAs
SAMPLES = 96


def _rows(batch, seed):
    rng = np.random.default_rng(seed)
    return sorted({0, 1, 2, 31, 32, 33, batch // 2, batch - 3, batch - 2, batch - 1} | set(rng.integers(0, batch, SAMPLES).tolist()))


def _torch():
    import torch
    return torch, torch.device("cuda:0")


@pytest.mark.parametrize("layout", ["split", "interleaved"])
@pytest.mark.parametrize("n", [16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
def test_c2c_f32_many_tiles(wf, oracle, n, layout):
    torch, dev = _torch()
    C = wf._cabi
    batch = TARGET_BYTES // (8 * n) + 3
    g = torch.Generator(device=dev); g.manual_seed(1000 + n)
    flags = C.PLAN_NO_HOST_BUFFERS | C.PLAN_NO_DEVICE_BUFFERS
    if layout == "split":
        re = torch.rand(batch * n, device=dev, generator=g) * 2 - 1
        im = torch.rand(batch * n, device=dev, generator=g) * 2 - 1
        ore, oim = torch.empty_like(re), torch.empty_like(im)
        plan = wf.Plan(C.C2C, C.F32, C.SPLIT, n, batch, 0, flags)
        torch.cuda.synchronize()      # torch filled the inputs on ITS stream; the plan's stream is non-blocking
        plan.exec_device(C.FORWARD, (re.data_ptr(), im.data_ptr()), (ore.data_ptr(), oim.data_ptr()))
        plan.sync()
        o2r, o2i = torch.empty_like(re), torch.empty_like(im)
        for _ in range(REPEATS):
            plan.exec_device(C.FORWARD, (re.data_ptr(), im.data_ptr()), (o2r.data_ptr(), o2i.data_ptr()))
            plan.sync()
            assert torch.equal(o2r, ore) and torch.equal(o2i, oim), (n, "forward launches differ bitwise")
        del o2r, o2i
        for r in _rows(batch, n):
            a, b = re[r * n:(r + 1) * n].cpu().numpy(), im[r * n:(r + 1) * n].cpu().numpy()
            er, ei = oracle.fft_split_f32(a, b)
            got = np.r_[ore[r * n:(r + 1) * n].cpu().numpy(), oim[r * n:(r + 1) * n].cpu().numpy()]
            assert rel_err(got, np.r_[er, ei], np.r_[a, b]) <= f32_bound(n), (n, r)
        bre, bim = torch.empty_like(re), torch.empty_like(im)
        plan.exec_device(C.INVERSE, (ore.data_ptr(), oim.data_ptr()), (bre.data_ptr(), bim.data_ptr()))
        plan.sync()
        assert float((bre - re).abs().max()) < 1e-4 and float((bim - im).abs().max()) < 1e-4
    else:
        z = torch.rand(batch * 2 * n, device=dev, generator=g) * 2 - 1
        o = torch.empty_like(z)
        plan = wf.Plan(C.C2C, C.F32, C.INTERLEAVED, n, batch, 0, flags)
        torch.cuda.synchronize()      # torch filled the inputs on ITS stream; the plan's stream is non-blocking
        plan.exec_device(C.FORWARD, (z.data_ptr(), None), (o.data_ptr(), None))
        plan.sync()
        for r in _rows(batch, n):
            a = z[r * 2 * n:(r + 1) * 2 * n].cpu().numpy()
            assert rel_err(o[r * 2 * n:(r + 1) * 2 * n].cpu().numpy(), oracle.fft_interleaved_f32(a), a) <= f32_bound(n), (n, r)
        plan.exec_device(C.INVERSE, (o.data_ptr(), None), (o.data_ptr(), None))     # in place
        plan.sync()
        assert float((o - z).abs().max()) < 1e-4
    plan.destroy()



@pytest.mark.parametrize("layout", ["split", "interleaved"])
@pytest.mark.parametrize("n,tag", [(2048, "_rpf"), (4096, "_rpf"), (128, "_ts_g"), (256, "_ts_g"), (512, "_ts_g")])
def test_c2c_f32_alternate_persistent_kernels_many_tiles(wf, oracle, n, tag, layout):
    """Persistent alternates of the default kernel on the same plan: the register-prefetch kernels (k_c2c_rpf: the next
    transform is loaded into a second register set while the current one is computed; results leave through two alternating
    shared-memory tiles) and the grouped row copies of k_c2c_pipe (several rows per bulk copy, thread groups of a warp in
    different copy groups).  Many tiles per CTA, a ragged last tile, both directions.  Same plan and twiddles as the default
    kernel, so the whole batch must match it BITWISE; sampled rows against the oracle; repeated launches identical."""
    torch, dev = _torch()
    C = wf._cabi
    batch = TARGET_BYTES // (8 * n) + 3
    g = torch.Generator(device=dev); g.manual_seed(4000 + n)
    flags = C.PLAN_NO_HOST_BUFFERS | C.PLAN_NO_DEVICE_BUFFERS
    split = layout == "split"
    a0 = torch.rand(batch * n * (1 if split else 2), device=dev, generator=g) * 2 - 1
    a1 = torch.rand(batch * n, device=dev, generator=g) * 2 - 1 if split else None
    ptr = lambda t: t.data_ptr() if t is not None else None
    plan = wf.Plan(C.C2C, C.F32, C.SPLIT if split else C.INTERLEAVED, n, batch, 0, flags)
    torch.cuda.synchronize()      # torch filled the inputs on ITS stream; the plan's stream is non-blocking
    names = plan.variants()
    rpf = [i for i, v in enumerate(names) if tag in v]
    base = next(i for i, v in enumerate(names) if tag not in v)      # the established kernel on the same plan
    assert rpf and ("_ts" in names[base]), names
    for direction in (C.FORWARD, C.INVERSE):
        ref0, ref1 = torch.empty_like(a0), (torch.empty_like(a1) if split else None)
        plan.set_variant(base)
        plan.exec_device(direction, (ptr(a0), ptr(a1)), (ptr(ref0), ptr(ref1)))
        plan.sync()
        for vi in rpf:
            plan.set_variant(vi)
            o0, o1 = torch.empty_like(a0), (torch.empty_like(a1) if split else None)
            for rep in range(3):
                o0.zero_()
                torch.cuda.synchronize()
                plan.exec_device(direction, (ptr(a0), ptr(a1)), (ptr(o0), ptr(o1)))
                plan.sync()
                assert torch.equal(o0, ref0) and (not split or torch.equal(o1, ref1)), (names[vi], direction, rep)
        if direction == C.FORWARD:
            for r in _rows(batch, n)[:24]:
                if split:
                    x, y = a0[r * n:(r + 1) * n].cpu().numpy(), a1[r * n:(r + 1) * n].cpu().numpy()
                    er, ei = oracle.fft_split_f32(x, y)
                    got = np.r_[o0[r * n:(r + 1) * n].cpu().numpy(), o1[r * n:(r + 1) * n].cpu().numpy()]
                    assert rel_err(got, np.r_[er, ei], np.r_[x, y]) <= f32_bound(n), (n, r)
                else:
                    z = a0[2 * r * n:2 * (r + 1) * n].cpu().numpy()
                    assert rel_err(o0[2 * r * n:2 * (r + 1) * n].cpu().numpy(), oracle.fft_interleaved_f32(z), z) <= f32_bound(n), (n, r)
    # in place (the contexts' call shape): the loads of the NEXT row are in flight while this row's results are stored
    plan.set_variant(rpf[0])
    b0, b1 = a0.clone(), (a1.clone() if split else None)
    torch.cuda.synchronize()
    plan.exec_device(C.FORWARD, (ptr(b0), ptr(b1)), (ptr(b0), ptr(b1)))
    plan.exec_device(C.INVERSE, (ptr(b0), ptr(b1)), (ptr(b0), ptr(b1)))
    plan.sync()
    assert float((b0 - a0).abs().max()) < 1e-4
    plan.destroy()


@pytest.mark.parametrize("n", [32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384])
def test_real_f32_many_tiles(wf, oracle, n):
    torch, dev = _torch()
    C = wf._cabi
    batch = TARGET_BYTES // (4 * n) + 3          # odd tail: the last r2c/c2r tile cannot move as one 16-byte-sized bulk copy
    g = torch.Generator(device=dev); g.manual_seed(2000 + n)
    x = torch.rand(batch * n, device=dev, generator=g) * 2 - 1
    spec = torch.empty(batch * (n + 2), device=dev)
    back = torch.empty_like(x)
    plan = wf.Plan(C.R2C, C.F32, 0, n, batch, 0, C.PLAN_NO_HOST_BUFFERS | C.PLAN_NO_DEVICE_BUFFERS)
    torch.cuda.synchronize()      # torch filled the inputs on ITS stream; the plan's stream is non-blocking
    plan.exec_device(C.FORWARD, (x.data_ptr(), None), (spec.data_ptr(), None))
    plan.exec_device(C.INVERSE, (spec.data_ptr(), None), (back.data_ptr(), None))
    plan.sync()
    s2, b2 = torch.empty_like(spec), torch.empty_like(back)
    for _ in range(REPEATS):
        plan.exec_device(C.FORWARD, (x.data_ptr(), None), (s2.data_ptr(), None))
        plan.exec_device(C.INVERSE, (s2.data_ptr(), None), (b2.data_ptr(), None))
        plan.sync()
        assert torch.equal(s2, spec) and torch.equal(b2, back), (n, "launches differ bitwise")
    del s2, b2
    for r in _rows(batch, n):
        a = x[r * n:(r + 1) * n].cpu().numpy()
        s = spec[r * (n + 2):(r + 1) * (n + 2)].cpu().numpy()
        so = oracle.rfft_split_f32(a)
        assert rel_err(s, so, a) <= f32_bound(n), (n, r)
        assert rel_err(back[r * n:(r + 1) * n].cpu().numpy(), oracle.irfft_split_f32(s), s) <= f32_bound(n), (n, r)
    assert float((back - x).abs().max()) < 1e-4
    plan.destroy()


@pytest.mark.parametrize("n", [64, 256, 512, 1024, 2048, 4096])
def test_c2c_f64_many_tiles(wf, oracle, n):
    torch, dev = _torch()
    C = wf._cabi
    batch = TARGET_BYTES // (16 * n) + 3
    g = torch.Generator(device=dev); g.manual_seed(3000 + n)
    z = torch.rand(batch * 2 * n, device=dev, generator=g, dtype=torch.float64) * 2 - 1
    o = torch.empty_like(z)
    plan = wf.Plan(C.C2C, C.F64, C.INTERLEAVED, n, batch, 0, C.PLAN_NO_HOST_BUFFERS | C.PLAN_NO_DEVICE_BUFFERS)
    torch.cuda.synchronize()      # torch filled the inputs on ITS stream; the plan's stream is non-blocking
    plan.exec_device(C.FORWARD, (z.data_ptr(), None), (o.data_ptr(), None))
    plan.sync()
    o2 = torch.empty_like(o)
    for _ in range(REPEATS):
        plan.exec_device(C.FORWARD, (z.data_ptr(), None), (o2.data_ptr(), None))
        plan.sync()
        assert torch.equal(o2, o), (n, "forward launches differ bitwise")
    del o2
    for r in _rows(batch, n)[::3]:
        a = z[r * 2 * n:(r + 1) * 2 * n].cpu().numpy()
        assert rel_err(o[r * 2 * n:(r + 1) * 2 * n].cpu().numpy(), oracle.fft_f64(a), a) <= f64_bound(n), (n, r)
    plan.exec_device(C.INVERSE, (o.data_ptr(), None), (o.data_ptr(), None))
    plan.sync()
    # the reference's f64 twiddles carry a 6.5e-11 Taylor truncation error: fft and ifft are inverse to ~1e-10
    assert float((o - z).abs().max()) < 2e-9
    plan.destroy()


@pytest.mark.parametrize("n", [128, 256, 512, 1024, 2048, 4096])
def test_real_f64_many_tiles(wf, oracle, n):
    torch, dev = _torch()
    C = wf._cabi
    batch = TARGET_BYTES // (8 * n) + 3
    g = torch.Generator(device=dev); g.manual_seed(4000 + n)
    x = torch.rand(batch * n, device=dev, generator=g, dtype=torch.float64) * 2 - 1
    spec = torch.empty(batch * (n + 2), device=dev, dtype=torch.float64)
    back = torch.empty_like(x)
    plan = wf.Plan(C.R2C, C.F64, 0, n, batch, 0, C.PLAN_NO_HOST_BUFFERS | C.PLAN_NO_DEVICE_BUFFERS)
    torch.cuda.synchronize()      # torch filled the inputs on ITS stream; the plan's stream is non-blocking
    plan.exec_device(C.FORWARD, (x.data_ptr(), None), (spec.data_ptr(), None))
    plan.exec_device(C.INVERSE, (spec.data_ptr(), None), (back.data_ptr(), None))      # f64 c2r: extension, round trip only
    plan.sync()
    s2, b2 = torch.empty_like(spec), torch.empty_like(back)
    for _ in range(REPEATS):
        plan.exec_device(C.FORWARD, (x.data_ptr(), None), (s2.data_ptr(), None))
        plan.exec_device(C.INVERSE, (s2.data_ptr(), None), (b2.data_ptr(), None))
        plan.sync()
        assert torch.equal(s2, spec) and torch.equal(b2, back), (n, "launches differ bitwise")
    del s2, b2
    for r in _rows(batch, n)[::3]:
        a = x[r * n:(r + 1) * n].cpu().numpy()
        assert rel_err(spec[r * (n + 2):(r + 1) * (n + 2)].cpu().numpy(), oracle.rfft_f64(a), a) <= f64_bound(n), (n, r)
    assert float((back - x).abs().max()) < 2e-9
    plan.destroy()
