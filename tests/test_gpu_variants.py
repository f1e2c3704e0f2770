"""Every compiled kernel variant (not only each plan's default) against the CPU oracle, with odd
batch sizes so ragged tiles and the packed-lane tail are exercised."""
import math

import numpy as np
import pytest

from conftest import f32_bound, f64_bound, rel_err

pytestmark = pytest.mark.gpu


def _batch(n):
    return 37 if n <= 1024 else 7


@pytest.mark.parametrize("n", [4, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
def test_c2c_f32_all_variants(wf, oracle, n):
    C = wf._cabi
    b = _batch(n)
    rng = np.random.default_rng(n)
    re = rng.uniform(-1, 1, (b, n)).astype(np.float32)
    im = rng.uniform(-1, 1, (b, n)).astype(np.float32)
    il = np.empty((b, 2 * n), np.float32)
    il[:, 0::2], il[:, 1::2] = re, im
    rows = sorted({0, 1, b // 2, b - 2, b - 1})
    for layout in (C.SPLIT, C.INTERLEAVED):
        plan = wf.Plan(C.C2C, C.F32, layout, n, b)
        names = plan.variants()
        assert len(names) >= 1
        for vi, vn in enumerate(names):
            plan.set_variant(vi)
            for inv in (False, True):
                if layout == C.SPLIT:
                    plan.host(0)[:] = re.ravel()
                    plan.host(1)[:] = im.ravel()
                else:
                    plan.host(0)[:] = il.ravel()
                plan.exec(C.INVERSE if inv else C.FORWARD)
                for r in rows:
                    if layout == C.SPLIT:
                        o = np.r_[oracle.fft_split_f32(re[r], im[r], inv)]
                        g = np.r_[plan.host(0).reshape(b, n)[r], plan.host(1).reshape(b, n)[r]]
                        err = rel_err(g, o, np.r_[re[r], im[r]])
                    else:
                        err = rel_err(plan.host(0).reshape(b, 2 * n)[r], oracle.fft_interleaved_f32(il[r], inv), il[r])
                    assert err <= f32_bound(n), (vn, layout, inv, r, err)
        plan.destroy()


@pytest.mark.parametrize("n", [8, 16, 32, 64, 128, 256, 1024, 4096, 16384])
def test_real_f32_all_variants(wf, oracle, n):
    C = wf._cabi
    b = _batch(n)
    rng = np.random.default_rng(n + 1)
    x = rng.uniform(-1, 1, (b, n)).astype(np.float32)
    spec = np.stack([oracle.rfft_split_f32(x[r]) for r in range(b)])
    plan = wf.Plan(C.R2C, C.F32, 0, n, b)
    for vi, vn in enumerate(plan.variants()):
        plan.set_variant(vi)
        plan.host(C.BUF_TIME)[:] = x.ravel()
        plan.exec(C.FORWARD)
        g = plan.host(C.BUF_SPECTRUM).reshape(b, n + 2)
        for r in range(b):
            assert rel_err(g[r], spec[r], x[r]) <= f32_bound(n), (vn, "r2c", r)
        plan.host(C.BUF_SPECTRUM)[:] = spec.ravel()
        plan.exec(C.INVERSE)
        t = plan.host(C.BUF_TIME).reshape(b, n)
        for r in range(b):
            assert rel_err(t[r], oracle.irfft_split_f32(spec[r]), spec[r]) <= f32_bound(n), (vn, "c2r", r)
    plan.destroy()


@pytest.mark.parametrize("n", [16, 256, 512, 2048, 4096])
def test_f64_all_variants(wf, oracle, n):
    C = wf._cabi
    b = 7
    rng = np.random.default_rng(n + 2)
    d = rng.uniform(-1, 1, (b, 2 * n))
    plan = wf.Plan(C.C2C, C.F64, C.INTERLEAVED, n, b)
    for vi, vn in enumerate(plan.variants()):
        plan.set_variant(vi)
        for inv in (False, True):
            plan.host(0)[:] = d.ravel()
            plan.exec(C.INVERSE if inv else C.FORWARD)
            g = plan.host(0).reshape(b, 2 * n)
            for r in range(b):
                assert rel_err(g[r], oracle.fft_f64(d[r], inv), d[r]) <= f64_bound(n), (vn, inv, r)
    plan.destroy()
    x = rng.uniform(-1, 1, (b, n))
    plan = wf.Plan(C.R2C, C.F64, 0, n, b)
    for vi, vn in enumerate(plan.variants()):
        plan.set_variant(vi)
        plan.host(C.BUF_TIME)[:] = x.ravel()
        plan.exec(C.FORWARD)
        g = plan.host(C.BUF_SPECTRUM).reshape(b, n + 2).copy()
        for r in range(b):
            assert rel_err(g[r], oracle.rfft_f64(x[r]), x[r]) <= f64_bound(n), (vn, r)
        plan.exec(C.INVERSE)
        assert np.max(np.abs(plan.host(C.BUF_TIME).reshape(b, n) - x)) < 1e-9
    plan.destroy()
