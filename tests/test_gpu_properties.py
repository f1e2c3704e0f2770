"""Size-independent properties of the DFT, checked on the GPU path through the public contexts (the C ABI) for every
supported size: linearity, the shift theorem, impulses, Parseval, Hermitian symmetry, and the consistency of the real
transforms with the complex one.  These do not need the oracle: they hold for the transform itself, so they catch an
error the oracle and the kernels would share (a mis-read stage formula), and they run at every size in milliseconds.

Tolerances are the parity bound of the precision (relative to the 2-norm of the input), times a small constant where a
property combines several transforms."""
import numpy as np
import pytest

from backends import GpuBackend
from conftest import f32_bound, f64_bound

pytestmark = pytest.mark.gpu

C2C_SIZES = [4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192]
REAL_SIZES_F32 = [32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384]
REAL_SIZES_F64 = [8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096]
B = 19            # rows per call: odd, spans several thread-per-row groups


@pytest.fixture(scope="module")
def gpu(wf):
    return GpuBackend(wf)


def _cplx(rng, b, n):
    return rng.uniform(-1, 1, (b, n)) + 1j * rng.uniform(-1, 1, (b, n))


def _fft(gpu, z, kind, inverse=False):
    """complex rows -> complex rows through one of the three complex paths"""
    if kind == "split":
        re, im = gpu.fft_split_f32(z.real.astype(np.float32), z.imag.astype(np.float32), inverse)
        return re.astype(np.float64) + 1j * im
    il = np.empty((z.shape[0], 2 * z.shape[1]), np.float64)
    il[:, 0::2], il[:, 1::2] = z.real, z.imag
    out = gpu.fft_interleaved_f32(il, inverse) if kind == "il" else gpu.fft_f64(il, inverse)
    out = out.astype(np.float64)
    return out[:, 0::2] + 1j * out[:, 1::2]


def _bound(kind, n):
    return f64_bound(n) if kind == "f64" else f32_bound(n)


def _err(a, b, ref):
    return float(np.max(np.abs(a - b), axis=-1).max() / np.linalg.norm(ref, axis=-1).min())


@pytest.mark.parametrize("kind", ["split", "il", "f64"])
@pytest.mark.parametrize("n", C2C_SIZES)
def test_c2c_linearity_shift_parseval(gpu, n, kind):
    rng = np.random.default_rng(n)
    x, y = _cplx(rng, B, n), _cplx(rng, B, n)
    if kind != "f64":                                  # inputs exactly representable in the transform's precision
        x = x.astype(np.complex64).astype(np.complex128); y = y.astype(np.complex64).astype(np.complex128)
    a, b = 0.75, -1.25                                 # exact in binary: a*x + b*y rounds once
    fx, fy = _fft(gpu, x, kind), _fft(gpu, y, kind)
    fz = _fft(gpu, a * x + b * y, kind)
    tol = _bound(kind, n)
    assert _err(fz, a * fx + b * fy, x) <= 4 * tol
    # agreement with the definition (numpy's f64 FFT as the O(N log N) stand-in for the DFT sum); the reference's Taylor
    # twiddles put the f64 modules ~1e-10 from the true DFT, the f32 ones ~1e-6
    true = np.fft.fft(x, axis=-1)
    assert _err(fx, true, x) <= (2e-9 if kind == "f64" else 4 * tol)
    # shift theorem: rotating the input by s multiplies bin k by exp(-2 pi i k s / n)
    s = 3 % n
    fs = _fft(gpu, np.roll(x, s, axis=-1), kind)
    k = np.arange(n)
    assert _err(fs, fx * np.exp(-2j * np.pi * k * s / n), x) <= (4e-9 if kind == "f64" else 6 * tol)
    # Parseval
    assert np.allclose(np.sum(np.abs(fx) ** 2, axis=-1), n * np.sum(np.abs(x) ** 2, axis=-1), rtol=1e-5 if kind != "f64" else 1e-9)
    # inverse: normalised by 1/n
    back = _fft(gpu, fx.astype(np.complex64).astype(np.complex128) if kind != "f64" else fx, kind, inverse=True)
    assert _err(back, x, x) <= (2e-9 if kind == "f64" else 1e-5)


@pytest.mark.parametrize("kind", ["split", "il", "f64"])
@pytest.mark.parametrize("n", [4, 16, 64, 256, 1024, 4096])
def test_c2c_impulses_and_constants(gpu, n, kind):
    # row r: unit impulse at position p_r -> bin k = exp(-2 pi i k p / n); last row: all ones -> n at bin 0
    pos = [0, 1, n // 2, n - 1]
    x = np.zeros((len(pos) + 1, n), np.complex128)
    for r, p_ in enumerate(pos):
        x[r, p_] = 1.0
    x[-1, :] = 1.0
    f = _fft(gpu, x, kind)
    k = np.arange(n)
    # an impulse (2-norm 1) reads the twiddle products of every pass directly: the f32 parity bound applies as it stands
    tol = 2e-9 if kind == "f64" else f32_bound(n)
    for r, p_ in enumerate(pos):
        assert np.max(np.abs(f[r] - np.exp(-2j * np.pi * k * p_ / n))) <= tol, (n, p_)
    assert abs(f[-1, 0] - n) <= tol * n and np.max(np.abs(f[-1, 1:])) <= tol * n


@pytest.mark.parametrize("prec,sizes", [("f32", REAL_SIZES_F32), ("f64", REAL_SIZES_F64)])
def test_real_transforms_agree_with_complex(gpu, prec, sizes):
    for n in sizes:
        rng = np.random.default_rng(n + 7)
        x = rng.uniform(-1, 1, (B, n))
        if prec == "f32":
            x = x.astype(np.float32).astype(np.float64)
            spec = gpu.rfft_f32(x).astype(np.float64)
        else:
            spec = gpu.rfft_f64(x)
        s = spec[:, 0::2] + 1j * spec[:, 1::2]                      # n/2 + 1 bins
        assert s.shape[1] == n // 2 + 1
        assert np.all(spec[:, 1] == 0.0) and np.all(spec[:, n + 1] == 0.0)      # DC and Nyquist are stored as exact reals
        true = np.fft.rfft(x, axis=-1)
        tol = 2e-9 if prec == "f64" else 4 * f32_bound(n)
        assert _err(s, true, x) <= tol, n
        if n <= 8192:                                               # the complex path on the same rows: Hermitian, same bins
            f = _fft(gpu, x.astype(np.complex128), "f64" if prec == "f64" else "split")
            assert _err(f[:, : n // 2 + 1], s, x) <= tol
            assert _err(f[:, 1:][:, ::-1].conj()[:, : n // 2 - 1], f[:, 1: n // 2], x) <= tol
        if prec == "f32":                                           # f32 inverse real transform: irfft(rfft(x)) = x
            back = gpu.irfft_f32(spec.astype(np.float32)).astype(np.float64)
            assert np.max(np.abs(back - x)) <= 1e-5, n
