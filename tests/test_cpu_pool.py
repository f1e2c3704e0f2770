"""The CPU baseline harness (oracle/watref_threads.c): the persistent pool must produce exactly what a single call of the
transpiled reference produces, for every transform bench.py times with it, and the spectrogram loop must equal the
numpy restatement of playground/src/spectrogram.js."""
import numpy as np
import pytest

import oracle as om


@pytest.fixture(scope="module")
def pool(watref):
    p = om.WatRefPool(watref, 3)
    yield p
    p.close()


def test_pool_matches_single_calls(watref, pool):
    rng = np.random.default_rng(0)
    n, b = 256, 41
    re, im = rng.uniform(-1, 1, (b, n)).astype(np.float32), rng.uniform(-1, 1, (b, n)).astype(np.float32)
    o0, o1 = np.zeros_like(re), np.zeros_like(im)
    pool.prepare("fft_split_native_f32", "precompute_twiddles_split", n)
    t = pool.run("fft_split_native_f32", "fft_split", n, re, 0, im, om.SPLIT_IMAG_OFFSET, o0, 0, o1, om.SPLIT_IMAG_OFFSET)
    assert t > 0
    for r in (0, 13, 14, 27, 40):                      # rows on both sides of the workers' range boundaries
        a, c = watref.fft_split_f32(re[r], im[r])
        assert np.array_equal(a, o0[r]) and np.array_equal(c, o1[r])
    x = rng.uniform(-1, 1, (b, n))
    out = np.zeros((b, n + 2))
    pool.prepare("fft_real_combined", "precompute_rfft_twiddles", n)
    pool.run("fft_real_combined", "rfft", n, x, 0, out0=out, src0=0)
    assert np.array_equal(out[17], watref.rfft_f64(x[17]))
    # switching size re-targets every worker's instance
    n2 = 64
    x2 = rng.uniform(-1, 1, (b, n2)).astype(np.float32)
    out2 = np.zeros((b, n2 + 2), np.float32)
    pool.prepare("fft_split_native_f32", "precompute_rfft_twiddles_split", n2)
    pool.run("fft_split_native_f32", "rfft_split", n2, x2, 0, out0=out2, src0=0, reps=2)
    assert np.array_equal(out2[40], watref.rfft_split_f32(x2[40]))


def test_pool_spectrogram_matches_restatement(watref, pool):
    rng = np.random.default_rng(1)
    n, hop = 256, 64
    x = rng.uniform(-1, 1, 5000).astype(np.float32)
    w = om.window_function("hann", n)
    frames = (len(x) - n) // hop + 1
    out = np.zeros((frames, n // 2 + 1), np.float32)
    pool.prepare("fft_split_native_f32", "precompute_rfft_twiddles_split", n)
    pool.run_stft(x, n, hop, w, -6.0, 70.0, out)
    ref = om.spectrogram_reference(x, n, hop, "hann", gain=-6.0, range_db=70.0, rfft=watref.rfft_split_f32)
    assert np.max(np.abs(out - ref)) < 1e-6


def test_one_shot_wrapper(watref):
    rng = np.random.default_rng(2)
    re, im = rng.uniform(-1, 1, (9, 64)).astype(np.float32), rng.uniform(-1, 1, (9, 64)).astype(np.float32)
    o0, o1 = np.zeros_like(re), np.zeros_like(im)
    watref.run_batch("fft_split_native_f32", "precompute_twiddles_split", "ifft_split", 64, re, 0, im, om.SPLIT_IMAG_OFFSET,
                     o0, 0, o1, om.SPLIT_IMAG_OFFSET, threads=2)
    a, c = watref.fft_split_f32(re[8], im[8], inverse=True)
    assert np.array_equal(a, o0[8]) and np.array_equal(c, o1[8])
