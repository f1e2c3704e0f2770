"""Batched spectrogram / STFT: the reference's batched caller, as one GPU launch.

Mirror of `generateSpectrogram({samples, sampleRate, fftContext, hopSize, windowType, zeroPadding,
gain, range})` in playground/src/spectrogram.js:270-372: same arguments (the FFT context is
replaced by `fftSize`), same returned fields, same numerics to f32 tolerance -- but the per-frame
JavaScript loop (slice, window, zero-pad, copy, run, magnitude, dB) is a single kernel over all
frames (csrc/wfb_kernels.cuh k_stft)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _cabi as C
from .contexts import _view


class Spectrogram:
    """Reusable plan: fixed signal length / FFT size / hop / window."""

    def __init__(self, num_samples, fftSize, hopSize, windowType="hann", zeroPadding=1, gain=0.0, range=80.0,
                 mode="db", device=0, flags=0):
        self._lib = C.lib()
        if windowType not in C.WINDOWS:
            windowType = "hann"              # the JS falls back to hann for unknown names (:31)
        err = ctypes.c_int(0)
        self._mode = C.STFT_COMPLEX if mode == "complex" else C.STFT_DB
        self._p = self._lib.wfb_stft_create(int(fftSize), int(zeroPadding), int(hopSize), C.WINDOWS[windowType],
                                            int(num_samples), self._mode, float(gain), float(range), int(device),
                                            int(flags), ctypes.byref(err))
        if not self._p:
            raise C.WatFFTError(err.value)
        self.fftSize, self.hopSize, self.windowSize = fftSize, hopSize, fftSize // zeroPadding
        self.numFrames = self._lib.wfb_stft_frames(self._p)
        self.numBins = self._lib.wfb_stft_bins(self._p)
        self.num_samples = num_samples

    def getInputBuffer(self):
        return _view(self._lib.wfb_stft_host_samples(self._p), 4 * self.num_samples, np.float32)

    def getOutputBuffer(self):
        out = _view(self._lib.wfb_stft_host_output(self._p), self._lib.wfb_stft_output_bytes(self._p), np.float32)
        shape = (self.numFrames, self.numBins, 2) if self._mode == C.STFT_COMPLEX else (self.numFrames, self.numBins)
        return out.reshape(shape)

    def run(self):
        C.check(self._lib.wfb_stft_exec(self._p, C.EXEC_DEFAULT))

    def run_device(self, d_samples, d_out, stream=None):
        C.check(self._lib.wfb_stft_exec_device(self._p, ctypes.c_void_p(d_samples), ctypes.c_void_p(d_out),
                                               ctypes.c_void_p(stream) if stream else None))

    def algorithmic_bytes(self):
        return self._lib.wfb_stft_algorithmic_bytes(self._p)

    def dispose(self):
        if getattr(self, "_p", None):
            self._lib.wfb_stft_destroy(self._p)
            self._p = None

    def __del__(self):
        try:
            self.dispose()
        except Exception:
            pass


def generateSpectrogram(samples, sampleRate, fftSize, hopSize, windowType="hann", zeroPadding=1, gain=0.0, range=80.0,
                        device=0):
    """Same result object as the reference's generateSpectrogram (spectrogram.js:362-372)."""
    samples = np.ascontiguousarray(samples, np.float32)
    window = fftSize // zeroPadding
    if (len(samples) - window) // hopSize + 1 <= 0 or len(samples) < window:
        raise ValueError("Audio too short for the given FFT size")          # spectrogram.js:299-301
    sp = Spectrogram(len(samples), fftSize, hopSize, windowType, zeroPadding, gain, range, device=device)
    sp.getInputBuffer()[:] = samples
    sp.run()
    data = sp.getOutputBuffer().copy().ravel()
    out = {"data": data, "numFrames": sp.numFrames, "numBins": sp.numBins, "fftSize": fftSize,
           "windowSize": window, "hopSize": hopSize, "sampleRate": sampleRate}
    sp.dispose()
    return out
