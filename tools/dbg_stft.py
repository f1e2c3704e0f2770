import os, sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import watfft_b200 as wf, oracle as om
om.build(); orc = om.Oracle()
def sig(n, seed=0):
    t = np.arange(n) / 16000.0
    rng = np.random.default_rng(seed)
    return (0.6*np.sin(2*np.pi*(200+1500*t)*t) + 0.3*np.sin(2*np.pi*3000*t) + 0.05*rng.uniform(-1,1,n)).astype(np.float32)
for n_fft in (2048, 4096):
    hop = n_fft // 4
    for frames in (9, 3001):
        x = sig((frames-1)*hop + n_fft + 3, n_fft)
        res = {}
        for s in ("1", "0"):
            os.environ["WFB_STFT_SPAN"] = s
            got = wf.generateSpectrogram(x, 16000.0, n_fft, hop, "hann", 1, gain=-3.0, range=75.0)
            res[s] = got["data"].reshape(got["numFrames"], got["numBins"])
        for fr in (0, 1, frames//2, frames-1):
            ref = om.spectrogram_reference(x[fr*hop: fr*hop+n_fft], n_fft, hop, "hann", 1, gain=-3.0, range_db=75.0, rfft=orc.rfft_split_f32)[0]
            for s in ("1", "0"):
                e = np.abs(res[s][fr] - ref)
                bad = np.nonzero(e > 2e-4)[0]
                print(n_fft, frames, "span" if s == "1" else "direct", "frame", fr, "maxerr %.3g" % e.max(), "bad bins", bad[:12], len(bad))
