"""The reference's 16-bit wav writer for Tacotron-side audio (`tacotron/datasets/audio.py`: dc_notch_filter :17-23,
save_wav :25-34), the function `tacotron_synthesize.py:112` saves its preview wav with.  Host-side numpy / scipy: the file is
written once per utterance, after the samples have left the GPU.

What the chain does to a float waveform, in order:
  1. speex's DC notch: the biquad  H(z) = r (1 - z^-1)^2 / (1 - 2 r z^-1 + d z^-2),  r = 0.982,  d = r^2 + 0.7 (1 - r)^2;
  2. peak-normalise to 0.999;
  3. soften the dynamics: sign(x) |x|^0.95 (the `f1` gain of :29 cancels in step 4);
  4. scale the peak to 32767 (never by more than 32767 / 0.01) and truncate towards zero to int16.

Here the preview wav is vocoded by WaveRNN instead of Griffin-Lim (feature inversion is out of scope); the writer is the same.
"""
from __future__ import annotations

import numpy as np

NOTCH_RADIUS = 0.982          # audio.py:19


def dc_notch_coefficients(radius=NOTCH_RADIUS):
    """(b, a) of the notch biquad (audio.py:19-22)."""
    r = float(radius)
    b = r * np.array([1.0, -2.0, 1.0])
    a = np.array([1.0, -2.0 * r, r * r + 0.7 * (1.0 - r) ** 2])
    return b, a


def dc_notch_filter(wav):
    """Remove DC / sub-sonic drift; float64 out, same length (direct-form IIR with zero initial state like lfilter, :23)."""
    from scipy import signal
    b, a = dc_notch_coefficients()
    return signal.lfilter(b, a, np.asarray(wav, dtype=np.float64))


def to_int16(wav):
    """Steps 1-4 of the module docstring -> int16 samples (what save_wav writes)."""
    y = dc_notch_filter(wav)
    peak = np.abs(y).max() if y.size else 0.0
    if not peak > 0.0:                                    # silence: the reference would divide 0 by 0; write zeros
        return np.zeros(y.shape, dtype=np.int16)
    y = y / peak * 0.999
    gain = 0.5 * 32767 / max(0.01, float(np.abs(y).max()))          # :29 (kept so the roundings match the reference's)
    y = gain * (np.sign(y) * np.power(np.abs(y), 0.95))
    y = y * (32767 / max(0.01, float(np.abs(y).max())))
    return y.astype(np.int16)


def save_wav(wav, path, sr):
    """Write `wav` (float, any scale) as a 16-bit PCM file at `sr` Hz the way the reference's `audio.save_wav` does."""
    from scipy.io import wavfile
    wavfile.write(str(path), int(sr), to_int16(wav))
