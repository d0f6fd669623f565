"""The slice of the reference's `wavernn/utils/dsp.py` that the generation path touches:
label <-> float (:8-15), mu-law (:92-103) and wav writing (:22-23).  Feature extraction
(STFT / mel / Griffin-Lim, :26-117) is preprocessing and out of scope.
"""
from __future__ import annotations

import math

import numpy as np

from . import hparams as hp


def label_2_float(x, bits):
    return 2 * x / (2 ** bits - 1.) - 1.


def float_2_label(x, bits):
    assert abs(x).max() <= 1.0
    x = (x + 1.) * (2 ** bits - 1) / 2
    return x.clip(0, 2 ** bits - 1)


def encode_mu_law(x, mu):
    mu = mu - 1
    fx = np.sign(x) * np.log(1 + mu * np.abs(x)) / np.log(1 + mu)
    return np.floor((fx + 1) / 2 * mu + 0.5)


def decode_mu_law(y, mu, from_labels=True):
    if from_labels:
        y = label_2_float(y, math.log2(mu))
    mu = mu - 1
    return np.sign(y) / mu * ((1 + mu) ** np.abs(y) - 1)


def save_wav(x, path, sample_rate=None):
    """float32 PCM wav at hp.sample_rate.  The reference calls librosa.output.write_wav (dsp.py:23), which no
    longer exists in any current librosa; scipy writes the same float32 samples."""
    from scipy.io import wavfile
    sr = sample_rate if sample_rate is not None else hp.sample_rate
    wavfile.write(str(path), int(sr), np.asarray(x).astype(np.float32))
