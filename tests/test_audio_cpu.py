"""DC-notch 16-bit wav writer (reference tacotron/datasets/audio.py:17-34) against vectors written by the
reference's own statements (oracle/make_golden_audio.py).  Integer output: bit-exact."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from tacotronv2_wavernn_chinese_b200.tacotron import audio

G = np.load(os.path.join(GOLDEN, 'audio_save_wav_from_reference.npz'))
NAMES = sorted(k[3:] for k in G.files if k.startswith('in_'))


def test_fixture_covers_the_cases():
    assert NAMES == ['loud', 'mulaw_like', 'noise', 'quiet', 'short', 'tone_dc']


@pytest.mark.parametrize('name', NAMES)
def test_notch_matches_reference(name):
    y = audio.dc_notch_filter(G['in_' + name])
    assert y.dtype == np.float64 and np.array_equal(y, G['notch_' + name])


@pytest.mark.parametrize('name', NAMES)
def test_save_wav_is_bit_exact(name, tmp_path):
    from scipy.io import wavfile
    x = G['in_' + name]
    keep = x.copy()
    p = tmp_path / (name + '.wav')
    audio.save_wav(x, p, sr=22050)
    assert np.array_equal(x, keep)                       # the caller's buffer is not modified
    sr, y = wavfile.read(p)
    assert sr == 22050 and y.dtype == np.int16 and np.array_equal(y, G['pcm_' + name])
    assert np.abs(y.astype(np.int32)).max() in (32766, 32767)        # peak-normalised whatever the input scale


def test_notch_removes_dc():
    y = audio.dc_notch_filter(np.full(20000, 0.7))
    assert abs(y[-1]) < 1e-6 and abs(y[0] - 0.7 * audio.NOTCH_RADIUS) < 1e-12


def test_silence_writes_zeros(tmp_path):
    from scipy.io import wavfile
    audio.save_wav(np.zeros(100, np.float32), tmp_path / 'z.wav', sr=22050)
    sr, y = wavfile.read(tmp_path / 'z.wav')
    assert y.dtype == np.int16 and not y.any() and y.shape == (100,)


def test_preview_wav_names_and_writes_like_the_reference(tmp_path, monkeypatch):
    """tacotron_synthesize.py --voc_weights: `step-..-mel-pred.npy` -> `step-..-wav-from-mel.wav` (reference :110-112) through
    the DC-notch writer; the vocoder itself is stubbed here (its GPU path is covered by tests/test_cli_gpu.py)."""
    import sys
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    import tacotron_synthesize as ts
    import wavernn_gen
    from scipy.io import wavfile

    seen = {}

    class Stub:
        def load(self, path):
            seen['weights'] = path

        def generate(self, mels, save_path, batched, target, overlap, mu_law, seed=None):
            seen['shape'], seen['save_path'], seen['batched'], seen['seed'] = tuple(mels.shape), save_path, batched, seed
            n = (mels.shape[-1] - 1) * 275
            return 0.25 * np.sin(np.arange(n) * 0.05) + 0.1

    monkeypatch.setattr(wavernn_gen, 'build_model', lambda: Stub())
    mel_path = str(tmp_path / 'step-206500-abc-mel-pred.npy')
    np.save(mel_path, np.random.RandomState(0).uniform(0, 1, (12, 80)).astype(np.float32))     # shorter than the 21-frame minimum
    out = ts.preview_wav(mel_path, 'w.pyt', os.path.join(ROOT, 'wavernn_hparams.py'), 22050, seed=4)
    assert out == str(tmp_path / 'step-206500-abc-wav-from-mel.wav')
    assert seen == dict(weights='w.pyt', shape=(1, 80, 21), save_path=None, batched=False, seed=4)
    sr, y = wavfile.read(out)
    assert sr == 22050 and y.dtype == np.int16 and y.shape == (20 * 275,)
    assert np.array_equal(y, audio.to_int16(0.25 * np.sin(np.arange(20 * 275) * 0.05) + 0.1))
    assert abs(float(y[2000:].astype(np.float64).mean())) < 200         # the 0.1 offset (~9000 counts at this scale) is gone
