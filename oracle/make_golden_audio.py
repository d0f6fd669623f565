#!/usr/bin/env python
"""Golden vectors for the DC-notch 16-bit wav writer, produced by the reference's OWN statements
(tacotron/datasets/audio.py: dc_notch_filter :17-23, save_wav :25-34) -- container only (needs /root/reference).

    python oracle/make_golden_audio.py        ->  tests/golden/audio_save_wav_from_reference.npz

The reference module imports librosa and tensorflow at the top (neither is installed here, neither is used by these two
functions), so the two function definitions are lifted out of the file's syntax tree and compiled unmodified against
numpy / scipy.signal / scipy.io.wavfile; the wav files they write are read back as the expected int16 samples.
Test infrastructure: nothing in the product imports this.
"""
import ast
import os
import tempfile

import numpy as np
from scipy import signal
from scipy.io import wavfile

REF = os.environ.get('B200TTS_REFERENCE', '/root/reference')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(REF, 'tacotron', 'datasets', 'audio.py')


def reference_functions():
    tree = ast.parse(open(SRC, encoding='utf-8').read(), SRC)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ('dc_notch_filter', 'save_wav')]
    assert [n.name for n in keep] == ['dc_notch_filter', 'save_wav']
    ns = {'np': np, 'signal': signal, 'wavfile': wavfile}
    exec(compile(ast.Module(body=keep, type_ignores=[]), SRC, 'exec'), ns)
    return ns['dc_notch_filter'], ns['save_wav']


def cases():
    rs = np.random.RandomState(7)
    t = np.arange(6000) / 22050.0
    out = {
        'tone_dc': (0.3 * np.sin(2 * np.pi * 220 * t) + 0.2).astype(np.float32),                # a tone riding on a DC offset
        'noise': rs.uniform(-1, 1, 4000).astype(np.float32),
        'quiet': (1e-3 * rs.standard_normal(3000)).astype(np.float64),                          # far below full scale
        'loud': (40.0 * rs.standard_normal(2500)).astype(np.float64),                           # far above full scale
        'mulaw_like': np.tanh(3 * rs.standard_normal(5000)).astype(np.float32),
        'short': np.array([0.5, -0.25, 0.125], dtype=np.float64),
    }
    return out


def main():
    notch, save = reference_functions()
    blob = {}
    with tempfile.TemporaryDirectory() as d:
        for name, x in cases().items():
            p = os.path.join(d, name + '.wav')
            save(x.copy(), p, sr=22050)
            sr, y = wavfile.read(p)
            assert sr == 22050 and y.dtype == np.int16
            blob['in_' + name] = x
            blob['pcm_' + name] = y
            blob['notch_' + name] = np.asarray(notch(x.copy()), dtype=np.float64)
    out = os.path.join(ROOT, 'tests', 'golden', 'audio_save_wav_from_reference.npz')
    np.savez_compressed(out, **blob)
    print(out, sorted(blob))


if __name__ == '__main__':
    main()
