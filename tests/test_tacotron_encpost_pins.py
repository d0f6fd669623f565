"""NUMERIC pins of the Tacotron-2 encoder and postnet restatements (CPU suite).

tests/golden/taco_encpost_from_graph.npz: outputs of the reference's own serialized encoder-convolution, encoder-LSTM-step
and postnet sub-graphs (`tacotron_model.ckpt-206500.meta`), executed on the shipped weights by the numpy op interpreter
oracle/tf_graph_eval.py with the inference substitutions described in oracle/make_golden_taco_encpost.py (moving
statistics for the batch norms, dropout as identity).  Sentence: train.txt line 241.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import tacotron_oracle as to
from taco_common import real_taco_weights


def _close(o, g, tol=2e-6):
    g = np.asarray(g)
    o = np.asarray(o, dtype=np.float32).reshape(g.shape)
    err = float(np.abs(o - g).max() / max(1.0, float(np.abs(g).max())))
    assert err <= tol, err


@pytest.fixture(scope='module')
def ctx():
    w = real_taco_weights()
    if w is None:
        pytest.skip('shipped Tacotron checkpoint not available on this box')
    return w, np.load(os.path.join(GOLDEN, 'taco_encpost_from_graph.npz'))


def test_encoder_convolutions(ctx):
    w, z = ctx
    convs = to.encoder_convs(w, z['ids'])
    for i in (1, 2, 3):
        _close(convs[i - 1], z[f'graph_enc_conv{i}'])


def test_encoder_lstm_steps(ctx):
    w, z = ctx
    for d in ('fw', 'bw'):
        k = w[f'encoder_LSTM/bidirectional_rnn/{d}/encoder_{d}_LSTM/kernel']
        b = w[f'encoder_LSTM/bidirectional_rnn/{d}/encoder_{d}_LSTM/bias']
        for n in (0, 1, 25):
            c, h = to.lstm_cell(z[f'lstm_{d}_{n}_x'], z[f'lstm_{d}_{n}_c'], z[f'lstm_{d}_{n}_h'], k, b)
            _close(c, z[f'graph_lstm_{d}_{n}_new_c'])
            _close(h, z[f'graph_lstm_{d}_{n}_new_h'])


def test_postnet(ctx):
    w, z = ctx
    dec = z['dec_frames']
    _close(to.postnet(w, dec), z['graph_mel'])
    assert float(np.abs(z['graph_mel']).max()) <= 4.1 + 1e-6            # the clip is part of the executed graph
