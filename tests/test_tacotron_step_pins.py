"""NUMERIC pin of the Tacotron-2 decoder-step restatement (CPU suite).

tests/golden/taco_step_from_graph.npz holds, for five loop states of a real sentence (train.txt line 241, steps 0, 1, 7,
60, 200), the values obtained by EXECUTING the reference's own serialized decoder-step graph (the `CustomDecoderStep`
body inside `tacotron_model.ckpt-206500.meta`) on the shipped checkpoint with a numpy op interpreter
(oracle/tf_graph_eval.py, oracle/make_golden_taco_step.py -- no TensorFlow).  oracle.tacotron_oracle.decoder_step must
reproduce every intermediate: prenet output, both LSTM cells, query, energies, softmax, cumulated alignments, forward
recursion, context, the transition probability mu, the frame and the stop logit.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import tacotron_oracle as to
from taco_common import real_taco_weights

KEYS = ('prenet', 'new_c1', 'new_h1', 'new_c2', 'new_h2', 'query', 'energy', 'softmax', 'cum', 'forward_raw', 'alignments',
        'context', 'mu', 'frame', 'stop_logit')


def test_decoder_step_reproduces_the_serialized_reference_graph():
    w = real_taco_weights()
    if w is None:
        pytest.skip('shipped Tacotron checkpoint not available on this box')
    z = np.load(os.path.join(GOLDEN, 'taco_step_from_graph.npz'))
    memory = z['memory']
    np.testing.assert_allclose(to.encoder(w, z['ids']), memory, rtol=0, atol=1e-5)     # the fixture's memory is the oracle's own
    keys = (memory @ w['memory_layer/kernel']).astype(np.float32)
    for s in z['steps']:
        st = {k[len(f's{s}_state_'):]: z[k] for k in z.files if k.startswith(f's{s}_state_')}
        st['mu'] = np.float32(st['mu'])
        out, _ = to.decoder_step(w, memory, keys, z[f's{s}_x'], z[f's{s}_m'], st)
        for k in KEYS:
            g = z[f's{s}_graph_{k}']
            o = np.asarray(out[k], dtype=np.float32).reshape(g.shape)
            tol = 1e-6 * max(1.0, float(np.abs(g).max()))
            assert np.abs(o - g).max() <= tol, (int(s), k, float(np.abs(o - g).max()))
        # sanity of the fixture itself: a proper distribution that has moved along the sentence by step 200
        al = z[f's{s}_graph_alignments'][0]
        assert abs(al.sum() - 1.0) < 1e-5 and al.min() >= 0
    assert z['s200_graph_alignments'][0].argmax() > z['s7_graph_alignments'][0].argmax()


def test_step_fixture_is_sensitive_to_the_assumptions(monkeypatch):
    """The comparison is live: a wrong forget bias (0 instead of 1) is off by O(1), not by rounding."""
    w = real_taco_weights()
    if w is None:
        pytest.skip('shipped Tacotron checkpoint not available on this box')
    z = np.load(os.path.join(GOLDEN, 'taco_step_from_graph.npz'))
    memory = z['memory']
    keys = (memory @ w['memory_layer/kernel']).astype(np.float32)
    s = 60
    st = {k[len(f's{s}_state_'):]: z[k] for k in z.files if k.startswith(f's{s}_state_')}
    st['mu'] = np.float32(st['mu'])
    orig = to.lstm_cell
    monkeypatch.setattr(to, 'lstm_cell', lambda x, c, h, k, b, forget_bias=0.0: orig(x, c, h, k, b, forget_bias=0.0))
    out, _ = to.decoder_step(w, memory, keys, z[f's{s}_x'], z[f's{s}_m'], st)
    assert np.abs(out['frame'] - z[f's{s}_graph_frame']).max() > 1e-2
