"""Pin of the forward-attention step WITH the inference window against the reference's own code.

tests/golden/taco_window_from_reference.npz: 60 loop states of a windowed run (train.txt line 241, shipped checkpoint) and what
the reference's unmodified `ForwardLocationSensitiveAttention.__call__` (forward_attention.py:119-231) returns for them when its
statements are executed on numpy arrays (oracle/ref_harness_taco_attention.py, oracle/make_golden_taco_window.py).
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import tacotron_oracle as to
from taco_common import real_taco_weights


@pytest.fixture(scope='module')
def ctx():
    w = real_taco_weights()
    if w is None:
        pytest.skip('shipped Tacotron checkpoint not available on this box')
    return w, np.load(os.path.join(GOLDEN, 'taco_window_from_reference.npz'))


def test_window_fixture_covers_every_branch(ctx):
    """The fixture exercises every branch of the window (hold, advance, forced advance after 10 steps, early hold)."""
    _, z = ctx
    # coverage of the fixture: every branch of the window occurs
    held = (z['ref_max_att'] == z['max_att'])
    advanced = (z['ref_max_att'] == z['max_att'] + 1)
    assert held.any() and advanced.any() and (held | advanced).all()                   # never more than one token per step
    pushed = advanced & (z['pos_rec'] == 9)                                            # 10 steps on one token -> forced advance
    assert pushed.any() and (z['ref_pos_rec'][pushed] == 1).all()
    assert ((z['pos_rec'] < 5) & (z['max_att'] >= 2)).any()                            # the early `short_mask` hold is exercised
    assert (z['ref_pos_rec'][held & ~pushed] == z['pos_rec'][held & ~pushed] + 1).all()


def test_windowed_attention_step_matches_reference_statements(ctx):
    w, z = ctx
    memory = z['memory']
    keys = (memory @ w['memory_layer/kernel']).astype(np.float32)
    Wq = w['decoder/Location_Sensitive_Attention/query_layer/kernel']
    v_a = w['decoder/Location_Sensitive_Attention/attention_variable_projection']
    b_a = w['decoder/Location_Sensitive_Attention/attention_bias']
    for i in range(len(z['step'])):
        alpha, cum, mu = z['alpha'][i], z['cum'][i], np.float32(z['mu'][i])
        h2 = z['query'][i][None]
        # the attention half of decoder_step, driven by the stored query (the LSTM half is pinned in test_tacotron_step_pins.py)
        q = (h2 @ Wq).astype(np.float32)
        energy = (np.tanh(keys + q + to.location_features(w, cum) + b_a, dtype=np.float32) * v_a).sum(axis=1).astype(np.float32)
        e = np.exp(energy - energy.max(), dtype=np.float32)
        a = (e / e.sum(dtype=np.float32)).astype(np.float32)
        shift = np.concatenate([[np.float32(0)], alpha[:-1]]).astype(np.float32)
        raw = (((np.float32(1) - mu) * alpha + mu * shift + np.float32(1e-10)) * a).astype(np.float32)
        al, new_max, new_pos = to.attention_window(raw, int(z['max_att'][i]), int(z['pos_rec'][i]))
        al = (al / al.sum(dtype=np.float32)).astype(np.float32)
        ctxv = (al[None, :] @ memory).astype(np.float32)
        new_mu = to._sigmoid((np.concatenate([ctxv, h2], axis=1) @ w['decoder/dense/kernel'] + w['decoder/dense/bias']).astype(np.float32))[0, 0]
        assert (new_max, new_pos) == (int(z['ref_max_att'][i]), int(z['ref_pos_rec'][i])), int(z['step'][i])
        np.testing.assert_allclose(al, z['ref_alignments'][i], rtol=0, atol=1e-6)
        np.testing.assert_allclose((cum + a).astype(np.float32), z['ref_cum'][i], rtol=0, atol=1e-6)
        np.testing.assert_allclose(ctxv[0], z['ref_context'][i], rtol=0, atol=1e-5)
        assert abs(float(new_mu) - float(z['ref_mu'][i])) <= 1e-6


def test_stop_rule_matches_reference_helper(ctx):
    """decode() stops when `stop > 0.5`: that is tf.round (half-to-even) == 1 in the reference's TacoTestHelper.next_inputs."""
    _, z = ctx
    ours = z['helper_stop_probability'] > 0.5
    np.testing.assert_array_equal(ours, z['helper_finished'])
    assert not ours[3] and ours[4]                                  # exactly 0.5 does not stop; the next float above does


def test_zoneout_cell_matches_reference_class(ctx):
    """ZoneoutLSTMCell at inference (modules.py:114-142, the reference's real class around a stand-in LSTMCell): the cell
    OUTPUT is the un-zoned h, the carried state is 0.9 * new + 0.1 * prev."""
    w, z = ctx
    k1 = w['decoder/decoder_LSTM/multi_rnn_cell/cell_0/decoder_LSTM_1/kernel']
    b1 = w['decoder/decoder_LSTM/multi_rnn_cell/cell_0/decoder_LSTM_1/bias']
    out, c, h = to.zoneout_lstm(z['zoneout_x'], z['zoneout_c'], z['zoneout_h'], k1, b1)
    np.testing.assert_allclose(out, z['zoneout_ref_output'], rtol=0, atol=1e-6)
    np.testing.assert_allclose(c, z['zoneout_ref_c'], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(h, z['zoneout_ref_h'], rtol=1e-6, atol=1e-6)
    assert np.abs(out - h).max() > 1e-3                              # output and carried h really differ
