"""Structural pins of the Tacotron-2 restatement against the reference's own serialized graph (CPU suite).

tests/golden/taco_graph_facts.json was read out of `logs-Tacotron-2/taco_pretrained/tacotron_model.ckpt-206500.meta` (the
MetaGraphDef the reference saved next to its checkpoint) by oracle/make_golden_taco_graph.py, without TensorFlow.  The
numeric pins are test_tacotron_step_pins.py / test_tacotron_encpost_pins.py (the same graph executed in numpy); what is pinned HERE is every
assumption it makes about arithmetic that lives inside TensorFlow: gate order, forget bias, zoneout, dropout scaling,
batch-norm epsilon, and the wiring of the forward-attention step.
"""
import json
import os

import numpy as np
import pytest

from oracle import tacotron_oracle as to

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def facts():
    with open(os.path.join(HERE, 'golden', 'taco_graph_facts.json')) as f:
        return json.load(f)


def test_lstm_cells_match_the_graph(facts):
    A = to.GRAPH_ASSUMPTIONS
    for key in ('decoder_lstm_1', 'decoder_lstm_2', 'encoder_lstm_fw', 'encoder_lstm_bw'):
        g = facts[key]
        assert g['gate_order'] == A['lstm_gate_order'], key
        assert g['forget_bias'] == A['lstm_forget_bias'], key
        assert g['kernel_input_order'] == A['lstm_kernel_input_order'], key
        assert g['unzoned_h_is_cell_output'] is A['cell_output_is_unzoned_h'], key
        # training form in the graph: prev + (1 - z) * dropout(new - prev, rate z); its expectation is the inference form
        # (1 - z) * new + z * prev that ZoneoutLSTMCell uses when is_training is False (modules.py:137-138)
        assert g['zoneout_cell'] == pytest.approx(A['zoneout'], rel=1e-6)
        assert g['zoneout_hidden'] == pytest.approx(A['zoneout'], rel=1e-6)
        assert g['zoneout_keep_scale'] == pytest.approx(1.0 - A['zoneout'], rel=1e-6)
    assert facts['decoder_lstm_2_reads'] == 'unzoned_h1'


def test_lstm_cell_code_follows_the_table():
    """The table is not decoration: lstm_cell really splits i,j,f,o and adds the forget bias to f."""
    rs = np.random.RandomState(0)
    n, nin = 4, 3
    x, c, h = rs.randn(1, nin).astype(np.float32), rs.randn(1, n).astype(np.float32), rs.randn(1, n).astype(np.float32)
    k, b = rs.randn(nin + n, 4 * n).astype(np.float32), rs.randn(4 * n).astype(np.float32)
    z = np.concatenate([x, h], -1).astype(np.float64) @ k + b
    i, j, f, o = (z[:, q * n:(q + 1) * n] for q in range(4))
    sg = lambda v: 1 / (1 + np.exp(-v))
    c_ref = sg(f + 1.0) * c + sg(i) * np.tanh(j)
    h_ref = sg(o) * np.tanh(c_ref)
    new_c, new_h = to.lstm_cell(x, c, h, k, b)
    np.testing.assert_allclose(new_c, c_ref, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(new_h, h_ref, rtol=1e-5, atol=1e-6)
    out, zc, zh = to.zoneout_lstm(x, c, h, k, b)
    np.testing.assert_array_equal(out, new_h)                                  # the un-zoned h is what the next layer sees
    np.testing.assert_allclose(zc, 0.9 * new_c + 0.1 * c, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(zh, 0.9 * new_h + 0.1 * h, rtol=1e-6, atol=1e-7)


def test_prenet_and_batch_norm_constants(facts):
    A = to.GRAPH_ASSUMPTIONS
    p = facts['prenet']
    assert p['dropout_rate'] == [A['prenet_dropout_rate']] * 2
    assert p['activation'] == ['Relu', 'Relu'] and all(p['keep_if_uniform_ge_rate']) and all(p['scale_is_one_over_keep'])
    assert p['second_layer_reads_dropped_first']
    assert facts['batch_norm_layers'] == 8                                    # 3 encoder convolutions + 5 postnet convolutions
    assert facts['batch_norm_epsilon'] == [pytest.approx(A['batch_norm_epsilon'], rel=1e-6)]
    lo, hi = A['output_clip']
    assert facts['output_clip'] == [pytest.approx(lo, rel=1e-6), hi, pytest.approx(lo, rel=1e-6), hi]


def test_attention_step_wiring(facts):
    A = to.GRAPH_ASSUMPTIONS
    a = facts['attention']
    assert a['energy_ops'] == ['Add', 'Add', 'Add', 'Tanh', 'Mul', 'Sum']    # v . tanh(keys + query + location + bias)
    assert a['softmax_input_masked_with'] == '-inf'
    assert a['cumulated_alignments_add'] == A['attention_cumulates']
    assert a['cumulated_state_feeds_location_conv'] and a['cumulated_state_updated_with_sum']
    assert A['attention_location_input'] == 'cumulated'
    # ((1 - mu) * alpha + mu * shift(alpha) + eps) * softmax, normalised by its sum
    assert a['forward_term'] == ['Sub', 'Mul', 'Mul', 'Add', 'Add', 'Mul', 'Sum', 'RealDiv']
    assert a['forward_epsilon'] == pytest.approx(A['attention_forward_epsilon'], rel=1e-6)
    assert a['shifted_alpha_is_zero_then_alpha_without_last'] and a['alpha_state_is_normalised']
    assert a['context_is_alpha_times_memory'] and a['mu_state_is_sigmoid_dense']
    assert a['query_is'] == 'h2'
    assert a['mu_dense_input'] == A['attention_mu_dense_input']
    assert a['projection_input'] == A['projection_input']
    assert a['lstm_input'] == A['lstm_input']
    assert a['frame_and_stop_share_input']


def test_conv_blocks(facts):
    """conv('same') + bias -> activation -> batch norm, as conv_block() in the oracle (modules.py:379-391)."""
    cb = facts['conv_blocks']
    for i in (1, 2, 3):
        assert cb[f'conv_layer_{i}_encoder_convolutions']['before_batch_norm'] == 'Relu'
    for i in (1, 2, 3, 4):
        assert cb[f'conv_layer_{i}_postnet_convolutions']['before_batch_norm'] == 'Tanh'
    assert cb['conv_layer_5_postnet_convolutions']['before_batch_norm'] == 'linear'
    for k, v in cb.items():
        assert v['padding'] == 'SAME', k
        if k != 'location_features_convolution':
            assert v['bias_before_activation'] and v['scale_is_gamma_rsqrt_var_plus_eps'], k


def test_encoder_bilstm_wiring(facts):
    """memory = concat([forward outputs, backward outputs re-reversed]) as in oracle.encoder (modules.py:207-217)."""
    e = facts['encoder_bilstm']
    assert e['bw_reads_reversed_conv_output'] and e['bw_output_is_reversed_back']
    assert e['memory_concat'] == ['fw', 'bw_reversed']
