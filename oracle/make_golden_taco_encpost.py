"""Numeric golden vectors for the Tacotron-2 encoder and postnet, from the reference's own serialized graph.

TEST INFRASTRUCTURE; run HERE (needs /root/reference).  Same method as make_golden_taco_step.py: the sub-graphs of
`tacotron_model.ckpt-206500.meta` are executed op by op in numpy (oracle/tf_graph_eval.py) on the shipped weights.
The serialized graph is the TRAINING graph, so three things are fed to obtain the inference computation the reference's
`Synthesizer` runs (is_training = False):
  * batch normalisation: the `moments/Squeeze{,_1}` nodes (batch mean / variance) are fed with the layer's
    `moving_mean` / `moving_variance` variables -- what tf.layers.batch_normalization(training=False) reads -- so the
    serialized formula x * (gamma * rsqrt(var + eps)) + (beta - mean * gamma * rsqrt(var + eps)) is executed as is;
  * dropout after each conv block: keep-mask fed with ones and the 1/(1-rate) factor with 1 (identity at inference);
  * zoneout of the encoder LSTMs: recorded are the un-zoned new_c / new_h of single loop iterations (the loop body is
    straight-line), which do not depend on it.

    python oracle/make_golden_taco_encpost.py        ->  tests/golden/taco_encpost_from_graph.npz
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import tacotron_oracle as to                               # noqa: E402
import tf_graph_eval as E                                   # noqa: E402
from make_golden_taco_graph import DEFAULT_META, P         # noqa: E402
from tacotronv2_wavernn_chinese_b200.tacotron import ckpt  # noqa: E402

CKPT_DIR = os.path.dirname(DEFAULT_META)
ENC = [f'encoder_convolutions/conv_layer_{i}_encoder_convolutions/' for i in (1, 2, 3)]
POST = [f'postnet_convolutions/conv_layer_{i}_postnet_convolutions/' for i in (1, 2, 3, 4, 5)]


def inference_feeds(variables, blocks):
    feeds = {}
    for b in blocks:
        bn = P + b + 'batch_normalization/'
        feeds[bn + 'moments/Squeeze'] = variables[bn + 'moving_mean']
        feeds[bn + 'moments/Squeeze_1'] = variables[bn + 'moving_variance']
        d = P + b + 'dropout_' + b.split('/')[1] + '/dropout/'
        feeds[d + 'truediv'] = np.float32(1.0)
        feeds[d + 'Cast'] = np.float32(1.0)
    return feeds


def encoder_lstm_step(nodes, variables, direction, x_t, c, h):
    sc = P + f'encoder_LSTM/bidirectional_rnn/{direction}/{direction}/while/'
    inner = sc + f'encoder_{direction}_LSTM/'
    h_var = [i for i in nodes[inner + 'concat']['input'] if 'Identity' in i][0]          # loop variable holding h
    c_var = nodes[inner + 'mul']['input'][1]                                             # sigmoid(f + 1) * c_prev
    ev = E.Evaluator(nodes, variables, {sc + 'TensorArrayReadV3': x_t, h_var: h, c_var: c})
    return np.asarray(ev.get(inner + 'add_1')), np.asarray(ev.get(inner + 'mul_2'))


def main():
    nodes = E.load_graph(DEFAULT_META)
    variables = ckpt.load_bundle(CKPT_DIR)
    w = ckpt.load_tacotron_weights(CKPT_DIR)
    ids = np.array(json.load(open(os.path.join(ROOT, 'tests', 'golden', 'taco_symbols.json')))['sentences']['241']['ids'])
    out = {'ids': ids}
    report = {}

    def record(name, graph_value, oracle_value):
        g = np.asarray(graph_value, dtype=np.float32)
        out['graph_' + name] = g
        o = np.asarray(oracle_value, dtype=np.float32).reshape(g.shape)
        report[name] = float(np.abs(o - g).max() / max(1.0, float(np.abs(g).max())))

    # ---- encoder convolutions -------------------------------------------------------------------------------------
    feeds = inference_feeds(variables, ENC)
    feeds['datafeeder/input_queue_Dequeue'] = ids[None].astype(np.int32)
    ev = E.Evaluator(nodes, variables, feeds)
    convs = to.encoder_convs(w, ids)
    for i, b in enumerate(ENC):
        record(f'enc_conv{i + 1}', ev.get(P + b + 'batch_normalization/batchnorm/add_1')[0], convs[i])

    # ---- encoder BiLSTM: single iterations of the two loop bodies on states of the oracle's own run -------------------
    x = convs[-1]
    Tx = x.shape[0]
    for direction, order in (('fw', list(range(Tx))), ('bw', list(range(Tx - 1, -1, -1)))):
        k = w[f'encoder_LSTM/bidirectional_rnn/{direction}/encoder_{direction}_LSTM/kernel']
        b = w[f'encoder_LSTM/bidirectional_rnn/{direction}/encoder_{direction}_LSTM/bias']
        c = h = np.zeros((1, k.shape[1] // 4), dtype=np.float32)
        for n, t in enumerate(order):
            if n in (0, 1, 25):
                gc, gh = encoder_lstm_step(nodes, variables, direction, x[t:t + 1], c, h)
                oc, oh = to.lstm_cell(x[t:t + 1], c, h, k, b)
                out[f'lstm_{direction}_{n}_x'], out[f'lstm_{direction}_{n}_c'], out[f'lstm_{direction}_{n}_h'] = x[t:t + 1], c, h
                record(f'lstm_{direction}_{n}_new_c', gc, oc)
                record(f'lstm_{direction}_{n}_new_h', gh, oh)
            _, c, h = to.zoneout_lstm(x[t:t + 1], c, h, k, b)

    # ---- postnet on the first 64 decoder frames of the oracle's run (seed-1238 masks) ---------------------------------
    memory = to.encoder(w, ids)
    dec = to.decode(w, memory, seed=1238, max_iters=64)['frames']
    out['dec_frames'] = dec
    feeds = inference_feeds(variables, POST)
    feeds[P + 'Reshape_3'] = dec[None]                                                   # decoder output, before the clip
    ev = E.Evaluator(nodes, variables, feeds)
    record('postnet_conv1', ev.get(P + POST[0] + 'batch_normalization/batchnorm/add_1')[0],
           to.conv_block(np.clip(dec, np.float32(-4.1), np.float32(4.0)), w, POST[0][:-1], lambda v: np.tanh(v, dtype=np.float32)))
    record('mel', ev.get(P + 'Minimum_1')[0], to.postnet(w, dec))

    path = os.path.join(ROOT, 'tests', 'golden', 'taco_encpost_from_graph.npz')
    np.savez_compressed(path, **out)
    print(f'wrote {path} ({os.path.getsize(path)} bytes); oracle vs serialized graph, max error relative to max(1, |value|):')
    for k, v in report.items():
        print(f'  {k:22s} {v:.3e}')


if __name__ == '__main__':
    main()
