"""Numeric golden vectors for the Tacotron-2 decoder step, produced by EXECUTING the reference's own serialized graph.

TEST INFRASTRUCTURE; run HERE (needs /root/reference).  TensorFlow cannot run in this container, but the reference ships
`tacotron_model.ckpt-206500.meta`, the op-level graph its model code built, next to the weights.  The body of the decoder
while-loop (`decoder/while/CustomDecoderStep/*`: prenet, two LSTM cells, location-sensitive forward attention, frame and
stop projections) is a straight-line sub-graph of ~250 basic ops.  oracle/tf_graph_eval.py evaluates it in numpy, op by
op as serialized (MatMul, BiasAdd, Split, Sigmoid, Conv2D, Softmax, StridedSlice with its masks, the Tensordot expansion,
...), on the shipped checkpoint's variables, for loop states taken from a real sentence (train.txt line 241).  What is
fed: the loop variables, the encoder memory / keys, and the prenet dropout keep-masks (the graph draws them with
RandomUniform).  What is recorded: every intermediate the restatement also computes.

The serialized graph is the TRAINING graph, so zoneout appears in its dropout form; the recorded LSTM values are the
un-zoned new_c / new_h (what the inference code zones with fixed factors, modules.py:137-138), which do not depend on it.

    python oracle/make_golden_taco_step.py        ->  tests/golden/taco_step_from_graph.npz
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import tacotron_oracle as to                               # noqa: E402
import tf_graph_eval as E                                   # noqa: E402
from make_golden_taco_graph import DEFAULT_META, P, STEP   # noqa: E402
from tacotronv2_wavernn_chinese_b200.tacotron import ckpt  # noqa: E402

CKPT_DIR = os.path.dirname(DEFAULT_META)
LOOP = P + 'decoder/while/'
CELL = STEP + 'decoder_LSTM/decoder_LSTM/multi_rnn_cell/'
TARGETS = {
    'prenet': STEP + 'decoder_prenet/dropout_2decoder_prenet/dropout/mul_1',
    'new_c1': CELL + 'cell_0/decoder_LSTM_1/add_1',
    'new_h1': CELL + 'cell_0/decoder_LSTM_1/mul_2',
    'new_c2': CELL + 'cell_1/decoder_LSTM_2/add_1',
    'new_h2': CELL + 'cell_1/decoder_LSTM_2/mul_2',
    'query': STEP + 'Location_Sensitive_Attention/query_layer/MatMul',
    'energy': STEP + 'Location_Sensitive_Attention/Sum',
    'softmax': STEP + 'Softmax',
    'cum': STEP + 'add',
    'forward_raw': STEP + 'mul_3',
    'alignments': STEP + 'truediv',
    'context': STEP + 'Squeeze',
    'mu': STEP + 'dense/Sigmoid',
    'frame': STEP + 'linear_transform_projection/projection_linear_transform_projection/BiasAdd',
    'stop_logit': STEP + 'stop_token_projection/projection_stop_token_projection/BiasAdd',
}


def run_graph_step(nodes, variables, memory, keys, x, m, st):
    """Feeds one loop state into the serialized step and returns {name: value} for TARGETS (batch dimension 1)."""
    Tx = memory.shape[0]
    feeds = {
        LOOP + 'Identity_17': x,                                        # next_inputs (previous frame)
        LOOP + 'Identity_4': st['c1'], LOOP + 'Identity_5': st['h1'],
        LOOP + 'Identity_6': st['c2'], LOOP + 'Identity_7': st['h2'],
        LOOP + 'Identity_8': st['ctx'],
        LOOP + 'Identity_13': np.reshape(st['mu'], (1, 1)).astype(np.float32),
        LOOP + 'Identity_14': st['alpha'][None, :], LOOP + 'Identity_15': st['cum'][None, :],
        P + 'ForwardLocationSensitiveAttention/memory_layer/Tensordot': keys[None],      # keys = memory_layer(memory)
        P + 'ForwardLocationSensitiveAttention/mul': memory[None],                       # values (mask is all ones at B = 1)
        'datafeeder/input_queue_Dequeue:1': np.array([Tx], dtype=np.int32),             # input_lengths
        STEP + 'decoder_prenet/dropout_1decoder_prenet/dropout/Cast': m[0][None, :],     # keep flags instead of RandomUniform
        STEP + 'decoder_prenet/dropout_2decoder_prenet/dropout/Cast': m[1][None, :],
    }
    ev = E.Evaluator(nodes, variables, feeds)
    return {k: np.asarray(ev.get(v)) for k, v in TARGETS.items()}


def main():
    nodes = E.load_graph(DEFAULT_META)
    variables = ckpt.load_bundle(CKPT_DIR)                                               # graph variable name -> array
    w = ckpt.load_tacotron_weights(CKPT_DIR)
    import json
    sent = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'taco_symbols.json')))['sentences']['241']
    ids = np.array(sent['ids'] if isinstance(sent, dict) else sent)
    memory = to.encoder(w, ids)
    keys = (memory @ w['memory_layer/kernel']).astype(np.float32)
    steps = (0, 1, 7, 60, 200)
    d = to.decode(w, memory, seed=1238, max_iters=max(steps) + 1, capture_states=steps)
    out = {'ids': ids, 'memory': memory, 'steps': np.array(steps)}
    worst = {}
    for s in steps:
        x, st, m, _ = d['states'][s]
        g = run_graph_step(nodes, variables, memory, keys, x, m, st)
        o, _ = to.decoder_step(w, memory, keys, x, m, st)
        out[f's{s}_x'], out[f's{s}_m'] = x, m
        for k, v in st.items():
            out[f's{s}_state_{k}'] = np.asarray(v, dtype=np.float32)
        for k, v in g.items():
            v = np.asarray(v, dtype=np.float32)
            out[f's{s}_graph_{k}'] = v
            ov = np.asarray(o[k], dtype=np.float32).reshape(v.shape)
            err = float(np.abs(ov - v).max() / max(1.0, float(np.abs(v).max())))
            worst[k] = max(worst.get(k, 0.0), err)
    path = os.path.join(ROOT, 'tests', 'golden', 'taco_step_from_graph.npz')
    np.savez_compressed(path, **out)
    print(f'wrote {path} ({os.path.getsize(path)} bytes); oracle vs serialized graph, max error relative to max(1, |value|):')
    for k, v in worst.items():
        print(f'  {k:12s} {v:.3e}')


if __name__ == '__main__':
    main()
