"""Structural pins for the Tacotron-2 restatement, read out of the reference's own serialized graph.

TEST INFRASTRUCTURE (like everything under oracle/): run HERE, where /root/reference exists; the result is the small
fixture tests/golden/taco_graph_facts.json, which travels.

TensorFlow 1.14 cannot run in this container, so oracle/tacotron_oracle.py cannot be compared with a TensorFlow run.
What the reference does ship is `logs-Tacotron-2/taco_pretrained/tacotron_model.ckpt-206500.meta`, the
MetaGraphDef written next to the checkpoint: the op-level graph that the reference's model code (tacotron/models/*.py)
built.  It is the TRAINING graph (teacher-forcing helper, zoneout in its dropout form, batch statistics), but the LSTM
cells, the attention step, the prenet and the projections are the same code in both modes.  This script walks that
protobuf WITHOUT TensorFlow (schema-less wire-format decoder below) and extracts every fact the restatement had to
assume about third-party arithmetic: LSTM gate order and forget bias, zoneout rates and which h is passed on, dropout
rate / scaling of the prenet, batch-norm epsilon, the attention recursion (what is cumulated, what feeds the location
convolution, the 1e-10, the mask value, operand orders of the concats), the clip range.

    python oracle/make_golden_taco_graph.py [--meta PATH] [--out tests/golden/taco_graph_facts.json]
"""
from __future__ import annotations

import argparse
import json
import math
import os
import struct

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_META = '/root/reference/logs-Tacotron-2/taco_pretrained/tacotron_model.ckpt-206500.meta'
P = 'Tacotron_model/inference/'
STEP = P + 'decoder/while/CustomDecoderStep/'


# ---- protobuf wire format (no schema needed: only field numbers of MetaGraphDef / GraphDef / NodeDef / AttrValue) ----
def _varint(b, p):
    r = s = 0
    while True:
        c = b[p]
        p += 1
        r |= (c & 0x7F) << s
        s += 7
        if not c & 0x80:
            return r, p


def _fields(b):
    p, n = 0, len(b)
    while p < n:
        key, p = _varint(b, p)
        f, w = key >> 3, key & 7
        if w == 0:
            v, p = _varint(b, p)
        elif w == 1:
            v, p = b[p:p + 8], p + 8
        elif w == 2:
            ln, p = _varint(b, p)
            v, p = b[p:p + ln], p + ln
        elif w == 5:
            v, p = b[p:p + 4], p + 4
        else:
            raise ValueError(f'unsupported wire type {w}')
        yield f, w, v


def load_graph(path):
    """{node name: dict(op, input [names], attr {key: raw AttrValue bytes})} of MetaGraphDef.graph_def (field 2)."""
    data = open(path, 'rb').read()
    graph_def = next(v for f, w, v in _fields(data) if f == 2 and w == 2)
    nodes = {}
    for f, w, nb in _fields(graph_def):
        if f != 1:
            continue
        d = {'input': [], 'attr': {}}
        for ff, ww, v in _fields(nb):
            if ff == 1:
                d['name'] = v.decode()
            elif ff == 2:
                d['op'] = v.decode()
            elif ff == 3:
                d['input'].append(v.decode())
            elif ff == 5:                                   # map<string, AttrValue> entry
                k = val = None
                for f3, w3, v3 in _fields(v):
                    if f3 == 1:
                        k = v3.decode()
                    elif f3 == 2:
                        val = v3
                d['attr'][k] = val
        nodes[d['name']] = d
    return nodes


def const_float(nodes, name):
    """Scalar float value of a Const node (AttrValue.tensor = field 8; TensorProto.float_val = 5, tensor_content = 4)."""
    for f, w, v in _fields(nodes[name]['attr']['value']):
        if f == 8:
            for ff, ww, x in _fields(v):
                if ff == 5:
                    return struct.unpack('<f', x if ww == 5 else x[:4])[0]
                if ff == 4 and len(x) == 4:
                    return struct.unpack('<f', x)[0]
    raise KeyError(name)


def consumers(nodes, ref):
    return [n for n, d in nodes.items() if ref in d['input'] and 'optimizer' not in n]


# ---- facts ------------------------------------------------------------------------------------------------------------
def lstm_facts(nodes, cell_scope, lstm_name):
    """cell_scope: the ZoneoutLSTMCell scope; lstm_name: the inner tf LSTMCell scope (holds concat/MatMul/split)."""
    inner = cell_scope + lstm_name + '/'
    split = inner + 'split'
    roles, forget_bias = {}, None
    for k in range(4):
        ref = split if k == 0 else f'{split}:{k}'
        (c,) = consumers(nodes, ref)
        op = nodes[c]['op']
        if op == 'Add':                                     # f + forget_bias -> sigmoid -> * c_prev
            other = [i for i in nodes[c]['input'] if i != ref][0]
            forget_bias = const_float(nodes, other)
            (sg,) = consumers(nodes, c)
            assert nodes[sg]['op'] == 'Sigmoid'
            roles[k] = 'f'
        elif op == 'Tanh':
            roles[k] = 'j'
        else:
            assert op == 'Sigmoid', op
            (mul,) = consumers(nodes, c)
            other = [i for i in nodes[mul]['input'] if i != c][0]
            src = nodes[other]['input'][0]                  # Tanh(split:1) -> input gate ; Tanh(new_c) -> output gate
            roles[k] = 'i' if src.startswith(split) else 'o'
    concat = nodes[inner + 'concat']['input'][:2]
    new_h = inner + 'mul_2'                                 # o * tanh(new_c)
    # consumers of the un-zoned h other than the zoneout arithmetic itself (`sub_1` = new_h - prev_h): the next layer / the
    # attention query / dynamic_rnn's output Select.  The zoned h only ever reaches the recurrent state.
    outside = [c for c in consumers(nodes, new_h) if not c.startswith(inner) and nodes[c]['op'] != 'Sub']
    facts = {
        'gate_order': [roles[k] for k in range(4)],
        'forget_bias': forget_bias,
        'kernel_input_order': ['x', 'h'] if '/while/Identity' in concat[1] or 'Identity' in concat[1] else ['h', 'x'],
        'unzoned_h_is_cell_output': bool(outside),
    }
    if cell_scope + 'dropout/rate' in nodes:                # zoneout, training form: prev + (1-z)*dropout(new-prev, rate z)
        facts['zoneout_cell'] = const_float(nodes, cell_scope + 'dropout/rate')
        facts['zoneout_hidden'] = const_float(nodes, cell_scope + 'dropout_1/rate')
        facts['zoneout_keep_scale'] = const_float(nodes, cell_scope + 'mul/x')
    return facts


def attention_facts(nodes):
    S = STEP
    loop = P + 'decoder/while/'

    def nxt(i):                                             # what the loop variable Identity_i receives for the next step
        return nodes[loop + f'NextIteration_{i}']['input'][0]

    softmax = S + 'Softmax'
    cum_add = nodes[S + 'add']['input']
    cum_state = [i for i in cum_add if i != softmax][0]     # loop variable holding the cumulated alignments
    loc_src = nodes[S + 'Location_Sensitive_Attention/ExpandDims_1']['input'][0]
    energy_chain = [nodes[S + 'Location_Sensitive_Attention/' + n]['op'] for n in ('add', 'add_1', 'add_2', 'Tanh', 'mul', 'Sum')]
    alpha_state = nodes[S + 'mul_1']['input'][1]
    mu_state = nodes[S + 'mul_2']['input'][0]
    shifted = nodes[S + 'concat_1']['input'][:2]
    h2 = S + 'decoder_LSTM/decoder_LSTM/multi_rnn_cell/cell_1/decoder_LSTM_2/mul_2'
    ctx = S + 'Squeeze'

    def names(inputs):
        m = {h2: 'h2', ctx: 'context', S + 'decoder_prenet/dropout_2decoder_prenet/dropout/mul_1': 'prenet'}
        return [m.get(i, 'prev_context' if i == loop + 'Identity_8' else i.replace(P, '')) for i in inputs[:2]]

    return {
        'energy_ops': energy_chain,                                            # keys+query, +location, +bias, tanh, *v, sum
        'softmax_input_masked_with': const_float(nodes, S + 'mul/x'),
        'cumulated_alignments_add': 'softmax' if softmax in cum_add else 'other',
        'cumulated_state_feeds_location_conv': loc_src == cum_state,
        'cumulated_state_updated_with_sum': nxt(int(cum_state.rsplit('_', 1)[1])) == S + 'add',
        'forward_term': [nodes[S + n]['op'] for n in ('sub', 'mul_1', 'mul_2', 'add_1', 'add_2', 'mul_3', 'Sum', 'truediv')],
        'forward_epsilon': const_float(nodes, S + 'add_2/y'),
        'shifted_alpha_is_zero_then_alpha_without_last': nodes[shifted[0]]['op'] == 'Reshape' and
        nodes[shifted[1]]['op'] == 'StridedSlice' and nodes[shifted[1]]['input'][0] == alpha_state,
        'alpha_state_is_normalised': nxt(int(alpha_state.rsplit('_', 1)[1])) == S + 'truediv',
        'mu_state_is_sigmoid_dense': nxt(int(mu_state.rsplit('_', 1)[1])) == S + 'dense/Sigmoid',
        'context_is_alpha_times_memory': nodes[S + 'MatMul']['op'] == 'BatchMatMulV2' and
        nodes[S + 'ExpandDims']['input'][0] == S + 'truediv',
        'query_is': names(nodes[S + 'Location_Sensitive_Attention/query_layer/MatMul']['input'])[0],
        'mu_dense_input': names(nodes[S + 'concat_2']['input']),
        'projection_input': names(nodes[S + 'concat_3']['input']),
        'lstm_input': names(nodes[S + 'concat']['input']),
        'frame_and_stop_share_input': nodes[S + 'stop_token_projection/projection_stop_token_projection/MatMul']['input'][0] ==
        nodes[S + 'linear_transform_projection/projection_linear_transform_projection/MatMul']['input'][0],
    }


def prenet_facts(nodes):
    S = STEP + 'decoder_prenet/'
    out = {'activation': [nodes[S + f'dense_{i}/Relu']['op'] for i in (1, 2)], 'dropout_rate': [], 'keep_if_uniform_ge_rate': [],
           'scale_is_one_over_keep': []}
    for i in (1, 2):
        d = S + f'dropout_{i}decoder_prenet/dropout/'
        out['dropout_rate'].append(const_float(nodes, d + 'rate'))
        out['keep_if_uniform_ge_rate'].append(nodes[d + 'GreaterEqual']['input'][1] == d + 'rate')
        out['scale_is_one_over_keep'].append(nodes[d + 'truediv']['op'] == 'RealDiv' and nodes[d + 'sub']['input'] == [d + 'sub/x', d + 'rate']
                                             and const_float(nodes, d + 'truediv/x') == 1.0 and const_float(nodes, d + 'sub/x') == 1.0)
    out['second_layer_reads_dropped_first'] = nodes[S + 'dense_2/MatMul']['input'][0] == S + 'dropout_1decoder_prenet/dropout/mul_1'
    return out


def conv_block_facts(nodes):
    """Per conv block: what sits between the convolution and the batch norm, and the convolution's padding."""
    out = {}
    blocks = [f'encoder_convolutions/conv_layer_{i}_encoder_convolutions/' for i in (1, 2, 3)] + \
             [f'postnet_convolutions/conv_layer_{i}_postnet_convolutions/' for i in (1, 2, 3, 4, 5)]
    for b in blocks:
        bn_in = nodes[P + b + 'batch_normalization/batchnorm/mul_1']['input'][0]
        op = nodes[bn_in]['op']
        conv = nodes[P + b + 'conv1d/conv1d']
        out[b.split('/')[1]] = {
            'before_batch_norm': op if op in ('Relu', 'Tanh') else 'linear',
            'bias_before_activation': nodes[P + b + 'conv1d/BiasAdd']['input'][0].endswith('conv1d/conv1d/Squeeze'),
            'padding': conv['attr']['padding'].split(b'\x04', 1)[1].decode(),
            'scale_is_gamma_rsqrt_var_plus_eps': nodes[P + b + 'batch_normalization/batchnorm/Rsqrt']['input'][0].endswith('batchnorm/add')
            and nodes[P + b + 'batch_normalization/batchnorm/mul']['input'][1].endswith('gamma/read'),
        }
    loc = nodes[STEP + 'Location_Sensitive_Attention/location_features_convolution/conv1d']
    out['location_features_convolution'] = {'padding': loc['attr']['padding'].split(b'\x04', 1)[1].decode()}
    return out


def collect(meta_path):
    nodes = load_graph(meta_path)
    cell = STEP + 'decoder_LSTM/decoder_LSTM/multi_rnn_cell/'
    bn = sorted(n for n in nodes if n.startswith(P) and n.endswith('batch_normalization/batchnorm/add/y'))
    facts = {
        'source': 'MetaGraphDef ' + os.path.basename(meta_path) + ' (training graph of the reference model code), '
                  f'{len(nodes)} nodes, walked without TensorFlow by oracle/make_golden_taco_graph.py',
        'decoder_lstm_1': lstm_facts(nodes, cell + 'cell_0/', 'decoder_LSTM_1'),
        'decoder_lstm_2': lstm_facts(nodes, cell + 'cell_1/', 'decoder_LSTM_2'),
        'decoder_lstm_2_reads': 'unzoned_h1' if nodes[cell + 'cell_1/decoder_LSTM_2/concat']['input'][0] ==
        cell + 'cell_0/decoder_LSTM_1/mul_2' else 'other',
        'encoder_lstm_fw': lstm_facts(nodes, P + 'encoder_LSTM/bidirectional_rnn/fw/fw/while/', 'encoder_fw_LSTM'),
        'encoder_lstm_bw': lstm_facts(nodes, P + 'encoder_LSTM/bidirectional_rnn/bw/bw/while/', 'encoder_bw_LSTM'),
        'encoder_bilstm': {
            'bw_reads_reversed_conv_output': nodes[P + 'encoder_LSTM/bidirectional_rnn/bw/ReverseSequence']['input'][0].endswith(
                'conv_layer_3_encoder_convolutions/dropout/mul_1'),
            'bw_output_is_reversed_back': nodes[P + 'encoder_LSTM/ReverseSequence']['input'][0] ==
            P + 'encoder_LSTM/bidirectional_rnn/bw/bw/transpose_1',
            'memory_concat': ['fw' if '/fw/' in i else 'bw_reversed' if i.endswith('encoder_LSTM/ReverseSequence') else i
                              for i in nodes[P + 'encoder_LSTM/concat']['input'][:2]],
        },
        'prenet': prenet_facts(nodes),
        'conv_blocks': conv_block_facts(nodes),
        'attention': attention_facts(nodes),
        'batch_norm_epsilon': sorted({round(const_float(nodes, n), 9) for n in bn}),
        'batch_norm_layers': len(bn),
        'output_clip': [const_float(nodes, P + 'Maximum/y'), const_float(nodes, P + 'Minimum/y'),
                        const_float(nodes, P + 'Maximum_1/y'), const_float(nodes, P + 'Minimum_1/y')],
    }
    return facts


def _jsonable(x):
    if isinstance(x, float):
        if math.isinf(x):
            return '-inf' if x < 0 else 'inf'
        return float(repr(struct.unpack('<f', struct.pack('<f', x))[0]))
    if isinstance(x, dict):
        return {k: _jsonable(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_jsonable(v) for v in x]
    return x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--meta', default=DEFAULT_META)
    ap.add_argument('--out', default=os.path.join(ROOT, 'tests', 'golden', 'taco_graph_facts.json'))
    a = ap.parse_args()
    facts = _jsonable(collect(a.meta))
    with open(a.out, 'w') as f:
        json.dump(facts, f, indent=1, sort_keys=True)
        f.write('\n')
    print(f'wrote {a.out}: {len(json.dumps(facts))} bytes, keys {sorted(facts)}')


if __name__ == '__main__':
    main()
