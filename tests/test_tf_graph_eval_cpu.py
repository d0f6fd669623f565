"""The numpy op interpreter behind the Tacotron golden vectors (oracle/tf_graph_eval.py) checked on its own: the ops whose
TensorFlow semantics are easy to get wrong (SAME-padded cross-correlation, StridedSlice masks, Split / Squeeze attributes,
packed-constant decoding) against plain numpy / scipy formulations, and a hand-built three-node graph end to end."""
import os
import struct
import sys

import numpy as np
import pytest
from scipy.signal import correlate

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import tf_graph_eval as E  # noqa: E402


def test_conv2d_same_is_cross_correlation_with_tf_padding():
    rs = np.random.RandomState(0)
    for k in (5, 31, 4):                                    # even kernel: TF pads (k-1)//2 on the left, the rest on the right
        x = rs.randn(1, 1, 20, 3).astype(np.float32)
        f = rs.randn(1, k, 3, 2).astype(np.float32)
        y = E._conv2d_same_nhwc(x, f)
        left = (k - 1) // 2
        xp = np.zeros((20 + k - 1, 3), dtype=np.float64)
        xp[left:left + 20] = x[0, 0]
        ref = np.stack([sum(correlate(xp[:, c], f[0, :, c, o].astype(np.float64), mode='valid') for c in range(3)) for o in range(2)], 1)
        np.testing.assert_allclose(y[0, 0], ref, rtol=1e-5, atol=1e-5)


def _attr_i(v):
    return b'\x18' + (bytes([v]) if v < 128 else None)    # AttrValue.i (field 3, varint)


def test_strided_slice_masks():
    x = np.arange(24).reshape(2, 3, 4)
    node = {'attr': {'begin_mask': _attr_i(3), 'end_mask': _attr_i(1), 'shrink_axis_mask': _attr_i(0)}}
    np.testing.assert_array_equal(E._strided_slice(x[0], [0, 0], [0, -1], [1, 1], node), x[0][:, :-1])        # alpha[:, :-1]
    node = {'attr': {'begin_mask': _attr_i(1), 'end_mask': _attr_i(1), 'shrink_axis_mask': _attr_i(6)}}
    np.testing.assert_array_equal(E._strided_slice(x, [0, 0, 0], [0, 1, 1], [1, 1, 1], node), x[:, 0, 0])      # keys[:, 0, 0]
    node = {'attr': {}}
    np.testing.assert_array_equal(E._strided_slice(np.arange(5), [1], [2], [1], node), np.arange(5)[1:2])


def test_tensor_proto_decoding():
    # TensorProto{dtype=float, shape=[], float_val=[1e-10]} and {dtype=int32, shape=[2], tensor_content=<0,-1>}
    scalar = b'\x08\x01' + b'\x12\x00' + b'\x2d' + struct.pack('<f', 1e-10)
    assert E._parse_tensor(scalar) == pytest.approx(1e-10)
    dims = b''.join(b'\x12\x02\x08' + bytes([d]) for d in (2,))
    vec = b'\x08\x03' + b'\x12' + bytes([len(dims)]) + dims + b'\x22\x08' + struct.pack('<ii', 0, -1)
    np.testing.assert_array_equal(E._parse_tensor(vec), [0, -1])
    fill = b'\x08\x01' + b'\x12' + bytes([len(dims)]) + dims + b'\x2d' + struct.pack('<f', 0.5)               # scalar broadcast
    np.testing.assert_array_equal(E._parse_tensor(fill), [0.5, 0.5])


def test_small_graph_end_to_end():
    """x -> MatMul(W) -> BiasAdd(b) -> Split(2)[1] -> Sigmoid, with a fed placeholder and 'variables'."""
    def const_i(v):
        return {'op': 'Const', 'input': [], 'attr': {'value': b'\x42\x07' + b'\x08\x03\x12\x00\x3a\x01' + bytes([v])}}
    nodes = {
        'W': {'op': 'VariableV2', 'input': [], 'attr': {}}, 'b': {'op': 'VariableV2', 'input': [], 'attr': {}},
        'mm': {'op': 'MatMul', 'input': ['x', 'W'], 'attr': {}},
        'ba': {'op': 'BiasAdd', 'input': ['mm', 'b', '^ctl'], 'attr': {}},
        'dim': const_i(1),
        'sp': {'op': 'Split', 'input': ['dim', 'ba'], 'attr': {'num_split': b'\x18\x02'}},
        'sg': {'op': 'Sigmoid', 'input': ['sp:1'], 'attr': {}},
    }
    rs = np.random.RandomState(1)
    W, b, x = rs.randn(3, 4).astype(np.float32), rs.randn(4).astype(np.float32), rs.randn(2, 3).astype(np.float32)
    out = E.Evaluator(nodes, {'W': W, 'b': b}, {'x': x}).get('sg')
    ref = 1 / (1 + np.exp(-((x @ W + b)[:, 2:])))
    np.testing.assert_allclose(out, ref, rtol=1e-6, atol=1e-6)
