"""Minimal numpy evaluator for straight-line pieces of a serialized TensorFlow 1.x GraphDef -- TEST INFRASTRUCTURE.

Used by oracle/make_golden_taco_step.py to EXECUTE the reference's own decoder-step sub-graph (as serialized in
`tacotron_model.ckpt-206500.meta`) on the shipped weights, so that oracle/tacotron_oracle.py gets numeric golden vectors
that come from the reference's graph rather than from a second reading of its source.  No TensorFlow: the graph is
decoded with the schema-less protobuf walker of make_golden_taco_graph.py and every op below is restated from its
documented TF semantics in a few lines of numpy (fp32 in, fp32 out).  Control flow (Enter / Merge / Switch /
NextIteration) is NOT interpreted: the caller feeds the loop variables (`.../while/Identity_k`) and anything random
(`.../dropout/Cast`), and asks for nodes of the loop body.
"""
from __future__ import annotations

import struct

import numpy as np

from make_golden_taco_graph import _fields, load_graph  # noqa: F401  (same directory; re-exported)

_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 10: np.bool_}


def _parse_shape(b):
    dims = []
    for f, w, v in _fields(b):
        if f == 2:                                   # Dim
            size = 0
            for ff, ww, x in _fields(v):
                if ff == 1:
                    size = x if x < (1 << 63) else x - (1 << 64)
            dims.append(size)
    return dims


def _parse_tensor(b):
    dtype, shape, content, vals = 1, [], None, []
    for f, w, v in _fields(b):
        if f == 1:
            dtype = v
        elif f == 2:
            shape = _parse_shape(v)
        elif f == 4:
            content = v
        elif f == 5:                                 # float_val (packed or not)
            vals += list(struct.unpack('<%df' % (len(v) // 4), v)) if w in (2, 5) else [v]
        elif f in (7, 10, 11):                       # int_val / int64_val / bool_val (varints, maybe packed)
            if w == 0:
                vals.append(v)
            else:
                p = 0
                while p < len(v):
                    r = s = 0
                    while True:
                        c = v[p]; p += 1
                        r |= (c & 0x7F) << s; s += 7
                        if not c & 0x80:
                            break
                    vals.append(r)
    np_dt = _DTYPES[dtype]
    if content is not None and len(content):
        arr = np.frombuffer(content, dtype=np_dt).copy()
    else:
        if dtype in (3, 9):
            vals = [x - (1 << 64) if x >= (1 << 63) else x for x in vals]
        arr = np.array(vals if vals else [0], dtype=np_dt)
    n = int(np.prod(shape)) if shape else 1
    if arr.size == 1 and n != 1:
        arr = np.full(n, arr[0], dtype=np_dt)
    return arr.reshape(shape)


def _attr(node, key, kind, default=None):
    raw = node['attr'].get(key)
    if raw is None:
        return default
    for f, w, v in _fields(raw):
        if kind == 'i' and f == 3:
            return v - (1 << 64) if v >= (1 << 63) else v
        if kind == 'b' and f == 5:
            return bool(v)
        if kind == 's' and f == 2:
            return v.decode()
        if kind == 'type' and f == 6:
            return v
        if kind == 'tensor' and f == 8:
            return _parse_tensor(v)
        if kind == 'list_i' and f == 1:
            out = []
            for ff, ww, x in _fields(v):
                if ff == 3:
                    if ww == 0:
                        out.append(x)
                    else:
                        p = 0
                        while p < len(x):
                            r = s = 0
                            while True:
                                c = x[p]; p += 1
                                r |= (c & 0x7F) << s; s += 7
                                if not c & 0x80:
                                    break
                            out.append(r)
            return out
    return default


def _strided_slice(x, begin, end, strides, node):
    bm, em = _attr(node, 'begin_mask', 'i', 0), _attr(node, 'end_mask', 'i', 0)
    sm = _attr(node, 'shrink_axis_mask', 'i', 0)
    assert _attr(node, 'ellipsis_mask', 'i', 0) == 0 and _attr(node, 'new_axis_mask', 'i', 0) == 0
    idx = []
    for d in range(len(begin)):
        if sm >> d & 1:
            idx.append(int(begin[d]))
            continue
        b = None if bm >> d & 1 else int(begin[d])
        e = None if em >> d & 1 else int(end[d])
        idx.append(slice(b, e, int(strides[d])))
    return x[tuple(idx)]


def _conv2d_same_nhwc(x, f):
    """tf.nn.conv2d(padding='SAME', strides 1): cross-correlation, x [B,H,W,Cin], f [kh,kw,Cin,Cout]."""
    B, H, W, Cin = x.shape
    kh, kw, _, Cout = f.shape
    ph, pw = kh - 1, kw - 1
    xp = np.zeros((B, H + ph, W + pw, Cin), dtype=np.float32)
    xp[:, ph // 2:ph // 2 + H, pw // 2:pw // 2 + W] = x
    y = np.zeros((B, H, W, Cout), dtype=np.float32)
    for i in range(kh):
        for j in range(kw):
            y += np.tensordot(xp[:, i:i + H, j:j + W], f[i, j], axes=([3], [0])).astype(np.float32)
    return y


class Evaluator:
    def __init__(self, nodes, variables, feeds):
        self.nodes, self.variables, self.memo = nodes, variables, {}
        for k, v in feeds.items():
            self.memo[k if ':' in k else k + ':0'] = np.asarray(v)

    def get(self, ref):
        if ref.startswith('^'):
            raise ValueError('control input')
        name, _, idx = ref.partition(':')
        key = f'{name}:{idx or 0}'
        if key not in self.memo:
            outs = self._run(name)
            if not isinstance(outs, (list, tuple)):
                outs = [outs]
            for i, o in enumerate(outs):
                self.memo[f'{name}:{i}'] = o
        return self.memo[key]

    def _run(self, name):
        n = self.nodes[name]
        op = n['op']
        ins = [i for i in n['input'] if not i.startswith('^')]
        g = self.get
        f32 = np.float32
        if op == 'Const':
            return _attr(n, 'value', 'tensor')
        if op == 'VariableV2':
            return self.variables[name]
        if op in ('Identity', 'Enter', 'StopGradient'):
            return g(ins[0])
        if op == 'MatMul':
            a, b = g(ins[0]), g(ins[1])
            if _attr(n, 'transpose_a', 'b', False):
                a = a.T
            if _attr(n, 'transpose_b', 'b', False):
                b = b.T
            return (a @ b).astype(f32)
        if op == 'BatchMatMulV2':
            assert not _attr(n, 'adj_x', 'b', False) and not _attr(n, 'adj_y', 'b', False)
            return np.matmul(g(ins[0]), g(ins[1])).astype(f32)
        if op == 'BiasAdd':
            return (g(ins[0]) + g(ins[1])).astype(f32)
        if op in ('Add', 'AddV2'):
            a, b = g(ins[0]), g(ins[1])
            return (a + b).astype(np.result_type(a, b))
        if op == 'Sub':
            a, b = g(ins[0]), g(ins[1])
            return (a - b).astype(np.result_type(a, b))
        if op == 'Mul':
            a, b = g(ins[0]), g(ins[1])
            with np.errstate(invalid='ignore'):
                return (a * b).astype(np.result_type(a, b))
        if op == 'RealDiv':
            return (g(ins[0]) / g(ins[1])).astype(f32)
        if op == 'Rsqrt':
            return (f32(1) / np.sqrt(g(ins[0]), dtype=f32)).astype(f32)
        if op == 'Maximum':
            return np.maximum(g(ins[0]), g(ins[1]))
        if op == 'Minimum':
            return np.minimum(g(ins[0]), g(ins[1]))
        if op == 'Relu':
            return np.maximum(g(ins[0]), f32(0))
        if op == 'Sigmoid':
            x = g(ins[0])
            return (f32(1) / (f32(1) + np.exp(-x, dtype=f32))).astype(f32)
        if op == 'Tanh':
            return np.tanh(g(ins[0]), dtype=f32)
        if op == 'Softmax':
            x = g(ins[0])
            e = np.exp(x - x.max(axis=-1, keepdims=True), dtype=f32)
            return (e / e.sum(axis=-1, keepdims=True, dtype=f32)).astype(f32)
        if op == 'ConcatV2':
            return np.concatenate([g(i) for i in ins[:-1]], axis=int(g(ins[-1])))
        if op == 'Split':
            return np.split(g(ins[1]), _attr(n, 'num_split', 'i'), axis=int(g(ins[0])))
        if op == 'ExpandDims':
            return np.expand_dims(g(ins[0]), int(g(ins[1])))
        if op == 'Squeeze':
            dims = _attr(n, 'squeeze_dims', 'list_i', [])
            return np.squeeze(g(ins[0]), axis=tuple(dims) if dims else None)
        if op in ('Sum', 'Prod'):
            axis = g(ins[1])
            axis = tuple(int(a) for a in np.atleast_1d(axis))
            fn = np.sum if op == 'Sum' else np.prod
            x = g(ins[0])
            return fn(x, axis=axis, keepdims=_attr(n, 'keep_dims', 'b', False)).astype(x.dtype)
        if op == 'ArgMax':
            return np.argmax(g(ins[0]), axis=int(g(ins[1]))).astype(np.int64)
        if op == 'Select':
            return np.where(g(ins[0]), g(ins[1]), g(ins[2]))
        if op == 'Fill':
            return np.full(tuple(int(d) for d in g(ins[0])), g(ins[1]))
        if op == 'Shape':
            return np.array(g(ins[0]).shape, dtype=np.int32)
        if op == 'ZerosLike':
            return np.zeros_like(g(ins[0]))
        if op == 'Reshape':
            return g(ins[0]).reshape(tuple(int(d) for d in g(ins[1])))
        if op == 'Transpose':
            return np.transpose(g(ins[0]), tuple(int(d) for d in g(ins[1])))
        if op == 'GatherV2':
            return np.take(g(ins[0]), g(ins[1]), axis=int(g(ins[2])))
        if op == 'Pack':
            return np.stack([g(i) for i in ins], axis=_attr(n, 'axis', 'i', 0))
        if op == 'StridedSlice':
            return _strided_slice(g(ins[0]), g(ins[1]), g(ins[2]), g(ins[3]), n)
        if op == 'Range':
            return np.arange(int(g(ins[0])), int(g(ins[1])), int(g(ins[2])), dtype=np.int32)
        if op == 'Cast':
            return g(ins[0]).astype(_DTYPES[_attr(n, 'DstT', 'type')])
        if op == 'Less':
            return g(ins[0]) < g(ins[1])
        if op == 'GreaterEqual':
            return g(ins[0]) >= g(ins[1])
        if op == 'Conv2D':
            assert _attr(n, 'padding', 's') == 'SAME' and _attr(n, 'data_format', 's', 'NHWC') == 'NHWC'
            return _conv2d_same_nhwc(g(ins[0]), g(ins[1]))
        raise NotImplementedError(f'{op} ({name})')
