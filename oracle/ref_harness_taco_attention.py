"""Execute the reference's OWN `ForwardLocationSensitiveAttention.__call__` (tacotron/models/forward_attention.py:119-231)
on numpy arrays -- container-only test tooling, the Tacotron counterpart of oracle/ref_harness.py.

TensorFlow 1.14 is not installable here, but that method only uses a few dozen element-wise / shape ops.  The unmodified
reference file is imported from /root/reference with a stand-in `tensorflow` package whose ops are numpy one-liners
(`tf.where` -> np.where, `tf.sequence_mask` -> arange < length, `tf.argmax`, `tf.concat`, `tf.reduce_sum`, ...), so every
STATEMENT of the reference's forward recursion and of its inference window (the `if not self.is_training:` block,
:171-215) runs exactly as written.  Nothing is copied from the reference; the class is instantiated without its
constructor (which needs tf.contrib's BahdanauAttention) and given the attributes `__call__` reads: keys, values, the
query / location layers as numpy closures over the checkpoint weights, `_probability_fn` = softmax.
Used by oracle/make_golden_taco_window.py.
"""
from __future__ import annotations

import collections
import contextlib
import importlib.util
import io
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get('B200TTS_REFERENCE', '/root/reference')
REF_FILE = os.path.join(REF_ROOT, 'tacotron/models/forward_attention.py')
F32 = np.float32

State = collections.namedtuple('State', 'alignments cumulated_alignments alpha mu max_attentions pos_rec')


class _Dim(int):
    """tf.Dimension look-alike: the reference reads `W_keys.shape[-1].value` (:32)."""
    @property
    def value(self):
        return int(self)


class _TFArray(np.ndarray):
    """ndarray whose .shape yields _Dim entries; numpy itself never looks at this Python-level property."""
    @property
    def shape(self):
        return tuple(_Dim(d) for d in np.ndarray.shape.__get__(self))


def available() -> bool:
    return os.path.isfile(REF_FILE)


def _np_dtype(dt):
    return {'int32': np.int32, 'float32': np.float32, None: None}.get(dt, dt)


def _make_tf(variables):
    """A `tensorflow` look-alike with exactly the ops forward_attention.py touches, all numpy."""
    tf = types.ModuleType('tensorflow')
    tf.int32, tf.float32 = 'int32', 'float32'
    tf.expand_dims = lambda x, axis: np.expand_dims(x, axis)
    tf.zeros_like = lambda x, dtype=None: np.zeros_like(x, dtype=_np_dtype(dtype))
    tf.ones_like = lambda x, dtype=None: np.ones_like(x, dtype=_np_dtype(dtype))
    tf.reshape = lambda x, shape: np.reshape(x, shape)
    tf.concat = lambda xs, axis: np.concatenate(xs, axis=axis)
    tf.clip_by_value = lambda x, lo, hi: np.clip(x, lo, hi)
    tf.argmax = lambda x, axis, output_type=None: np.argmax(x, axis=axis).astype(_np_dtype(output_type) or np.int64)
    tf.shape = lambda x: np.array(np.shape(x), dtype=np.int32)
    tf.where = lambda c, a, b: np.where(c, a, b)
    tf.less_equal, tf.less, tf.equal = np.less_equal, np.less, np.equal
    tf.logical_and, tf.logical_or, tf.logical_not = np.logical_and, np.logical_or, np.logical_not
    tf.sequence_mask = lambda lengths, maxlen: np.arange(int(maxlen))[None, :] < np.asarray(lengths)[:, None]
    tf.reduce_sum = lambda x, axis=None, keepdims=False: np.sum(x, axis=tuple(axis) if isinstance(axis, list) else axis,
                                                                keepdims=keepdims, dtype=x.dtype)
    tf.squeeze = lambda x, axis: np.squeeze(x, axis=axis)
    tf.tanh = lambda x: np.tanh(x, dtype=F32)
    tf.convert_to_tensor = lambda x, dtype=None: np.asarray(x, dtype=_np_dtype(dtype))
    tf.zeros_initializer = lambda: None
    tf.get_variable = lambda name, shape=None, dtype=None, initializer=None: variables[name]
    tf.nn = types.SimpleNamespace(sigmoid=lambda x: (F32(1) / (F32(1) + np.exp(-x, dtype=F32))).astype(F32))
    tf.contrib = types.SimpleNamespace(layers=types.SimpleNamespace(xavier_initializer=lambda: None))

    def dense(x, units, activation=None, use_bias=True):          # the one tf.layers.dense of __call__: new_mu (:227)
        y = (x @ variables['dense/kernel'] + variables['dense/bias']).astype(F32)
        return activation(y) if activation is not None else y
    tf.layers = types.SimpleNamespace(dense=dense, Conv1D=object, Dense=object)
    return tf


def import_reference(variables):
    """-> the reference module `forward_attention`, executed against the numpy stand-in for tensorflow."""
    tf = _make_tf(variables)
    aw = types.ModuleType('tensorflow.contrib.seq2seq.python.ops.attention_wrapper')
    aw.BahdanauAttention = type('BahdanauAttention', (), {})
    core = types.ModuleType('tensorflow.python.layers.core')
    ops = types.ModuleType('tensorflow.python.ops')
    ops.array_ops = types.SimpleNamespace(shape=tf.shape)
    ops.math_ops = types.SimpleNamespace(matmul=lambda a, b: np.matmul(a, b).astype(F32))
    ops.nn_ops = types.SimpleNamespace()
    ops.variable_scope = types.SimpleNamespace(variable_scope=lambda *a, **k: contextlib.nullcontext())
    layers_pkg = types.ModuleType('tensorflow.python.layers')
    layers_pkg.core = core
    python_pkg = types.ModuleType('tensorflow.python')
    python_pkg.layers, python_pkg.ops = layers_pkg, ops
    fake = {'tensorflow': tf, 'tensorflow.contrib': types.ModuleType('tensorflow.contrib'),
            'tensorflow.contrib.seq2seq': types.ModuleType('x'), 'tensorflow.contrib.seq2seq.python': types.ModuleType('x'),
            'tensorflow.contrib.seq2seq.python.ops': types.ModuleType('x'),
            'tensorflow.contrib.seq2seq.python.ops.attention_wrapper': aw,
            'tensorflow.python': python_pkg, 'tensorflow.python.layers': layers_pkg,
            'tensorflow.python.layers.core': core, 'tensorflow.python.ops': ops}
    saved = {k: sys.modules.get(k) for k in fake}
    sys.modules.update(fake)
    try:
        spec = importlib.util.spec_from_file_location('_ref_forward_attention', REF_FILE)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


class ReferenceAttention:
    """The reference's attention object for ONE sentence, built from the checkpoint weights (`w`, prefix-stripped names)."""

    def __init__(self, w, memory, is_training=False):
        P = 'decoder/'
        A = P + 'Location_Sensitive_Attention/'
        variables = {'attention_variable_projection': w[A + 'attention_variable_projection'],
                     'attention_bias': w[A + 'attention_bias'],
                     'dense/kernel': w[P + 'dense/kernel'], 'dense/bias': w[P + 'dense/bias']}
        self.mod = import_reference(variables)
        cls = self.mod.ForwardLocationSensitiveAttention
        obj = object.__new__(cls)                                                         # no constructor: it needs tf.contrib
        K, kb = w[A + 'location_features_convolution/kernel'], w[A + 'location_features_convolution/bias']

        def location_convolution(x):                                                      # tf.layers.Conv1D(31, 'same'), [B,T,1]
            k = K.shape[0]
            T = x.shape[1]
            xp = np.zeros((x.shape[0], T + k - 1, x.shape[2]), dtype=F32)
            xp[:, (k - 1) // 2:(k - 1) // 2 + T] = x
            y = np.zeros((x.shape[0], T, K.shape[2]), dtype=F32)
            for j in range(k):
                y += (xp[:, j:j + T] @ K[j]).astype(F32)
            return (y + kb).astype(F32)

        def softmax(e, _prev):
            z = np.exp(e - e.max(axis=-1, keepdims=True), dtype=F32)
            return (z / z.sum(axis=-1, keepdims=True, dtype=F32)).astype(F32)

        obj.query_layer = lambda q: (q @ w[A + 'query_layer/kernel']).astype(F32)
        obj.location_convolution = location_convolution
        obj.location_layer = lambda f: (f @ w[A + 'location_features_layer/kernel']).astype(F32)
        obj.keys = (memory @ w['memory_layer/kernel']).astype(F32)[None].view(_TFArray)
        obj.values = memory[None].astype(F32)
        obj._probability_fn = softmax
        obj.is_training = is_training
        self.obj, self.cls = obj, cls

    def __call__(self, query, state):
        """query [1,256]; state fields batched [1,...].  Returns the reference's 6-tuple."""
        with contextlib.redirect_stdout(io.StringIO()):                                   # the reference prints a banner (:172-174)
            return self.cls.__call__(self.obj, query, state)


def reference_test_helper_next_inputs(stop_probabilities, outputs):
    """Runs the reference's own `TacoTestHelper.next_inputs` (tacotron/models/helpers.py:42-66, hparams of the shipped model:
    outputs_per_step = 1, stop_at_any = True) on numpy values.  stop_probabilities [B, r], outputs [B, r * num_mels].
    Returns (finished: bool, next_inputs)."""
    tf = types.ModuleType('tensorflow')
    tf.bool = 'bool'
    tf.name_scope = lambda *a, **k: contextlib.nullcontext()
    tf.round = lambda x: np.round(x)                                  # half-to-even, like tf.round
    tf.cast = lambda x, dt: np.asarray(x).astype(np.bool_ if dt == 'bool' else dt)
    tf.reduce_all = lambda x, axis=None: np.all(x, axis=axis)
    tf.reduce_any = lambda x, axis=None: np.any(x, axis=axis)
    tf.tile = lambda x, m: np.tile(np.asarray(x), m)
    tf.TensorShape = lambda x: tuple(x)
    seq2seq = types.ModuleType('tensorflow.contrib.seq2seq')
    seq2seq.Helper = type('Helper', (), {})
    fake = {'tensorflow': tf, 'tensorflow.contrib': types.ModuleType('tensorflow.contrib'), 'tensorflow.contrib.seq2seq': seq2seq}
    saved = {k: sys.modules.get(k) for k in fake}
    sys.modules.update(fake)
    try:
        spec = importlib.util.spec_from_file_location('_ref_helpers', os.path.join(REF_ROOT, 'tacotron/models/helpers.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    hp = types.SimpleNamespace(outputs_per_step=1, use_all_outputs=False, num_mels=80, stop_at_any=True)
    helper = mod.TacoTestHelper(batch_size=np.shape(outputs)[0], hparams=hp, input_seq_lengths=None)
    _, go = helper.initialize()
    finished, nxt, _ = helper.next_inputs(0, np.asarray(outputs), None, None, np.asarray(stop_probabilities, dtype=np.float32))
    return bool(finished), np.asarray(nxt), np.asarray(go)


def reference_zoneout_lstm(x, c, h, kernel, bias, lstm_cell, zoneout=0.1):
    """Runs the reference's own `ZoneoutLSTMCell` (tacotron/models/modules.py:81-142; real constructor, real `__call__`,
    is_training = False) with the inner `tf.nn.rnn_cell.LSTMCell` replaced by `lstm_cell(x, c, h, kernel, bias) ->
    (new_c, new_h)` (that arithmetic is pinned separately against the serialized graph).  What this executes is the
    reference's zoneout statement at inference and which h it returns as the cell OUTPUT.  Returns (output, c, h)."""
    StateTuple = collections.namedtuple('LSTMStateTuple', 'c h')

    class LSTMCell:
        def __init__(self, num_units, state_is_tuple=True, name=None):
            self._num_units, self._num_proj = num_units, None

        def __call__(self, inputs, state, scope=None):
            new_c, new_h = lstm_cell(inputs, state[0], state[1], kernel, bias)
            return new_h, StateTuple(new_c, new_h)                    # tf LSTMCell: output = new_h, state = (new_c, new_h)

    tf = types.ModuleType('tensorflow')
    tf.nn = types.SimpleNamespace(rnn_cell=types.SimpleNamespace(RNNCell=object, LSTMCell=LSTMCell, LSTMStateTuple=StateTuple),
                                  relu=None, sigmoid=None, tanh=None)
    tf.layers = types.SimpleNamespace(Dense=lambda *a, **k: None)
    tf.constant_initializer = lambda *a, **k: None
    saved = sys.modules.get('tensorflow')
    sys.modules['tensorflow'] = tf
    try:
        spec = importlib.util.spec_from_file_location('_ref_modules', os.path.join(REF_ROOT, 'tacotron/models/modules.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is None:
            sys.modules.pop('tensorflow', None)
        else:
            sys.modules['tensorflow'] = saved
    cell = mod.ZoneoutLSTMCell(kernel.shape[1] // 4, is_training=False, zoneout_factor_cell=zoneout, zoneout_factor_output=zoneout)
    out, new_state = cell(x, StateTuple(c, h))
    return np.asarray(out), np.asarray(new_state.c), np.asarray(new_state.h)
