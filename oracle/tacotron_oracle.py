"""CPU oracle for the Tacotron-2 forward-attention inference path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The reference runs on TensorFlow 1.14 (`tf.contrib.*`), which is not installable here and
the reference ships no tests / golden mels for this path.  This is therefore a numpy (float32)
restatement written from the reference's source plus the documented semantics of the TF 1.14 ops it calls
(`tf.nn.rnn_cell.LSTMCell`: gates i,j,f,o, forget_bias 1.0, kernel [in+units, 4*units]; `tf.layers.batch_normalization`
epsilon 1e-3 with moving statistics; `tf.layers.conv1d` 'same' cross-correlation with kernel [k,in,out];
`tf.layers.dropout` keep-scaling; `BahdanauAttention` bias-free memory/query layers, -inf score masking, softmax).
*** Pinning status: the DECODER STEP (rows a-7, a-8 of SURVEY section 8: prenet, both LSTM cells, forward attention,
projections), the ENCODER (convolution blocks, single BiLSTM iterations, the fw/bw wiring) and the POSTNET are pinned
numerically AND structurally against the reference's own serialized graph, executed op by op in numpy on the shipped
weights, and the WHOLE decoder run of a sentence (405 steps of config 4) is pinned against the trajectory obtained by carrying
the loop state through that serialized step graph from the first step to the stop token (bit-identical).  What has no
reference vectors: outputs of a TensorFlow session itself (TF's own random masks). ***  Pins that ARE checked:
  * tests/test_tacotron_run_pins.py: `decode` reproduces every frame, every attended position and the stop step of the
    whole-sentence trajectory driven through the serialized graph (oracle/make_golden_taco_run.py ->
    tests/golden/taco_run_from_graph.npz; measured difference 0.0 over 405 steps); the same for a 17-token and a 52-token sentence
    and for the run WITH the inference window (418 steps; attention step = the reference's own class executed on numpy);
  * tests/test_tacotron_window_pins.py: the attention step WITH the optional inference window (row a-9) against the reference's
    own `ForwardLocationSensitiveAttention.__call__` (forward_attention.py:119-231), whose statements are executed unmodified
    on numpy arrays through a stand-in for the few tensorflow ops they use (oracle/ref_harness_taco_attention.py ->
    tests/golden/taco_window_from_reference.npz, 60 loop states covering every branch; measured difference 0.0); the same
    file pins the stop rule (the reference's TacoTestHelper) and zoneout at inference (its ZoneoutLSTMCell class);
  * tests/test_tacotron_encpost_pins.py: encoder_convs / lstm_cell on the encoder kernels / postnet against the serialized
    sub-graphs (oracle/make_golden_taco_encpost.py -> tests/golden/taco_encpost_from_graph.npz; 2e-7 relative);
  * tests/test_tacotron_step_pins.py: `decoder_step` reproduces (to 1e-6; measured 0.0) every intermediate obtained by
    EXECUTING the `CustomDecoderStep` sub-graph of `tacotron_model.ckpt-206500.meta` with a numpy op interpreter
    (oracle/tf_graph_eval.py) on the shipped weights, for five loop states of a real sentence
    (oracle/make_golden_taco_step.py -> tests/golden/taco_step_from_graph.npz);
  * tests/test_tacotron_graph_pins.py: every assumption about TF-internal arithmetic (GRAPH_ASSUMPTIONS below: LSTM gate
    order / forget bias / concat order, zoneout and which h is passed on, prenet dropout rate and scaling, batch-norm
    epsilon, the forward-attention step's op sequence, what is cumulated, what feeds the location convolution, the 1e-10,
    the -inf mask, the operand order of every concat, the clip range) against tests/golden/taco_graph_facts.json, which
    oracle/make_golden_taco_graph.py read out of the reference's own serialized graph (`tacotron_model.ckpt-206500.meta`);
  * tests/test_tacotron_oracle.py: variable names/shapes of the shipped checkpoint, and that on a real training sentence
    (train.txt line 241, 444 ground-truth frames) the decoder attends monotonically and its stop token fires near the
    ground-truth length.

Only tests/, __graft_entry__.smoke() and bench.py may import this module.  Reference citations are relative to the
reference root.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
P = 'decoder/'

# Everything this restatement assumes about arithmetic that lives in TensorFlow 1.14 rather than in the reference's own
# files.  The functions below take their constants from this table, and tests/test_tacotron_graph_pins.py compares the table
# with tests/golden/taco_graph_facts.json -- facts read out of the reference's serialized graph (`*.meta` next to the
# shipped checkpoint) by oracle/make_golden_taco_graph.py.  The numeric pins (the same graph EXECUTED in numpy) are
# tests/test_tacotron_step_pins.py and tests/test_tacotron_encpost_pins.py.
GRAPH_ASSUMPTIONS = {
    'lstm_gate_order': ['i', 'j', 'f', 'o'],        # np.split order in lstm_cell
    'lstm_forget_bias': 1.0,
    'lstm_kernel_input_order': ['x', 'h'],          # concat([x, h]) @ kernel
    'zoneout': 0.1,                                 # both cell and hidden state, encoder and decoder
    'cell_output_is_unzoned_h': True,
    'batch_norm_epsilon': 1e-3,
    'prenet_dropout_rate': 0.5,                     # keep-mask scaled by 1 / (1 - rate)
    'attention_forward_epsilon': 1e-10,
    'attention_cumulates': 'softmax',               # cum += softmax output BEFORE the forward modulation
    'attention_location_input': 'cumulated',
    'attention_mu_dense_input': ['context', 'h2'],
    'projection_input': ['h2', 'context'],
    'lstm_input': ['prenet', 'prev_context'],
    'output_clip': [-4.1, 4.0],
}


def _sigmoid(x):
    return (F32(1) / (F32(1) + np.exp(-x, dtype=F32))).astype(F32)


def lstm_cell(x, c, h, kernel, bias, forget_bias=GRAPH_ASSUMPTIONS['lstm_forget_bias']):
    """tf.nn.rnn_cell.LSTMCell step (used at tacotron/models/modules.py:100,118): returns (new_c, new_h)."""
    z = (np.concatenate([x, h], axis=-1) @ kernel + bias).astype(F32)
    i, j, f, o = np.split(z, 4, axis=-1)
    new_c = (_sigmoid(f + F32(forget_bias)) * c + _sigmoid(i) * np.tanh(j, dtype=F32)).astype(F32)
    new_h = (_sigmoid(o) * np.tanh(new_c, dtype=F32)).astype(F32)
    return new_c, new_h


def zoneout_lstm(x, c, h, kernel, bias, zoneout=GRAPH_ASSUMPTIONS['zoneout']):
    """ZoneoutLSTMCell.__call__ at inference, modules.py:114-142: the OUTPUT is the un-zoned new_h (:118,:142), the
    carried state is (1-z)*new + z*prev (:137-138)."""
    new_c, new_h = lstm_cell(x, c, h, kernel, bias)
    zc = ((F32(1) - F32(zoneout)) * new_c + F32(zoneout) * c).astype(F32)
    zh = ((F32(1) - F32(zoneout)) * new_h + F32(zoneout) * h).astype(F32)
    return new_h, zc, zh


def conv1d_same(x, kernel, bias):
    """tf.layers.conv1d(padding='same') on [T, Cin] with kernel [k, Cin, Cout] (modules.py:382-387)."""
    k = kernel.shape[0]
    T = x.shape[0]
    left = (k - 1) // 2
    xp = np.zeros((T + k - 1, x.shape[1]), dtype=F32)
    xp[left:left + T] = x
    y = np.zeros((T, kernel.shape[2]), dtype=F32)
    for j in range(k):
        y += (xp[j:j + T] @ kernel[j]).astype(F32)
    return (y + bias).astype(F32)


def batch_norm(x, w, prefix, eps=GRAPH_ASSUMPTIONS['batch_norm_epsilon']):
    """tf.layers.batch_normalization(training=False): moving statistics, epsilon 1e-3 (modules.py:388)."""
    g, b = w[prefix + '/gamma'], w[prefix + '/beta']
    m, v = w[prefix + '/moving_mean'], w[prefix + '/moving_variance']
    return ((x - m) * (g / np.sqrt(v + F32(eps))) + b).astype(F32)


def conv_block(x, w, scope, activation):
    """modules.py:379-391 with batch_norm_position='after' (tacotron_hparams.py:127): conv -> activation -> BN; dropout off."""
    y = conv1d_same(x, w[scope + '/conv1d/kernel'], w[scope + '/conv1d/bias'])
    y = activation(y)
    return batch_norm(y, w, scope + '/batch_normalization')


def encoder_convs(w, ids):
    """Embedding lookup (tacotron.py:44-47) and the 3 encoder convolution blocks (modules.py:168-174); returns the output of
    every block, [Tx, 256] each (pinned against the serialized graph by tests/test_tacotron_encpost_pins.py)."""
    x = w['inputs_embedding'][np.asarray(ids)].astype(F32)
    outs = []
    for i in (1, 2, 3):
        x = conv_block(x, w, f'encoder_convolutions/conv_layer_{i}_encoder_convolutions', lambda v: np.maximum(v, F32(0)))
        outs.append(x)
    return outs


def encoder(w, ids):
    """Embedding lookup (tacotron.py:44-47) -> 3 x conv k5/256 ReLU -> BN (modules.py:168-174) -> BiLSTM 2x256 with zoneout
    (modules.py:207-217).  ids: int [Tx].  Returns memory [Tx, 512]."""
    x = encoder_convs(w, ids)[-1]
    Tx = x.shape[0]
    outs = []
    for direction, order in (('fw', range(Tx)), ('bw', range(Tx - 1, -1, -1))):
        k = w[f'encoder_LSTM/bidirectional_rnn/{direction}/encoder_{direction}_LSTM/kernel']
        b = w[f'encoder_LSTM/bidirectional_rnn/{direction}/encoder_{direction}_LSTM/bias']
        U = k.shape[1] // 4
        c = np.zeros((1, U), dtype=F32)
        h = np.zeros((1, U), dtype=F32)
        o = np.zeros((Tx, U), dtype=F32)
        for t in order:
            out, c, h = zoneout_lstm(x[t:t + 1], c, h, k, b)
            o[t] = out[0]
        outs.append(o)
    return np.concatenate(outs, axis=1).astype(F32)


def location_features(w, cum):
    """location_convolution (attention.py:100-102) + location_layer (:104-105) on cumulated alignments [Tx]."""
    K = w[P + 'Location_Sensitive_Attention/location_features_convolution/kernel']          # [31, 1, 32]
    f = conv1d_same(cum[:, None].astype(F32), K, w[P + 'Location_Sensitive_Attention/location_features_convolution/bias'])
    return (f @ w[P + 'Location_Sensitive_Attention/location_features_layer/kernel']).astype(F32)   # [Tx, 128]


def attention_window(al, max_att, pos_rec):
    """The optional inference window of forward_attention.py:171-215 for one sentence: the attended position may advance
    by at most one token per step, may not leave token <= 2 before 5 steps have been spent there (`short_mask`), is pushed
    on after 10 steps on the same token, and the raw forward term `al` [Tx] is zeroed outside [new_max-2, new_max+3) with
    the entry at new_max set to twice the remaining mass.  Returns (al, new_max, new_pos_rec).
    tests/test_tacotron_window_pins.py holds this against the reference's own statements executed on numpy arrays."""
    Tx = al.shape[0]
    new_max = int(np.argmax(al))
    new_max = max_att if new_max <= max_att else max_att + 1
    if pos_rec < 5 and 2 < new_max:
        new_max = max_att
    if new_max == max_att:
        pos_rec = pos_rec + 1
    else:
        pos_rec = 1
    if not pos_rec < 10:                                                                     # :191-195
        new_max, pos_rec = new_max + 1, 1
    idx = np.arange(Tx)
    keep = (idx >= new_max - 2) & (idx < new_max + 3)
    al = np.where(keep, al, F32(0)).astype(F32)
    peak = idx == min(max(new_max, 0), Tx - 1)
    tot = al.sum(dtype=F32)
    tot = F32(1.0) if tot < F32(1e-10) else tot                                              # :209-213
    al = np.where(peak & (idx < new_max + 1), tot * F32(2.0), al).astype(F32)                # :215
    return al, new_max, pos_rec


def decoder_step(w, memory, keys, x, m, st, zoneout=GRAPH_ASSUMPTIONS['zoneout'], modulate=None):
    """ONE iteration of the decoder loop: TacotronDecoderCell.__call__ (Architecture_wrappers.py:175-218) with
    Prenet (modules.py:240-251), 2 x ZoneoutLSTMCell (:114-142), ForwardLocationSensitiveAttention.__call__
    (attention.py:119-231) and the two projections (modules.py:304,334-342).

    x [1,80] previous frame; m [2,256] prenet keep flags; st = dict(c1,h1,c2,h2 [1,256], ctx [1,512], alpha, cum [Tx], mu);
    `modulate(al) -> al` lets decode() apply the optional inference window before the normalisation.
    Returns (intermediates, new state).  tests/test_tacotron_step_pins.py checks the intermediates against the reference's
    own serialized decoder-step graph executed on the shipped weights (oracle/make_golden_taco_step.py).
    """
    A = GRAPH_ASSUMPTIONS
    Wq = w[P + 'Location_Sensitive_Attention/query_layer/kernel']
    v_a = w[P + 'Location_Sensitive_Attention/attention_variable_projection']
    b_a = w[P + 'Location_Sensitive_Attention/attention_bias']
    k1 = w[P + 'decoder_LSTM/multi_rnn_cell/cell_0/decoder_LSTM_1/kernel']
    b1 = w[P + 'decoder_LSTM/multi_rnn_cell/cell_0/decoder_LSTM_1/bias']
    k2 = w[P + 'decoder_LSTM/multi_rnn_cell/cell_1/decoder_LSTM_2/kernel']
    b2 = w[P + 'decoder_LSTM/multi_rnn_cell/cell_1/decoder_LSTM_2/bias']
    z = F32(zoneout)
    # -- prenet, dropout always on (modules.py:240-251)
    p = x
    for li in (1, 2):
        p = np.maximum((p @ w[P + f'decoder_prenet/dense_{li}/kernel'] + w[P + f'decoder_prenet/dense_{li}/bias']).astype(F32), 0)
        p = (p * m[li - 1] * F32(1.0 / (1.0 - A['prenet_dropout_rate']))).astype(F32)       # keep-scaling 1/(1-rate)
    # -- 2 x zoneout LSTM (Architecture_wrappers.py:180-183): the un-zoned h goes on, the zoned pair is the carried state
    nc1, nh1 = lstm_cell(np.concatenate([p, st['ctx']], axis=1), st['c1'], st['h1'], k1, b1)
    nc2, nh2 = lstm_cell(nh1, st['c2'], st['h2'], k2, b2)
    zoned = lambda new, prev: ((F32(1) - z) * new + z * prev).astype(F32)                    # modules.py:137-138
    # -- attention (attention.py:132-167)
    q = (nh2 @ Wq).astype(F32)                                                               # [1,128]
    loc = location_features(w, st['cum'])
    energy = (np.tanh(keys + q + loc + b_a, dtype=F32) * v_a).sum(axis=1).astype(F32)        # [Tx]
    e = np.exp(energy - energy.max(), dtype=F32)
    a = (e / e.sum(dtype=F32)).astype(F32)                                                   # softmax (probability_fn)
    cum = (st['cum'] + a).astype(F32)                                                        # :154 (pre-modulation)
    shift = np.concatenate([[F32(0)], st['alpha'][:-1]]).astype(F32)
    al_raw = (((F32(1) - st['mu']) * st['alpha'] + st['mu'] * shift + F32(A['attention_forward_epsilon'])) * a).astype(F32)   # :167
    al = modulate(al_raw) if modulate is not None else al_raw
    al = (al / al.sum(dtype=F32)).astype(F32)                                                # :220
    ctx = (al[None, :] @ memory).astype(F32)                                                 # :222
    mu = _sigmoid((np.concatenate([ctx, nh2], axis=1) @ w[P + 'dense/kernel'] + w[P + 'dense/bias']).astype(F32))[0, 0]
    # -- projections (Architecture_wrappers.py:196-199)
    pin = np.concatenate([nh2, ctx], axis=1)
    frame = (pin @ w[P + 'linear_transform_projection/projection_linear_transform_projection/kernel']
             + w[P + 'linear_transform_projection/projection_linear_transform_projection/bias']).astype(F32)
    stop_logit = (pin @ w[P + 'stop_token_projection/projection_stop_token_projection/kernel']
                  + w[P + 'stop_token_projection/projection_stop_token_projection/bias']).astype(F32)
    out = dict(prenet=p, new_c1=nc1, new_h1=nh1, new_c2=nc2, new_h2=nh2, query=q, energy=energy, softmax=a, cum=cum,
               forward_raw=al_raw, alignments=al, context=ctx, mu=mu, frame=frame, stop_logit=stop_logit)
    new = dict(c1=zoned(nc1, st['c1']), h1=zoned(nh1, st['h1']), c2=zoned(nc2, st['c2']), h2=zoned(nh2, st['h2']),
               ctx=ctx, alpha=al, cum=cum, mu=mu)
    return out, new


def decode(w, memory, dropout_masks=None, seed=0, max_iters=2000, window=False, zoneout=GRAPH_ASSUMPTIONS['zoneout'],
           capture_states=()):
    """The decoder while-loop for ONE sentence: dynamic_decode(CustomDecoder(TacotronDecoderCell, TacoTestHelper))
    (tacotron.py:99-103; custom_decoder.py:105-135; helpers.py:36-66); `window=True` adds forward_attention.py:171-215.

    memory [Tx, 512]; dropout_masks optional [steps, 2, 256] of {0,1} keep flags for the two prenet layers
    (drawn from RandomState(seed) when None; prenet dropout is ON at inference, modules.py:249).
    Returns dict(frames [n,80] raw decoder outputs, stop [n], alignments [n,Tx], n_steps, masks
    [, states {step: (x, state, masks, window state)} for the steps listed in `capture_states`]).
    """
    Tx = memory.shape[0]
    keys = (memory @ w['memory_layer/kernel']).astype(F32)                                   # BahdanauAttention ctor
    U = w[P + 'decoder_LSTM/multi_rnn_cell/cell_0/decoder_LSTM_1/kernel'].shape[1] // 4
    rs = np.random.RandomState(seed)
    zeros = np.zeros((1, U), dtype=F32)
    alpha0 = np.zeros(Tx, dtype=F32); alpha0[0] = 1                                          # init_alpha (attention.py:112)
    st = dict(c1=zeros, h1=zeros, c2=zeros, h2=zeros,
              ctx=np.zeros((1, memory.shape[1]), dtype=F32),                                 # zero_state attention (:164)
              alpha=alpha0, cum=alpha0.copy(), mu=F32(0.5))                                  # init_cumulated_alignments, init_mu
    win = dict(max_att=0, pos_rec=0)
    x = np.zeros((1, 80), dtype=F32)                                                         # _go_frames (helpers.py:149)
    frames, stops, aligns, masks, captured = [], [], [], [], {}

    def window_modulate(al):
        al, win['max_att'], win['pos_rec'] = attention_window(al, win['max_att'], win['pos_rec'])
        return al

    for step in range(max_iters):
        m = dropout_masks[step] if dropout_masks is not None else (rs.uniform(size=(2, 256)) >= GRAPH_ASSUMPTIONS['prenet_dropout_rate'])
        m = np.asarray(m, dtype=F32)
        masks.append(m)
        if step in capture_states:
            captured[step] = (x.copy(), {k: np.array(v, copy=True) for k, v in st.items()}, m.copy(), dict(win))
        out, st = decoder_step(w, memory, keys, x, m, st, zoneout, window_modulate if window else None)
        frame = out['frame']
        stop = _sigmoid(out['stop_logit'])[0, 0]
        frames.append(frame[0]); stops.append(stop); aligns.append(out['alignments'])
        x = frame                                                                            # helpers.py:64 (r = 1)
        if stop > 0.5:                                                                       # tf.round, half-to-even (:45)
            break
    return dict(frames=np.stack(frames), stop=np.array(stops, dtype=F32), alignments=np.stack(aligns),
                n_steps=len(frames), masks=np.stack(masks), states=captured)


def postnet(w, dec):
    """Clip (tacotron.py:111-112) -> 5 x conv k5 (tanh x4, linear) -> BN (modules.py:368-376) -> projection -> residual add
    -> clip (tacotron.py:115-129).  dec [n,80] -> mel [n,80]."""
    lo, hi = F32(GRAPH_ASSUMPTIONS['output_clip'][0]), F32(GRAPH_ASSUMPTIONS['output_clip'][1])
    d = np.clip(dec, lo, hi).astype(F32)
    x = d
    for i in (1, 2, 3, 4):
        x = conv_block(x, w, f'postnet_convolutions/conv_layer_{i}_postnet_convolutions', lambda v: np.tanh(v, dtype=F32))
    x = conv_block(x, w, 'postnet_convolutions/conv_layer_5_postnet_convolutions', lambda v: v)
    r = (x @ w['postnet_projection/projection_postnet_projection/kernel']
         + w['postnet_projection/projection_postnet_projection/bias']).astype(F32)
    return np.clip(d + r, lo, hi).astype(F32)


def synthesize(w, ids, **kw):
    """Synthesizer.synthesize numerics (tacotron_synthesize.py:97-116): returns (mel_for_wavernn [n,80] in [0,1], dict)."""
    mem = encoder(w, ids)
    d = decode(w, mem, **kw)
    mel = postnet(w, d['frames'])
    rounded = np.round(d['stop'])
    target = int(np.argmax(rounded == 1)) if (rounded == 1).any() else len(rounded)          # :104-105
    mel = np.clip(mel[:target], -4.0, 4.0)                                                   # :107-108
    return np.clip((mel + 4.0) / 8.0, 0, 1).astype(F32), dict(decode=d, memory=mem, target_length=target)   # :115
