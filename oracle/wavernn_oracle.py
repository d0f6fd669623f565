"""CPU oracle for the WaveRNN generation hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy (float32) restatement of the reference algorithm
(lturing/tacotronv2_wavernn_chinese, wavernn/models/fatchord_version.py and
wavernn/utils/dsp.py).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import this module; the product
package never does (it fails loudly when the CUDA library is missing).

Parity pin: `tests/golden/wavernn_*.npz` were produced by running the
*reference itself* (imported from /root/reference by oracle/make_golden_wavernn.py)
and `tests/test_oracle_golden.py` checks this restatement against them.

Every function cites the reference lines it follows (paths relative to the
reference root).
"""
from __future__ import annotations

import time
import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------------------
def as_params(state_dict) -> dict:
    """state_dict (torch tensors or numpy arrays) -> dict of contiguous numpy arrays."""
    out = {}
    for k, v in state_dict.items():
        if hasattr(v, 'detach'):
            v = v.detach().cpu().numpy()
        out[k] = np.ascontiguousarray(v)
    return out


def _dims(p):
    rnn = p['rnn1.weight_hh_l0'].shape[1]
    aux = p['rnn2.weight_ih_l0'].shape[1] - rnn
    feat = p['I.weight'].shape[1] - aux - 1
    ncls = p['fc3.weight'].shape[0]
    scales = []
    j = 0
    while f'upsample.up_layers.{2 * j + 1}.weight' in p:
        scales.append((p[f'upsample.up_layers.{2 * j + 1}.weight'].shape[-1] - 1) // 2)
        j += 1
    pad = (p['upsample.resnet.conv_in.weight'].shape[2] - 1) // 2
    nblocks = 0
    while f'upsample.resnet.layers.{nblocks}.conv1.weight' in p:
        nblocks += 1
    return dict(rnn=rnn, aux=aux, feat=feat, ncls=ncls, scales=tuple(scales), pad=pad,
                hop=int(np.prod(scales)), nblocks=nblocks)


# ----------------------------------------------------------------------------------------------
# upsample network  (fatchord_version.py:13-89)
# ----------------------------------------------------------------------------------------------
def pad_tensor(x, pad, side='both'):
    """fatchord_version.py:281-291; x is [B, T, C]."""
    b, t, c = x.shape
    total = t + 2 * pad if side == 'both' else t + pad
    out = np.zeros((b, total, c), dtype=x.dtype)
    if side in ('before', 'both'):
        out[:, pad:pad + t, :] = x
    elif side == 'after':
        out[:, :t, :] = x
    return out


def _bn_eval(x, p, prefix, eps=1e-5):
    """nn.BatchNorm1d in eval mode on [B, C, L] (fatchord_version.py:18-19,36; torch default eps 1e-5)."""
    mean = p[prefix + '.running_mean'].astype(F32)[None, :, None]
    var = p[prefix + '.running_var'].astype(F32)[None, :, None]
    w = p[prefix + '.weight'].astype(F32)[None, :, None]
    b = p[prefix + '.bias'].astype(F32)[None, :, None]
    inv = (F32(1.0) / np.sqrt(var + F32(eps))).astype(F32)
    return ((x - mean) * inv * w + b).astype(F32)


def mel_resnet(p, m):
    """MelResNet.forward, fatchord_version.py:42-48 (ResBlock.forward :21-28). m: [B, feat, Tp] -> [B, O, Tp-2*pad]."""
    w = p['upsample.resnet.conv_in.weight'].astype(F32)              # [C, feat, k]
    C, feat, k = w.shape
    B, _, Tp = m.shape
    T = Tp - (k - 1)
    x = np.zeros((B, C, T), dtype=F32)
    for j in range(k):                                               # valid conv, no bias (:35)
        x += np.einsum('ci,bit->bct', w[:, :, j], m[:, :, j:j + T]).astype(F32)
    x = np.maximum(_bn_eval(x, p, 'upsample.resnet.batch_norm'), F32(0))
    i = 0
    while f'upsample.resnet.layers.{i}.conv1.weight' in p:
        pre = f'upsample.resnet.layers.{i}'
        res = x
        y = np.einsum('co,bot->bct', p[pre + '.conv1.weight'][:, :, 0].astype(F32), x).astype(F32)
        y = np.maximum(_bn_eval(y, p, pre + '.batch_norm1'), F32(0))
        y = np.einsum('co,bot->bct', p[pre + '.conv2.weight'][:, :, 0].astype(F32), y).astype(F32)
        y = _bn_eval(y, p, pre + '.batch_norm2')
        x = (y + res).astype(F32)
        i += 1
    x = np.einsum('co,bot->bct', p['upsample.resnet.conv_out.weight'][:, :, 0].astype(F32), x).astype(F32)
    x = x + p['upsample.resnet.conv_out.bias'].astype(F32)[None, :, None]
    return x.astype(F32)


def stretch_conv(p, m, scales):
    """The three Stretch2d + Conv2d(1,1,(1,2s+1),padding=(0,s)) stages, fatchord_version.py:57-61,73-80,86-87.

    m: [B, feat, Tp] -> [B, feat, Tp*prod(scales)] (before the indent trim of :88).
    """
    x = m.astype(F32)
    for j, s in enumerate(scales):
        w = p[f'upsample.up_layers.{2 * j + 1}.weight'].reshape(-1).astype(F32)   # [2s+1]
        x = np.repeat(x, s, axis=2)                                              # Stretch2d(s, 1)
        L = x.shape[2]
        xp = np.zeros(x.shape[:2] + (L + 2 * s,), dtype=F32)
        xp[:, :, s:s + L] = x
        y = np.zeros_like(x)
        for kk in range(2 * s + 1):                                              # cross-correlation, zero padded
            y += w[kk] * xp[:, :, kk:kk + L]
        x = y.astype(F32)
    return x


def upsample(p, mel_padded):
    """UpsampleNetwork.forward, fatchord_version.py:82-89.

    mel_padded: [B, feat, T+2*pad] -> (mels [B, T*hop, feat], aux [B, T*hop, res_out]).
    """
    d = _dims(p)
    aux = mel_resnet(p, mel_padded.astype(F32))                      # [B, O, T]
    aux = np.repeat(aux, d['hop'], axis=2)                           # resnet_stretch (:84)
    m = stretch_conv(p, mel_padded, d['scales'])
    indent = d['pad'] * d['hop']
    m = m[:, :, indent:-indent]
    return np.ascontiguousarray(m.transpose(0, 2, 1)), np.ascontiguousarray(aux.transpose(0, 2, 1))


def aux_frames(p, mel_padded):
    """Frame-rate output of MelResNet, [B, T, res_out] (aux is constant within a hop)."""
    return np.ascontiguousarray(mel_resnet(p, mel_padded.astype(F32)).transpose(0, 2, 1))


# ----------------------------------------------------------------------------------------------
# per-sample step  (fatchord_version.py:201-237)
# ----------------------------------------------------------------------------------------------
def _sigmoid(x):
    return (F32(1.0) / (F32(1.0) + np.exp(-x, dtype=F32))).astype(F32)


def gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    """torch.nn.GRUCell (gate rows r,z,n), the cell `get_gru_cell` builds at fatchord_version.py:273-279."""
    H = h.shape[1]
    gi = (x @ w_ih.T + b_ih).astype(F32)
    gh = (h @ w_hh.T + b_hh).astype(F32)
    r = _sigmoid(gi[:, :H] + gh[:, :H])
    z = _sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = np.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:], dtype=F32)
    return ((F32(1.0) - z) * n + z * h).astype(F32)


def step_logits(p, x, m_t, a1, a2, a3, a4, h1, h2):
    """One iteration of the hot loop up to the logits, fatchord_version.py:208-223.

    x [B,1], m_t [B,feat], a* [B,aux], h* [B,rnn] -> (logits [B,ncls], h1', h2').
    """
    xi = np.concatenate([x, m_t, a1], axis=1).astype(F32)
    xi = (xi @ p['I.weight'].T + p['I.bias']).astype(F32)                                   # :209
    h1 = gru_cell(xi, h1, p['rnn1.weight_ih_l0'], p['rnn1.weight_hh_l0'],
                  p['rnn1.bias_ih_l0'], p['rnn1.bias_hh_l0'])                                # :210
    xx = (xi + h1).astype(F32)                                                              # :212
    h2 = gru_cell(np.concatenate([xx, a2], axis=1), h2, p['rnn2.weight_ih_l0'], p['rnn2.weight_hh_l0'],
                  p['rnn2.bias_ih_l0'], p['rnn2.bias_hh_l0'])                                # :213-214
    xx = (xx + h2).astype(F32)                                                              # :216
    f = np.maximum((np.concatenate([xx, a3], axis=1) @ p['fc1.weight'].T + p['fc1.bias']).astype(F32), F32(0))
    f = np.maximum((np.concatenate([f, a4], axis=1) @ p['fc2.weight'].T + p['fc2.bias']).astype(F32), F32(0))
    logits = (f @ p['fc3.weight'].T + p['fc3.bias']).astype(F32)                             # :223
    return logits, h1, h2


def softmax(logits):
    """F.softmax(logits, dim=1), fatchord_version.py:232."""
    e = np.exp(logits - logits.max(axis=1, keepdims=True), dtype=F32)
    return (e / e.sum(axis=1, keepdims=True, dtype=F32)).astype(F32)


def sample_race(posterior, q):
    """`Categorical(posterior).sample()` (fatchord_version.py:233-235) == torch.multinomial(p, 1, True)
    == argmax(p / q) with q ~ Exp(1); `q` is supplied so every implementation shares the noise."""
    return np.argmax((posterior / q).astype(F32), axis=1)


def label_to_float(label, ncls):
    """`2 * label.float() / (n_classes - 1.) - 1.` in float32 (fatchord_version.py:235)."""
    return (F32(2.0) * label.astype(F32) / F32(ncls - 1.0) - F32(1.0)).astype(F32)


def decode_mu_law(y, mu):
    """wavernn/utils/dsp.py:98-103 with from_labels=False; y float64 in [-1, 1]."""
    mu = mu - 1
    return np.sign(y) / mu * ((1 + mu) ** np.abs(y) - 1)


# ----------------------------------------------------------------------------------------------
# generate  (fatchord_version.py:169-264, unbatched branch)
# ----------------------------------------------------------------------------------------------
def generate(p, mels, q=None, teacher=None, keep_logits=None, mu_law=True, seed=0, max_steps=None, cond=None):
    """WaveRNN.generate restated for a batch of independent utterances.

    mels      [B, feat, T] float32 in [0,1]
    q         [S, B, ncls] Exp(1) noise; None -> drawn from numpy RandomState(seed)
    teacher   optional [B, S] labels fed back instead of the sampled ones
    keep_logits  None | 'all' | iterable of step indices -> logits kept for those steps
    cond      optional (m_up [B,S,feat], aux [B,S,4*aux]) from a previous `upsample` call: skips the conditioning network
              (bench.py's CPU arm times many short runs on the same conditioning)
    Returns dict(labels [B,S] int16, wave [B, wave_len] float64, logits {step: [B,ncls]}, seconds, loop_seconds).
    The reference returns only utterance 0 (:253); this returns every row.
    """
    d = _dims(p)
    pf = {k: (v.astype(F32) if v.dtype.kind == 'f' else v) for k, v in p.items()}
    mels = np.asarray(mels, dtype=F32)
    B, _, T = mels.shape
    hop = d['hop']
    wave_len = (T - 1) * hop                                                                # :184
    t0 = time.perf_counter()
    if cond is None:
        mp = pad_tensor(mels.transpose(0, 2, 1), d['pad'], 'both').transpose(0, 2, 1)       # :185
        m_up, aux = upsample(pf, mp)                                                        # :186
    else:
        m_up, aux = cond
    t_loop = time.perf_counter()
    S = m_up.shape[1] if max_steps is None else min(max_steps, m_up.shape[1])
    A = d['aux']
    h1 = np.zeros((B, d['rnn']), dtype=F32)                                                 # :194-196
    h2 = np.zeros((B, d['rnn']), dtype=F32)
    x = np.zeros((B, 1), dtype=F32)
    labels = np.zeros((B, S), dtype=np.int16)
    rs = np.random.RandomState(seed) if q is None else None
    keep = None if keep_logits is None else (set(range(S)) if keep_logits == 'all' else set(keep_logits))
    kept = {}
    for i in range(S):                                                                      # :201
        a = aux[:, i, :]
        logits, h1, h2 = step_logits(pf, x, m_up[:, i, :], a[:, :A], a[:, A:2 * A],
                                     a[:, 2 * A:3 * A], a[:, 3 * A:4 * A], h1, h2)
        if keep is not None and i in keep:
            kept[i] = logits.copy()
        qi = q[i] if q is not None else np.maximum(
            rs.standard_exponential(size=logits.shape).astype(F32), F32(1e-30))
        lab = sample_race(softmax(logits), qi)                                              # :232-235
        labels[:, i] = lab
        fb = teacher[:, i].astype(np.int64) if teacher is not None else lab
        x = label_to_float(fb, d['ncls'])[:, None]                                          # :235-237
    seconds = time.perf_counter() - t0
    loop_seconds = time.perf_counter() - t_loop
    wave = finish_wave(labels, d['ncls'], wave_len, hop, mu_law) if S == m_up.shape[1] else None
    return dict(labels=labels, wave=wave, logits=kept, seconds=seconds, loop_seconds=loop_seconds, steps=S)


def fold_with_overlap(x, target, overlap):
    """fatchord_version.py:293-340.  x [1, S, F] -> [num_folds, target + 2*overlap, F]."""
    _, total_len, feats = x.shape
    num_folds = (total_len - overlap) // (target + overlap)
    extended_len = num_folds * (overlap + target) + overlap
    remaining = total_len - extended_len
    if remaining != 0:
        num_folds += 1
        padding = target + 2 * overlap - remaining
        x = pad_tensor(x, padding, side='after')
    folded = np.zeros((num_folds, target + 2 * overlap, feats), dtype=x.dtype)
    for i in range(num_folds):
        start = i * (target + overlap)
        folded[i] = x[0, start:start + target + 2 * overlap, :]
    return folded


def xfade_and_unfold(y, target, overlap):
    """fatchord_version.py:342-405.  y float64 [num_folds, target + 2*overlap] (modified in place like the reference)."""
    num_folds, length = y.shape
    target = length - 2 * overlap
    total_len = num_folds * (target + overlap) + overlap
    silence_len = overlap // 2
    fade_len = overlap - silence_len
    t = np.linspace(-1, 1, fade_len, dtype=np.float64)
    fade_in = np.concatenate([np.zeros(silence_len), np.sqrt(0.5 * (1 + t))])
    fade_out = np.concatenate([np.ones(silence_len), np.sqrt(0.5 * (1 - t))])
    y[:, :overlap] *= fade_in
    y[:, -overlap:] *= fade_out
    unfolded = np.zeros(total_len, dtype=np.float64)
    for i in range(num_folds):
        start = i * (target + overlap)
        unfolded[start:start + target + 2 * overlap] += y[i]
    return unfolded


def generate_batched(p, mel, target, overlap, q=None, mu_law=True, seed=0, teacher=None, keep_logits=(), max_steps=None):
    """WaveRNN.generate(batched=True), fatchord_version.py:169-264: ONE utterance mel [1, feat, T]; q [L, num_folds, ncls].
    Returns dict(labels [num_folds, L], wave [wave_len] float64, seconds)."""
    d = _dims(p)
    pf = {k: (v.astype(F32) if v.dtype.kind == 'f' else v) for k, v in p.items()}
    mel = np.asarray(mel, dtype=F32)
    assert mel.shape[0] == 1
    T = mel.shape[2]
    hop = d['hop']
    wave_len = (T - 1) * hop
    t0 = time.perf_counter()
    mp = pad_tensor(mel.transpose(0, 2, 1), d['pad'], 'both').transpose(0, 2, 1)
    m_up, aux = upsample(pf, mp)
    m_up = fold_with_overlap(m_up, target, overlap)                                          # :189
    aux = fold_with_overlap(aux, target, overlap)                                            # :190
    B, S, _ = m_up.shape
    A = d['aux']
    h1 = np.zeros((B, d['rnn']), dtype=F32)
    h2 = np.zeros((B, d['rnn']), dtype=F32)
    x = np.zeros((B, 1), dtype=F32)
    labels = np.zeros((B, S), dtype=np.int16)
    rs = np.random.RandomState(seed) if q is None else None
    kept = {}
    for i in range(S if max_steps is None else min(S, max_steps)):
        a = aux[:, i, :]
        logits, h1, h2 = step_logits(pf, x, m_up[:, i, :], a[:, :A], a[:, A:2 * A], a[:, 2 * A:3 * A], a[:, 3 * A:4 * A], h1, h2)
        if i in keep_logits:
            kept[i] = logits.copy()
        qi = q[i] if q is not None else np.maximum(rs.standard_exponential(size=logits.shape).astype(F32), F32(1e-30))
        lab = sample_race(softmax(logits), qi)
        labels[:, i] = lab
        fb = teacher[:, i].astype(np.int64) if teacher is not None else lab
        x = label_to_float(fb, d['ncls'])[:, None]
    out = label_to_float(labels, d['ncls']).astype(np.float64)
    if mu_law:
        out = decode_mu_law(out, d['ncls'])
    out = xfade_and_unfold(out, target, overlap)                                             # :251
    fade = np.linspace(1, 0, 20 * hop)
    out = out[:wave_len].copy()
    out[-20 * hop:] *= fade
    return dict(labels=labels, wave=out, seconds=time.perf_counter() - t0, logits=kept)


def finish_wave(labels, ncls, wave_len, hop, mu_law=True):
    """generate() epilogue, fatchord_version.py:243-258: float64, mu-law decode, truncate, 20-hop fade-out."""
    out = label_to_float(np.asarray(labels), ncls).astype(np.float64)                      # :243-245
    if mu_law:
        out = decode_mu_law(out, ncls)                                                      # :248
    fade = np.linspace(1, 0, 20 * hop)                                                      # :256
    out = out[:, :wave_len].copy()                                                          # :257
    out[:, -20 * hop:] *= fade                                                              # :258 (needs T >= 21)
    return out


def teacher_forced_logits(p, x, mels):
    """WaveRNN.forward (fatchord_version.py:131-167): logits for a GIVEN sample sequence.

    x [B, S] float32 samples fed at each step (x[:,0] is the first input, normally 0), mels [B, feat, T+2*pad]
    already padded (forward() does not pad).  Returns [B, S, ncls].
    """
    d = _dims(p)
    pf = {k: (v.astype(F32) if v.dtype.kind == 'f' else v) for k, v in p.items()}
    m_up, aux = upsample(pf, np.asarray(mels, dtype=F32))
    B, S = x.shape
    A = d['aux']
    h1 = np.zeros((B, d['rnn']), dtype=F32)
    h2 = np.zeros((B, d['rnn']), dtype=F32)
    out = np.zeros((B, S, d['ncls']), dtype=F32)
    for i in range(S):
        a = aux[:, i, :]
        logits, h1, h2 = step_logits(pf, x[:, i:i + 1].astype(F32), m_up[:, i, :], a[:, :A], a[:, A:2 * A],
                                     a[:, 2 * A:3 * A], a[:, 3 * A:4 * A], h1, h2)
        out[:, i] = logits
    return out
