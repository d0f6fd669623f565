"""Import and drive the UNMODIFIED reference WaveRNN from /root/reference -- container-only test tooling.

Used by oracle/make_golden_wavernn.py (golden fixtures) and by tests that are
skipped when /root/reference is absent (it does not exist on the GPU box).
Nothing here is copied from the reference: the reference package is imported
in-process with three shims: stub `matplotlib` and
`librosa` (absent here, unused on this path), and `np.cumproduct` (removed in
numpy 2, used at wavernn/models/fatchord_version.py:68).
"""
from __future__ import annotations

import os
import sys
import types
import contextlib

import numpy as np
import torch

REF_ROOT = os.environ.get('B200TTS_REFERENCE', '/root/reference')
REF_CKPT = os.path.join(REF_ROOT, 'logs_wavernn/checkpoints/latest_weights.pyt')


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, 'wavernn/models/fatchord_version.py'))


_fv = None


def import_reference():
    """Returns the reference module `wavernn.models.fatchord_version` (imported once)."""
    global _fv
    if _fv is not None:
        return _fv
    if not available():
        raise RuntimeError(f'reference not found under {REF_ROOT}')
    mpl = types.ModuleType('matplotlib')
    mpl.use = lambda *a, **k: None
    mpl.interactive = lambda *a, **k: None
    plt = types.ModuleType('matplotlib.pyplot')
    mpl.pyplot = plt
    for name, mod in (('matplotlib', mpl), ('matplotlib.pyplot', plt), ('librosa', types.ModuleType('librosa'))):
        sys.modules.setdefault(name, mod)
    if not hasattr(np, 'cumproduct'):
        np.cumproduct = np.cumprod
    # our own drop-in package is also called `wavernn` at some call sites; make sure the reference wins here
    for k in [k for k in sys.modules if k == 'wavernn' or k.startswith('wavernn.')]:
        del sys.modules[k]
    sys.path.insert(0, REF_ROOT)
    try:
        from wavernn.utils import hparams as hp
        if not hp.is_configured():
            hp.configure(os.path.join(REF_ROOT, 'wavernn_hparams.py'))
        from wavernn.models import fatchord_version as fv
    finally:
        sys.path.remove(REF_ROOT)
    fv.save_wav = lambda x, p: None          # librosa.output.write_wav no longer exists (dsp.py:23)
    fv.stream = lambda *a, **k: None         # silence the \r progress line (display.py:17)
    fv._hp = hp
    _fv = fv
    return fv


def build_model(state_dict=None):
    """Reference WaveRNN with the shipped hparams; loads `state_dict` (numpy/torch) or the shipped checkpoint."""
    fv = import_reference()
    hp = fv._hp
    with contextlib.redirect_stdout(open(os.devnull, 'w')):
        m = fv.WaveRNN(hp.voc_rnn_dims, hp.voc_fc_dims, hp.bits, hp.voc_pad, hp.voc_upsample_factors, hp.num_mels,
                       hp.voc_compute_dims, hp.voc_res_out_dims, hp.voc_res_blocks, hp.hop_length,
                       hp.sample_rate, hp.voc_mode)
    if state_dict is None:
        m.load(REF_CKPT)
    else:
        sd = {k: torch.as_tensor(np.asarray(v)) for k, v in state_dict.items()}
        missing = m.load_state_dict(sd, strict=False)
        assert not missing.missing_keys, missing
    return m


class _RaceCategorical:
    """Stand-in for torch.distributions.Categorical whose sample() is `argmax(p / q_step)` with injected q.

    torch.multinomial(p, 1, replacement=True) on CPU is exactly `argmax(p / Exp(1))`
    (SURVEY.md 7.3-2); feeding q from a tensor lets the reference, the oracle and the
    CUDA kernels share the noise.
    """
    q = None          # torch [S, B, ncls]
    step = 0
    labels = None     # list of [B] tensors

    def __init__(self, probs):
        self.p = probs

    def sample(self):
        cls = type(self)
        lab = (self.p / cls.q[cls.step]).argmax(-1)
        cls.step += 1
        cls.labels.append(lab.clone())
        return lab


def reference_generate(model, mels, q, capture_logits_at=(), batched=False, target=None, overlap=None):
    """Runs the reference `generate` (fatchord_version.py:169) unbatched with injected noise.

    Returns dict(wave0 float64 [wave_len] (utterance 0 only -- :253), labels [B,S] int16, logits {step: [B,ncls]}).
    """
    fv = import_reference()
    want = set(int(s) for s in capture_logits_at)
    kept = {}
    counter = {'i': 0}

    def hook(_mod, _inp, out):
        if counter['i'] in want:
            kept[counter['i']] = out.detach().numpy().copy()
        counter['i'] += 1

    h = model.fc3.register_forward_hook(hook)
    _RaceCategorical.q = torch.as_tensor(q)
    _RaceCategorical.step = 0
    _RaceCategorical.labels = []
    orig = torch.distributions.Categorical
    torch.distributions.Categorical = _RaceCategorical
    try:
        hp = fv._hp
        wave = model.generate(torch.as_tensor(mels), '/dev/null', bool(batched), target if target is not None else hp.voc_target,
                              overlap if overlap is not None else hp.voc_overlap, hp.mu_law)
    finally:
        torch.distributions.Categorical = orig
        h.remove()
    labels = torch.stack(_RaceCategorical.labels, 1).numpy().astype(np.int16)
    return dict(wave0=wave, labels=labels, logits=kept)


def reference_upsample(model, mels):
    """pad_tensor + UpsampleNetwork.forward exactly as generate() calls them (fatchord_version.py:185-186)."""
    model.eval()
    with torch.no_grad():
        m = torch.as_tensor(mels)
        mp = model.pad_tensor(m.transpose(1, 2), pad=model.pad, side='both')
        up, aux = model.upsample(mp.transpose(1, 2))
    return up.numpy(), aux.numpy()


def reference_forward_logits(model, x, mels_padded):
    """Teacher-forced `WaveRNN.forward` (fatchord_version.py:131-167); restores the `step` side effect (:139)."""
    model.eval()
    step = model.step.clone()
    with torch.no_grad():
        out = model(torch.as_tensor(x), torch.as_tensor(mels_padded)).numpy()
    model.step.copy_(step)
    return out
