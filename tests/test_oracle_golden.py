"""The numpy oracle (oracle/wavernn_oracle.py) against fixtures produced by the REFERENCE ITSELF
(oracle/make_golden_wavernn.py).  CPU only."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import wavernn_oracle as wo
from tacotronv2_wavernn_chinese_b200 import synth


def _load(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'))


def _params(case, ckpt_state_dict=None):
    if case.startswith('wavernn_synth'):
        return wo.as_params(synth.synth_state_dict(11))
    return wo.as_params(ckpt_state_dict)


def _padded(mels):
    B, F, T = mels.shape
    mp = np.zeros((B, F, T + 4), dtype=np.float32)
    mp[:, :, 2:-2] = mels
    return mp


@pytest.mark.parametrize('case', ['wavernn_synth_T24', 'wavernn_ckpt_T24'])
def test_upsample_matches_reference(case, request):
    g = _load(case)
    p = _params(case, request.getfixturevalue('ckpt_state_dict') if 'ckpt' in case else None)
    mels = synth.synth_mels(int(g['mel_seed']), int(g['B']), int(g['T']))
    up, aux = wo.upsample(p, _padded(mels))
    st = int(g['up_stride'])
    # tolerance: fp32 re-association only (the reference runs the same fp32 ops through ATen)
    np.testing.assert_allclose(up[:, ::st], g['mels_up_sub'], rtol=0, atol=2e-6)
    np.testing.assert_allclose(aux[:, ::275], g['aux_frames'], rtol=0, atol=5e-5)
    np.testing.assert_allclose(wo.aux_frames(p, _padded(mels)), g['aux_frames'], rtol=0, atol=5e-5)


@pytest.mark.parametrize('case', ['wavernn_synth_T24', 'wavernn_ckpt_T24'])
def test_generate_matches_reference(case, request):
    g = _load(case)
    p = _params(case, request.getfixturevalue('ckpt_state_dict') if 'ckpt' in case else None)
    B, T = int(g['B']), int(g['T'])
    mels = synth.synth_mels(int(g['mel_seed']), B, T)
    q = synth.synth_exponential_noise(int(g['noise_seed']), T * 275, B)
    steps = [int(s) for s in g['logit_steps']]
    # (1) teacher-forced on the reference's own label sequence: logits must agree at every recorded step.
    #     Floor: the reference's nn.GRU-vs-nn.GRUCell disagreement is ~1e-3 abs at |logit|~450.
    r = wo.generate(p, mels, q=q, teacher=g['labels'], keep_logits=steps)
    got = np.stack([r['logits'][s] for s in steps])
    scale = max(1.0, float(np.abs(g['gen_logits']).max()))
    assert np.abs(got - g['gen_logits']).max() <= 5e-6 * scale + 1e-5
    assert np.abs(got - g['fwd_logits']).max() <= 1e-5 * scale + 1e-5
    # sampled labels under teacher forcing == reference labels (same logits, same noise)
    mism = np.nonzero(r['labels'] != g['labels'])
    assert mism[0].size <= 2, f'{mism[0].size} label mismatches under teacher forcing'
    # (2) free running with the shared noise: identical label sequence, identical wave for utterance 0
    r = wo.generate(p, mels, q=q)
    assert np.array_equal(r['labels'], g['labels'])
    np.testing.assert_allclose(r['wave'][0], g['wave0'], rtol=0, atol=1e-12)


def test_generate_config1_shape_labels(ckpt_state_dict):
    """BASELINE config 1 shape (80 frames, 22 000 steps) on the shipped checkpoint: free-running labels."""
    g = _load('wavernn_ckpt_T80')
    p = wo.as_params(ckpt_state_dict)
    mels = synth.synth_mels(int(g['mel_seed']), 1, 80)
    q = synth.synth_exponential_noise(int(g['noise_seed']), 80 * 275, 1)
    r = wo.generate(p, mels, q=q)
    same = r['labels'] == g['labels']
    first = int(np.argmin(same[0])) if not same.all() else same.shape[1]
    assert first == same.shape[1], f'diverged from the reference at step {first}'
    np.testing.assert_allclose(r['wave'][0], g['wave0'], rtol=0, atol=1e-12)


@pytest.mark.parametrize('case', ['wavernn_synth_batched_T30', 'wavernn_ckpt_batched_T30'])
def test_batched_generate_matches_reference(case, request):
    """generate(batched=True): fold_with_overlap + xfade_and_unfold (fatchord_version.py:293-405) vs the reference's own run."""
    g = _load(case)
    p = _params('wavernn_synth' if 'synth' in case else 'ckpt', request.getfixturevalue('ckpt_state_dict') if 'ckpt' in case else None)
    T, target, overlap = int(g['T']), int(g['target']), int(g['overlap'])
    mel = synth.synth_mels(int(g['mel_seed']), 1, T)
    nf, L = g['labels'].shape
    q = synth.synth_exponential_noise(int(g['noise_seed']), L, nf)
    r = wo.generate_batched(p, mel, target, overlap, q=q)
    assert np.array_equal(r['labels'], g['labels'])
    np.testing.assert_allclose(r['wave'], g['wave'], rtol=0, atol=1e-12)


def test_finish_wave_requires_21_frames():
    with pytest.raises(ValueError):
        wo.finish_wave(np.zeros((1, 20 * 275), dtype=np.int16), 1024, 19 * 275, 275)
