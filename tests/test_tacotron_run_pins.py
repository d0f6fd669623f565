"""WHOLE-RUN pin of the Tacotron-2 decoder restatement (CPU suite).

tests/golden/taco_run_from_graph.npz is the trajectory obtained by carrying the loop state through the reference's own
serialized decoder-step graph (`CustomDecoderStep` inside `tacotron_model.ckpt-206500.meta`, evaluated op by op in numpy)
for the WHOLE config-4 sentence (train.txt line 241, 405 steps, prenet keep-masks of seed 1238) -- never passing through
oracle.tacotron_oracle.decoder_step (oracle/make_golden_taco_run.py).  oracle.tacotron_oracle.decode must reproduce every
frame, every attended position and the stop step.  The decoder is chaotic (test_tacotron_oracle.py), so a single differing
rounding anywhere would show as an O(1) difference a hundred steps later: the comparison is over all 405 steps.
"""
import os
import warnings

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import tacotron_oracle as to
from taco_common import real_taco_weights


def _fixture():
    z = np.load(os.path.join(GOLDEN, 'taco_run_from_graph.npz'))
    n = int(z['n_steps'])
    masks = np.unpackbits(z['masks'], axis=1)[:, :512].reshape(n, 2, 256).astype(np.float32)
    return z, n, masks


def test_fixture_is_a_full_sentence():
    z, n, masks = _fixture()
    assert z['frames'].shape == (n, 80) and n == 405 and z['stop'].shape == (n,)
    assert (z['stop'][:-1] <= 0.5).all() and z['stop'][-1] > 0.5                 # stops exactly once, at the last step
    assert 0.4 < masks.mean() < 0.6
    am = z['align_argmax'].astype(int)
    assert am[0] <= 1 and am[-1] >= len(z['ids']) - 3 and (np.diff(am) >= -1).all()      # attention walks the sentence to its end


def test_whole_run_reproduces_the_serialized_reference_graph():
    w = real_taco_weights()
    if w is None:
        pytest.skip('shipped Tacotron checkpoint not available on this box')
    z, n, masks = _fixture()
    memory = to.encoder(w, z['ids'])
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)                          # exp overflow inside the sigmoid of large logits
        d = to.decode(w, memory, dropout_masks=masks, max_iters=n + 50)
    m = min(n, d['n_steps'])
    err = np.abs(d['frames'][:m] - z['frames'][:m]).max(axis=1)
    # the first 40 steps are robust against a BLAS that rounds differently from the one the fixture was made with
    assert err[:40].max() <= 1e-5, float(err[:40].max())
    if err[:40].max() > 0:                                                       # another BLAS: chaos amplifies its roundings later on
        warnings.warn(f'numpy matmul differs from the fixture\'s in the last bit (max {err[:40].max():.1e} in 40 steps); '
                      'whole-run identity not asserted on this host')
        assert abs(d['n_steps'] - n) <= 0.05 * n
        return
    assert d['n_steps'] == n                                                     # identical stop step
    assert err.max() <= 1e-6, (int(err.argmax()), float(err.max()))              # measured: 0 over all 405 steps
    assert np.array_equal(d['alignments'].argmax(1), z['align_argmax'])
    np.testing.assert_allclose(d['alignments'].max(1), z['align_peak'], rtol=0, atol=1e-6)
    np.testing.assert_allclose(d['stop'], z['stop'], rtol=0, atol=1e-6)
    assert np.array_equal(d['masks'], masks)


def test_whole_run_fixture_is_sensitive():
    """Live comparison: the same run with the zoneout factor 0 instead of 0.1 leaves the fixture's trajectory within a few steps."""
    w = real_taco_weights()
    if w is None:
        pytest.skip('shipped Tacotron checkpoint not available on this box')
    z, n, masks = _fixture()
    memory = to.encoder(w, z['ids'])
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        d = to.decode(w, memory, dropout_masks=masks, max_iters=60, zoneout=0.0)
    assert np.abs(d['frames'][:60] - z['frames'][:60]).max() > 1e-2


def test_whole_encoder_and_whole_postnet_through_the_graph():
    """The run-once neighbours over the WHOLE sentence: all 51 tokens through both serialized encoder-LSTM loop bodies (inference
    zoneout between iterations), all 405 decoder frames through the serialized postnet (measured 3.9e-7 / 9.5e-7 absolute)."""
    w = real_taco_weights()
    if w is None:
        pytest.skip('shipped Tacotron checkpoint not available on this box')
    z, n, _ = _fixture()
    mem = to.encoder(w, z['ids'])
    assert mem.shape == z['memory_graph'].shape == (len(z['ids']), 512)
    assert np.abs(mem - z['memory_graph']).max() <= 2e-6 * max(1.0, float(np.abs(z['memory_graph']).max()))
    mel = to.postnet(w, z['frames'])
    assert mel.shape == z['mel_graph'].shape == (n, 80)
    assert np.abs(mel - z['mel_graph']).max() <= 2e-6 * max(1.0, float(np.abs(z['mel_graph']).max()))
    assert z['mel_graph'].min() >= -4.1 - 1e-6 and z['mel_graph'].max() <= 4.0 + 1e-6          # tacotron.py:126-129 clip


@pytest.mark.parametrize('key', ['48', '42'])
def test_whole_run_other_sentence_lengths(key):
    """The shortest (17 tokens, 133 steps) and the longest (52 tokens, 411 steps) sentence of train.txt 1-64, other mask streams:
    same method, trajectories only (tests/golden/taco_run_from_graph_more.npz; the masks are RandomState(seed) draws)."""
    w = real_taco_weights()
    if w is None:
        pytest.skip('shipped Tacotron checkpoint not available on this box')
    z = np.load(os.path.join(GOLDEN, 'taco_run_from_graph_more.npz'))
    assert key in list(z['sentences'])
    frames, ids = z[f's{key}_frames'], z[f's{key}_ids']
    n = frames.shape[0]
    masks = (np.random.RandomState(int(z[f's{key}_seed'])).uniform(size=(700, 2, 256)) >= 0.5).astype(np.float32)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        d = to.decode(w, to.encoder(w, ids), dropout_masks=masks, max_iters=700)
    m = min(n, d['n_steps'])
    err = np.abs(d['frames'][:m] - frames[:m]).max(axis=1)
    assert err[:40].max() <= 1e-5
    if err[:40].max() > 0:                      # a BLAS rounding differently from the fixture's: see the config-4 test above
        assert abs(d['n_steps'] - n) <= 0.05 * n
        return
    assert d['n_steps'] == n and err.max() <= 1e-6
    assert np.array_equal(d['alignments'].argmax(1), z[f's{key}_align_argmax'])
    np.testing.assert_allclose(d['stop'], z[f's{key}_stop'], rtol=0, atol=1e-6)


def test_whole_run_with_the_inference_window():
    """Config-4 sentence WITH the optional inference window (SURVEY a-9; forward_attention.py:171-215), 418 steps: cells and
    projections from the serialized graph, the attention step from the reference's OWN ForwardLocationSensitiveAttention.__call__
    executed on numpy arrays, state carried from step to step (oracle/make_golden_taco_run.py: drive_graph_window)."""
    w = real_taco_weights()
    if w is None:
        pytest.skip('shipped Tacotron checkpoint not available on this box')
    z, n0, masks = _fixture()
    zw = np.load(os.path.join(GOLDEN, 'taco_run_from_graph_more.npz'))
    frames = zw['w241_frames']
    n = frames.shape[0]
    assert n == 418 and np.abs(frames[:n0] - z['frames'][:min(n, n0)]).max() > 1.0      # the window changes the trajectory
    masks = (np.random.RandomState(int(z['seed'])).uniform(size=(700, 2, 256)) >= 0.5).astype(np.float32)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        d = to.decode(w, to.encoder(w, z['ids']), dropout_masks=masks, max_iters=700, window=True)
    m = min(n, d['n_steps'])
    err = np.abs(d['frames'][:m] - frames[:m]).max(axis=1)
    assert err[:40].max() <= 1e-5
    if err[:40].max() > 0:
        assert abs(d['n_steps'] - n) <= 0.05 * n
        return
    assert d['n_steps'] == n and err.max() <= 1e-6
    assert np.array_equal(d['alignments'].argmax(1), zw['w241_align_argmax'])
    mx = zw['w241_max_att'].astype(int)
    assert (np.diff(mx) >= 0).all() and (np.diff(mx) <= 1).all()                        # monotone, at most one token per step
