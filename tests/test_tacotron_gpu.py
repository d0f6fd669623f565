"""sm_100a Tacotron-2 decoder loop (taco_decoder_kernel, through the C ABI) against the oracle with shared dropout masks.
The oracle itself is parity-UNPINNED against the TF reference (see oracle/tacotron_oracle.py); tolerance from the north
star: mel within 1e-4 abs, identical stop step."""
import numpy as np
import pytest

from oracle import tacotron_oracle as to
from taco_common import real_taco_weights, sentences, synth_taco_weights

pytestmark = pytest.mark.gpu


def _engine(w):
    from tacotronv2_wavernn_chinese_b200.tacotron.engine import TacoDecoderEngine
    return TacoDecoderEngine(w)


def _masks(seed, B, steps):
    return (np.random.RandomState(seed).uniform(size=(B, steps, 2, 256)) >= 0.5).astype(np.uint8)


@pytest.mark.parametrize('window', [False, True])
def test_decoder_vs_oracle_synthetic_weights(window):
    w = synth_taco_weights(7)
    eng = _engine(w)
    rs = np.random.RandomState(11)
    B, Tx, steps = 3, 37, 48
    mem = rs.uniform(-1, 1, (B, Tx, 512)).astype(np.float32)
    lengths = np.array([37, 20, 29], dtype=np.int32)
    masks = _masks(5, B, steps)
    out = eng.decode(mem, lengths, masks=masks, max_steps=steps, window=window)
    fr, st, al, ns = (out[k].cpu().numpy() for k in ('frames', 'stop', 'align', 'nsteps'))
    for b in range(B):
        ref = to.decode(w, mem[b, :lengths[b]], dropout_masks=masks[b], max_iters=steps, window=window)
        n = ref['n_steps']
        assert ns[b] == n
        np.testing.assert_allclose(fr[b, :n], ref['frames'], rtol=0, atol=1e-4)
        np.testing.assert_allclose(st[b, :n], ref['stop'], rtol=0, atol=1e-5)
        np.testing.assert_allclose(al[b, :n, :lengths[b]], ref['alignments'], rtol=0, atol=1e-5)
        assert np.all(al[b, :n, lengths[b]:] == 0)


HORIZON = 60     # steps over which fp32 evaluations of the shipped checkpoint still agree to 1e-4 (CPU test
                 # test_real_checkpoint_decoder_is_chaotic: fp32-vs-fp64 ORACLE error 2e-5 @80, 3.5e-4 @120, O(1) by 300;
                 # measured on B200, tools/taco_err_profile.py: kernel-vs-fp64 <= 3.4e-5 to step 60, 5e-4 @79, 3.8e-4 @150)
HORIZON2 = 150   # ... and to 2e-3


def test_decoder_vs_oracle_real_checkpoint_config4():
    """BASELINE config 4: 50-token pinyin sentence (train.txt line 241), shipped checkpoint, seed-1238 dropout masks.
    The north star asks for mel within 1e-4 and an identical stop step; the shipped decoder amplifies rounding noise
    exponentially after ~150 steps (the oracle disagrees with ITSELF in float64 by then), so 1e-4 is asserted over the
    first HORIZON steps, 2e-3 over the first HORIZON2, and the full run is checked qualitatively (monotone alignment, stop step within 5 %)."""
    w = real_taco_weights()
    if w is None:
        pytest.skip('shipped Tacotron checkpoint not available on this box')
    ids = sentences()['sentences']['241']['ids']
    mem = to.encoder(w, ids)                       # encoder is a 'next' row: the oracle prepares the decoder's INPUT here
    steps = 700
    masks = _masks(1238, 1, steps)
    ref = to.decode(w, mem, dropout_masks=masks[0], max_iters=steps)
    out = _engine(w).decode(mem[None], masks=masks, max_steps=steps)
    n = int(out['nsteps'][0])
    fr = out['frames'].cpu().numpy()[0, :n]
    al = out['align'].cpu().numpy()[0, :n]
    err = np.abs(fr[:HORIZON] - ref['frames'][:HORIZON]).max()
    assert err <= 1e-4, err
    np.testing.assert_allclose(out['stop'].cpu().numpy()[0, :HORIZON], ref['stop'][:HORIZON], rtol=0, atol=1e-5)
    assert np.array_equal(al[:HORIZON].argmax(1), ref['alignments'][:HORIZON].argmax(1))
    mel_gpu, mel_ref = to.postnet(w, fr[:HORIZON]), to.postnet(w, ref['frames'][:HORIZON])
    assert np.abs(mel_gpu[:HORIZON - 4] - mel_ref[:HORIZON - 4]).max() <= 1e-4
    assert np.abs(fr[:HORIZON2] - ref['frames'][:HORIZON2]).max() <= 2e-3
    assert np.array_equal(al[:HORIZON2].argmax(1), ref['alignments'][:HORIZON2].argmax(1))
    # beyond the horizon: same qualitative behaviour
    assert abs(n - ref['n_steps']) <= 0.05 * ref['n_steps'], (n, ref['n_steps'])
    path = al.argmax(1)
    assert path[-1] >= 48 and (np.diff(path) >= -1).all()
    assert out['stop'].cpu().numpy()[0, n - 1] > 0.5


def test_philox_dropout_replay_and_batch_invariance():
    w = synth_taco_weights(7)
    eng = _engine(w)
    rs = np.random.RandomState(3)
    B, Tx, steps = 4, 30, 32
    mem = rs.uniform(-1, 1, (B, Tx, 512)).astype(np.float32)
    masks = eng.philox_masks(77, 10, B, steps).cpu().numpy()
    assert 0.45 < masks.mean() < 0.55
    a = eng.decode(mem, seed=77, utterance_offset=10, max_steps=steps)
    b = eng.decode(mem, masks=masks, max_steps=steps)
    assert np.array_equal(a['frames'].cpu().numpy(), b['frames'].cpu().numpy())
    solo = eng.decode(mem[2:3], seed=77, utterance_offset=12, max_steps=steps)      # keyed by the GLOBAL sentence index
    assert np.array_equal(solo['frames'].cpu().numpy()[0], a['frames'].cpu().numpy()[2])


def test_decoder_error_behaviour():
    from tacotronv2_wavernn_chinese_b200._lib import B200TTSError
    from tacotronv2_wavernn_chinese_b200.tacotron.engine import TacoDecoderEngine
    w = synth_taco_weights(7)
    bad = dict(w)
    del bad['decoder/dense/kernel']
    with pytest.raises(B200TTSError) as e:
        TacoDecoderEngine(bad)
    assert e.value.code == -4
    eng = _engine(w)
    with pytest.raises(ValueError):
        eng.decode(np.zeros((1, 5, 100), np.float32))
    with pytest.raises(B200TTSError):
        eng.decode(np.zeros((1, 600, 512), np.float32), max_steps=4)        # Tx_max > 512
