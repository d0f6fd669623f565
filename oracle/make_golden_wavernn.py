"""Generate tests/golden/wavernn_*.npz by running the UNMODIFIED reference from /root/reference.

Run in the build container (the reference is not present on the GPU box):

    python oracle/make_golden_wavernn.py

Cases
  wavernn_synth_T24.npz   synthetic weights (tacotronv2_wavernn_chinese_b200.synth, seed 11) -- fully portable
  wavernn_ckpt_T24.npz    the shipped checkpoint latest_weights.pyt, 24-frame mel, B=2
  wavernn_ckpt_T80.npz    the shipped checkpoint, BASELINE config 1 shape (80-frame mel, 22 000 steps), B=1

Each file holds inputs' SEEDS (inputs are regenerated from numpy RandomState) and the
reference's outputs: sub-sampled upsample outputs, free-running labels under injected
Exp(1) race noise, fc3 logits at selected steps (from inside generate()), the
teacher-forced logits of WaveRNN.forward at the same steps, and the returned wave.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh                                   # noqa: E402
from tacotronv2_wavernn_chinese_b200 import synth                      # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
UP_STRIDE = 13       # time stride for the stored upsampled mels (coprime with 275 -> covers every phase)


def check_race_equals_multinomial(model):
    """The injected-noise stand-in must reproduce the stock Categorical.sample() stream bit for bit."""
    mel = synth.synth_mels(5, 1, 21)
    hp = rh.import_reference()._hp
    torch.manual_seed(1234)
    stock = model.generate(torch.as_tensor(mel), '/dev/null', False, hp.voc_target, hp.voc_overlap, hp.mu_law)
    torch.manual_seed(1234)
    # generate() first builds two nn.GRUCell objects (fatchord_version.py:178-179) whose random
    # initialisation consumes the global generator before any sampling noise is drawn
    model.get_gru_cell(model.rnn1), model.get_gru_cell(model.rnn2)
    S = 21 * 275
    q = torch.stack([torch.empty(1, 1024).exponential_(1) for _ in range(S)])
    got = rh.reference_generate(model, mel, q)
    assert np.array_equal(stock, got['wave0']), 'race sampler != torch.multinomial stream'
    print('  race-sampler == stock multinomial: OK')


def make_case(name, model, mel_seed, noise_seed, B, T, logit_steps, store_upsample=True):
    mels = synth.synth_mels(mel_seed, B, T)
    S = T * 275
    q = synth.synth_exponential_noise(noise_seed, S, B)
    res = rh.reference_generate(model, mels, q, capture_logits_at=logit_steps)
    labels = res['labels']
    out = dict(mel_seed=mel_seed, noise_seed=noise_seed, B=B, T=T,
               labels=labels, wave0=res['wave0'], logit_steps=np.array(logit_steps, dtype=np.int64),
               gen_logits=np.stack([res['logits'][s] for s in logit_steps]))          # [n, B, 1024]
    # teacher-forced forward() on the same sample sequence: x_t = float(label_{t-1}), x_0 = 0
    x = np.zeros((B, S), dtype=np.float32)
    x[:, 1:] = (2.0 * labels[:, :-1].astype(np.float32) / np.float32(1023.0) - 1.0).astype(np.float32)
    mp = np.zeros((B, 80, T + 4), dtype=np.float32)
    mp[:, :, 2:-2] = mels
    fwd = rh.reference_forward_logits(model, x, mp)                                    # [B, S, 1024]
    out['fwd_logits'] = np.stack([fwd[:, s] for s in logit_steps])
    if store_upsample:
        up, aux = rh.reference_upsample(model, mels)
        out['up_stride'] = UP_STRIDE
        out['mels_up_sub'] = up[:, ::UP_STRIDE].copy()
        out['aux_frames'] = aux[:, ::275].copy()
        assert np.array_equal(np.repeat(out['aux_frames'], 275, axis=1), aux), 'aux not constant within a hop'
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **out)
    d = np.abs(out['gen_logits'] - out['fwd_logits']).max()
    print(f'  {name}: labels {labels.shape} range [{labels.min()},{labels.max()}] '
          f'max|gen-fwd logit|={d:.2e} max|logit|={np.abs(out["gen_logits"]).max():.1f} '
          f'-> {os.path.getsize(path) / 1e3:.0f} kB')


def make_batched_case(name, model, mel_seed, noise_seed, T, target, overlap):
    """generate(batched=True): fold-with-overlap of ONE utterance (fatchord_version.py:188-190, :250-251)."""
    mels = synth.synth_mels(mel_seed, 1, T)
    S = T * 275
    nf = (S - overlap) // (target + overlap)
    if S - (nf * (target + overlap) + overlap) != 0:
        nf += 1
    L = target + 2 * overlap
    q = synth.synth_exponential_noise(noise_seed, L, nf)
    res = rh.reference_generate(model, mels, q, batched=True, target=target, overlap=overlap)
    assert res['labels'].shape == (nf, L), res['labels'].shape
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, mel_seed=mel_seed, noise_seed=noise_seed, T=T, target=target, overlap=overlap,
                        labels=res['labels'], wave=res['wave0'])
    print(f'  {name}: folds {nf} x {L}, wave {res["wave0"].shape} -> {os.path.getsize(path) / 1e3:.0f} kB')


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    print('reference checkpoint model')
    ck = rh.build_model(None)
    check_race_equals_multinomial(ck)
    steps24 = [0, 1, 2, 3, 274, 275, 276] + list(range(500, 24 * 275, 487))
    make_case('wavernn_ckpt_T24', ck, mel_seed=101, noise_seed=201, B=2, T=24, logit_steps=steps24)
    steps80 = [0, 1, 2, 275, 5000, 10000, 15000, 21999]
    make_case('wavernn_ckpt_T80', ck, mel_seed=1234, noise_seed=202, B=1, T=80, logit_steps=steps80,
              store_upsample=False)
    make_batched_case('wavernn_ckpt_batched_T30', ck, mel_seed=103, noise_seed=204, T=30, target=2700, overlap=500)
    print('synthetic-weights model (portable)')
    sy = rh.build_model(synth.synth_state_dict(11))
    make_case('wavernn_synth_T24', sy, mel_seed=102, noise_seed=203, B=2, T=24, logit_steps=steps24)
    make_batched_case('wavernn_synth_batched_T30', sy, mel_seed=104, noise_seed=205, T=30, target=2750, overlap=550)


if __name__ == '__main__':
    main()
