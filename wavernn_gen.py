#!/usr/bin/env python
"""Vocode a mel `.npy` with WaveRNN on a B200 -- drop-in for the reference's `wavernn_gen.py`.

Same flags (reference wavernn_gen.py:49-61), same input contract (float32 `(T, 80)` mel in [0, 1],
:22-28), same output name `./wavernn_inference_output/{stem}_gen_NOT_BATCHED_step={k}k.wav` (:35-39, :124).
Differences, all deliberate:
  * the model runs on the GPU through libb200tts.so; the reference pins `device = cpu` (:93).  `--force_cpu`
    is rejected because this build has no CPU path.
  * `.wav` input is refused: in the reference that branch references undefined names (:19, :30) and cannot run.
  * extra flags: `--seed` (Philox stream), `--kernel {auto,utterance,grid}`.
"""
from __future__ import annotations

import argparse
import os
import time

import numpy as np
import torch

from tacotronv2_wavernn_chinese_b200.wavernn.models.fatchord_version import WaveRNN
from tacotronv2_wavernn_chinese_b200.wavernn.utils import hparams as hp
from tacotronv2_wavernn_chinese_b200.wavernn.utils.display import simple_table
from tacotronv2_wavernn_chinese_b200.wavernn.utils.paths import Paths


def gen_from_file(model: WaveRNN, load_path, save_path, batched, target, overlap, seed=None, kernel='auto'):
    k = model.get_step() // 1000
    load_path = str(load_path)
    if load_path.endswith('.npy'):
        mel = np.load(load_path).T
        if mel.ndim != 2 or mel.shape[0] != hp.num_mels:
            raise ValueError(f'Expected a numpy array shaped (n_hops, n_mels={hp.num_mels}), but got {mel.T.shape}!')
        _max, _min = np.max(mel), np.min(mel)
        if _max >= 1.01 or _min <= -0.01:
            raise ValueError(f'Expected spectrogram range in [0,1] but was instead [{_min}, {_max}]')
    else:
        raise ValueError(f'Expected a .npy mel spectrogram, but got {os.path.splitext(load_path)[1] or load_path}!')
    mel = torch.tensor(mel, dtype=torch.float32).unsqueeze(0)
    batch_str = f'gen_batched_target{target}_overlap{overlap}' if batched else 'gen_NOT_BATCHED'
    idx = load_path.split('/')[-1].strip().split('.')[0]
    save_str = os.path.join(str(save_path), f'{idx}_{batch_str}_step={k}k.wav')
    t0 = time.time()
    wav = model.generate(mel, save_str, batched, target, overlap, hp.mu_law, seed=seed, kernel=kernel)
    dt = time.time() - t0
    print(f'\n\nstep = {k * 1000}')
    print(f'{len(wav)} samples in {dt:.2f} s  ({len(wav) / dt / 1000:.1f} kHz, RTF {dt / (len(wav) / hp.sample_rate):.3f}) -> {save_str}')
    return save_str


def main(argv=None):
    parser = argparse.ArgumentParser(description='Generate WaveRNN Samples')
    parser.add_argument('--batched', '-b', dest='batched', action='store_true', help='Fast Batched Generation')
    parser.add_argument('--unbatched', '-u', dest='batched', action='store_false', help='Slow Unbatched Generation')
    parser.add_argument('--samples', '-s', type=int, help='[int] number of utterances to generate')
    parser.add_argument('--target', '-t', type=int, help='[int] number of samples in each batch index')
    parser.add_argument('--overlap', '-o', type=int, help='[int] number of crossover samples')
    parser.add_argument('--file', '-f', type=str, help='[string/path] mel spectrogram .npy to vocode')
    parser.add_argument('--voc_weights', '-w', type=str, help='[string/path] Load in different WaveRNN weights')
    parser.add_argument('--gta', '-g', dest='gta', action='store_true', help='Generate from GTA testset')
    parser.add_argument('--force_cpu', '-c', action='store_true', help='(rejected: this build has no CPU path)')
    parser.add_argument('--hp_file', metavar='FILE', default='wavernn_hparams.py', help='The file to use for the hyperparameters')
    parser.add_argument('--seed', type=int, default=None, help='[int] Philox seed of the sampling noise')
    parser.add_argument('--kernel', choices=('auto', 'utterance', 'grid'), default='auto', help='generation kernel')
    parser.set_defaults(batched=None)
    args = parser.parse_args(argv)

    hp.configure(args.hp_file)
    target = hp.voc_target if args.target is None else args.target
    overlap = hp.voc_overlap if args.overlap is None else args.overlap
    # the reference parses --batched/-b but then forces False (wavernn_gen.py:77); here the flag is honoured and the
    # default stays hp.voc_gen_batched (False)
    batched = hp.voc_gen_batched if args.batched is None else args.batched
    if args.force_cpu:
        raise SystemExit('--force_cpu: this build runs the generation loop on sm_100a only; there is no CPU fallback')
    if not torch.cuda.is_available():
        raise SystemExit('no CUDA device visible; the B200 WaveRNN path has no CPU fallback')
    device = torch.device('cuda')
    print('Using device:', device)
    print('\nInitialising Model...\n')
    model = WaveRNN(rnn_dims=hp.voc_rnn_dims, fc_dims=hp.voc_fc_dims, bits=hp.bits, pad=hp.voc_pad,
                    upsample_factors=hp.voc_upsample_factors, feat_dims=hp.num_mels,
                    compute_dims=hp.voc_compute_dims, res_out_dims=hp.voc_res_out_dims,
                    res_blocks=hp.voc_res_blocks, hop_length=hp.hop_length, sample_rate=hp.sample_rate,
                    mode=hp.voc_mode)
    paths = Paths(hp.voc_model_id)
    voc_weights = args.voc_weights if args.voc_weights else paths.voc_latest_weights
    print(voc_weights)
    model.load(voc_weights)
    simple_table([('Generation Mode', 'Batched' if batched else 'Unbatched'),
                  ('Target Samples', target if batched else 'N/A'),
                  ('Overlap Samples', overlap if batched else 'N/A')])
    if args.file:
        out_dir = './wavernn_inference_output'
        os.makedirs(out_dir, exist_ok=True)
        gen_from_file(model, args.file, out_dir, batched, target, overlap, seed=args.seed, kernel=args.kernel)
    print('\n\nExiting...\n')


if __name__ == '__main__':
    main()
