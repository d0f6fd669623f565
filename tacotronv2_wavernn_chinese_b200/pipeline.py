"""Text -> mel -> audio in one process on the GPU (BASELINE config 5).

The reference couples its two halves only through a `.npy` file on disk (tacotron_synthesize.py:114-116 ->
wavernn_gen.py:22, float32 (T, 80) = clip((mel + 4) / 8, 0, 1)).  Here the same array goes straight from the Tacotron
postnet to the WaveRNN conditioning network; ragged sentences are zero-padded to the longest mel, which leaves every
utterance's samples identical to a batch-1 run (the conditioning network sees zero frames past the end either way), and each
row is truncated / faded at its own length.
"""
from __future__ import annotations

import numpy as np
import torch


def synthesize_batch(synth, wavernn_engine, texts, seed=0, utterance_offset=0, kernel='auto', min_frames=21):
    """synth: tacotron.synthesizer.Synthesizer (loaded); wavernn_engine: engine.WaveRNNEngine.
    Returns (list of float64 waves, list of mels [T_b, 80])."""
    mels, _ = synth.mels(texts, seed=seed, utterance_offset=utterance_offset)
    keep = [m for m in mels]
    T = max(max(m.shape[0] for m in keep), min_frames)
    B = len(keep)
    batch = np.zeros((B, keep[0].shape[1], T), dtype=np.float32)
    frames = np.zeros(B, dtype=np.int32)
    for b, m in enumerate(keep):
        batch[b, :, :m.shape[0]] = m.T
        frames[b] = max(m.shape[0], min_frames)     # shorter than the 20-hop fade-out cannot be faded (reference raises)
    out = wavernn_engine.generate(torch.as_tensor(batch), seed=seed, utterance_offset=utterance_offset, kernel=kernel,
                                  utt_frames=frames)
    wave = out['wave'].cpu().numpy()
    wavernn_engine.check()
    hop = wavernn_engine.hop
    return [wave[b, :(frames[b] - 1) * hop].copy() for b in range(B)], keep
