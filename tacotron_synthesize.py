#!/usr/bin/env python
"""Text (pinyin tokens) -> mel `.npy` with Tacotron-2 on a B200 -- drop-in for the reference's `tacotron_synthesize.py`.

    python tacotron_synthesize.py --text 'm ao2 h a2 d eng3 b ei4 l ei4 。'

Same output contract: `./tacotron_inference_output/step-{step}-{md5(text)}-mel-pred.npy`, float32 (T, 80) =
clip((mel + 4) / 8, 0, 1) (reference :114-116, :187-191), which `wavernn_gen.py --file` consumes.  Differences:
  * encoder / decoder loop / postnet run on the GPU through libb200tts.so; the TF checkpoint is read without TensorFlow.
  * `--text` takes the space-separated pinyin the reference obtains from `get_pyin` (:187); Hanzi input needs the
    reference's pure-Python `tacotron/pinyin` front-end on PYTHONPATH (out of scope here) -- it is used when importable.
  * no Griffin-Lim preview wav, no PNG plots (the alignment is saved as `...-align.npy` instead).
"""
from __future__ import annotations

import argparse
import hashlib
import os
import time

from tacotron_hparams import hparams
from tacotronv2_wavernn_chinese_b200.tacotron.synthesizer import Synthesizer


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--text', required=True, help='pinyin tokens separated by spaces (or Hanzi if the pinyin front-end is importable)')
    ap.add_argument('--checkpoint', default='logs-Tacotron-2/taco_pretrained', help='TF checkpoint prefix or directory')
    ap.add_argument('--train_txt', default=hparams.tacotron_input, help='training list the symbol table is rebuilt from')
    ap.add_argument('--symbols_json', default=None, help='JSON file with a "symbols" list (instead of scanning --train_txt)')
    ap.add_argument('--hparams', default='', help='comma separated name=value overrides')
    ap.add_argument('--out_dir', default='tacotron_inference_output')
    ap.add_argument('--seed', type=int, default=0, help='Philox seed of the (always on) prenet dropout')
    args = ap.parse_args(argv)
    hp = hparams.parse(args.hparams)
    text = args.text
    if not any(ch.isascii() and ch.isalnum() for ch in text):
        try:
            from tacotron.pinyin.parse_text_to_pyin import get_pyin          # the reference's front-end, if present
            _, text = get_pyin(text)
        except Exception as e:
            raise SystemExit(f'Hanzi input needs the reference pinyin front-end on PYTHONPATH ({e}); pass pinyin tokens instead')
    symbols = None
    if args.symbols_json:
        import json
        symbols = json.load(open(args.symbols_json, encoding='utf-8'))['symbols']
    synth = Synthesizer().load(args.checkpoint, hp, symbols=symbols, train_txt=args.train_txt)
    idx = hashlib.md5(text.encode('utf8')).hexdigest()
    t0 = time.time()
    mel_path, align_path = synth.synthesize(text, args.out_dir, idx, seed=args.seed)
    print(f'pred_mel_path: {mel_path}')
    print(f'last: {time.time() - t0} seconds')
    return mel_path


if __name__ == '__main__':
    main()
