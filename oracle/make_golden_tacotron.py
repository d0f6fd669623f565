"""Container-only: derive the Tacotron test inputs that need /root/reference.

  tests/golden/taco_symbols.json     the 191-entry symbol table the reference builds at import by scanning train.txt
                                     (tacotron/utils/symbols.py:12-28), and the token ids of train.txt lines 241/378/407
                                     (the 50-token sentences of BASELINE config 4) incl. EOS (tacotron/utils/text.py:18-31)
  oracle/_ref/tacotron_weights.npz   the inference variables of the shipped TF checkpoint (git-ignored travel copy, 20 MB)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tacotronv2_wavernn_chinese_b200.tacotron import ckpt  # noqa: E402

REF = '/root/reference'


def main():
    lines = open(os.path.join(REF, 'train.txt'), encoding='utf-8').read().splitlines()
    chars = set()
    for line in lines:
        for tok in line.strip().split('|')[-1].strip().split(' '):
            chars.add(tok)
    symbols = ['_', '~'] + sorted(chars)
    s2i = {s: i for i, s in enumerate(symbols)}
    sent = {}
    for ln in list(range(1, 65)) + [241, 378, 407]:          # lines 1-64 = BASELINE config 5, 241/378/407 = config 4
        cols = lines[ln - 1].split('|')
        toks = cols[-1].strip().split(' ')
        sent[str(ln)] = dict(ids=[s2i[t] for t in toks if t in s2i] + [s2i['~']], frames=int(cols[3]), n_tokens=len(toks))
    os.makedirs(os.path.join(ROOT, 'tests', 'golden'), exist_ok=True)
    json.dump(dict(symbols=symbols, sentences=sent), open(os.path.join(ROOT, 'tests', 'golden', 'taco_symbols.json'), 'w'),
              ensure_ascii=False)
    w = ckpt.load_tacotron_weights(os.path.join(REF, 'logs-Tacotron-2/taco_pretrained'))
    os.makedirs(os.path.join(ROOT, 'oracle', '_ref'), exist_ok=True)
    np.savez(os.path.join(ROOT, 'oracle', '_ref', 'tacotron_weights.npz'), **w)
    print(len(symbols), {k: (len(v['ids']), v['frames']) for k, v in sent.items() if k in ('241', '378', '407')}, len(sent), len(w))


if __name__ == '__main__':
    main()
