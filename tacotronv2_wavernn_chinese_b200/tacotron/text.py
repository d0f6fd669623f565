"""Symbol table and text -> ids, TF-free (reference tacotron/utils/symbols.py:12-28 and tacotron/utils/text.py:18-31).

The reference builds its table at import time by scanning the last column of train.txt; the frozen copy it serves from
(`website/app/text.py:1`) has the same 191 entries.  `build_symbols` repeats the scan; `Symbols` maps tokens to ids and
appends EOS `~` like `text_to_sequence`.
"""
from __future__ import annotations

PAD, EOS = '_', '~'


def build_symbols(train_txt):
    chars = set()
    with open(train_txt, 'r', encoding='utf-8') as f:
        for line in f:
            for w in line.strip().split('|')[-1].strip().split(' '):
                chars.add(w)
    return [PAD, EOS] + sorted(chars)


class Symbols:
    def __init__(self, symbols):
        self.symbols = list(symbols)
        self._to_id = {s: i for i, s in enumerate(self.symbols)}

    def __len__(self):
        return len(self.symbols)

    def text_to_sequence(self, text):
        """text: space separated pinyin tokens (or a list); unknown tokens are dropped like the reference; EOS appended."""
        toks = text.split(' ') if isinstance(text, str) else list(text)
        return [self._to_id[w] for w in toks if w in self._to_id] + [self._to_id[EOS]]

    def sequence_to_text(self, seq):
        return ' '.join(self.symbols[i] for i in seq if 0 <= i < len(self.symbols))
