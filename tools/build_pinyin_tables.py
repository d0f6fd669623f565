#!/usr/bin/env python
"""Digest the two public pinyin dictionaries into tacotronv2_wavernn_chinese_b200/tacotron/data/pinyin_tables.txt.gz.

    python tools/build_pinyin_tables.py [DIR_WITH pinyin.txt AND large_pinyin.txt]     (default: /root/reference/tacotron/pinyin)

Sources: mozillazg/pinyin-data 0.8.1 (`pinyin.txt`, "U+4E2D: zhōng,zhòng  # 中") and mozillazg/phrase-pinyin-data 0.10.2
(`large_pinyin.txt`, "数星星: shǔ xīng xīng"), both MIT licensed.  The parse follows how the reference reads them
(tacotron/pinyin/parse_text_to_pyin.py:15-43: two header lines skipped, whitespace stripped, first reading of a character,
phrases grouped by first character in file order); every syllable is stored converted with tone_to_digit.
"""
import gzip
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tacotronv2_wavernn_chinese_b200.tacotron.pinyin import tone_to_digit  # noqa: E402

src = sys.argv[1] if len(sys.argv) > 1 else '/root/reference/tacotron/pinyin'
out = os.path.join(ROOT, 'tacotronv2_wavernn_chinese_b200', 'tacotron', 'data', 'pinyin_tables.txt.gz')
chars, phrases = {}, []
with open(os.path.join(src, 'pinyin.txt'), encoding='utf-8') as f:
    for i, line in enumerate(f):
        if i < 2:
            continue
        line = re.sub(r'\s+', '', line.strip())
        py, word = line.split(':')[1].split('#')[:2]
        chars[word.strip()] = tone_to_digit(py.strip().split(',')[0])
with open(os.path.join(src, 'large_pinyin.txt'), encoding='utf-8') as f:
    for i, line in enumerate(f):
        if i < 2:
            continue
        parts = line.strip().split(':')
        pz, py = parts[0].strip(), parts[1].strip().split(' ')
        phrases.append((pz, [tone_to_digit(p) for p in py if p]))
os.makedirs(os.path.dirname(out), exist_ok=True)
with gzip.GzipFile(out, 'wb', mtime=0) as g:
    w = lambda s: g.write(s.encode('utf-8'))
    w('# digest of mozillazg/pinyin-data 0.8.1 + mozillazg/phrase-pinyin-data 0.10.2 (MIT); written by tools/build_pinyin_tables.py\n')
    w('#section chars\n')
    for k, v in chars.items():
        w(f'{k}\t{v}\n')
    w('#section phrases\n')
    for pz, py in phrases:
        w(f'{pz}\t{"|".join(py)}\n')
print(out, os.path.getsize(out), 'bytes;', len(chars), 'characters,', len(phrases), 'phrases')
