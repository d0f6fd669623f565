"""Hanzi -> space separated pinyin tokens, TF-free and stand-alone: the text front-end `tacotron_synthesize.py --text` needs
(SURVEY.md 8f rank 4).  Restates the reference's `get_pyin` (tacotron/pinyin/parse_text_to_pyin.py:164-237, with
`preprocess` :105-141, `tone_to_digit` :154-162, `split_pyin` :143-152, `int_to_words` :46-87) over the same two public
dictionaries the reference reads (mozillazg/pinyin-data 0.8.1 `pinyin.txt`, mozillazg/phrase-pinyin-data 0.10.2
`large_pinyin.txt`, both MIT), which ship here as ONE pre-digested table, `data/pinyin_tables.txt.gz`, written by
`tools/build_pinyin_tables.py`: every syllable is stored already converted to the model's token form
("zhōng" -> "zh ong1"), only the first reading of a character is kept (the only one the reference uses, :231), phrases keep
the dictionary's order because the reference takes the FIRST phrase that matches at a position (:219-227).

`tests/test_pinyin_cpu.py` checks exact equality with the reference's own output on 300 lines of its train.txt.
"""
from __future__ import annotations

import gzip
import os
import re

_TABLE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'pinyin_tables.txt.gz')

# parse_text_to_pyin.py:4-7 (tone mark -> base letter + tone digit)
TONE_MAP = {'ā': 'a1', 'á': 'a2', 'ǎ': 'a3', 'à': 'a4', 'ō': 'o1', 'ó': 'o2', 'ǒ': 'o3', 'ò': 'o4', 'ē': 'e1', 'é': 'e2',
            'ě': 'e3', 'è': 'e4', 'ī': 'i1', 'í': 'i2', 'ǐ': 'i3', 'ì': 'i4', 'ū': 'u1', 'ú': 'u2', 'ǔ': 'u3', 'ù': 'u4',
            'ü': 'v0', 'ǖ': 'v1', 'ǘ': 'v2', 'ǚ': 'v3', 'ǜ': 'v4', 'ń': 'n2', 'ň': 'n3'}


def split_pyin(pyin: str) -> str:
    """initial / final split (:143-152): 'zhong1' -> 'zh ong1', 'an1' -> 'an1', 'e2' -> 'e2', 'ma3' -> 'm a3'."""
    if pyin[:2] in ('ch', 'sh', 'zh'):
        return pyin[:2] + ' ' + pyin[2:]
    if pyin[0] in ('a', 'e', 'o'):
        return pyin
    if len(pyin) == 2 and pyin[-1].isdigit():
        return pyin
    return pyin[0] + ' ' + pyin[1:]


def tone_to_digit(pyin: str) -> str:
    """first tone-marked letter -> plain letter, its tone digit appended at the END of the syllable; then split (:154-162)."""
    for i, ch in enumerate(pyin):
        if ch in TONE_MAP:
            pyin = pyin[:i] + TONE_MAP[ch][0] + pyin[i + 1:] + TONE_MAP[ch][1]
            break
    return split_pyin(pyin)


def int_to_words(astr: str) -> str:
    """digits -> Chinese numerals with 万/亿 groups (:46-87; a '，' follows every group marker like the reference)."""
    units_big = ['', '万', '亿']
    units = ['', '十', '百', '千']
    digit = dict(zip('0123456789', '零一二三四五六七八九'))
    res, zero_seen = '', False
    for i, ch in enumerate(astr):
        j = len(astr) - 1 - i
        if ch == '0':
            zero_seen = True
        else:
            if zero_seen:
                res += '零'
            zero_seen = False
            if not (ch == '1' and len(astr) == 2 and j % 4 == 1):
                res += digit[ch]
            res += units[j % 4]
        if j % 4 == 0 and j // 4 > 0:
            res += units_big[j // 4] + '，'
            zero_seen = False
    return res


# The reference's punctuation normalisation (:105-141) is a fixed sequence of literal and regex rewrites; kept here as data,
# applied in order.  ('lit', a, b) = str.replace, ('re', pattern, b) = re.sub.
_REWRITES = (
    ('re', r'[）（]', ''),
    ('lit', '：“', '，'), ('lit', '：', '，'), ('lit', '”！', '！'), ('lit', '”。', '。'),
    ('lit', '……”', '。'), ('lit', '……', '。'), ('lit', '…。', '。'), ('lit', '…”', '。'), ('lit', '…', '。'), ('lit', '.', '。'),
    ('lit', '”', ''), ('lit', '“', ''), ('lit', '、', '，'), ('lit', '-', '，'),
    ('lit', '—', '，'), ('lit', '-', '，'), ('lit', '；', '。'),
    ('re', r'，[，\s]+', '，'), ('re', r'。[。，\s]+', '。'), ('re', r'，。+', '。'),
    ('re', r'？[？\s]+', '？'), ('re', r'，？+', '？'),
    ('re', r'！[！\s]+', '！'), ('re', r'，！+', '！'),
    ('re', r'\.+', '。'), ('re', r',+', '，'), ('re', r'!+', '！'), ('re', r'\?+', '？'),
    ('re', r'\s+', ' '), ('lit', '|', ''),
)
_REWRITES_COMPILED = tuple((k, re.compile(a) if k == 're' else a, b) for k, a, b in _REWRITES)


def preprocess(text: str, tone: bool = False) -> str:
    """punctuation normalisation (:105-141): prosody marks dropped unless `tone`, lower-cased, `_REWRITES` in order, stripped."""
    if not tone:
        text = re.sub(r'#\d+', '', text)
    text = text.lower()
    for kind, a, b in _REWRITES_COMPILED:
        text = a.sub(b, text) if kind == 're' else text.replace(a, b)
    return text.strip()


class PinyinTables:
    """char -> token string of its first reading; first char -> [(phrase, [token string per character])] in dictionary order."""

    def __init__(self, path: str = _TABLE):
        if not os.path.isfile(path):
            raise FileNotFoundError(f'{path} is missing: run `python tools/build_pinyin_tables.py` (needs the two dictionaries)')
        self.hz, self.phrase = {}, {}
        with gzip.open(path, 'rt', encoding='utf-8') as f:
            section = None
            for line in f:
                line = line.rstrip('\n')
                if line.startswith('#'):
                    if line.startswith('#section '):
                        section = line.split(' ', 1)[1]
                    continue
                key, val = line.split('\t')
                if section == 'chars':
                    self.hz[key] = val
                else:
                    self.phrase.setdefault(key[0], []).append((key, val.split('|')))


_tables = None


def tables() -> PinyinTables:
    global _tables
    if _tables is None:
        _tables = PinyinTables()
    return _tables


def get_pyin(text: str, tone: bool = False, tabs: PinyinTables | None = None):
    """-> (space separated tokens, normalised text): the pair the reference's get_pyin returns (:164-237).  The first
    feeds `Synthesizer.synthesize`, the md5 of the second names the output files (tacotron_synthesize.py:187-196)."""
    tabs = tabs or tables()
    text = preprocess(text, tone)
    res, i, n = [], 0, len(text)
    while i < n:
        ch = text[i]
        if text[i:i + 3] in ('pi1', 'bi1'):                       # :170-181 (kept as ONE token by the reference)
            res.append(text[i:i + 3])
            i += 3
            continue
        if ch == '#':                                             # prosody marks #1..#4 (:183-189)
            i += 1
            if i < n and text[i] in '1234':
                res.append('#' + text[i])
                i += 1
            continue
        if 'a' <= ch <= 'z':                                      # pinyin typed directly, space separated (:191-201)
            j = i
            while i < n and 'a' <= text[i] <= 'z':
                i += 1
            if i < n and text[i] in '1234':
                i += 1
            res.append(text[j:i])
            if i < n and text[i] == ' ':                          # (the reference indexes text[i] unguarded and raises at the end)
                i += 1
            continue
        if ch.isdigit():                                          # :203-211
            j = i
            while i < n and text[i].isdigit():
                i += 1
            res.extend(get_pyin(int_to_words(text[j:i]), False, tabs)[0].split(' '))
            continue
        hit = False
        for pz, py in tabs.phrase.get(ch, ()):                    # first dictionary phrase that matches here (:213-227)
            if text.startswith(pz, i) and len(py) >= len(pz):
                res.extend(py[:len(pz)])
                i += len(pz)
                hit = True
                break
        if hit:
            continue
        res.append(tabs.hz.get(ch, ch))                           # single character, first reading; unknown symbols pass through
        i += 1
    return ' '.join(res), text
