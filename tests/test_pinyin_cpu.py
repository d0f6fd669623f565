"""Hanzi -> pinyin front-end (tacotronv2_wavernn_chinese_b200/tacotron/pinyin.py) against vectors the reference's own
get_pyin produced (oracle/make_golden_pinyin.py): exact string equality, bit for bit."""
import json
import os

from conftest import GOLDEN
from tacotronv2_wavernn_chinese_b200.tacotron import pinyin


def test_get_pyin_equals_reference_on_train_txt():
    g = json.load(open(os.path.join(GOLDEN, 'pinyin_from_reference.json'), encoding='utf-8'))
    assert len(g['cases']) >= 300
    tabs = pinyin.tables()
    assert len(tabs.hz) > 40000 and sum(len(v) for v in tabs.phrase.values()) > 390000
    for c in g['cases']:
        got = pinyin.get_pyin(c['text'], c['tone'])
        assert got == (c['pinyin'], c['normalized']), c['text']


def test_pieces():
    assert pinyin.tone_to_digit('zhōng') == 'zh ong1' and pinyin.tone_to_digit('ān') == 'an1' and pinyin.tone_to_digit('de') == 'd e'
    assert pinyin.tone_to_digit('lǜ') == 'l v4' and pinyin.tone_to_digit('ér') == 'er2' and pinyin.tone_to_digit('ó') == 'o2'
    assert pinyin.int_to_words('10086') == '一万，零八十六' and pinyin.int_to_words('12') == '十二' and pinyin.int_to_words('400') == '四百'
    assert pinyin.preprocess('他说：“好……”') == '他说，好。'
    # inline pinyin at the very end of the text: the reference indexes one past the end there; this must not raise
    assert pinyin.get_pyin('你好 h ao3')[0].endswith('h ao3')
