#!/usr/bin/env python
"""Golden vectors for the Hanzi -> pinyin front-end, produced by the UNMODIFIED reference `get_pyin`
(tacotron/pinyin/parse_text_to_pyin.py:164) -- container only (needs /root/reference).

    python oracle/make_golden_pinyin.py        ->  tests/golden/pinyin_from_reference.json

The reference module reads its dictionaries relative to the working directory (:12-13), so it is imported with
cwd = the reference root; its debug prints are swallowed.  Cases: the Hanzi column of train.txt lines 1-300 plus
sentences exercising digits, inline pinyin, prosody marks and the punctuation rewrites.
"""
import contextlib
import io
import json
import os
import sys

REF = os.environ.get('B200TTS_REFERENCE', '/root/reference')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXTRA = ['现在是凌晨零点二十七分，帮您订好上午八点的闹钟。', '如果打穿地球，那么从一头到另一头h ui4发生什么？', '价格是12345元，电话10086。',
         '中邮消费金融来电是想提醒您，您的贷款已逾期——请致电客服400；再见！', '他说：“你好……”然后走了。', '2020年3月15日，气温-5度']
TONE = ['卡尔普#1陪外孙#2玩滑梯#4。']


def main():
    os.chdir(REF)
    sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):
        from tacotron.pinyin.parse_text_to_pyin import get_pyin
    lines = open(os.path.join(REF, 'train.txt'), encoding='utf-8').read().splitlines()
    cases = []
    for tx, tone in [(l.split('|')[4], False) for l in lines[:300]] + [(t, False) for t in EXTRA + TONE] + [(t, True) for t in TONE]:
        with contextlib.redirect_stdout(io.StringIO()):
            pyin, norm = get_pyin(tx, tone)
        cases.append({'text': tx, 'tone': tone, 'pinyin': pyin, 'normalized': norm})
    out = os.path.join(ROOT, 'tests', 'golden', 'pinyin_from_reference.json')
    json.dump({'generator': 'oracle/make_golden_pinyin.py', 'cases': cases}, open(out, 'w', encoding='utf-8'), ensure_ascii=False, indent=0)
    print(out, len(cases), 'cases')


if __name__ == '__main__':
    main()
