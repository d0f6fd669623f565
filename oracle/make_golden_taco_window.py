"""Golden vectors for the forward-attention step WITH the inference window, produced by running the
reference's own `ForwardLocationSensitiveAttention.__call__` (tacotron/models/forward_attention.py:119-231) statement by
statement on numpy arrays (oracle/ref_harness_taco_attention.py: the unmodified file imported against a numpy stand-in
for the few tensorflow ops it uses).  TEST INFRASTRUCTURE; run HERE (needs /root/reference).

States come from the oracle's own windowed run of train.txt line 241 on the shipped checkpoint (seed-1238 masks); the
selection covers every branch of the window: position held / advanced, the `pos_rec < 5` hold near the start, the forced
advance after 10 steps on one token, the clamp at the last token.

    python oracle/make_golden_taco_window.py        ->  tests/golden/taco_window_from_reference.npz
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_harness_taco_attention as R                     # noqa: E402
import tacotron_oracle as to                               # noqa: E402
from tacotronv2_wavernn_chinese_b200.tacotron import ckpt  # noqa: E402

CKPT_DIR = os.path.join(R.REF_ROOT, 'logs-Tacotron-2/taco_pretrained')


def main():
    w = ckpt.load_tacotron_weights(CKPT_DIR)
    ids = np.array(json.load(open(os.path.join(ROOT, 'tests', 'golden', 'taco_symbols.json')))['sentences']['241']['ids'])
    memory = to.encoder(w, ids)
    keys = (memory @ w['memory_layer/kernel']).astype(np.float32)
    n_run = 460
    d = to.decode(w, memory, seed=1238, max_iters=n_run, window=True, capture_states=range(n_run))
    att = R.ReferenceAttention(w, memory)
    rows, seen, worst = [], set(), {}
    for s in sorted(d['states']):
        x, st, m, win = d['states'][s]
        o, _ = to.decoder_step(w, memory, keys, x, m, st,
                               modulate=lambda al: to.attention_window(al, win['max_att'], win['pos_rec'])[0])
        _, new_max, new_pos = to.attention_window(o['forward_raw'], win['max_att'], win['pos_rec'])
        raw_max = int(np.argmax(o['forward_raw']))
        branch = (raw_max <= win['max_att'], win['pos_rec'] < 5 and 2 < min(raw_max, win['max_att'] + 1), new_max == win['max_att'],
                  new_pos == 1 and new_max > win['max_att'] + (0 if raw_max > win['max_att'] else -1), new_max >= memory.shape[0] - 1)
        if s >= 40 and branch in seen and s % 23:
            continue
        seen.add(branch)
        state = R.State(alignments=st['alpha'][None], cumulated_alignments=st['cum'][None], alpha=st['alpha'][None],
                        mu=np.reshape(st['mu'], (1, 1)).astype(np.float32), max_attentions=np.array([win['max_att']], np.int32),
                        pos_rec=np.array([win['pos_rec']], np.int32))
        al, mu, ctx, cum, mx, pr = att(o['new_h2'], state)
        rows.append(dict(step=s, query=o['new_h2'][0], alpha=st['alpha'], cum=st['cum'], mu=np.float32(st['mu']),
                         max_att=win['max_att'], pos_rec=win['pos_rec'], ref_alignments=np.asarray(al)[0], ref_mu=np.asarray(mu)[0, 0],
                         ref_context=np.asarray(ctx)[0], ref_cum=np.asarray(cum)[0], ref_max_att=int(mx[0]), ref_pos_rec=int(pr[0])))
        for k, (a, b) in dict(alignments=(o['alignments'], al[0]), mu=(o['mu'], mu[0, 0]), context=(o['context'][0], ctx[0]),
                              cum=(o['cum'], cum[0]), max_att=(new_max, mx[0]), pos_rec=(new_pos, pr[0])).items():
            worst[k] = max(worst.get(k, 0.0), float(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)).max()))
    out = {'ids': ids, 'memory': memory}
    for k in rows[0]:
        out[k] = np.stack([np.asarray(r[k]) for r in rows])
    # the loop driver's stop rule and next input (helpers.py:42-66), from the reference's own TacoTestHelper
    probs = np.array([0.0, 0.2, 0.49999997, 0.5, 0.50000006, 0.7, 1.0], dtype=np.float32)
    frame = np.arange(80, dtype=np.float32)[None]
    res = [R.reference_test_helper_next_inputs(np.array([[p]]), frame) for p in probs]
    out['helper_stop_probability'] = probs
    out['helper_finished'] = np.array([r[0] for r in res])
    assert all(np.array_equal(r[1], frame) for r in res) and not res[0][2].any() and res[0][2].shape == (1, 80)
    # zoneout at inference (modules.py:137-138) and the cell output, from the reference's own ZoneoutLSTMCell
    rs = np.random.RandomState(5)
    k1 = w['decoder/decoder_LSTM/multi_rnn_cell/cell_0/decoder_LSTM_1/kernel']
    b1 = w['decoder/decoder_LSTM/multi_rnn_cell/cell_0/decoder_LSTM_1/bias']
    zx, zc, zh = (rs.randn(1, n).astype(np.float32) for n in (768, 256, 256))
    zo = R.reference_zoneout_lstm(zx, zc, zh, k1, b1, to.lstm_cell)
    out.update(zoneout_x=zx, zoneout_c=zc, zoneout_h=zh, zoneout_ref_output=zo[0], zoneout_ref_c=zo[1], zoneout_ref_h=zo[2])
    mine = to.zoneout_lstm(zx, zc, zh, k1, b1)
    print('zoneout cell: oracle vs reference statements', [float(np.abs(a - b).max()) for a, b in zip(mine, zo)])
    path = os.path.join(ROOT, 'tests', 'golden', 'taco_window_from_reference.npz')
    np.savez_compressed(path, **out)
    print(f'wrote {path} ({os.path.getsize(path)} bytes, {len(rows)} steps of {n_run}, run stopped at {d["n_steps"]}); '
          f'oracle vs the reference statements, max abs difference:')
    for k, v in worst.items():
        print(f'  {k:11s} {v:.3e}')
    print('  steps kept:', [r['step'] for r in rows][:60], '...')
    print('  forced advances (pos_rec reset with max+1):', sum(1 for r in rows if r['ref_pos_rec'] == 1 and r['ref_max_att'] > r['max_att']))


if __name__ == '__main__':
    main()
