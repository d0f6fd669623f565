"""WHOLE-RUN golden for the Tacotron-2 decoder, produced by driving the reference's own serialized decoder-step graph through
the complete sentence (config 4: train.txt line 241, ~400 steps) -- container only (needs /root/reference).

TEST INFRASTRUCTURE.  oracle/make_golden_taco_step.py pins single steps from states the restatement itself produced; here the
loop state is carried from graph step to graph step, so the restatement's whole trajectory (every frame, every alignment, the
stop step) is compared with a trajectory that never passes through oracle.tacotron_oracle.decoder_step.  All per-step arithmetic
is the serialized `decoder/while/CustomDecoderStep/*` sub-graph evaluated by oracle/tf_graph_eval.py on the shipped checkpoint.
The glue between two steps is the inference loop's own and is pinned separately by executing the reference's classes:
  * zoneout at inference, state <- 0.9 new + 0.1 previous (modules.py:137-138; tests/test_tacotron_graph_pins.py runs
    ZoneoutLSTMCell itself) -- the serialized graph is the TRAINING graph, so this line is applied here, outside it;
  * next input = the frame just produced, stop when round(sigmoid(stop logit)) == 1 (helpers.py:42-66; TacoTestHelper itself
    is executed in the same test file);
  * initial state (attention.py:112-117, helpers.py:149).
The prenet keep-masks are the ones oracle.decode draws for seed 1238 (RandomState stream, [steps, 2, 256]).

    python oracle/make_golden_taco_run.py        ->  tests/golden/taco_run_from_graph.npz
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import tacotron_oracle as to                               # noqa: E402
import tf_graph_eval as E                                   # noqa: E402
from make_golden_taco_graph import DEFAULT_META            # noqa: E402
from make_golden_taco_step import CKPT_DIR, run_graph_step  # noqa: E402
from make_golden_taco_encpost import ENC, POST, P, encoder_lstm_step, inference_feeds   # noqa: E402
from tacotronv2_wavernn_chinese_b200.tacotron import ckpt  # noqa: E402

F32 = np.float32
SEED = 1238
MAX_ITERS = 700
MORE = (('48', 11), ('42', 12))                  # (train.txt line, RandomState seed of the keep-masks): 17 and 52 tokens


def drive_graph(nodes, variables, memory, keys, masks, zoneout=0.1, max_iters=MAX_ITERS):
    Tx, U = memory.shape[0], 256
    z = F32(zoneout)
    zeros = np.zeros((1, U), dtype=F32)
    alpha0 = np.zeros(Tx, dtype=F32)
    alpha0[0] = 1
    st = dict(c1=zeros, h1=zeros, c2=zeros, h2=zeros, ctx=np.zeros((1, memory.shape[1]), dtype=F32), alpha=alpha0, cum=alpha0.copy(),
              mu=F32(0.5))
    x = np.zeros((1, 80), dtype=F32)
    frames, stops, aligns = [], [], []
    zoned = lambda new, prev: ((F32(1) - z) * new + z * prev).astype(F32)
    for step in range(max_iters):
        g = run_graph_step(nodes, variables, memory, keys, x, masks[step].astype(F32), st)
        st = dict(c1=zoned(g['new_c1'], st['c1']), h1=zoned(g['new_h1'], st['h1']), c2=zoned(g['new_c2'], st['c2']),
                  h2=zoned(g['new_h2'], st['h2']), ctx=np.asarray(g['context'], dtype=F32).reshape(1, -1),
                  alpha=np.asarray(g['alignments'], dtype=F32).reshape(-1), cum=np.asarray(g['cum'], dtype=F32).reshape(-1),
                  mu=F32(np.asarray(g['mu']).reshape(-1)[0]))
        frame = np.asarray(g['frame'], dtype=F32).reshape(1, 80)
        logit = F32(np.asarray(g['stop_logit']).reshape(-1)[0])
        stop = F32(1) / (F32(1) + np.exp(-logit, dtype=F32))
        frames.append(frame[0]); stops.append(stop); aligns.append(st['alpha'].copy())
        x = frame
        if stop > 0.5:
            break
    return np.stack(frames), np.array(stops, dtype=F32), np.stack(aligns)


def drive_graph_window(nodes, variables, w, memory, keys, masks, zoneout=0.1, max_iters=MAX_ITERS):
    """The same loop with the optional INFERENCE WINDOW (SURVEY a-9), which the serialized graph does not contain: prenet, both LSTM
    cells and the two projections come from the serialized step graph, the attention step -- forward recursion, window, context,
    transition probability -- from the reference's OWN `ForwardLocationSensitiveAttention.__call__` (forward_attention.py:119-231)
    executed on numpy arrays (oracle/ref_harness_taco_attention.py), whose context is fed back into the graph's projections."""
    import ref_harness_taco_attention as R
    from make_golden_taco_step import STEP, TARGETS, LOOP
    att = R.ReferenceAttention(w, memory)
    Tx = memory.shape[0]
    z = F32(zoneout)
    zeros = np.zeros((1, 256), dtype=F32)
    alpha0 = np.zeros(Tx, dtype=F32)
    alpha0[0] = 1
    st = dict(c1=zeros, h1=zeros, c2=zeros, h2=zeros, ctx=np.zeros((1, memory.shape[1]), dtype=F32), alpha=alpha0, cum=alpha0.copy(),
              mu=F32(0.5))
    max_att, pos_rec = np.zeros(1, np.int32), np.zeros(1, np.int32)
    x = np.zeros((1, 80), dtype=F32)
    frames, stops, aligns, maxes = [], [], [], []
    zoned = lambda new, prev: ((F32(1) - z) * new + z * prev).astype(F32)
    for step in range(max_iters):
        m = masks[step].astype(F32)
        g = run_graph_step(nodes, variables, memory, keys, x, m, st)                       # cells: independent of this step's attention
        state = R.State(alignments=st['alpha'][None], cumulated_alignments=st['cum'][None], alpha=st['alpha'][None],
                        mu=np.reshape(st['mu'], (1, 1)).astype(F32), max_attentions=max_att, pos_rec=pos_rec)
        al, mu, ctx, cum, max_att, pos_rec = att(np.asarray(g['new_h2'], dtype=F32).reshape(1, -1), state)
        max_att, pos_rec = np.asarray(max_att, np.int32).reshape(1), np.asarray(pos_rec, np.int32).reshape(1)
        ctx = np.asarray(ctx, dtype=F32).reshape(1, -1)
        # projections of [new_h2, context] as serialized, with the windowed context fed in place of the graph's own
        feeds = {
            LOOP + 'Identity_17': x, LOOP + 'Identity_4': st['c1'], LOOP + 'Identity_5': st['h1'], LOOP + 'Identity_6': st['c2'],
            LOOP + 'Identity_7': st['h2'], LOOP + 'Identity_8': st['ctx'],
            STEP + 'decoder_prenet/dropout_1decoder_prenet/dropout/Cast': m[0][None, :],
            STEP + 'decoder_prenet/dropout_2decoder_prenet/dropout/Cast': m[1][None, :],
            TARGETS['context']: ctx,
        }
        ev = E.Evaluator(nodes, variables, feeds)
        frame = np.asarray(ev.get(TARGETS['frame']), dtype=F32).reshape(1, 80)
        logit = F32(np.asarray(ev.get(TARGETS['stop_logit'])).reshape(-1)[0])
        st = dict(c1=zoned(g['new_c1'], st['c1']), h1=zoned(g['new_h1'], st['h1']), c2=zoned(g['new_c2'], st['c2']),
                  h2=zoned(g['new_h2'], st['h2']), ctx=ctx, alpha=np.asarray(al, dtype=F32).reshape(-1),
                  cum=np.asarray(cum, dtype=F32).reshape(-1), mu=F32(np.asarray(mu).reshape(-1)[0]))
        stop = F32(1) / (F32(1) + np.exp(-logit, dtype=F32))
        frames.append(frame[0]); stops.append(stop); aligns.append(st['alpha'].copy()); maxes.append(int(max_att[0]))
        x = frame
        if stop > 0.5:
            break
    return np.stack(frames), np.array(stops, dtype=F32), np.stack(aligns), np.array(maxes, dtype=np.int16)


def encoder_through_graph(nodes, variables, ids, zoneout=0.1):
    """The whole encoder of one sentence: conv blocks as serialized (inference feeds), then the two LSTM loop bodies iterated
    over all tokens with the inference zoneout between iterations (modules.py:137-138); the un-zoned h is the output (:142)."""
    feeds = inference_feeds(variables, ENC)
    feeds['datafeeder/input_queue_Dequeue'] = ids[None].astype(np.int32)
    x = np.asarray(E.Evaluator(nodes, variables, feeds).get(P + ENC[-1] + 'batch_normalization/batchnorm/add_1')[0], dtype=F32)
    Tx = x.shape[0]
    z = F32(zoneout)
    outs = []
    for direction, order in (('fw', range(Tx)), ('bw', range(Tx - 1, -1, -1))):
        c = h = np.zeros((1, 256), dtype=F32)
        o = np.zeros((Tx, 256), dtype=F32)
        for t in order:
            nc, nh = encoder_lstm_step(nodes, variables, direction, x[t:t + 1], c, h)
            nc, nh = np.asarray(nc, dtype=F32).reshape(1, -1), np.asarray(nh, dtype=F32).reshape(1, -1)
            o[t] = nh[0]
            c, h = ((F32(1) - z) * nc + z * c).astype(F32), ((F32(1) - z) * nh + z * h).astype(F32)
        outs.append(o)
    return np.concatenate(outs, axis=1).astype(F32)


def postnet_through_graph(nodes, variables, frames):
    feeds = inference_feeds(variables, POST)
    feeds[P + 'Reshape_3'] = frames[None]                                                # decoder output, before the clip
    return np.asarray(E.Evaluator(nodes, variables, feeds).get(P + 'Minimum_1')[0], dtype=F32)


def main():
    nodes = E.load_graph(DEFAULT_META)
    variables = ckpt.load_bundle(CKPT_DIR)
    w = ckpt.load_tacotron_weights(CKPT_DIR)
    sent = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'taco_symbols.json')))['sentences']['241']
    ids = np.array(sent['ids'])
    memory = to.encoder(w, ids)
    keys = (memory @ w['memory_layer/kernel']).astype(F32)
    masks = (np.random.RandomState(SEED).uniform(size=(MAX_ITERS, 2, 256)) >= 0.5)      # the stream oracle.decode(seed=SEED) draws
    t0 = time.time()
    frames, stops, aligns = drive_graph(nodes, variables, memory, keys, masks)
    n = len(frames)
    print(f'graph-driven run: {n} steps in {time.time() - t0:.1f} s')
    d = to.decode(w, memory, dropout_masks=masks.astype(F32), max_iters=MAX_ITERS)
    print('oracle run:', d['n_steps'], 'steps')
    m = min(n, d['n_steps'])
    err = np.abs(d['frames'][:m] - frames[:m]).max(axis=1)
    for s in (0, 10, 50, 100, 200, 300, m - 1):
        if s < m:
            print(f'  step {s:4d}: max |frame diff| {err[s]:.3e}')
    print('  overall', float(err.max()), ' alignment argmax identical:', bool((d['alignments'][:m].argmax(1) == aligns[:m].argmax(1)).all()))
    # the run-once neighbours over the whole sentence: encoder (all 51 tokens through both LSTM loop bodies), postnet (all frames)
    mem_g = encoder_through_graph(nodes, variables, ids)
    mel_g = postnet_through_graph(nodes, variables, frames)
    print(f'  encoder memory: graph vs oracle {np.abs(mem_g - memory).max():.3e} (max |memory| {np.abs(mem_g).max():.2f});'
          f'  postnet mel: {np.abs(mel_g - to.postnet(w, frames)).max():.3e} (max |mel| {np.abs(mel_g).max():.2f})')
    path = os.path.join(ROOT, 'tests', 'golden', 'taco_run_from_graph.npz')
    np.savez_compressed(path, memory_graph=mem_g, mel_graph=mel_g, ids=ids, seed=np.array(SEED), n_steps=np.array(n), frames=frames, stop=stops,
                        align_argmax=aligns.argmax(1).astype(np.int16), align_peak=aligns.max(1).astype(F32),
                        masks=np.packbits(masks[:n].reshape(n, -1), axis=1))
    print(f'wrote {path} ({os.path.getsize(path)} bytes)')
    # two more sentences of other lengths (the shortest and the longest of train.txt 1-64), other mask streams: trajectories only
    more = {}
    for key, seed in MORE:
        ids2 = np.array(json.load(open(os.path.join(ROOT, 'tests', 'golden', 'taco_symbols.json')))['sentences'][key]['ids'])
        mem2 = to.encoder(w, ids2)
        keys2 = (mem2 @ w['memory_layer/kernel']).astype(F32)
        masks2 = (np.random.RandomState(seed).uniform(size=(MAX_ITERS, 2, 256)) >= 0.5)
        f2, s2, a2 = drive_graph(nodes, variables, mem2, keys2, masks2)
        d2 = to.decode(w, mem2, dropout_masks=masks2.astype(F32), max_iters=MAX_ITERS)
        print(f'  sentence {key} ({len(ids2)} tokens, seed {seed}): graph {len(f2)} steps, oracle {d2["n_steps"]} steps, '
              f'max |frame diff| {np.abs(d2["frames"][:len(f2)] - f2[:d2["n_steps"]]).max():.3e}')
        more.update({f's{key}_ids': ids2, f's{key}_seed': np.array(seed), f's{key}_frames': f2, f's{key}_stop': s2,
                     f's{key}_align_argmax': a2.argmax(1).astype(np.int16)})
    # config-4 sentence once more WITH the inference window
    fw, sw, aw, mw = drive_graph_window(nodes, variables, w, memory, keys, masks)
    dw = to.decode(w, memory, dropout_masks=masks.astype(F32), max_iters=MAX_ITERS, window=True)
    mm = min(len(fw), dw['n_steps'])
    print(f'  windowed run of sentence 241: graph + reference attention class {len(fw)} steps, oracle {dw["n_steps"]} steps, '
          f'max |frame diff| {np.abs(dw["frames"][:mm] - fw[:mm]).max():.3e}; differs from the un-windowed run by '
          f'{np.abs(fw[:min(len(fw), n)] - frames[:min(len(fw), n)]).max():.2f}')
    more.update(w241_frames=fw, w241_stop=sw, w241_align_argmax=aw.argmax(1).astype(np.int16), w241_max_att=mw)
    path2 = os.path.join(ROOT, 'tests', 'golden', 'taco_run_from_graph_more.npz')
    np.savez_compressed(path2, sentences=np.array([k for k, _ in MORE]), **more)
    print(f'wrote {path2} ({os.path.getsize(path2)} bytes)')


if __name__ == '__main__':
    main()
