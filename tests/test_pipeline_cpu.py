"""Host logic of the ragged / sharded text->audio pipeline (tacotronv2_wavernn_chinese_b200/pipeline.py) with stand-in engines on
CPU: chunk planning, round-robin dealing, and -- world_size 2 over gloo -- that every rank ends up with every sentence's wave
and that a sentence's wave depends only on (seed, GLOBAL sentence index, its mel), never on the chunking or the rank count."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tacotronv2_wavernn_chinese_b200 import pipeline as pl

HOP = 275


class FakeSynth:
    def __init__(self, frames):
        self.frames = frames

    def mels(self, texts, seed=0, utterance_offset=0, **kw):
        out = []
        for b, _t in enumerate(texts):
            rs = np.random.RandomState(1000 + utterance_offset + b)
            out.append(rs.uniform(0, 1, (self.frames[utterance_offset + b], 80)).astype(np.float32))
        return out, {}


class FakeVoc:
    """wave row = deterministic function of (seed, global id, mel content, own length); zero beyond the row's own length."""
    hop = HOP

    def __init__(self):
        self.calls = []

    def generate(self, mels, seed=0, utterance_ids=None, kernel='auto', utt_frames=None, pack=None):
        B, _, T = mels.shape
        self.calls.append((B, T) if pack is None else ('pack', pack['rows'], B, T))
        wave = torch.zeros(B, (T - 1) * HOP, dtype=torch.float64)
        for r in range(B):
            n = (int(utt_frames[r]) - 1) * HOP
            rs = np.random.RandomState((seed * 7919 + int(utterance_ids[r])) % (2 ** 31))
            wave[r, :n] = torch.as_tensor(rs.uniform(-1, 1, n) + float(np.ascontiguousarray(mels[r, :, :int(utt_frames[r])].numpy(), dtype=np.float64).sum()) * 1e-3)
        return {'wave': wave}

    def check(self):
        pass


FRAMES = [120, 33, 410, 25, 77, 300, 15, 64, 200, 51, 90]          # one sentence shorter than the 21-frame minimum


def test_plan_chunks_and_deal():
    chunks = pl.plan_chunks(FRAMES, max_rows=4)
    assert sorted(i for c in chunks for i in c) == list(range(len(FRAMES))) and all(len(c) <= 4 for c in chunks)
    flat = [FRAMES[i] for c in chunks for i in c]
    assert flat == sorted(FRAMES, reverse=True)
    done, need = pl.padded_lockstep_rows(FRAMES, chunks)
    one = pl.padded_lockstep_rows(FRAMES, [list(range(len(FRAMES)))])[0]
    assert need <= done < 0.5 * one                         # length-sorted chunks compute far fewer padded lock-steps
    shares = pl.deal_round_robin(FRAMES, 3)
    assert sorted(i for s in shares for i in s) == list(range(len(FRAMES)))
    assert [FRAMES[s[0]] for s in shares] == sorted(FRAMES, reverse=True)[:3]
    assert max(len(s) for s in shares) - min(len(s) for s in shares) <= 1


def test_pack_schedule_is_a_valid_lpt_plan():
    for rows in (2, 8):
        sch = pl.pack_schedule(FRAMES, rows)
        utt, start = sch['utt'], sch['start']
        assert utt.shape[0] == rows and start.shape == (rows, utt.shape[1] + 1) and utt.dtype == np.int32
        seen = []
        for r in range(rows):
            assert start[r, 0] == 0
            for k in range(utt.shape[1]):
                if utt[r, k] >= 0:
                    seen.append(int(utt[r, k]))
                    assert start[r, k + 1] - start[r, k] == max(FRAMES[utt[r, k]], 21) * HOP       # a whole utterance, back to back
                else:
                    assert (utt[r, k:] < 0).all() and (start[r, k + 1:] == 2 ** 31 - 1).all()
        assert sorted(seen) == list(range(len(FRAMES)))
        ends = [max(int(start[r, k + 1]) for k in range(utt.shape[1]) if utt[r, k] >= 0) for r in range(rows) if utt[r, 0] >= 0]
        assert sch['steps'] == max(ends)
        # list scheduling bound: makespan <= mean load + longest job
        total = sum(max(f, 21) for f in FRAMES) * HOP
        assert sch['steps'] <= total / rows + max(FRAMES) * HOP
    kind, plan = pl.plan_ragged(FRAMES, 32)                 # 11 ragged sentences: 8 packed rows beat one 16-row padded launch
    assert kind == 'pack' and plan['rows'] == 8
    assert pl.plan_ragged([300] * 8, 32)[0] == 'chunks'     # nothing to gain when every row holds one equally long utterance


def test_vocode_ragged_is_independent_of_chunking():
    mels, _ = FakeSynth(FRAMES).mels([''] * len(FRAMES))
    ids = list(range(len(FRAMES)))
    a = pl.vocode_ragged(FakeVoc(), mels, ids, seed=3, max_rows=32, allow_pack=False)
    pk = FakeVoc()
    p2 = pl.vocode_ragged(pk, mels, ids, seed=3, max_rows=32)
    assert pk.calls == [('pack', 8, len(FRAMES), 410)]
    voc = FakeVoc()
    b = pl.vocode_ragged(voc, mels, ids, seed=3, max_rows=3)
    assert all(np.array_equal(x, y) for x, y in zip(a, p2))
    assert len(voc.calls) == 4 and voc.calls[0] == (3, 410)
    for i, (x, y) in enumerate(zip(a, b)):
        assert x.shape == ((max(FRAMES[i], 21) - 1) * HOP,) and np.array_equal(x, y)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        texts = [''] * len(FRAMES)
        waves, mels = pl.synthesize_sharded(FakeSynth(FRAMES), FakeVoc(), texts, seed=5, max_rows=3)
        ref, _ = pl.synthesize_batch(FakeSynth(FRAMES), FakeVoc(), texts, seed=5)
        out[rank] = bool(len(waves) == len(FRAMES) and all(np.array_equal(a, b) for a, b in zip(waves, ref)))
    finally:
        dist.destroy_process_group()


def test_synthesize_sharded_gloo_world2():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}
