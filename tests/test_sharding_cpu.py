"""world_size-2 gloo test of the multi-GPU plumbing (tacotronv2_wavernn_chinese_b200/dist.py): shard bounds, ragged
all-gather, and the property the real path relies on -- a row's result depends only on (seed, GLOBAL row index)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tacotronv2_wavernn_chinese_b200.dist import all_gather_rows, generate_sharded, shard_bounds


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 5, 64, 255, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _fake_generate(mels, seed, utterance_offset, **kw):
    """Stand-in for WaveRNNEngine.generate on CPU: labels are a pure function of (seed, GLOBAL row, mel content)."""
    B, _, T = mels.shape
    rows = []
    for b in range(B):
        rs = np.random.RandomState((seed * 1000003 + utterance_offset + b) % (2 ** 31))
        rows.append((rs.randint(0, 1024, size=T * 5) + int(mels[b].sum() * 100) % 7).astype(np.int16))
    return {'labels': torch.as_tensor(np.stack(rows)) if rows else torch.zeros(0, T * 5, dtype=torch.int16)}


def _worker(rank, world, port, n, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        mels = torch.as_tensor(np.random.RandomState(7).uniform(0, 1, (n, 80, 6)).astype(np.float32))
        labels, (lo, hi) = generate_sharded(_fake_generate, mels, seed=11)
        ref = _fake_generate(mels, 11, 0)['labels']
        ok = torch.equal(labels, ref) and (lo, hi) == shard_bounds(n, world, rank)
        t = all_gather_rows(torch.full((hi - lo, 3), float(rank)), n)
        ok = ok and t.shape == (n, 3) and (hi == lo or float(t[lo, 0]) == rank)
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n', [1, 5, 8])       # n = 1 < world: rank 1 owns an EMPTY shard and must not hang the gather
def test_generate_sharded_gloo_world2(n):
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, n, out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}
