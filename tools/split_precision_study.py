"""Would a split-operand tensor-core GEMM keep the free-running label sequence?  (round-2 design study, CPU only)

The grid kernel computes in fp32 on the CUDA cores because bf16 / TF32 operands derail the sampled label sequence
(DESIGN.md section 4).  The tensor-core alternative is operand splitting: every fp32 value v = v0 + v1 + v2 with bf16
parts (or v0 + v1 with TF32 parts), the GEMM becomes 3-9 low-precision products accumulated in fp32.  This script runs
the numpy oracle's generate() with the matrix products replaced by such emulations, under the SAME sampling noise, and
reports where each variant's label sequence first departs from the plain-fp32 one (and from float64 arithmetic, which
shows how far fp32 itself is from "exact").  Products of bf16/TF32 parts are exact in fp32, so only the accumulation
order of the real tensor core is not modelled.

    python tools/split_precision_study.py [--steps 22000] [--ckpt oracle/_ref/latest_weights.pyt]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import wavernn_oracle as wo            # noqa: E402  (checker-side tool, like bench.py's cpu_baseline)
from tacotronv2_wavernn_chinese_b200 import synth  # noqa: E402

F32 = np.float32


def round_mantissa(a, drop):
    """Round-to-nearest-even away the low `drop` bits of fp32 values (16 -> bf16, 13 -> TF32)."""
    u = np.ascontiguousarray(a, dtype=F32).view(np.uint32).astype(np.uint64)
    half = np.uint64(1 << (drop - 1))
    lsb = (u >> np.uint64(drop)) & np.uint64(1)
    u = ((u + half - np.uint64(1) + lsb) >> np.uint64(drop)) << np.uint64(drop)
    return u.astype(np.uint32).view(F32).reshape(np.shape(a))


def split(a, drop, parts):
    out, rest = [], np.asarray(a, dtype=F32)
    for _ in range(parts):
        p = round_mantissa(rest, drop)
        out.append(p)
        rest = (rest - p).astype(F32)
    return out


class Weight:
    """Stands in for a weight ndarray inside the oracle: `x @ W.T` is routed to the chosen emulation."""
    __array_ufunc__ = None

    def __init__(self, w, mode, transposed=False):
        self.w = np.asarray(w, dtype=F32)
        self.mode, self.transposed = mode, transposed
        self.dtype, self.shape, self.ndim = self.w.dtype, self.w.shape, self.w.ndim
        drop, parts = {'bf16': (16, 1), 'bf16x2': (16, 2), 'bf16x3_6': (16, 3), 'bf16x3_9': (16, 3), 'tf32': (13, 1), 'tf32x3': (13, 2)}.get(
            mode, (0, 0))
        self.drop, self.nparts = drop, parts
        self.parts = [np.ascontiguousarray(p.T) for p in split(self.w, drop, parts)] if parts else None
        self.wT64 = np.ascontiguousarray(self.w.T.astype(np.float64)) if mode == 'fp64' else None
        self.wT = np.ascontiguousarray(self.w.T)
        self.rng = np.random.RandomState(abs(hash(mode)) % (2 ** 31))

    def astype(self, _):
        return self

    @property
    def T(self):
        return Weight.__new__(Weight)._alias(self)

    def _alias(self, other):
        self.__dict__.update(other.__dict__)
        self.transposed = not other.transposed
        return self

    def __rmatmul__(self, x):
        assert self.transposed
        if self.mode == 'fp32':
            return x @ self.wT
        if self.mode.startswith('noise'):          # fp32 result + gaussian error of sigma * max|result| (a stand-in for
            y = x @ self.wT                        # an accumulator that is less exact than IEEE fp32, e.g. the tensor core's)
            sigma = float(self.mode[5:])
            return (y + self.rng.standard_normal(y.shape).astype(F32) * F32(sigma * np.abs(y).max())).astype(F32)
        if self.mode == 'fp64':
            return (x.astype(np.float64) @ self.wT64).astype(F32)
        xs = split(x, self.drop, self.nparts)
        n = self.nparts
        if self.mode == 'bf16x3_6':
            terms = [(i, j) for i in range(n) for j in range(n) if i + j <= 2]
        elif self.mode in ('tf32x3', 'bf16x2'):
            terms = [(0, 0), (0, 1), (1, 0)]
        else:
            terms = [(i, j) for i in range(n) for j in range(n)]
        acc = None
        for i, j in sorted(terms, key=lambda t: -(t[0] + t[1])):      # small terms first
            t = xs[i] @ self.parts[j]
            acc = t if acc is None else (acc + t).astype(F32)
        return acc


def wrap(params, mode):
    names = ('I.weight', 'rnn1.weight_ih_l0', 'rnn1.weight_hh_l0', 'rnn2.weight_ih_l0', 'rnn2.weight_hh_l0',
             'fc1.weight', 'fc2.weight', 'fc3.weight')
    out = dict(params)
    for k in names:
        out[k] = Weight(params[k], mode)
    return out


def first_diff(a, b):
    d = np.nonzero(a != b)[1] if a.ndim == 2 else np.nonzero(a != b)[0]
    return int(d.min()) if d.size else None


def study(name, params, mels, steps, seed, modes):
    print(f'== {name}: {mels.shape[0]} utterance(s), {steps} steps, shared noise seed {seed}')
    ref = None
    for mode in modes:
        t0 = time.time()
        lab = wo.generate(wrap(params, mode), mels, seed=seed, max_steps=steps)['labels']
        if ref is None:
            ref = lab
        per_utt = [first_diff(ref[b], lab[b]) for b in range(lab.shape[0])]
        same = sum(d is None for d in per_utt)
        print(f'  {mode:9s} identical to fp32 on {same}/{len(per_utt)} utterances; first divergence per utterance: '
              f'{per_utt}   ({time.time() - t0:.0f} s)', flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=22000)
    ap.add_argument('--utterances', type=int, default=4)
    ap.add_argument('--modes', default='fp32,fp64,bf16x3_9,bf16x3_6,tf32x3,tf32,bf16', help='first entry is the comparison base')
    ap.add_argument('--skip-synth', action='store_true')
    ap.add_argument('--ckpt', default=os.path.join(ROOT, 'oracle', '_ref', 'latest_weights.pyt'))
    a = ap.parse_args()
    frames = (a.steps + 274) // 275 + 1
    mels = synth.synth_mels(1, a.utterances, max(frames, 21))
    modes = a.modes.split(',')
    if not a.skip_synth:
        study('synthetic weights (synth_state_dict(0))', wo.as_params(synth.synth_state_dict(0)), mels, a.steps, 7, modes)
    if os.path.isfile(a.ckpt):
        import torch
        sd = torch.load(a.ckpt, map_location='cpu', weights_only=False)
        study('shipped checkpoint', wo.as_params(sd), mels, a.steps, 7, modes)


if __name__ == '__main__':
    main()
