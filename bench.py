#!/usr/bin/env python
"""Benchmark of the WaveRNN generation hot path (BASELINE.json metric: audio samples/sec, batched utterances).

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...       # the reference algorithm on the host cores (oracle port)

One "step" = one full pass of the hot path (conditioning network + all T*hop autoregressive sample steps +
mu-law decode/fade) over this rank's batch of synthetic 80-frame mels.  Workload = BASELINE config 3
(batch=256 utterances of 80 frames); per-GPU work is fixed as N grows ("weak"), utterances are sharded with no
data-path collective, and for N > 1 every timed step ends with the NCCL all-gather of the int16 labels.

Prints ONE JSON line (rank 0).  `value` is measured with inputs resident in HBM; `e2e` goes through the C-ABI host
entry point (b200tts_wavernn_generate_host: pinned H2D of the mels, generation, D2H of labels + float64 wave).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HOP, FEAT, NCLS = 275, 80, 1024
STEP_WEIGHT_BYTES = 17_371_136          # all per-step weights + biases, fp32, touched once per lock-step
COND_BYTES_PER_UTT = 836                # 80 mel + 128 aux fp32 in, 4 B out, per utterance-sample
FLOP_PER_SAMPLE = 8_668_160             # 2 * 4 334 080 MAC per utterance-sample


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', choices=('b200', 'reference'), default='b200')
    ap.add_argument('--batch', type=int, default=256, help='utterances per GPU')
    ap.add_argument('--frames', type=int, default=80)
    ap.add_argument('--kernel', default='auto', choices=('auto', 'grid', 'utterance'))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self._stop_evt, self.max_mhz = index, [], set(), threading.Event(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8): 'hw_slowdown',
                 getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
                 getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
                 getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4): 'sw_power_cap'}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {'sm_mhz': med, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons), 'samples': len(self.samples)}


_BEST_THREADS = None


_CPU_CACHE = {}


def _cpu_workload(batch):
    """Weights + conditioning of the CPU arm, built once per process: `batch` utterances (the arm's own batch size, so the
    host runs [batch x K] GEMMs like the reference's loop would on a [batch, 80, T] mel), minimum-length mels (21 frames:
    the per-step cost does not depend on the utterance length), 8 distinct utterances tiled to `batch` rows (it does not
    depend on the values either; the numpy conditioning network would otherwise take ~0.3 s per utterance)."""
    if batch not in _CPU_CACHE:
        from oracle import wavernn_oracle as wo
        from tacotronv2_wavernn_chinese_b200 import synth
        p = {k: (v.astype(np.float32) if v.dtype.kind == 'f' else v) for k, v in wo.as_params(synth.synth_state_dict(0)).items()}
        distinct = min(8, batch)
        mels = synth.synth_mels(1234, distinct, 21)
        d = wo._dims(p)
        mp = wo.pad_tensor(mels.transpose(0, 2, 1), d['pad'], 'both').transpose(0, 2, 1)
        m_up, aux = wo.upsample(p, mp)
        reps = (batch + distinct - 1) // distinct
        tile = lambda a: np.ascontiguousarray(np.tile(a, (reps, 1, 1))[:batch])
        _CPU_CACHE[batch] = (wo, p, tile(mels), (tile(m_up), tile(aux)))
    return _CPU_CACHE[batch]


def _pick_threads(batch):
    """Probe a few BLAS pool sizes on the actual workload and keep the fastest -- "all the host threads it can USE"
    (at batch 1 the matvecs are tiny and more than ~16 threads only adds contention; at batch 256 the GEMMs scale further)."""
    global _BEST_THREADS
    if _BEST_THREADS is not None:
        return _BEST_THREADS
    ncpu = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
    except Exception:
        _BEST_THREADS = ncpu
        return ncpu
    wo, p, mels, cond = _cpu_workload(batch)
    n = 30 if batch > 1 else 150
    best, best_rate = 1, 0.0
    for nt in sorted({1, 4, 8, 16, 32, 64, ncpu} & set(range(1, ncpu + 1))):
        with threadpool_limits(limits=nt):
            wo.generate(p, mels, max_steps=3, seed=0, cond=cond)
            r = wo.generate(p, mels, max_steps=n, seed=0, cond=cond)
        rate = n / r['loop_seconds']
        if rate > best_rate:
            best, best_rate = nt, rate
    _BEST_THREADS = best
    return best


def cpu_oracle_rate(batch, max_seconds=25.0):
    """The reference algorithm (numpy oracle port, oracle/wavernn_oracle.py) on the host cores, on the arm's batch size:
    as many lock-steps of `batch` utterances as fit the time bound; the one-shot conditioning network is outside the timed
    loop (see _cpu_workload).  Returns (samples/s, lock-steps run, threads used)."""
    import contextlib
    wo, p, mels, cond = _cpu_workload(batch)
    threads = _pick_threads(batch)
    try:
        from threadpoolctl import threadpool_limits
        ctx = threadpool_limits(limits=threads)
    except Exception:
        ctx = contextlib.nullcontext()
    with ctx:
        probe = wo.generate(p, mels, max_steps=10, seed=0, cond=cond)
        per_step = max(probe['loop_seconds'] / 10, 1e-6)
        steps = int(min(mels.shape[2] * HOP, max(10, max_seconds / per_step)))
        r = wo.generate(p, mels, max_steps=steps, seed=0, cond=cond)
    return batch * steps / r['loop_seconds'], steps, threads


def workload_config(args, N):
    """The `config` object both arms print (BASELINE config 3)."""
    B, T = args.batch, args.frames
    return {'workload': f'BASELINE config 3: WaveRNN generate(), batch={B} utterances/GPU of {T}-frame synthetic '
                        f'mels (voc_mode=RAW bits=10 hop=275), random-init weights of the shipped architecture',
            'utterances_per_gpu': B, 'global_batch': N * B, 'frames': T, 'steps_per_utterance': T * HOP}


def cpu_sample_text(batch, steps_run, frames):
    return (f'{batch} utterances x {steps_run} of {frames * HOP} lock-steps per bench step, sampling loop only (the one-shot '
            f'conditioning network is outside the timed region); numpy oracle port of generate(), fp32, BLAS threads chosen by probe')


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores, on the b200 arm's workload (the Python reference itself
    cannot travel to the GPU box; this is the oracle port, pinned bit-for-bit to the reference's labels by
    tests/test_oracle_golden.py).  Each bench step is a time-bounded sample of the 22 000 lock-steps of the batch."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    per_step_budget = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    rates, steps_run, threads = [], 0, 1
    for i in range(args.warmup + args.steps):
        rate, steps_run, threads = cpu_oracle_rate(args.batch, max_seconds=per_step_budget)
        if i >= args.warmup:
            rates.append(rate)
    v = float(np.mean(rates))
    cfg = workload_config(args, 1)
    cfg['parallelism'] = f'host CPU, {threads} BLAS threads (rank 0 only)'
    line = {
        'impl': 'reference', 'metric': 'wavernn_audio_samples_per_sec', 'value': v, 'unit': 'samples/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * args.batch * steps_run / v,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': cfg,
        'cpu_baseline': {'value': v, 'unit': 'samples/s', 'cores': threads, 'kind': 'port',
                         'sample': cpu_sample_text(args.batch, steps_run, args.frames)},
        'e2e': {'value': v, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'rtf': 22050.0 / (v / args.batch),
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference(args)
    import torch
    import torch.distributed as dist
    from tacotronv2_wavernn_chinese_b200 import synth
    from tacotronv2_wavernn_chinese_b200.engine import WaveRNNEngine

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the B200 path has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    N = world
    B, T = args.batch, args.frames
    S, wave_len = T * HOP, (T - 1) * HOP

    eng = WaveRNNEngine(synth.synth_state_dict(0), synth.DEFAULT_DIMS, device=local)
    mels_host = synth.synth_mels(1236 + rank, B, T)                       # this rank's shard of the global batch
    mels_dev = torch.as_tensor(mels_host).to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)   # > 126 MB L2
    # NCCL has no int16: the labels travel as raw bytes
    gathered = [torch.empty(B, S * 2, device=dev, dtype=torch.uint8) for _ in range(N)] if N > 1 else None

    def sync_all():
        torch.cuda.synchronize(dev)
        if N > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def one_step(i):
        flush.fill_(float(i))                                              # evict L2 between iterations
        out = eng.generate(mels_dev, seed=20260923, utterance_offset=rank * B, kernel=args.kernel)
        if N > 1:
            dist.all_gather(gathered, out['labels'].view(torch.uint8))
        return out

    for i in range(args.warmup):
        one_step(i)
    sync_all()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    ev0.record()
    for i in range(args.steps):
        one_step(i)
        kernel_ms.append(None)
    ev1.record()
    sync_all()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    launches = eng.launch_count - launches0 + args.steps          # + the L2 flush fill per step
    # the dominant kernel's own duration (CUDA events recorded around it on the launch stream by the library)
    kms = []
    for i in range(min(3, max(1, args.steps))):
        one_step(i)
        kms.append(eng.last_kernel_ms())
    gen_ms = float(np.mean(kms))
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if N > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = N * B * S * args.steps / (ms / 1e3)

    # ---- end to end through the C-ABI host entry point --------------------------------------------------
    e2e = None
    if not args.no_e2e:
        eng.generate_host(mels_host, seed=1)                                # warm the pinned staging buffers
        sync_all()
        t0 = time.perf_counter()
        for i in range(args.steps):
            res = eng.generate_host(mels_host, seed=20260923 + i, utterance_offset=rank * B, kernel=args.kernel)
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        if N > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        assert res['wave'].shape == (B, wave_len)
        e2e = {'value': N * B * S * args.steps / dt, 'unit': 'samples/s',
               'h2d_bytes_per_step': int(N * B * FEAT * T * 4),
               'd2h_bytes_per_step': int(N * B * (S * 2 + wave_len * 8))}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
        peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s'
        alg_bytes = S * (STEP_WEIGHT_BYTES + B * COND_BYTES_PER_UTT)       # per launch of the generation kernel
        achieved = alg_bytes / (gen_ms / 1e3) / 1e9
        sm_mhz = clocks.get('sm_mhz') or 1965.0
        fp32_peak_tflops = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        flops = B * S * FLOP_PER_SAMPLE / (gen_ms / 1e3) / 1e12
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json'))).get('dram_bytes_per_launch')
        except Exception:
            pass
        line = {
            'metric': 'wavernn_audio_samples_per_sec', 'value': value, 'unit': 'samples/s', 'n_gpus': N,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': dict(workload_config(args, N), kernel=args.kernel, l2='flushed between timed iterations (256 MiB fill)',
                           parallelism=f'utterance-sharded x{N}, NCCL all-gather of labels' if N > 1 else 'single GPU'),
            'rtf': 22050.0 / (value / (N * B)),
            'us_per_lockstep': 1e3 * gen_ms / S,
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': achieved / hbm_peak,
                         'traffic': traffic, 'peak_source': peak_src,
                         'note': 'weights are SMEM-stationary, so the HBM form is small by construction; the binding '
                                 'resource is fp32 FMA issue (see flop_form)',
                         'kernel_ms': gen_ms, 'algorithmic_bytes_per_launch': alg_bytes,
                         'flop_form': {'achieved': flops, 'peak': fp32_peak_tflops, 'unit': 'TFLOP/s fp32 CUDA-core',
                                       'frac': flops / fp32_peak_tflops}},
            'clocks': clocks,
            'gpu_launches': int(launches),
            'e2e': e2e,
        }
        if not args.no_cpu_baseline and N == 1:        # reported baseline: rank 0, single-GPU runs only
            rate, steps_run, threads = cpu_oracle_rate(B, max_seconds=15.0)
            line['cpu_baseline'] = {'value': rate, 'unit': 'samples/s', 'cores': threads, 'kind': 'port',
                                    'sample': cpu_sample_text(B, steps_run, T)}
        print(json.dumps(line), flush=True)
    if N > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
