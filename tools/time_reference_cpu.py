"""BASELINE config 1 with the UNMODIFIED reference, in this container (checker-side tool; /root/reference does not exist on the
GPU box): `WaveRNN.generate()` of lturing/tacotronv2_wavernn_chinese on the shipped checkpoint, 80-frame uniform mel (seed 1234),
`torch.manual_seed(0)`, timed with perf_counter around generate() (conditioning included, wav write stubbed), for a few torch
thread counts -- next to the numpy oracle port on the same input.  -> profiles/r01_reference_cpu_container.json
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness, wavernn_oracle as wo  # noqa: E402


def main():
    model = ref_harness.build_model()
    hp = ref_harness.import_reference()._hp
    mel = torch.rand(1, 80, 80, generator=torch.Generator().manual_seed(1234))
    S = 80 * 275
    out = {'cpu_count': os.cpu_count(), 'frames': 80, 'steps': S, 'reference': [], 'oracle_port': []}
    for nt in (1, 4, os.cpu_count()):
        torch.set_num_threads(nt)
        torch.manual_seed(0)
        model.generate(mel[:, :, :21], os.devnull, False, hp.voc_target, hp.voc_overlap, hp.mu_law)      # warm-up
        torch.manual_seed(0)
        t0 = time.perf_counter()
        model.generate(mel, os.devnull, False, hp.voc_target, hp.voc_overlap, hp.mu_law)
        dt = time.perf_counter() - t0
        out['reference'].append({'torch_threads': nt, 'seconds': dt, 'samples_per_s': S / dt, 'rtf': dt / (S / 22050)})
        print(out['reference'][-1], flush=True)
    p = wo.as_params({k: v.numpy() for k, v in model.state_dict().items()})
    try:
        from threadpoolctl import threadpool_limits
    except Exception:
        threadpool_limits = None
    for nt in (1, 4, os.cpu_count()):
        ctx = threadpool_limits(limits=nt) if threadpool_limits else None
        if ctx:
            ctx.__enter__()
        r = wo.generate(p, mel.numpy(), seed=0)
        if ctx:
            ctx.__exit__(None, None, None)
        out['oracle_port'].append({'blas_threads': nt, 'seconds': r['seconds'], 'samples_per_s': S / r['seconds']})
        print(out['oracle_port'][-1], flush=True)
    json.dump(out, open(os.path.join(ROOT, 'profiles', 'r01_reference_cpu_container.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
