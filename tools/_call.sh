set -x
mkdir -p gpurun_out
( timeout 240 tools/exchange_bench.bin ) > gpurun_out/r02_exchange_bench.txt 2>&1
cat gpurun_out/r02_exchange_bench.txt
( timeout 600 python -m pytest tests/test_tacotron_gpu.py -q -x -k "teacher_forced" ) > gpurun_out/r02_c3_taco.log 2>&1
tail -15 gpurun_out/r02_c3_taco.log
python - <<'PY' > gpurun_out/r02_c3_e2e_diag.log 2>&1
import time, sys, numpy as np, torch
sys.path.insert(0, '.')
from tacotronv2_wavernn_chinese_b200 import synth
from tacotronv2_wavernn_chinese_b200.engine import WaveRNNEngine
eng = WaveRNNEngine(synth.synth_state_dict(0), synth.DEFAULT_DIMS, device=0)
m = synth.synth_mels(1, 256, 80)
for i in range(5):
    t0 = time.perf_counter(); r = eng.generate_host(m, seed=i); t1 = time.perf_counter()
    print('generate_host', i, t1 - t0, 'kernel ms', eng.last_kernel_ms(), flush=True)
md = torch.as_tensor(m).cuda()
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = eng.generate(md, seed=i); torch.cuda.synchronize(); t1 = time.perf_counter()
    print('generate dev', i, t1 - t0, flush=True)
PY
cat gpurun_out/r02_c3_e2e_diag.log
