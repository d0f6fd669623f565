set -x
mkdir -p gpurun_out
( timeout 120 tools/umma_split_bench.bin ) > gpurun_out/r02_umma_chunk.txt 2>&1
head -12 gpurun_out/r02_umma_chunk.txt
( timeout 1500 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/r02_c7_gputests.log 2>&1
tail -25 gpurun_out/r02_c7_gputests.log
( timeout 900 python bench.py --steps 4 --warmup 3 ) > gpurun_out/r02_c7_bench.json 2> gpurun_out/r02_c7_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_c7_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','us_per_lockstep')}); print(d['roofline']['flop_form']); print(d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
"
tail -3 gpurun_out/r02_c7_bench.err
