set -x
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -q -x -m gpu ) > gpurun_out/r02_gpu_suite.log 2>&1
tail -6 gpurun_out/r02_gpu_suite.log
( timeout 900 python bench.py --steps 3 --warmup 3 ) > gpurun_out/r02_bench_tc_n1.json 2> gpurun_out/r02_bench_tc_n1.err
tail -c 3500 gpurun_out/r02_bench_tc_n1.json; tail -2 gpurun_out/r02_bench_tc_n1.err
( timeout 300 env B200TTS_TC_PROF=1 python tools/quick_time.py tc 256 3000 ) > gpurun_out/r02_tc_prof.log 2>&1
grep "^tc prof" gpurun_out/r02_tc_prof.log | tail -5
