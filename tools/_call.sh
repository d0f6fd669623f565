set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_tacotron_gpu.py -q -x ) > gpurun_out/r02_c9_taco_tests.log 2>&1
tail -15 gpurun_out/r02_c9_taco_tests.log
( timeout 300 python bench.py --workload tacotron --steps 3 --warmup 2 --no-cpu-baseline ) > gpurun_out/r02_c9_taco.json 2> gpurun_out/r02_c9_taco.err
cat gpurun_out/r02_c9_taco.json; tail -3 gpurun_out/r02_c9_taco.err
( timeout 300 env B200TTS_GRID_PROF=1 python tools/quick_time.py grid 8,16,32 3000 ) > gpurun_out/r02_c9_push_time.log 2>&1
tail -6 gpurun_out/r02_c9_push_time.log
