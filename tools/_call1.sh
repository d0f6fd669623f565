set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_wavernn_gpu.py tests/test_tacotron_gpu.py -q -x -k "packed or pipeline or large_request" ) > gpurun_out/r02_c16_tests.log 2>&1
tail -12 gpurun_out/r02_c16_tests.log
( timeout 900 python bench.py --workload text2audio --steps 2 --warmup 1 ) > gpurun_out/r02_c16_t2a_n1.json 2> gpurun_out/r02_c16_t2a_n1.err
cat gpurun_out/r02_c16_t2a_n1.json; tail -3 gpurun_out/r02_c16_t2a_n1.err
( timeout 300 env B200TTS_GRID_PROF=1 python tools/quick_time.py grid 8,32 3000 ) > gpurun_out/r02_c16_push_time.log 2>&1
tail -4 gpurun_out/r02_c16_push_time.log
