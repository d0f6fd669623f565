set -x
mkdir -p gpurun_out
( timeout 300 env B200TTS_GRID_PROF=1 python tools/quick_time.py grid 1,4,8,16,32 3000 ) > gpurun_out/r02_c4_push_time.log 2>&1
tail -12 gpurun_out/r02_c4_push_time.log
( timeout 300 env B200TTS_PUSH_MIN_G=8 python tools/quick_time.py grid 1,4 3000 ) > gpurun_out/r02_c4_push_time_g8.log 2>&1
tail -3 gpurun_out/r02_c4_push_time_g8.log
( timeout 600 python -m pytest tests/test_wavernn_gpu.py -q -x -k "mapping and not 256 and not 300 and not 128 and not 100 and not 64 or fold or philox or invariance or golden or independent" ) > gpurun_out/r02_c4_tests.log 2>&1
tail -8 gpurun_out/r02_c4_tests.log
( timeout 600 python -m pytest tests/test_tacotron_gpu.py -q -x -k "pipeline" ) > gpurun_out/r02_c4_tests2.log 2>&1
tail -8 gpurun_out/r02_c4_tests2.log
( timeout 900 python bench.py --workload text2audio --steps 1 --warmup 1 ) > gpurun_out/r02_c4_t2a_n1.json 2> gpurun_out/r02_c4_t2a_n1.err
cat gpurun_out/r02_c4_t2a_n1.json; tail -5 gpurun_out/r02_c4_t2a_n1.err
( timeout 300 python bench.py --workload tacotron --steps 3 --warmup 2 ) > gpurun_out/r02_c4_taco.json 2> gpurun_out/r02_c4_taco.err
cat gpurun_out/r02_c4_taco.json; tail -5 gpurun_out/r02_c4_taco.err
