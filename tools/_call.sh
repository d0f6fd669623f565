set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
( timeout 300 env B200TTS_GRID_PROF=1 python tools/quick_time.py grid 1,4,8,16,32 3000 ) > gpurun_out/r02_c1_push_time.log 2>&1
tail -20 gpurun_out/r02_c1_push_time.log
( timeout 300 env B200TTS_PUSH=0 python tools/quick_time.py grid 1,8,32,64,128,256 3000 ) > gpurun_out/r02_c1_old_time.log 2>&1
tail -8 gpurun_out/r02_c1_old_time.log
( timeout 1200 python -m pytest tests/test_wavernn_gpu.py -q -x --durations=15 ) > gpurun_out/r02_c1_tests.log 2>&1
tail -60 gpurun_out/r02_c1_tests.log
