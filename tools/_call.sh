set -x
mkdir -p gpurun_out
( timeout 300 env B200TTS_GRID_PROF=1 python tools/quick_time.py grid 40,64,96,128,192,256 2000 ) > gpurun_out/r02_c6_mg_time.log 2>&1
tail -14 gpurun_out/r02_c6_mg_time.log
( timeout 900 python -m pytest tests/test_wavernn_gpu.py -q -x -k "mapping and (256 or 128 or 100 or 64) or tiles" ) > gpurun_out/r02_c6_tests.log 2>&1
tail -12 gpurun_out/r02_c6_tests.log
