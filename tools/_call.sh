set -x
mkdir -p gpurun_out
( timeout 300 env B200TTS_GRID_PROF=1 python tools/quick_time.py grid 1,8,16,32 3000 ) > gpurun_out/r02_c12_push_time.log 2>&1
tail -8 gpurun_out/r02_c12_push_time.log
( timeout 900 python -m pytest tests/test_wavernn_gpu.py tests/test_sharded_gpu.py -q -x -k "mapping and (32 or 20 or 12 or 7 or 3) or fold or independent or shards or golden or config2 or philox or invariance" ) > gpurun_out/r02_c12_tests.log 2>&1
tail -6 gpurun_out/r02_c12_tests.log
