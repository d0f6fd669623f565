set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_wavernn_gpu.py -q -x -k "auto_dispatch or large_request or test_tc_teacher" ) > gpurun_out/r02_tc_tests.log 2>&1
tail -8 gpurun_out/r02_tc_tests.log
( timeout 300 python tools/quick_time.py auto 512,300 2000 ) > gpurun_out/r02_tc_time.log 2>&1
tail -3 gpurun_out/r02_tc_time.log
( timeout 300 env B200TTS_TC=0 python tools/quick_time.py auto 512,300 2000 ) >> gpurun_out/r02_tc_time.log 2>&1
tail -2 gpurun_out/r02_tc_time.log
