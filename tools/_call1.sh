set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_wavernn_gpu.py -q -x -k "tc_teacher and 128" ) > gpurun_out/r02_tc_t1.log 2>&1
tail -15 gpurun_out/r02_tc_t1.log
( timeout 900 python -m pytest tests/test_wavernn_gpu.py -q -k "test_tc_" ) > gpurun_out/r02_tc_tests.log 2>&1
tail -25 gpurun_out/r02_tc_tests.log
( timeout 300 python tools/quick_time.py tc,grid 128,256 3000 ) > gpurun_out/r02_tc_time.log 2>&1
tail -6 gpurun_out/r02_tc_time.log
