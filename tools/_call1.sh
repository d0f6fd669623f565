set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_wavernn_gpu.py -q -x -k "test_tc_ or auto_dispatch" ) > gpurun_out/r02_tc_tests.log 2>&1
tail -2 gpurun_out/r02_tc_tests.log
( timeout 200 python tools/quick_time.py tc 128,256 3000 ) > gpurun_out/r02_tc_time.log 2>&1
tail -2 gpurun_out/r02_tc_time.log
