set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_wavernn_gpu.py -q -x -k "test_tc_" ) > gpurun_out/r02_tc_tests.log 2>&1
tail -4 gpurun_out/r02_tc_tests.log
( timeout 300 env B200TTS_TC_PROF=1 python tools/quick_time.py tc 128,256 3000 ) > gpurun_out/r02_tc_prof.log 2>&1
grep -v "^tc prof" gpurun_out/r02_tc_prof.log | tail -3; grep "^tc prof" gpurun_out/r02_tc_prof.log | tail -5
