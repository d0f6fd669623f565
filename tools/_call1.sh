set -x
mkdir -p gpurun_out
( timeout 300 env B200TTS_TC_PROF=2 python tools/quick_time.py tc 256 3000 ) > gpurun_out/r02_tc_chain.log 2>&1
tail -3 gpurun_out/r02_tc_chain.log
