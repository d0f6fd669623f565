set -x
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/r02_c15_gputests.log 2>&1
tail -14 gpurun_out/r02_c15_gputests.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r02_c15_smoke.log 2>&1
tail -6 gpurun_out/r02_c15_smoke.log
( timeout 300 python bench.py --workload tacotron --steps 3 --warmup 2 ) > gpurun_out/r02_c15_taco.json 2> gpurun_out/r02_c15_taco.err
cat gpurun_out/r02_c15_taco.json; tail -2 gpurun_out/r02_c15_taco.err
( timeout 300 env B200TTS_GRID_PROF=1 python tools/quick_time.py grid 1,8,16,32 3000 ) > gpurun_out/r02_c15_push_time.log 2>&1
tail -8 gpurun_out/r02_c15_push_time.log
