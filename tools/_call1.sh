set -x
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -q -x -m gpu ) > gpurun_out/r02_gpu_suite_final.log 2>&1
tail -4 gpurun_out/r02_gpu_suite_final.log
( timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/r02_smoke_final.log 2>&1
tail -2 gpurun_out/r02_smoke_final.log
