set -x
mkdir -p gpurun_out
( timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_tc_frames21.csv python bench.py --steps 1 --warmup 1 --frames 21 --no-e2e --no-cpu-baseline --no-strong ) > gpurun_out/ncu_b.log 2>&1
tail -3 gpurun_out/ncu_b.log; grep -c . gpurun_out/r02_launches_bench_tc_frames21.csv
( timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:wavernn_tc -c 1 --csv --log-file gpurun_out/r02_tc_b256_full_launch_dram.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-strong ) > gpurun_out/ncu_d.log 2>&1
tail -3 gpurun_out/ncu_d.log; tail -4 gpurun_out/r02_tc_b256_full_launch_dram.csv
( timeout 1200 ncu --set full --clock-control none --import-source on -k regex:wavernn_tc -c 1 -o gpurun_out/r02_tc_b256 python tools/quick_time.py tc 256 300 ) > gpurun_out/ncu_f.log 2>&1
tail -3 gpurun_out/ncu_f.log
ncu -i gpurun_out/r02_tc_b256.ncu-rep --page raw --csv > gpurun_out/r02_tc_b256_raw.csv
ncu -i gpurun_out/r02_tc_b256.ncu-rep --page source --csv > gpurun_out/r02_tc_b256_source.csv
ls -la gpurun_out/ | tail -8
