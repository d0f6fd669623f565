set -x
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 ) > gpurun_out/r02_bench_tc_n$N.json 2> gpurun_out/r02_bench_tc_n$N.err
tail -c 1800 gpurun_out/r02_bench_tc_n$N.json; tail -2 gpurun_out/r02_bench_tc_n$N.err
