set -x
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 ) > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -c 2600 gpurun_out/r02_bench_n$N.json; tail -2 gpurun_out/r02_bench_n$N.err
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --workload text2audio --steps 2 --warmup 1 ) > gpurun_out/r02_t2a_n$N.json 2> gpurun_out/r02_t2a_n$N.err
tail -c 1200 gpurun_out/r02_t2a_n$N.json; tail -2 gpurun_out/r02_t2a_n$N.err
