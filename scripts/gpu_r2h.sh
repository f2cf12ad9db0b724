#!/bin/bash
# round 2, call H: object-sharded bench at N GPUs (default steps, as the driver launches it)
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi -L | head -n 8
echo "== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 60 --warmup 6 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo rc=$?; tail -c 3000 gpurun_out/bench_n$N.json; tail -n 8 gpurun_out/bench_n$N.err | cut -c1-300
echo "== bench --impl reference N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo rc=$?; tail -c 1200 gpurun_out/bench_ref_n$N.json; tail -n 4 gpurun_out/bench_ref_n$N.err | cut -c1-300
