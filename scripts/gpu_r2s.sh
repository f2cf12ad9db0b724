#!/bin/bash
# round 2, call S (2 GPUs): per-rank stage table of the sharded frame + the N = 2 bench line
mkdir -p gpurun_out
echo "== shard profile (2 ranks)"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/shard_profile.py > gpurun_out/shard_prof.log 2>&1; echo rc=$?; tail -n 4 gpurun_out/shard_prof.log | cut -c1-2500
echo "== bench N=2"; MFB200_BENCH_LEGS=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?; tail -c 2500 gpurun_out/bench_n2.json; tail -n 3 gpurun_out/bench_n2.err
