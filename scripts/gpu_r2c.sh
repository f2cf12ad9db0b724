#!/bin/bash
# round 2, call C (gpurun --gpus 2): in-library NCCL exchange: sharded == single process (bit for bit), bench at N = 2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi2.txt 2>&1
echo "== sharded tests (2 GPUs: NCCL inside the library)"; timeout 900 python -X faulthandler -m pytest tests/test_gpu_sharded.py -q -m gpu -p no:cacheprovider --timeout 800 --tb=short > gpurun_out/pytest_sharded_nccl.log 2>&1; echo rc=$?; tail -n 25 gpurun_out/pytest_sharded_nccl.log | cut -c1-600
echo "== bench N=2"; NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 60 --warmup 6 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?; tail -c 4000 gpurun_out/bench_n2.json; tail -n 8 gpurun_out/bench_n2.err | cut -c1-300
