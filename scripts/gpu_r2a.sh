#!/bin/bash
# round 2, call A: all GPU tests file by file + quick bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
for f in tests/test_gpu_parity.py tests/test_gpu_ref.py tests/test_gpu_seg.py tests/test_gpu_multi.py tests/test_gpu_sharded.py tests/test_gpu_cnn.py; do
  b=$(basename $f .py)
  timeout 900 python -X faulthandler -m pytest $f -q -m gpu -p no:cacheprovider --timeout 800 --tb=short > gpurun_out/pytest_$b.log 2>&1
  echo "$b rc=$?"; tail -n 25 gpurun_out/pytest_$b.log | cut -c1-600
done
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
