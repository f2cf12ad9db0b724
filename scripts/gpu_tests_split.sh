#!/bin/bash
# GPU tests file by file (a crash in one file must not hide the others); logs -> gpurun_out/pytest_<file>.log
mkdir -p gpurun_out
for f in tests/test_gpu_*.py; do
  b=$(basename $f .py)
  timeout 900 python -X faulthandler -m pytest $f -x -q -m gpu -p no:cacheprovider --timeout 400 --tb=short > gpurun_out/pytest_$b.log 2>&1
  echo "$b rc=$?"; tail -25 gpurun_out/pytest_$b.log | cut -c1-400
done
