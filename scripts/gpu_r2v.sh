#!/bin/bash
# round 2, call V: last check of the final tree: smoke + sharded suite (two ranks on one GPU over gloo) + parity quick
mkdir -p gpurun_out
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
echo "== sharded + parity quick"; timeout 500 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider --timeout 300 --tb=short -x -k "shards or stagewise or bench_state" > gpurun_out/pytest_v.log 2>&1; echo rc=$?; tail -n 4 gpurun_out/pytest_v.log | cut -c1-300
