#!/bin/bash
# round 2, call K: in-place clean (ticketed tail compaction), warp-per-sub-block keep sums, shuffle scan, invz in the cloud map
mkdir -p gpurun_out
echo "== parity + multi (in-place clean default)"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider --timeout 300 --tb=short -x > gpurun_out/pytest_k.log 2>&1; echo rc=$?; tail -n 8 gpurun_out/pytest_k.log | cut -c1-300
for ip in 1 0; do
echo "== bench main line INPLACE=$ip"; MFB200_CLEAN_INPLACE=$ip MFB200_BENCH_LEGS=0 timeout 400 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_ip$ip.json 2> gpurun_out/bench_ip$ip.err; echo rc=$?; python -c "
import json; b=json.load(open('gpurun_out/bench_ip$ip.json')); print(b['value'], b['e2e']['value'], {k:v['avg_ms'] for k,v in b['roofline']['kernels'].items()})"; tail -n 3 gpurun_out/bench_ip$ip.err
done
