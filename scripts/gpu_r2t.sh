#!/bin/bash
# round 2, call T: adaptive in-place / ping-pong clean (choice from the previous frame's moved fraction), zero-depth cull removed
mkdir -p gpurun_out
echo "== parity + multi"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider --timeout 400 --tb=short -x > gpurun_out/pytest_t.log 2>&1; echo rc=$?; tail -n 5 gpurun_out/pytest_t.log | cut -c1-300
echo "== early frames"; MFB200_BENCH_LEGS=0 timeout 300 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_t_early.json 2> gpurun_out/bench_t_early.err; python -c "
import json; b=json.load(open('gpurun_out/bench_t_early.json')); print(b['value'], {k:v['avg_ms'] for k,v in b['roofline']['kernels'].items() if 'clean' in k}, {k:v for k,v in b['roofline']['time_shares'].items() if 'clean' in k})"
echo "== bench main line"; MFB200_BENCH_LEGS=0 timeout 400 python bench.py > gpurun_out/bench_t.json 2> gpurun_out/bench_t.err; echo rc=$?; python -c "
import json; b=json.load(open('gpurun_out/bench_t.json')); print(b['value'], b['e2e']['value'], b['timed_region']['passes_ms'], b['roofline']['kernel'], {k:v['avg_ms'] for k,v in b['roofline']['kernels'].items()})"; tail -n 2 gpurun_out/bench_t.err
