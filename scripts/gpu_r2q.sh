#!/bin/bash
# round 2, call Q: splat rasteriser at 16 blocks per SM (32 registers, 88 B of spills) vs 12 (40 registers)
mkdir -p gpurun_out
for t in splat16 splat12; do
  echo "== bench main line tag='$t'"; MFB200_TAG=$t MFB200_BENCH_LEGS=0 timeout 400 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_q_$t.json 2> gpurun_out/bench_q_$t.err; echo rc=$?; python -c "
import json; b=json.load(open('gpurun_out/bench_q_$t.json')); print(b['value'], b['e2e']['value'], b['timed_region']['passes_ms'], {k:v['avg_ms'] for k,v in b['roofline']['kernels'].items()})"; tail -n 2 gpurun_out/bench_q_$t.err
done
echo "== parity quick (splat16)"; MFB200_TAG=splat16 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider --timeout 250 -x -k "stagewise or bench_state" 2>&1 | tail -n 3
