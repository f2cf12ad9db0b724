#!/bin/bash
# round 2, call M: warp-private candidate staging, out-of-line shared exchange routine; full bench line with the configs[4] leg
mkdir -p gpurun_out
echo "== parity + multi"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider --timeout 300 --tb=short -x > gpurun_out/pytest_m.log 2>&1; echo rc=$?; tail -n 6 gpurun_out/pytest_m.log | cut -c1-300
echo "== track timing"; MFB200_TAG=timing timeout 300 python scripts/track_timing.py > gpurun_out/track_timing_m.json 2> gpurun_out/track_timing_m.err; echo rc=$?; python -c "
import json; t=json.load(open('gpurun_out/track_timing_m.json')); print(t['total_us']); [print(L, {k:(v['n'],v['avg_us']) for k,v in st.items()}) for L,st in t['stages_us'].items()]"; tail -n 3 gpurun_out/track_timing_m.err
echo "== bench (all legs)"; timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_m.json 2> gpurun_out/bench_m.err; echo rc=$?; python -c "
import json; b=json.load(open('gpurun_out/bench_m.json')); print(b['value'], b['e2e']['value'], b['timed_region']['passes_ms'], {k:v['avg_ms'] for k,v in b['roofline']['kernels'].items()})
for k in ('cpu_baseline','cpu_seg','ref_cuda','eight_objects','ate','configs4_720p_16_objects','multi_object','backbone'): print(k, json.dumps(b.get(k))[:700])"; tail -n 3 gpurun_out/bench_m.err
