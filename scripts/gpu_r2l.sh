#!/bin/bash
# round 2, call L: solve in four stages (register-fed LDLT), compile-time exchange, packed 16-byte window texels (nine loads at once), p1 prefetch
mkdir -p gpurun_out
echo "== parity + multi + seg"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_gpu_seg.py -q -m gpu -p no:cacheprovider --timeout 300 --tb=short -x > gpurun_out/pytest_l.log 2>&1; echo rc=$?; tail -n 8 gpurun_out/pytest_l.log | cut -c1-300
echo "== track timing"; MFB200_TAG=timing timeout 300 python scripts/track_timing.py > gpurun_out/track_timing_l.json 2> gpurun_out/track_timing_l.err; echo rc=$?; python -c "
import json; t=json.load(open('gpurun_out/track_timing_l.json')); print(t['total_us']); [print(L, {k:(v['n'],v['avg_us']) for k,v in st.items()}) for L,st in t['stages_us'].items()]"; tail -n 3 gpurun_out/track_timing_l.err
echo "== bench main line"; MFB200_BENCH_LEGS=0 timeout 400 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_l.json 2> gpurun_out/bench_l.err; echo rc=$?; python -c "
import json; b=json.load(open('gpurun_out/bench_l.json')); print(b['value'], b['e2e']['value'], b['timed_region']['passes_ms'], {k:v['avg_ms'] for k,v in b['roofline']['kernels'].items()})"; tail -n 3 gpurun_out/bench_l.err
