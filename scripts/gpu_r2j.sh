#!/bin/bash
# round 2, call J (re-entry): full GPU suite with the flagged exchange as default, stage clock LL vs counter barrier, main line A/B
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 800 --tb=short --durations=14 > gpurun_out/pytest_gpu_j.log 2>&1; echo rc=$?; tail -n 24 gpurun_out/pytest_gpu_j.log | cut -c1-300
for ll in 1 0; do
echo "== track timing LL=$ll"; MFB200_TRACK_LL=$ll MFB200_TAG=timing timeout 300 python scripts/track_timing.py > gpurun_out/track_timing_ll$ll.json 2> gpurun_out/track_timing_ll$ll.err; echo rc=$?; python -c "
import json; t=json.load(open('gpurun_out/track_timing_ll$ll.json')); print(t['total_us']); [print(L, {k:(v['n'],v['avg_us']) for k,v in st.items()}) for L,st in t['stages_us'].items()]"; tail -n 3 gpurun_out/track_timing_ll$ll.err
done
for ll in 1 0; do
echo "== bench main line LL=$ll"; MFB200_TRACK_LL=$ll MFB200_BENCH_LEGS=0 timeout 400 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_ll$ll.json 2> gpurun_out/bench_ll$ll.err; echo rc=$?; python -c "
import json; b=json.load(open('gpurun_out/bench_ll$ll.json')); print(b['value'], b['e2e']['value'], {k:v['avg_ms'] for k,v in b['roofline']['kernels'].items()})"; tail -n 3 gpurun_out/bench_ll$ll.err
done
