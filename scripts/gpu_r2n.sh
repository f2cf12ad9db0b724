#!/bin/bash
# round 2, call N: per-model shares of the tracker's grid (flat grid), inlined exchange again; full suite of the multi-model tests; full bench line
mkdir -p gpurun_out
echo "== parity + multi + sharded + seg"; timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_gpu_sharded.py tests/test_gpu_seg.py -q -m gpu -p no:cacheprovider --timeout 400 --tb=short -x > gpurun_out/pytest_n.log 2>&1; echo rc=$?; tail -n 6 gpurun_out/pytest_n.log | cut -c1-300
echo "== track timing"; MFB200_TAG=timing timeout 300 python scripts/track_timing.py > gpurun_out/track_timing_n.json 2> gpurun_out/track_timing_n.err; echo rc=$?; python -c "
import json; t=json.load(open('gpurun_out/track_timing_n.json')); print(t['total_us']); [print(L, {k:(v['n'],v['avg_us']) for k,v in st.items()}) for L,st in t['stages_us'].items()]"; tail -n 3 gpurun_out/track_timing_n.err
echo "== bench (all legs)"; SECONDS=0; timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_n.json 2> gpurun_out/bench_n.err; echo rc=$? seconds=$SECONDS; python -c "
import json; b=json.load(open('gpurun_out/bench_n.json')); print(b['value'], b['e2e']['value'], b['timed_region']['passes_ms'], {k:v['avg_ms'] for k,v in b['roofline']['kernels'].items()})
for k in ('cpu_baseline','cpu_seg','ref_cuda','eight_objects','ate','configs4_720p_16_objects','multi_object','backbone'): print(k, json.dumps(b.get(k))[:700])"; tail -n 3 gpurun_out/bench_n.err
echo "== eight objects, equal shares (A/B)"; cat > /tmp/multi_ab.py <<'PY'
import os, sys, json
sys.path.insert(0, os.getcwd())
import torch, bench
import maskfusion_b200 as mfb
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
fr, cls = bench.multi_frames(8, 72)
r = bench.single_process_multi(torch, mfb, stream, 0, fr, cls, timed_from=34)
fr3, cls3 = bench.multi_frames(3, 60)
r3 = bench.single_process_multi(torch, mfb, stream, 0, fr3, cls3, timed_from=20)
print(json.dumps({"shares": os.environ.get("MFB200_TRACK_SHARES", "default"), "eight": r["value"], "three": r3["value"]}))
PY
MFB200_TRACK_SHARES=0 timeout 300 python /tmp/multi_ab.py 2>&1 | tail -n 1
