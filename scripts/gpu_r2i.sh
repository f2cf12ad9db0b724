#!/bin/bash
# round 2, call I: flagged exchange of the partial rows in the tracker: exactness, stage clock (with solve sub-stages), bench A/B
mkdir -p gpurun_out
echo "== parity (LL default)"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider --timeout 800 --tb=short -x > gpurun_out/pytest_parity_ll.log 2>&1; echo rc=$?; tail -n 8 gpurun_out/pytest_parity_ll.log | cut -c1-400
echo "== multi + sharded (LL default)"; timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sharded.py -q -m gpu -p no:cacheprovider --timeout 800 --tb=short -x > gpurun_out/pytest_multi_ll.log 2>&1; echo rc=$?; tail -n 8 gpurun_out/pytest_multi_ll.log | cut -c1-400
echo "== track timing LL"; MFB200_TAG=timing timeout 300 python scripts/track_timing.py > gpurun_out/track_timing_ll.json 2> gpurun_out/track_timing_ll.err; echo rc=$?; python -c "
import json; t=json.load(open('gpurun_out/track_timing_ll.json')); print(t['total_us']); [print(L, {k:(v['n'],v['avg_us']) for k,v in st.items()}) for L,st in t['stages_us'].items()]"; tail -n 3 gpurun_out/track_timing_ll.err
echo "== track timing counter barrier"; MFB200_TRACK_LL=0 MFB200_TAG=timing timeout 300 python scripts/track_timing.py > gpurun_out/track_timing_noll.json 2> gpurun_out/track_timing_noll.err; echo rc=$?; python -c "
import json; t=json.load(open('gpurun_out/track_timing_noll.json')); print(t['total_us']); [print(L, {k:(v['n'],v['avg_us']) for k,v in st.items()}) for L,st in t['stages_us'].items()]"
for ll in 1 0; do
echo "== bench main line LL=$ll"; MFB200_TRACK_LL=$ll MFB200_BENCH_LEGS=0 timeout 400 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_ll$ll.json 2> gpurun_out/bench_ll$ll.err; echo rc=$?; python -c "
import json; b=json.load(open('gpurun_out/bench_ll$ll.json')); print(b['value'], b['e2e']['value'], {k:v['avg_ms'] for k,v in b['roofline']['kernels'].items()})"; tail -n 3 gpurun_out/bench_ll$ll.err
done
cat > /tmp/multi_ab.py <<'PY'
import os, sys, json
sys.path.insert(0, os.getcwd())
import torch, bench
import maskfusion_b200 as mfb
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
fr, cls = bench.multi_frames(8, 72)
r = bench.single_process_multi(torch, mfb, stream, 0, fr, cls, timed_from=34)
fr3, cls3 = bench.multi_frames(3, 60)
r3 = bench.single_process_multi(torch, mfb, stream, 0, fr3, cls3, timed_from=20)
print(json.dumps({"ll": os.environ.get("MFB200_TRACK_LL", "default"), "eight": r["value"], "three": r3["value"]}))
PY
echo "== multi-object throughput LL=1"; MFB200_TRACK_LL=1 timeout 300 python /tmp/multi_ab.py 2>&1 | tail -n 1
echo "== multi-object throughput LL=0"; MFB200_TRACK_LL=0 timeout 300 python /tmp/multi_ab.py 2>&1 | tail -n 1
