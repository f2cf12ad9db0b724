#!/bin/bash
# round 2, call R: exact cull of the zero-depth source pixels in the photometric phase (object models): exactness + multi-object throughput
mkdir -p gpurun_out
echo "== parity + multi + sharded"; timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_gpu_sharded.py -q -m gpu -p no:cacheprovider --timeout 400 --tb=short -x > gpurun_out/pytest_r.log 2>&1; echo rc=$?; tail -n 5 gpurun_out/pytest_r.log | cut -c1-300
cat > /tmp/multi_ab.py <<'PY'
import os, sys, json
sys.path.insert(0, os.getcwd())
import torch, bench
import maskfusion_b200 as mfb
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
fr, cls = bench.multi_frames(8, 72)
fr3, cls3 = bench.multi_frames(3, 60)
r = bench.single_process_multi(torch, mfb, stream, 0, fr, cls, timed_from=34)
r3 = bench.single_process_multi(torch, mfb, stream, 0, fr3, cls3, timed_from=20)
print(json.dumps({"eight": r["value"], "three": r3["value"]}))
PY
echo "== multi-object throughput"; timeout 300 python /tmp/multi_ab.py 2>&1 | tail -n 1
echo "== bench main line"; MFB200_BENCH_LEGS=0 timeout 400 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err; echo rc=$?; python -c "
import json; b=json.load(open('gpurun_out/bench_r.json')); print(b['value'], b['e2e']['value'], b['timed_region']['passes_ms'], {k:v['avg_ms'] for k,v in b['roofline']['kernels'].items()})"; tail -n 2 gpurun_out/bench_r.err
echo "== early frames (heavy compaction case): in place vs ping-pong"; for ip in 1 0; do MFB200_CLEAN_INPLACE=$ip MFB200_BENCH_LEGS=0 timeout 300 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r_early$ip.json 2> gpurun_out/bench_r_early$ip.err; python -c "
import json; b=json.load(open('gpurun_out/bench_r_early$ip.json')); print('inplace=$ip', b['value'], {k:v['avg_ms'] for k,v in b['roofline']['kernels'].items() if 'clean' in k})"; done
