#!/bin/bash
# round 2, call O: share ratio A/B, bulk-copy staged bilateral (sanitizer, parity, time), bench-state parity test
mkdir -p gpurun_out
echo "== bench-state parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider --timeout 400 --tb=short -x -k "bench_state" > gpurun_out/pytest_o1.log 2>&1; echo rc=$?; tail -n 12 gpurun_out/pytest_o1.log | cut -c1-400
echo "== bilateral bulk under compute-sanitizer"; MFB200_BILATERAL_BULK=1 timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider --timeout 250 -x -k "stagewise_icp" > gpurun_out/bulk_sanitizer.log 2>&1; echo rc=$?; tail -n 6 gpurun_out/bulk_sanitizer.log | cut -c1-300
echo "== bilateral bulk parity"; MFB200_BILATERAL_BULK=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider --timeout 300 --tb=short -x -k "stagewise or degenerate or 720p or sequence" > gpurun_out/pytest_o2.log 2>&1; echo rc=$?; tail -n 5 gpurun_out/pytest_o2.log | cut -c1-300
for bk in 0 1; do
echo "== bench main line BULK=$bk"; MFB200_BILATERAL_BULK=$bk MFB200_BENCH_LEGS=0 timeout 400 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_bulk$bk.json 2> gpurun_out/bench_bulk$bk.err; echo rc=$?; python -c "
import json; b=json.load(open('gpurun_out/bench_bulk$bk.json')); print(b['value'], b['e2e']['value'], b['timed_region']['passes_ms'], {k:v['avg_ms'] for k,v in b['roofline']['kernels'].items()})"; tail -n 3 gpurun_out/bench_bulk$bk.err
done
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
print(json.dumps({"ratio": os.environ.get("MFB200_TRACK_HEAVY_RATIO", "default"), "shares": os.environ.get("MFB200_TRACK_SHARES", "1"), "eight": r["value"], "three": r3["value"]}))
PY
for r in 2 3 5; do MFB200_TRACK_HEAVY_RATIO=$r timeout 300 python /tmp/multi_ab.py 2>&1 | tail -n 1; done
