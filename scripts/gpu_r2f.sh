#!/bin/bash
# round 2, call F: object-model validity bitmask in the tracker: exactness + multi-object throughput A/B
mkdir -p gpurun_out
echo "== multi exactness with the bitmask"; MFB200_TRACK_BITS=1 timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sharded.py -q -m gpu -p no:cacheprovider --timeout 800 --tb=short --durations=5 > gpurun_out/pytest_multi_bits.log 2>&1; echo rc=$?; tail -n 20 gpurun_out/pytest_multi_bits.log | cut -c1-500
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
print(json.dumps({"bits": os.environ.get("MFB200_TRACK_BITS", "0"), "eight": r["value"], "three": r3["value"]}))
PY
echo "== multi-object throughput, bitmask off"; MFB200_TRACK_BITS=0 timeout 300 python /tmp/multi_ab.py 2>&1 | tail -n 2
echo "== multi-object throughput, bitmask on"; MFB200_TRACK_BITS=1 timeout 300 python /tmp/multi_ab.py 2>&1 | tail -n 2
