#!/bin/bash
# round 2, call U: splat units of small stores shared by 8 block columns; share logic factored out: multi-model exactness + throughput
mkdir -p gpurun_out
echo "== multi + seg"; timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_seg.py -q -m gpu -p no:cacheprovider --timeout 300 --tb=short -x > gpurun_out/pytest_u.log 2>&1; echo rc=$?; tail -n 4 gpurun_out/pytest_u.log | cut -c1-300
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
