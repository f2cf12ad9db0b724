#!/bin/bash
# round 2, call B: tracker stage clock, new multi-model flow tests, reference tracking schedule timing
mkdir -p gpurun_out
echo "== track timing"; MFB200_TAG=timing timeout 300 python scripts/track_timing.py > gpurun_out/track_timing.json 2> gpurun_out/track_timing.err; echo rc=$?; head -c 3000 gpurun_out/track_timing.json; tail -n 3 gpurun_out/track_timing.err
echo "== ref track"; timeout 300 python scripts/time_ref_track.py > gpurun_out/ref_track.json 2> gpurun_out/ref_track.err; echo rc=$?; cat gpurun_out/ref_track.json; tail -n 3 gpurun_out/ref_track.err
for t in "tests/test_gpu_multi.py" "tests/test_gpu_seg.py::test_multi_model_with_both_closes" "tests/test_gpu_sharded.py"; do
  b=$(echo $t | tr '/:' '__')
  timeout 1200 python -X faulthandler -m pytest "$t" -q -m gpu -p no:cacheprovider --timeout 1000 --tb=short --durations=8 > gpurun_out/pytest_$b.log 2>&1
  echo "$t rc=$?"; tail -n 30 gpurun_out/pytest_$b.log | cut -c1-600
done
