#!/bin/bash
# round 2, call B: tracker stage clock (cluster on/off), TMA bilateral check, new multi-model flow tests, reference tracking schedule, bench
mkdir -p gpurun_out
echo "== track timing (cluster)"; MFB200_TRACK_CLUSTER=1 MFB200_TAG=timing timeout 300 python scripts/track_timing.py > gpurun_out/track_timing.json 2> gpurun_out/track_timing.err; echo rc=$?; head -c 2500 gpurun_out/track_timing.json | tr -d '\n ' ; echo; tail -n 3 gpurun_out/track_timing.err
echo "== track timing (no cluster)"; MFB200_TRACK_CLUSTER=0 MFB200_TAG=timing timeout 300 python scripts/track_timing.py > gpurun_out/track_timing_nocl.json 2> gpurun_out/track_timing_nocl.err; echo rc=$?; head -c 2500 gpurun_out/track_timing_nocl.json | tr -d '\n '; echo
echo "== parity quick (conservative defaults)"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider --timeout 500 --tb=short -k "sequence_ate" > gpurun_out/pytest_parity_defaults.log 2>&1; echo rc=$?; tail -n 5 gpurun_out/pytest_parity_defaults.log | cut -c1-300
echo "== parity quick (cluster tracker, fused index/clean, tile CC)"; MFB200_TRACK_CLUSTER=1 MFB200_FUSE_INDEX=1 MFB200_CC_TILE=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider --timeout 500 --tb=short -k "sequence_ate or stagewise_rgbd" > gpurun_out/pytest_parity_quick.log 2>&1; echo rc=$?; tail -n 15 gpurun_out/pytest_parity_quick.log | cut -c1-500
echo "== bilateral through TMA"; MFB200_TRACK_CLUSTER=1 MFB200_FUSE_INDEX=1 MFB200_BILATERAL_TMA=1 timeout 240 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider --timeout 200 --tb=short -k "stagewise_icp or degenerate" > gpurun_out/pytest_tma.log 2>&1; echo rc=$?; tail -n 8 gpurun_out/pytest_tma.log | cut -c1-400
echo "== ref track"; timeout 300 python scripts/time_ref_track.py > gpurun_out/ref_track.json 2> gpurun_out/ref_track.err; echo rc=$?; cat gpurun_out/ref_track.json; tail -n 3 gpurun_out/ref_track.err
export MFB200_TRACK_CLUSTER=1 MFB200_FUSE_INDEX=1 MFB200_CC_TILE=1
for t in "tests/test_gpu_multi.py" "tests/test_gpu_seg.py::test_multi_model_with_both_closes" "tests/test_gpu_sharded.py"; do
  b=$(echo $t | tr '/:' '__')
  timeout 1200 python -X faulthandler -m pytest "$t" -q -m gpu -p no:cacheprovider --timeout 1000 --tb=short --durations=8 > gpurun_out/pytest_$b.log 2>&1
  echo "$t rc=$?"; tail -n 30 gpurun_out/pytest_$b.log | cut -c1-600
done
echo "== bench"; timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 6000 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
echo "== bench (TMA bilateral)"; MFB200_BENCH_LEGS=0 MFB200_BILATERAL_TMA=1 timeout 400 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_tma.json 2> gpurun_out/bench_tma.err; echo "rc=$?"; head -c 400 gpurun_out/bench_tma.json
echo "== bench (conservative defaults: one cooperative tracking launch, separate index/clean)"; MFB200_TRACK_CLUSTER=0 MFB200_FUSE_INDEX=0 MFB200_CC_TILE=0 MFB200_BENCH_LEGS=0 timeout 400 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_defaults.json 2> gpurun_out/bench_defaults.err; echo "rc=$?"; head -c 400 gpurun_out/bench_defaults.json
