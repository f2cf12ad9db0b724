#!/bin/bash
# round 2, call E: bbox fix + tracker input cache: exactness, stage clock, bench A/B; TMA bilateral diagnosis
mkdir -p gpurun_out
echo "== TMA check"; timeout 200 python scripts/tma_check.py 2>&1 | tail -n 6
echo "== TMA under compute-sanitizer"; MFB200_BILATERAL_TMA=1 timeout 200 compute-sanitizer --tool memcheck --print-limit 5 python scripts/tma_check.py one /tmp/x.npy > gpurun_out/tma_sanitizer.log 2>&1; grep -E "Illegal|Invalid|error|at |Host Frame" gpurun_out/tma_sanitizer.log | head -n 12
echo "== track timing (cache)"; MFB200_TRACK_CACHE=1 MFB200_TAG=timing timeout 300 python scripts/track_timing.py > gpurun_out/track_timing_cache.json 2> gpurun_out/track_timing_cache.err; echo rc=$?; head -c 2500 gpurun_out/track_timing_cache.json | tr -d '\n '; echo; tail -n 3 gpurun_out/track_timing_cache.err
echo "== parity (cache)"; MFB200_TRACK_CACHE=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider --timeout 800 --tb=short > gpurun_out/pytest_parity_cache.log 2>&1; echo rc=$?; tail -n 12 gpurun_out/pytest_parity_cache.log | cut -c1-400
echo "== multi (cache, bbox fix)"; MFB200_TRACK_CACHE=1 timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_seg.py -q -m gpu -p no:cacheprovider --timeout 800 --tb=short --durations=6 > gpurun_out/pytest_multi_cache.log 2>&1; echo rc=$?; tail -n 25 gpurun_out/pytest_multi_cache.log | cut -c1-500
echo "== bench cache on"; MFB200_TRACK_CACHE=1 MFB200_BENCH_LEGS=0 timeout 400 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_cache.json 2> gpurun_out/bench_cache.err; echo rc=$?; python -c "
import json; b=json.load(open('gpurun_out/bench_cache.json')); print(b['value'], b['e2e']['value'], {k:v['avg_ms'] for k,v in b['roofline']['kernels'].items()})"
echo "== bench cache off"; MFB200_TRACK_CACHE=0 MFB200_BENCH_LEGS=0 timeout 400 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_nocache.json 2> gpurun_out/bench_nocache.err; echo rc=$?; python -c "
import json; b=json.load(open('gpurun_out/bench_nocache.json')); print(b['value'], b['e2e']['value'], {k:v['avg_ms'] for k,v in b['roofline']['kernels'].items()})"
echo "== multi overlap (single process) exactness"; MFB200_MULTI_OVERLAP=1 MFB200_TRACK_CACHE=1 timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider --timeout 500 --tb=short -k "three_tracked or static_objects or table_scene_eight" > gpurun_out/pytest_multi_overlap.log 2>&1; echo rc=$?; tail -n 8 gpurun_out/pytest_multi_overlap.log | cut -c1-400
