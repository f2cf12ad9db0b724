#!/bin/bash
# One gpurun call: smoke + GPU parity tests + bench + ncu.  Logs -> gpurun_out/.
# usage: scripts/gpu_round.sh [quick|full] [ncu kernel regexes...]
set -u
mkdir -p gpurun_out
MODE=${1:-full}; shift || true
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -2 gpurun_out/smoke.log
if [ "${CNN_FIRST:-0}" = "1" ]; then
  echo "== pytest cnn (isolated, short timeout: a hung tcgen05 pipeline must not take the box down)"
  timeout 240 python -m pytest tests/test_gpu_cnn.py -q -m gpu -p no:cacheprovider --timeout 120 > gpurun_out/pytest_cnn.log 2>&1 ; echo "pytest cnn rc=$?" ; tail -30 gpurun_out/pytest_cnn.log | cut -c1-600
fi
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 ${PYTEST_EXTRA:-} > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -12 gpurun_out/pytest_gpu.log | cut -c1-400
echo "== bench" ; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err ; echo "bench rc=$?" ; tail -c 3500 gpurun_out/bench.json ; tail -5 gpurun_out/bench.err
if [ "$MODE" = "full" ]; then
  echo "== bench reference arm" ; timeout 600 python bench.py --impl reference --steps 3 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ; tail -c 1200 gpurun_out/bench_ref.json
  echo "== ncu launch list" ; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 330 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 2 > gpurun_out/ncu_bench.log 2>&1 ; echo "ncu rc=$?"
  echo "== backbone" ; timeout 300 python scripts/bench_cnn.py 1024 10 > gpurun_out/bench_cnn.json 2> gpurun_out/bench_cnn.err ; tail -c 400 gpurun_out/bench_cnn.json
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm -s 40 -c 12 -f -o gpurun_out/prof_k_gemm python scripts/bench_cnn.py 1024 2 > gpurun_out/ncu_k_gemm.log 2>&1 ; echo "ncu gemm rc=$?"
  for K in "$@"; do
    echo "== ncu full $K" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -f -o gpurun_out/prof_$K python bench.py --steps 3 --warmup 2 > gpurun_out/ncu_$K.log 2>&1 ; echo "rc=$?"
  done
fi
echo done
