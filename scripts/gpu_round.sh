#!/bin/bash
# One gpurun call: smoke + GPU parity tests + bench + ncu launch list.  Logs -> gpurun_out/.
# usage: scripts/gpu_round.sh [quick|full]
set -u
mkdir -p gpurun_out
MODE=${1:-full}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -3 gpurun_out/smoke.log
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -25 gpurun_out/pytest_gpu.log
echo "== bench" ; timeout 600 python bench.py --steps 30 --warmup 4 > gpurun_out/bench.json 2> gpurun_out/bench.err ; echo "bench rc=$?" ; tail -c 3000 gpurun_out/bench.json ; tail -5 gpurun_out/bench.err
if [ "$MODE" = "full" ]; then
  echo "== bench reference arm" ; timeout 600 python bench.py --impl reference --steps 3 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ; tail -c 1500 gpurun_out/bench_ref.json
  echo "== ncu launch list" ; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 2 > gpurun_out/ncu_bench.log 2>&1 ; echo "ncu rc=$?"
fi
echo done
