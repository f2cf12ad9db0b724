#!/bin/bash
# round 2, final call: whole GPU suite, full bench line, reference arm, ncu launch list + one multi-kernel `--set full` capture of a bench frame
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo rc=$?; tail -n 2 gpurun_out/smoke.log
echo "== pytest gpu (all)"; SECONDS=0; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 800 --tb=short --durations=8 > gpurun_out/pytest_gpu_final.log 2>&1; echo rc=$? seconds=$SECONDS; tail -n 14 gpurun_out/pytest_gpu_final.log | cut -c1-300
echo "== bench (all legs)"; SECONDS=0; timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo rc=$? seconds=$SECONDS; python -c "
import json; b=json.load(open('gpurun_out/bench_final.json')); print(b['value'], b['e2e']['value'], b['timed_region']['passes_ms'], b['roofline']['frac'], {k:v['avg_ms'] for k,v in b['roofline']['kernels'].items()})
for k in ('cpu_baseline','eight_objects','ate','configs4_720p_16_objects','multi_object','backbone'): print(k, json.dumps(b.get(k))[:500])"; tail -n 3 gpurun_out/bench_final.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json
echo "== ncu launch list"; MFB200_BENCH_LEGS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 330 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 2 > gpurun_out/ncu_bench.log 2>&1; echo rc=$?
echo "== ncu full (one frame: tracker, index projection, clean p1/p2/compact, splat)"; MFB200_BENCH_LEGS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_track_persistent|k_index_project|k_clean_p1|k_clean_p2|k_clean_compact|k_splat_project" -s 18 -c 6 -f -o gpurun_out/prof_multi_r02 python bench.py --steps 3 --warmup 2 > gpurun_out/ncu_multi.log 2>&1; echo rc=$?; tail -n 3 gpurun_out/ncu_multi.log | cut -c1-200
echo done
