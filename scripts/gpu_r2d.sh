#!/bin/bash
# round 2, call D: ncu evidence -- launch list of the bench frames + full captures of the main kernels
mkdir -p gpurun_out
echo "== ncu launch list"; MFB200_BENCH_LEGS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 2 > gpurun_out/ncu_bench.log 2>&1 ; echo "rc=$?"
for K in k_track_cluster k_track_persistent k_splat_project k_clean_p1 k_clean_p2 k_bilateral; do
  echo "== ncu full $K"; MFB200_BENCH_LEGS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -f -o gpurun_out/prof_$K python bench.py --steps 3 --warmup 2 > gpurun_out/ncu_$K.log 2>&1 ; echo "rc=$?"
done
ls -la gpurun_out/*.ncu-rep | head
