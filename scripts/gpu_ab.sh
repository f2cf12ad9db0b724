#!/bin/bash
# A/B bench of library variants built with MFB200_TAG / MFB200_DEFINES (maskfusion_b200/build.py): usage gpu_ab.sh tag1 tag2 ...
mkdir -p gpurun_out
for t in "" "$@"; do
  echo "== variant '${t}'"
  MFB200_TAG=$t timeout 600 python bench.py --steps 150 --warmup 10 > gpurun_out/bench_ab_${t:-default}.json 2> gpurun_out/bench_ab_${t:-default}.err
  python - <<P
import json
try:
    j = json.loads(open("gpurun_out/bench_ab_${t:-default}.json").read().strip().splitlines()[-1])
    print("  fps", j["value"], "ms", j["ms_per_step"], "e2e", j["e2e"]["value"], "track_ms", j["roofline"]["avg_launch_ms"], {k: v for k, v in list(j["roofline"]["time_shares"].items())[:6]})
except Exception as e:
    print("  failed", e); print(open("gpurun_out/bench_ab_${t:-default}.err").read()[-1500:])
P
done
