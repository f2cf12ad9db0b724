#!/bin/bash
# short validation of the multi-model path: parity tests + the bench's multi-object leg alone
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -p no:cacheprovider --timeout 250 2>&1 | tail -4
timeout 120 python - <<'P'
import torch, bench
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
print(bench.multi_object_leg(torch, s))
P
