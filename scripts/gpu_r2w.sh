#!/bin/bash
# round 2, call W: the fallback path of the clean window (index-map images instead of packed texels)
timeout 150 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider --timeout 120 --tb=short -x -k "clean_with_other_threshold" 2>&1 | tail -n 8 | cut -c1-400
