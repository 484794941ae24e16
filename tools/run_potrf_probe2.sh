#!/bin/bash
# Cholesky (potrf_h3): chained launches on / off - time, accuracy, per-panel stamps; then the potrf_h3 tests in both modes
cd "$(dirname "$0")/.."
for c in 0 1; do
  echo "== GSMVI_POTRF_CHAIN=$c"
  GSMVI_POTRF_CHAIN=$c timeout 120 python tools/prof_potrf_h3.py 4096 20 check 2>&1 | grep -v cuSOLVER | tail -2
  GSMVI_POTRF_CHAIN=$c timeout 120 python tools/prof_potrf_h3.py 1024 20 check 2>&1 | grep -v cuSOLVER | tail -2
  GSMVI_POTRF_CHAIN=$c GSMVI_POTRF_TIMING=1 timeout 120 python tools/prof_potrf_h3.py 4096 2 2>&1 | grep "potrf_h3 panels" | tail -1
  GSMVI_POTRF_CHAIN=$c timeout 600 python -m pytest tests/test_gsm_gpu.py -q -x -k "potrf_h3" 2>&1 | tail -2
done
