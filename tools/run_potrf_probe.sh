#!/bin/bash
# Cholesky (potrf_h3) probe: time + accuracy at D = 4096 / 1024, phase stamps of panel 8, the potrf_h3 tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for D in 4096 1024; do
  echo "== D=$D"
  timeout 120 python tools/prof_potrf_h3.py $D 20 check 2>&1 | tail -3
done
GSMVI_POTRF_TIMING=1 timeout 120 python tools/prof_potrf_h3.py 4096 2 2>&1 | grep "panel 8" | tail -3
timeout 600 python -m pytest tests/test_gsm_gpu.py -q -x -k "potrf_h3" 2>&1 | tail -3
