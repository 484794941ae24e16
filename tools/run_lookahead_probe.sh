#!/bin/bash
# look-ahead Cholesky: timing with / without, phase stamps of panel 8 (warm: every launch uses the timing variant), tests
cd "$(dirname "$0")/.."
for D in 4096 1024; do
  for m in 1 0; do
    echo "== lookahead=$m D=$D"
    GSMVI_POTRF_LOOKAHEAD=$m timeout 120 python tools/prof_potrf_h3.py $D 20 check 2>&1 | tail -2
  done
done
GSMVI_POTRF_TIMING=1 timeout 120 python tools/prof_potrf_h3.py 4096 2 2>&1 | grep "panel 8" | tail -3
timeout 600 python -m pytest tests/test_gsm_gpu.py -q -x -k "potrf_h3" 2>&1 | tail -3
