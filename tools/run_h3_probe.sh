#!/bin/bash
# h3 (scaled 3xFP16) GEMM probe groups, each in its own process, bounded by timeout; plus the potrf phase timing
mkdir -p gpurun_out
for g in h3 h3mn h3perf; do
  echo "=== $g"
  timeout 300 python tools/gpu_gemm_probe.py $g 2>&1 | tail -60
  echo "exit=$?"
done 2>&1 | tee gpurun_out/h3_probe.log
echo "=== potrf timing" | tee -a gpurun_out/h3_probe.log
GSMVI_POTRF_TIMING=1 timeout 120 python tools/prof_potrf.py 2>&1 | tail -8 | tee -a gpurun_out/h3_probe.log
