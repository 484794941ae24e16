#!/bin/bash
# runs every probe group in its own process, bounded by timeout
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gemm_probe_smi.txt 2>&1
for g in basic1 basic3 round mn epi perf; do
  echo "=== $g"
  timeout 300 python tools/gpu_gemm_probe.py $g 2>&1 | tail -40
  echo "exit=$?"
done 2>&1 | tee gpurun_out/gemm_probe.log
