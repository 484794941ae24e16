#!/bin/bash
# final single-GPU check of the round: GPU test suite, bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r02G_pytest_gpu.log 2>&1; tail -6 gpurun_out/r02G_pytest_gpu.log
python bench.py > gpurun_out/r02G_bench_1gpu.json 2> gpurun_out/r02G_bench_1gpu.err; tail -c 300 gpurun_out/r02G_bench_1gpu.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02G_bench_1gpu.json"))
print("GSM", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["steps"], d["e2e"]["device_rng"]["value"], d["parity"]["relF_cov"], "frac", d["roofline"]["frac"], d["roofline"]["executed_frac"], d["roofline"]["peak"], d["roofline"]["launch_ms"], d["clocks"], d["gpu_launches"])
print("BaM", d["bam"]["value"], d["bam"]["e2e"]["value"], d["bam"]["parity"]["relF_cov"])
PY
