#!/bin/bash
# round-2 final single-GPU evidence: GPU test suite, bench line, ncu launch list of the bench, ncu --set full of the fused
# Cholesky panel kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r02F_pytest_gpu.log 2>&1; tail -4 gpurun_out/r02F_pytest_gpu.log
python bench.py > gpurun_out/r02F_bench_1gpu.json 2> gpurun_out/r02F_bench_1gpu.err; tail -c 300 gpurun_out/r02F_bench_1gpu.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02F_bench_1gpu.json"))
print("GSM", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["steps"], d["e2e"]["device_rng"]["value"], d["parity"]["relF_cov"], "frac", d["roofline"]["frac"], d["roofline"]["executed_frac"], d["roofline"]["peak"], d["roofline"]["launch_ms"], d["clocks"])
print("BaM", d["bam"]["value"], d["bam"]["e2e"]["value"], d["bam"]["parity"]["relF_cov"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02F_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-bam --no-cpu-baseline > gpurun_out/r02F_launches_bench.log 2>&1
python tools/parse_launches.py gpurun_out/r02F_launches_bench.csv > gpurun_out/r02F_launches_bench.txt 2>&1; head -12 gpurun_out/r02F_launches_bench.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:potrf_fused -s 40 -c 3 -o gpurun_out/r02F_prof_potrf_fused -f python tools/prof_potrf_h3.py 4096 3 > gpurun_out/r02F_ncu_potrf.log 2>&1; tail -2 gpurun_out/r02F_ncu_potrf.log
