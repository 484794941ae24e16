"""Secondary configurations of BASELINE.json (configs[1], configs[2]) on one GPU: iterations/s of the full fit loop.
Writes gpurun_out/bench_configs.json."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gsm-vi_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import gsmvi_oracle as orc
from gsmvi_b200.gsm import GSMEngine
from gsmvi_b200.bam import BaMEngine
from gsmvi_b200.targets import DenseGaussianTarget

out = {}


def timed(fn, n, warm):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(warm, warm + n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


# configs[1]: GSM D=512, B=64, dense-Gaussian target, 1000 iterations
mean_t, cov_t = orc.dense_gaussian_target(512, 0)
tgt = DenseGaussianTarget(mean_t, cov_t)
for npass in (4, 3):
    eng = GSMEngine(512, 64, tgt.lp_g, key=99, npass=npass)
    ms = timed(eng.step, 1000, 20)
    out["gsm_D512_B64_npass%d" % npass] = {"ms_per_iter": ms, "iters_per_s": 1e3 / ms, "launches_per_iter": eng.launches_per_step(),
                                           "algorithmic_gflop_per_iter": (9 * 64 * 512**2 + 512**3 / 3) / 1e9, "reverts": eng.n_reverts}
    print(npass, out["gsm_D512_B64_npass%d" % npass], flush=True)
# configs[2]: BaM D=1024, B=256, ill-conditioned Gaussian target kappa=1e2, schedule 100/(1+i), full and low-rank
mean_t, cov_t = orc.illcond_gaussian_target(1024, 1e2, 0)
tgt = DenseGaussianTarget(mean_t, cov_t)
for lowrank in (False, True):
    eng = BaMEngine(1024, 256, tgt.lp_g, key=99, use_lowrank=lowrank)
    ms = timed(lambda i: eng.step(i, 100.0 / (1 + i)), 20, 3)
    out["bam_D1024_B256_%s" % ("lowrank" if lowrank else "full")] = {
        "ms_per_iter": ms, "iters_per_s": 1e3 / ms, "ns_iters_mean": float(np.mean(eng.ns_iters[3:])), "reverts": eng.n_reverts}
    print(lowrank, out["bam_D1024_B256_%s" % ("lowrank" if lowrank else "full")], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_configs.json"), "w"), indent=1)
