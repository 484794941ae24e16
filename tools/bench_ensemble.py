"""BASELINE configs[4]: ensemble of 1024 independent GSM fits, D = 64, batch 32 on one GPU (they split across GPUs by
slicing the fits: no communication).  Writes gpurun_out/bench_ensemble.json."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gsm-vi_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import gsmvi_oracle as orc
from gsmvi_b200.ensemble import gsm_ensemble_fit
F, D, B, niter = 1024, 64, 32, 500
rng = np.random.RandomState(0)
means = rng.random_sample((F, D))
A = rng.normal(size=(F, D, D))
covs = A @ np.swapaxes(A, 1, 2) / D + 1e-3 * np.eye(D)
gsm_ensemble_fit(means[:8], covs[:8], key=1, batch_size=B, niter=10)  # warm-up / load
torch.cuda.synchronize()
t0 = time.perf_counter()
mu, S, rev = gsm_ensemble_fit(means, covs, key=1, batch_size=B, niter=niter)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
err = float(((S.double().cpu() - torch.as_tensor(covs)).flatten(1).norm(dim=1) / torch.as_tensor(covs).flatten(1).norm(dim=1)).median())
# CPU stand-in: the oracle loop on 2 fits, extrapolated (labelled as such)
_, lp_g, _ = orc.gaussian_score_fns(means[0], covs[0])
t1 = time.perf_counter()
orc.GSM(D, None, lp_g).fit(0, niter=niter, batch_size=B, sampler=orc.CholeskyTapeSampler(rng.normal(size=(niter + 1, B, D))))
cpu_one = time.perf_counter() - t1
out = {"fits": F, "D": D, "B": B, "niter": niter, "seconds_incl_setup_h2d": dt, "fit_iterations_per_s": F * (niter + 1) / dt,
       "median_relF_cov_vs_target": err, "max_reverts": int(rev.max()),
       "cpu_oracle_one_fit_s": cpu_one, "cpu_oracle_extrapolated_1024_fits_s": cpu_one * F}
print(out)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_ensemble.json"), "w"), indent=1)
