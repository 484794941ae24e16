"""2-GPU test of the batch-sharded path (one process per GPU, NCCL): the fit with the batch split over two ranks must
reproduce the single-GPU fit on the same z-tape.  Skipped unless at least two GPUs are visible."""
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
ROOT = sys.argv[1]
sys.path.insert(0, os.path.join(ROOT, "gsm-vi_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch, torch.distributed as dist
import gsmvi_oracle as orc
from gsmvi_b200.gsm import GSM
from gsmvi_b200.bam import BaM, Regularizers
from gsmvi_b200.targets import DenseGaussianTarget
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
relF = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
for (D, B, niter) in [(96, 32, 20), (512, 256, 6)]:
    mean_t, cov_t = orc.dense_gaussian_target(D, 0)
    tgt = DenseGaussianTarget(mean_t, cov_t)
    Z = np.random.RandomState(1).normal(size=(niter + 1, B, D)).astype(np.float32)
    m1, c1 = GSM(D, tgt.lp, tgt.lp_g).fit(0, niter=niter, batch_size=B, z_tape=Z, verbose=False)
    m2, c2 = GSM(D, tgt.lp, tgt.lp_g).fit(0, niter=niter, batch_size=B, z_tape=Z, verbose=False, process_group=dist.group.WORLD)
    e = (relF(c2, c1), relF(m2, m1))
    if rank == 0: print("GSM D=%d B=%d sharded-vs-single relF cov %.2e mean %.2e" % ((D, B) + e), flush=True)
    assert e[0] < 2e-5 and e[1] < 2e-5, e
    P = tgt.P.cpu().double().numpy(); c = tgt.c[:D].cpu().double().numpy()
    m_o, c_o = orc.GSM(D, None, lambda x: -(x @ P.T) + c).fit(0, niter=niter, batch_size=B, sampler=orc.CholeskyTapeSampler(Z.astype(np.float64)))
    assert relF(c2.cpu(), torch.as_tensor(c_o)) < 1e-4
    b1m, b1c = BaM(D, tgt.lp, tgt.lp_g).fit(0, Regularizers().custom(lambda i: 100 / (1 + i)), niter=niter, batch_size=B, z_tape=Z, verbose=False)
    b2m, b2c = BaM(D, tgt.lp, tgt.lp_g).fit(0, Regularizers().custom(lambda i: 100 / (1 + i)), niter=niter, batch_size=B, z_tape=Z, verbose=False, process_group=dist.group.WORLD)
    e = (relF(b2c, b1c), relF(b2m, b1m))
    if rank == 0: print("BaM D=%d B=%d sharded-vs-single relF cov %.2e mean %.2e" % ((D, B) + e), flush=True)
    assert e[0] < 2e-5 and e[1] < 2e-5, e
    # both ranks hold identical replicated state
    t = b2c.clone(); dist.broadcast(t, 0); assert torch.equal(t, b2c)
# ensemble of independent fits: each rank fits its slice (no communication), the gathered result equals the unsharded run
from gsmvi_b200.ensemble import gsm_ensemble_fit, shard_range
F, D, B, niter = 10, 24, 8, 30
rng = np.random.RandomState(0)
means = rng.random_sample((F, D)); A = rng.normal(size=(F, D, D)); covs = A @ np.swapaxes(A, 1, 2) / D + 1e-2 * np.eye(D)
mu_all, S_all, rev_all = gsm_ensemble_fit(means, covs, key=3, batch_size=B, niter=niter)
mu_loc, S_loc, rev_loc = gsm_ensemble_fit(means, covs, key=3, batch_size=B, niter=niter, process_group=dist.group.WORLD)
lo, hi = shard_range(F, rank, dist.get_world_size())
assert torch.equal(S_loc, S_all[lo:hi]) and torch.equal(mu_loc, mu_all[lo:hi]) and torch.equal(rev_loc, rev_all[lo:hi])
dist.destroy_process_group()
print("rank", rank, "ok", flush=True)
'''


def test_two_gpu_sharded_fit_matches_single_gpu(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29621", str(script), ROOT],
                       capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" ok") == 2
