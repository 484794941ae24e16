"""Ensembles of independent small GSM fits (BASELINE config 5: 1024 fits, D = 64, batch 32).

The reference has no ensemble API: a user would loop `GSM(D, lp, lp_g).fit(...)` over targets (gsmvi/gsm.py:79-133).
Here every fit runs its whole loop inside one CTA of a single kernel launch (csrc/gsm_ensemble.cu); fits are
independent, so ensembles split across GPUs by slicing the leading axis - no communication."""
import numpy as np
import torch

from . import _lib as L
from ._util import device, key_to_seed, to_dev


def shard_range(F, rank, world):
    """Contiguous slice [lo, hi) of F independent fits owned by `rank` of `world` (sizes differ by at most one)."""
    base, rem = divmod(F, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gsm_ensemble_fit(means_t, covs_t, key, mean=None, cov=None, batch_size=32, niter=1000, z_tape=None, process_group=None):
    """Fit F dense-Gaussian targets N(means_t[f], covs_t[f]) independently with GSM.

    means_t [F, D], covs_t [F, D, D] (numpy, fp64 used for the precision matrices); optional initial mean [F, D] /
    cov [F, D, D] (defaults 0 / I, gsm.py:100-103); z_tape optional [F, niter+1, batch_size, D].
    process_group: torch.distributed group (one process per GPU): this rank fits only its contiguous slice
    `shard_range(F, rank, world)` of the ensemble - replicas only, no communication - and returns that slice's results
    (fit f uses the Philox stream of its GLOBAL index, so the result does not depend on the number of ranks).
    Returns (mean [F_local, D], cov [F_local, D, D], reverts [F_local]) as CUDA tensors."""
    dev = device()
    means_t = np.asarray(means_t, dtype=np.float64)
    covs_t = np.asarray(covs_t, dtype=np.float64)
    first = 0
    if process_group is not None:
        import torch.distributed as dist
        lo, hi = shard_range(means_t.shape[0], dist.get_rank(process_group), dist.get_world_size(process_group))
        means_t, covs_t = means_t[lo:hi], covs_t[lo:hi]
        mean = None if mean is None else mean[lo:hi]
        cov = None if cov is None else cov[lo:hi]
        z_tape = None if z_tape is None else z_tape[lo:hi]
        first = lo
    F, D = means_t.shape
    if D > 64 or batch_size > 32:
        raise ValueError("the ensemble kernel supports D <= 64 and batch_size <= 32")
    P = np.linalg.inv(covs_t)
    P = (P + np.swapaxes(P, 1, 2)) / 2
    c = np.einsum("fij,fj->fi", P, means_t)
    Pd = torch.as_tensor(P, dtype=torch.float32).to(dev).contiguous()
    cd = torch.as_tensor(c, dtype=torch.float32).to(dev).contiguous()
    mu = torch.zeros(F, D, device=dev) if mean is None else to_dev(mean, dev).clone().contiguous()
    S = torch.eye(D, device=dev).repeat(F, 1, 1).contiguous() if cov is None else to_dev(cov, dev).clone().contiguous()
    zt = None
    if z_tape is not None:
        zt = to_dev(z_tape, dev).contiguous()
        assert tuple(zt.shape) == (F, niter + 1, batch_size, D)
    rev = torch.zeros(F, dtype=torch.int32, device=dev)
    L.gsm_ensemble_fit_raw(Pd, cd, mu, S, F, D, batch_size, niter, key_to_seed(key), zt, rev, first_fit=first)
    return mu, S, rev
