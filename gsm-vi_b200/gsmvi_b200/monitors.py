"""KLMonitor on B200 - drop-in for gsmvi/monitors.py of modichirag/GSM-VI.

Same dataclass fields (batch_size_kl, checkpoint, offset_evals, ref_samples), lists (rkl, fkl, nevals), `reset` and
`__call__(i, params, lp, key, nevals=1)` (gsmvi/monitors.py:43-125).  Sampling from q, its Cholesky factor and the
Gaussian log-density sums run on the device (libgsmvi_b200.so); `lp` is the user's callable (sum of log p over the
batch, examples/example_gsm.py:34) evaluated on CUDA tensors."""
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib as L
from ._util import device, key_to_seed, new_mat, new_vec, to_dev


def _as_float(v):
    if isinstance(v, torch.Tensor):
        return float(v.detach().double().sum().item())
    return float(np.sum(np.asarray(v, dtype=np.float64)))


class MonitorParams(list):
    """[mean, cov] as the reference passes them to a monitor (gsmvi/gsm.py:113), plus `chol`: the padded fp32 buffer of the
    lower Cholesky factor of `cov` that the fit's engine already holds (None: the monitor factors `cov` itself)."""
    chol = None


@dataclass
class KLMonitor:
    """Monitor reverse (and optionally forward) KL divergence during optimisation (gsmvi/monitors.py:43-67)."""
    batch_size_kl: int = 8
    checkpoint: int = 20
    offset_evals: int = 0
    ref_samples: object = None

    def __post_init__(self):
        self.rkl = []
        self.fkl = []
        self.nevals = []
        self._calls = 0
        self._bufs = {}  # work buffers of the last (N, D, device): the monitor is called every `checkpoint` iterations

    def _work(self, N, D, dev):
        key = (N, D, str(dev))
        w = self._bufs.get(key)
        if w is None:
            self._bufs.clear()
            w = self._bufs[key] = dict(
                S=new_mat(D, D, dev), L=new_mat(D, D, dev), mu=new_vec(D, dev), Z=new_mat(N, D, dev), X=new_mat(N, D, dev),
                bad=torch.zeros(1, dtype=torch.int32, device=dev), out=torch.zeros(1, dtype=torch.float64, device=dev),
                ws=torch.empty(L.workspace_bytes(L.WS_POTRF, N, D) // 4, dtype=torch.float32, device=dev))
        return w

    def reset(self, batch_size_kl=None, checkpoint=None, offset_evals=None, ref_samples=None):
        """gsmvi/monitors.py:69-81."""
        self.nevals = []
        self.rkl = []
        self.fkl = []
        if batch_size_kl is not None:
            self.batch_size_kl = batch_size_kl
        if checkpoint is not None:
            self.checkpoint = checkpoint
        if offset_evals is not None:
            self.offset_evals = offset_evals
        if ref_samples is not None:
            self.ref_samples = ref_samples
        print("offset evals reset to : ", self.offset_evals)

    def __call__(self, i, params, lp, key, nevals=1):
        """gsmvi/monitors.py:83-125.  Appends to rkl / fkl / nevals; exceptions are swallowed into NaN as in the
        reference.  Returns the key (callers ignore it, gsm.py:113)."""
        mu_in, cov_in = params
        try:
            dev = device()
            N = self.batch_size_kl
            mu_t, cov_t = to_dev(mu_in, dev), to_dev(cov_in, dev)
            D = mu_t.shape[0]
            w = self._work(N, D, dev)
            mu = w["mu"]
            mu[:D].copy_(mu_t)
            # the fit loops of this package hand over the factor they already hold (the engine's Cholesky of exactly this
            # covariance, its goodness check): no second factorisation per checkpoint.  Anyone else passes [mean, cov].
            Lb = getattr(params, "chol", None)
            if Lb is None:
                Sb, S = w["S"]
                S.copy_(cov_t)
                Lb, bad = w["L"][0], w["bad"]
                L.potrf_check(Sb, Lb, D, bad, w["ws"])
                if int(bad.item()) != 0:
                    raise FloatingPointError("covariance is not positive definite")
            Zb = w["Z"][0]
            Xb, X = w["X"]
            # monitors.py:101-106: q-samples; counter = (iteration, call index) so draws differ from the fit's
            L.philox_normal(Zb, N, D, key_to_seed(key) ^ 0x9E3779B97F4A7C15, (int(i) << 20) + self._calls)
            L.sample(mu, Lb, Zb, Xb, N, D)
            out = w["out"]
            L.gauss_logq_reduce(Zb, N, D, mu, Lb, out, from_z=True)
            logq = float(out.item())
            logl = _as_float(lp(X))
            self.rkl.append((logq - logl) / N)  # monitors.py:10-15,108
            if self.ref_samples is not None:  # monitors.py:110-113
                ref = self.ref_samples
                n_ref = ref.shape[0]
                g = torch.Generator().manual_seed((key_to_seed(key) + 7919 * (self._calls + 1)) % (2**63))
                idx = torch.randperm(n_ref, generator=g)[:N]
                ps = to_dev(ref, dev)[idx.to(dev)]
                Pb, P = new_mat(ps.shape[0], D, dev)
                P.copy_(ps)
                L.gauss_logq_reduce(Pb, ps.shape[0], D, mu, Lb, out, from_z=False)
                logq_p = float(out.item())
                logl_p = _as_float(lp(P))
                self.fkl.append((logl_p - logq_p) / ps.shape[0])  # monitors.py:17-22
            else:
                self.fkl.append(float("nan"))  # monitors.py:115
        except Exception as e:  # monitors.py:117-120
            print(f"Exception occured in monitor : {e}.\nAppending NaN")
            self.rkl.append(float("nan"))
            self.fkl.append(float("nan"))
        self._calls += 1
        self.nevals.append(self.offset_evals + nevals)  # monitors.py:122
        self.offset_evals = self.nevals[-1]  # monitors.py:123
        return key
