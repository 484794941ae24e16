"""Benchmark targets with a built-in device score.

The reference's examples build a dense Gaussian target and hand GSM/BaM a Python `lp_g` (examples/
example_gsm_numpy.py:8-31).  `DenseGaussianTarget` is that target with the score evaluated by the library's GEMM
(G = -(X - m) P); its `lp_g` / `lp` are also ordinary callables on CUDA tensors, so it can be passed wherever a user
callable is expected."""
import numpy as np
import torch

from . import _lib as L
from ._util import device, new_mat, new_vec


def dense_gaussian_target(D, seed=0):
    """(mean, cov) of the benchmark's dense-Gaussian family: the examples' generator (examples/example_gsm_numpy.py:11-14:
    mean ~ U(0,1)^D, A ~ N(0,1)^{DxD}) on a seeded RandomState with cov = A A^T / D + 1e-3 I, so that the spectrum is O(1)
    at any D.  numpy fp64."""
    rng = np.random.RandomState(seed)
    mean = rng.random_sample(D)
    A = rng.normal(size=(D, D))
    return mean, A @ A.T / D + np.eye(D) * 1e-3


def illcond_gaussian_target(D, kappa=1e2, seed=0):
    """(mean, cov) of the ill-conditioned family (BASELINE configs 3 / 4): cov = Q diag(logspace(0, -log10 kappa, D)) Q^T
    with Q from the QR factorisation of a seeded Gaussian matrix.  numpy fp64."""
    rng = np.random.RandomState(seed)
    mean = rng.random_sample(D)
    Q, _ = np.linalg.qr(rng.normal(size=(D, D)))
    cov = (Q * np.logspace(0, -np.log10(kappa), D)) @ Q.T
    return mean, (cov + cov.T) / 2


class DenseGaussianTarget:
    def __init__(self, mean, cov, dev=None):
        dev = dev or device()
        mean = np.asarray(mean, dtype=np.float64)
        cov = np.asarray(cov, dtype=np.float64)
        self.D = mean.shape[0]
        P = np.linalg.inv(cov)
        P = (P + P.T) / 2
        self.mean64, self.cov64, self.P64 = mean, cov, P
        self.Pb, self.P = new_mat(self.D, self.D, dev)
        self.P.copy_(torch.as_tensor(P, dtype=torch.float32))
        self.Phib, _ = new_mat(self.D, self.D, dev)  # pre-split (hi, lo) form of P for the score GEMM
        self.Plob, _ = new_mat(self.D, self.D, dev)
        L.tf32_split(self.Pb, self.Phib, self.Plob, self.D, self.D)
        self.Ph = L.HOperand(self.D, self.D, dev).split_from(self.P)  # fp16 (hi, lo) form for the scaled 3xFP16 engine
        Pf = np.asarray(P, dtype=np.float32).astype(np.float64)
        self.pnorm = float(np.abs(Pf).sum(axis=0).max())  # max_j sum_k |P_kj|: bound |x P| <= max|x| pnorm
        self.cmax = float(np.abs(P @ mean).max())
        self.c = new_vec(self.D, dev)
        self.c[: self.D].copy_(torch.as_tensor(P @ mean, dtype=torch.float32))
        self.m = torch.as_tensor(mean, dtype=torch.float32, device=dev)
        self._gsmvi_builtin_target = self

    def device_fp64(self):
        """(P [D, D], c = P m [D]) as contiguous fp64 device tensors: the operands of the fp64 small-D path (D <= 64)."""
        if getattr(self, "_P64d", None) is None:
            dev = self.P.device
            self._P64d = torch.as_tensor(self.P64, dtype=torch.float64).to(dev).contiguous()
            self._c64d = torch.as_tensor(self.P64 @ self.mean64, dtype=torch.float64).to(dev).contiguous()
        return self._P64d, self._c64d

    def lp_g(self, x):
        """Score -(x - m) P, [B, D] -> [B, D] (examples/example_gsm_numpy.py:24-29), via the device GEMM."""
        B = x.shape[0]
        Xb, X = new_mat(B, self.D, x.device)
        X.copy_(x)
        Gb, G = new_mat(B, self.D, x.device)
        L.gauss_score(Xb, self.Phib, self.c, Gb, B, self.D, P_lo=self.Plob)
        return G

    def lp(self, x):
        """Sum over the batch of the unnormalised log density (examples/example_gsm.py:34 convention)."""
        d = x.to(torch.float32) - self.m
        return -0.5 * torch.sum((d @ self.P) * d)
