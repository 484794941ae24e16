"""Benchmark targets with a built-in device score.

The reference's examples build a dense Gaussian target and hand GSM/BaM a Python `lp_g` (examples/
example_gsm_numpy.py:8-31).  `DenseGaussianTarget` is that target with the score evaluated by the library's GEMM
(G = -(X - m) P); its `lp_g` / `lp` are also ordinary callables on CUDA tensors, so it can be passed wherever a user
callable is expected."""
import numpy as np
import torch

from . import _lib as L
from ._util import device, new_mat, new_vec


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
