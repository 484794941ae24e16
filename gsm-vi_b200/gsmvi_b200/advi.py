"""ADVI on B200 - drop-in for gsmvi/advi.py of modichirag/GSM-VI: a full-rank Gaussian fitted by maximising the ELBO with
reparameterised samples, on the same device kernels GSM and BaM use (Philox / tape draws, x = mu + z L^T, scores, the
tensor-core GEMM, the monitor reductions).

Same public surface: `ADVI(D, lp).fit(key, opt, mean, cov, batch_size, niter, nprint, monitor) -> (mean, cov, losses)`
(gsmvi/advi.py:8-112).  The reference differentiates `lp` with jax.grad; here the score comes from `lp_g` if given, else
from torch.autograd on `lp` (a torch-differentiable callable returning the SUM of log p over the batch, as in
examples/example_advi.py:35).  The gradient of the loss with respect to (mean, scale factor) is assembled from the scores in
closed form (csrc/advi.cu) - it is what jax.value_and_grad(neg_elbo) returns - and the optimiser is Adam with optax's
defaults, fused into the same launch."""
import numpy as np
import torch

from . import _lib as L
from ._util import device, key_to_seed, ld_of, new_mat, new_vec, to_dev


def adam(learning_rate, b1=0.9, b2=0.999, eps=1e-8):
    """Stand-in for optax.adam(learning_rate) (examples/example_advi.py:46): the hyper-parameters of the fused device step."""
    return {"lr": float(learning_rate), "b1": float(b1), "b2": float(b2), "eps": float(eps)}


class ADVI:
    """Class for fitting a multivariate Gaussian distribution with dense covariance matrix by maximizing ELBO
    (gsmvi/advi.py:8-21)."""

    def __init__(self, D, lp, lp_g=None):
        """D: number of parameters; lp: log-probability, SUM over the batch, [B, D] CUDA tensor -> scalar (differentiable
        with torch.autograd unless lp_g is given); lp_g: optional score function [B, D] -> [B, D]."""
        self.D = D
        self.lp = lp
        self.lp_g = lp_g

    def _score(self, X):
        if self.lp_g is not None:
            return self.lp_g(X)
        x = X.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            val = self.lp(x)
            (g,) = torch.autograd.grad(val, x)
        return g

    def fit(self, key, opt, mean=None, cov=None, batch_size=8, niter=1000, nprint=10, monitor=None, *, z_tape=None,
            verbose=True, npass=3):
        """Main function to fit a multivariate Gaussian distribution to the target (gsmvi/advi.py:47-112).

        opt: `adam(lr)` of this module, a plain learning rate, or a dict with lr / b1 / b2 / eps (optax.adam's role).
        z_tape (keyword-only): optional [niter+1, batch_size, D] standard-normal draws instead of the Philox stream.
        Returns (mean [D], cov [D, D], losses) - losses[i] is the negative ELBO of iteration i (advi.py:31-45) as a float."""
        D, B, dev = self.D, batch_size, device()
        hp = adam(opt) if isinstance(opt, (int, float)) else dict(adam(opt["lr"]), **opt) if isinstance(opt, dict) else None
        if hp is None:
            raise TypeError("opt must be a learning rate, adam(lr) or a dict with lr / b1 / b2 / eps")
        seed = key_to_seed(key)
        mu = new_vec(D, dev)
        if mean is not None:
            mu[:D].copy_(to_dev(mean, dev))  # advi.py:77-78
        Sb, S = new_mat(D, D, dev)
        S.copy_(torch.eye(D, device=dev) if cov is None else to_dev(cov, dev))  # advi.py:79-80
        # optimisation is done on the (unconstrained) Cholesky factor of the covariance (advi.py:82-85)
        Lb, Lv = new_mat(D, D, dev)
        bad = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = torch.empty(L.workspace_bytes(L.WS_POTRF, B, D) // 4, dtype=torch.float32, device=dev)
        L.potrf_check(Sb, Lb, D, bad, ws)
        if int(bad.item()) != 0:
            raise ValueError("initial covariance is not positive definite")
        Zb, Z = new_mat(B, D, dev)
        Xb, X = new_mat(B, D, dev)
        Gb, G = new_mat(B, D, dev)
        GtZ, _ = new_mat(D, D, dev)
        mL, vL = new_mat(D, D, dev)[0], new_mat(D, D, dev)[0]
        gsum, m_mu, v_mu = new_vec(D, dev), new_vec(D, dev), new_vec(D, dev)
        logq = torch.zeros(1, dtype=torch.float64, device=dev)
        if z_tape is not None:
            z_tape = torch.as_tensor(np.asarray(z_tape) if not isinstance(z_tape, torch.Tensor) else z_tape,
                                     dtype=torch.float32)
            assert tuple(z_tape.shape[1:]) == (B, D) and z_tape.shape[0] >= niter + 1
        losses = []
        nevals = 1  # advi.py:90
        every = max(niter // max(nprint, 1), 1)

        def cov_now():  # scales_to_cov (advi.py:24-29)
            Lt = torch.tril(Lv)
            return Lt @ Lt.t()

        i = 0
        for i in range(niter + 1):  # advi.py:92
            if verbose and (i % every == 0):
                print(f"Iteration {i} of {niter}")
            if monitor is not None and (i % monitor.checkpoint) == 0:  # advi.py:95-100
                monitor(i, [mu[:D], cov_now()], self.lp, key, nevals=nevals)
                nevals = 0
            # ---- neg_elbo and its gradient (advi.py:31-45, 69-70): reparameterised samples, scores, log q from |z|^2
            if z_tape is not None:
                Z.copy_(z_tape[i], non_blocking=True)
            else:
                L.philox_normal(Zb, B, D, seed, i)
            L.sample(mu, Lb, Zb, Xb, B, D, npass)  # L is lower triangular: the sampler reads k <= i only
            G.copy_(to_dev(self._score(X), dev))
            L.gauss_logq_reduce(Zb, B, D, mu, Lb, logq, from_z=True)
            logl = self.lp(X)
            logl = logl.detach().double().sum() if isinstance(logl, torch.Tensor) else torch.tensor(float(np.sum(logl)),
                                                                                                   dtype=torch.float64)
            losses.append((logq[0] - logl.to(dev)).reshape(()))  # stays on the device: no read-back inside the loop
            # ---- Adam step on (mean, scales) (advi.py:71-73)
            L.advi_step(Lb, mu, Gb, Zb, GtZ, gsum, mL, vL, m_mu, v_mu, B, D, hp["lr"], hp["b1"], hp["b2"], hp["eps"], i + 1,
                        npass)
            nevals += batch_size  # advi.py:104
        mean_fit, cov_fit = mu[:D].clone(), cov_now()  # advi.py:108-109
        if monitor is not None:  # advi.py:110-111
            monitor(i, [mean_fit, cov_fit], self.lp, key, nevals=nevals)
        losses = torch.stack(losses).cpu().tolist() if losses else []
        return mean_fit, cov_fit, losses
