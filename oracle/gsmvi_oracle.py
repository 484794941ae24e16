"""CPU oracle for the GSM / BaM hot path of modichirag/GSM-VI.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (gsm-vi_b200/) imports this module; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do, and only as the checker or the timed
CPU baseline.  It is a plain numpy/scipy restatement of the reference's algorithm, function by function, with the
reference file:line each one follows (paths relative to /root/reference).

Pinning: the reference ships no tests or golden vectors.  The pins are created from the reference
itself: tests/golden/make_golden.py imports /root/reference/gsmvi/gsm_numpy.py directly, and gsmvi/gsm.py, gsmvi/bam.py
and gsmvi/monitors.py through a numpy-backed stand-in for the `jax` / `numpyro` modules they import (JAX is not
installable here), and stores the reference's outputs as fixtures under tests/golden/.  tests/test_oracle_golden.py
checks every function below against those fixtures.

All functions take a `dtype` (float64 default; float32 gives the "fp32 restatement" used to interpret tolerances).
"""
import numpy as np
import scipy.linalg
import scipy.sparse.linalg

# --------------------------------------------------------------------------------------------------------------
# GSM update
# --------------------------------------------------------------------------------------------------------------


def gsm_update_single(sample, v, mu0, S0):
    """One-sample GSM update; follows gsmvi/gsm.py:8-28 (numpy twin gsmvi/gsm_numpy.py:4-24) line by line."""
    S0v = S0 @ v  # gsm.py:11
    vSv = v @ S0v  # gsm.py:12
    mu_v = (mu0 - sample) @ v  # gsm.py:13
    rho = 0.5 * np.sqrt(1 + 4 * (vSv + mu_v**2)) - 0.5  # gsm.py:14
    eps0 = S0v - mu0 + sample  # gsm.py:15
    mu_vT = np.outer(mu0 - sample, v)  # gsm.py:18
    den = 1 + rho + mu_v  # gsm.py:19
    I = np.eye(sample.shape[0], dtype=S0.dtype)
    mu_update = 1 / (1 + rho) * ((I - mu_vT / den) @ eps0)  # gsm.py:21
    mu = mu0 + mu_update  # gsm.py:22
    S_update = np.outer(mu0 - sample, mu0 - sample) - np.outer(mu - sample, mu - sample)  # gsm.py:25-27
    return mu_update, S_update


def gsm_update_literal(samples, vs, mu0, S0, dtype=np.float64):
    """Batch GSM update as the reference computes it: per-sample updates, then the batch mean (gsm.py:31-58,
    gsm_numpy.py:27-55).  O(B D^2) memory; small cases only."""
    samples, vs, mu0, S0 = (np.asarray(a, dtype=dtype) for a in (samples, vs, mu0, S0))
    assert samples.ndim == 2 and vs.ndim == 2  # gsm.py:48-49
    B, D = samples.shape
    mu_up = np.zeros((B, D), dtype=dtype)
    S_up = np.zeros((B, D, D), dtype=dtype)
    for i in range(B):  # gsm_numpy.py:48-49 (jax.vmap at gsm.py:51-52)
        mu_up[i], S_up[i] = gsm_update_single(samples[i], vs[i], mu0, S0)
    mu = mu0 + np.mean(mu_up, axis=0)  # gsm.py:53,55
    S = S0 + np.mean(S_up, axis=0)  # gsm.py:54,56
    return mu, S


def gsm_update(samples, vs, mu0, S0, dtype=np.float64):
    """GEMM restatement of gsm.py:8-58 (what the device kernels compute).
    With d = mu0 - x, w = S0 g:  vSv = <w,g>, mu_v = <d,g>, rho as gsm.py:14, alpha = 1/(1+rho),
    g^T eps0 = vSv - mu_v  =>  mu_update u = alpha*w + beta*d,  beta = -alpha*(1 + (vSv - mu_v)/(1 + rho + mu_v)),
    e = mu - x = d + u,  S = S0 + (D^T D - E^T E)/B,  mu = mu0 + mean(u)."""
    X, G, mu0, S0 = (np.asarray(a, dtype=dtype) for a in (samples, vs, mu0, S0))
    B = X.shape[0]
    Dm = mu0[None, :] - X
    W = G @ S0.T
    vSv = np.sum(W * G, axis=1)
    mu_v = np.sum(Dm * G, axis=1)
    rho = 0.5 * np.sqrt(1 + 4 * (vSv + mu_v**2)) - 0.5
    alpha = 1 / (1 + rho)
    beta = -alpha * (1 + (vSv - mu_v) / (1 + rho + mu_v))
    U = alpha[:, None] * W + beta[:, None] * Dm
    E = Dm + U
    mu = mu0 + U.mean(axis=0)
    S = S0 + (Dm.T @ Dm - E.T @ E) / dtype(B)
    return mu, S


def check_goodness(cov):
    """gsm.py:136-150 / bam.py:219-233: host Cholesky; exception or NaN -> False.  (The reference's NaN branch
    raises NameError, swallowed by its bare except, so it also returns False.)"""
    try:
        if np.isnan(np.linalg.cholesky(cov)).any():
            return False
        return True
    except Exception:
        return False


# --------------------------------------------------------------------------------------------------------------
# Samplers (injectable, so identical draws can be fed to the device path)
# --------------------------------------------------------------------------------------------------------------


def numpy_mvn_sampler(mean, cov, batch_size, i):
    """The reference's sampler: np.random.multivariate_normal on the global RNG (gsm.py:119, gsm_numpy.py:119);
    internally an SVD factor of cov."""
    return np.random.multivariate_normal(mean=mean, cov=cov, size=batch_size)


class CholeskyTapeSampler:
    """x = mean + z L^T with L = chol(cov) and z taken from a fixed tape Z[niter+1, B, D] - the factor the device
    sampler uses (SURVEY.md section 8c parity protocol).  Same distribution as numpy_mvn_sampler, different x."""

    def __init__(self, Z):
        self.Z = np.asarray(Z)

    def __call__(self, mean, cov, batch_size, i):
        z = self.Z[i].astype(cov.dtype)
        assert z.shape[0] == batch_size
        L = np.linalg.cholesky(cov)
        return mean[None, :] + z @ L.T


# --------------------------------------------------------------------------------------------------------------
# GSM fit loop
# --------------------------------------------------------------------------------------------------------------


class GSM:
    """Follows gsmvi/gsm.py:62-133 (numpy twin gsm_numpy.py:60-129) with an injectable sampler and update."""

    def __init__(self, D, lp, lp_g):
        self.D, self.lp, self.lp_g = D, lp, lp_g

    def fit(self, key, mean=None, cov=None, batch_size=2, niter=5000, nprint=10, verbose=False, check_goodness_=True,
            monitor=None, sampler=None, update=gsm_update, dtype=np.float64, seed_numpy=True, trace=None):
        if mean is None:
            mean = np.zeros(self.D, dtype=dtype)  # gsm.py:100-101
        if cov is None:
            cov = np.identity(self.D, dtype=dtype)  # gsm.py:102-103
        mean, cov = np.asarray(mean, dtype=dtype), np.asarray(cov, dtype=dtype)
        if sampler is None:
            sampler = numpy_mvn_sampler
            if seed_numpy:
                np.random.seed(key)  # gsm_numpy.py:105 (seeded once)
        nevals = 1  # gsm.py:105
        self.n_reverts = 0
        for i in range(niter + 1):  # gsm.py:107
            if verbose and (i % (niter // nprint) == 0):  # gsm.py:108
                print(f"Iteration {i} of {niter}")
            if monitor is not None and (i % monitor.checkpoint) == 0:  # gsm.py:111-114
                monitor(i, [mean, cov], self.lp, key, nevals=nevals)
                nevals = 0
            samples = np.asarray(sampler(mean, cov, batch_size, i), dtype=dtype)  # gsm.py:119
            vs = np.asarray(self.lp_g(samples), dtype=dtype)  # gsm.py:121
            mean_new, cov_new = update(samples, vs, mean, cov, dtype=dtype)  # gsm.py:122
            nevals += batch_size  # gsm.py:123
            if check_goodness(cov_new):  # gsm.py:125-129
                mean, cov = mean_new, cov_new
            else:
                self.n_reverts += 1
                if verbose:
                    print("Bad update for covariance matrix. Revert")
            if trace is not None:
                trace.append((mean.copy(), cov.copy()))
        if monitor is not None:  # gsm.py:131-132
            monitor(i, [mean, cov], self.lp, key, nevals=nevals)
        return mean, cov


# --------------------------------------------------------------------------------------------------------------
# BaM update
# --------------------------------------------------------------------------------------------------------------


def bam_stats(samples, vs, mu0, S0, reg, dtype=np.float64, literal=False):
    """Batch statistics and the U, V matrices of bam.py:49-60 (same lines 90-101 in the low-rank variant).
    literal=True forms C and Gamma as the mean of stacked per-sample outer products, in the reference's operation
    order (the literal solve amplifies last-bit differences by ~1e10, so bit-faithful statistics matter there);
    otherwise as one GEMM each (what the device does)."""
    X, G, mu0, S0 = (np.asarray(a, dtype=dtype) for a in (samples, vs, mu0, S0))
    reg = dtype(reg)
    xbar = np.mean(X, axis=0)  # bam.py:50
    xdiff = X - xbar  # bam.py:52
    gbar = np.mean(G, axis=0)  # bam.py:55
    gdiff = G - gbar  # bam.py:56
    if literal:
        C = np.mean(np.stack([np.outer(a, a) for a in xdiff]), axis=0)  # bam.py:51,53
        Gm = np.mean(np.stack([np.outer(a, a) for a in gdiff]), axis=0)  # bam.py:57
    else:
        C = xdiff.T @ xdiff / dtype(X.shape[0])
        Gm = gdiff.T @ gdiff / dtype(X.shape[0])
    U = reg * Gm + reg / (1 + reg) * np.outer(gbar, gbar)  # bam.py:59
    V = S0 + reg * C + reg / (1 + reg) * np.outer(mu0 - xbar, mu0 - xbar)  # bam.py:60
    return xbar, gbar, U, V


def get_sqrt(M):
    """bam.py:19-28, GPU branch: scipy.linalg.sqrtm on the host, real part."""
    return np.real(scipy.linalg.sqrtm(M)).astype(M.dtype)


def bam_update_literal(samples, vs, mu0, S0, reg, dtype=np.float64):
    """bam.py:31-69 as written: S = 2 solve(I + sqrtm(I + 4UV)^T, V^T)."""
    mu0 = np.asarray(mu0, dtype=dtype)
    xbar, gbar, U, V = bam_stats(samples, vs, mu0, S0, reg, dtype, literal=True)
    reg = dtype(reg)
    I = np.identity(U.shape[0], dtype=dtype)  # bam.py:61
    mat = I + 4 * (U @ V)  # bam.py:63
    S = 2 * np.linalg.solve(I + get_sqrt(mat).T, V.T)  # bam.py:65
    mu = 1 / (1 + reg) * mu0 + reg / (1 + reg) * (S @ gbar + xbar)  # bam.py:67
    return mu, S


def bam_update(samples, vs, mu0, S0, reg, dtype=np.float64):
    """Symmetrised restatement of bam.py:59-67 (what the device solve computes; SURVEY.md section 7 hard part 1):
    V = L L^T, M = I + 4 L^T U L (SPD, eigenvalues >= 1), S = 2 L (I + M^{1/2})^{-1} L^T.  Algebraically identical to
    bam_update_literal (solves S U S + S = V) and symmetric PSD by construction."""
    mu0 = np.asarray(mu0, dtype=dtype)
    xbar, gbar, U, V = bam_stats(samples, vs, mu0, S0, reg, dtype)
    reg = dtype(reg)
    D = U.shape[0]
    L = np.linalg.cholesky((V + V.T) / 2)
    M = np.identity(D, dtype=dtype) + 4 * (L.T @ U @ L)
    M = (M + M.T) / 2
    w, Q = np.linalg.eigh(M)
    N = (Q * np.sqrt(np.maximum(w, 0))) @ Q.T
    R = np.linalg.cholesky(np.identity(D, dtype=dtype) + (N + N.T) / 2)
    T = scipy.linalg.solve_triangular(R, L.T, lower=True).T  # T = L R^{-T}
    S = 2 * (T @ T.T)  # = 2 L (I + N)^{-1} L^T, symmetric by construction
    mu = 1 / (1 + reg) * mu0 + reg / (1 + reg) * (S @ gbar + xbar)
    return mu, S


def compute_Q_host(U, B):
    """bam.py:10-13: truncated SVD factor, Q Q^T ~= U."""
    UU, DD, VV = scipy.sparse.linalg.svds(U, k=B)
    return UU * np.sqrt(DD)


def bam_lowrank_update_literal(samples, vs, mu0, S0, reg, dtype=np.float64):
    """bam.py:72-114 as written (svds factor of U, B x B sqrtm).  Requires B < D."""
    mu0 = np.asarray(mu0, dtype=dtype)
    xbar, gbar, U, V = bam_stats(samples, vs, mu0, S0, reg, dtype, literal=True)
    reg = dtype(reg)
    B = np.asarray(samples).shape[0]
    Q = compute_Q_host(U, B).astype(dtype)  # bam.py:104
    I = np.identity(B, dtype=dtype)
    VT = V.T
    A = VT @ Q  # bam.py:107
    BB = 0.5 * I + np.real(get_sqrt(A.T @ Q + 0.25 * I))  # bam.py:108
    BB = BB @ BB  # bam.py:109
    CC = np.linalg.solve(BB, A.T)  # bam.py:110
    S = VT - A @ CC  # bam.py:111
    mu = 1 / (1 + reg) * mu0 + reg / (1 + reg) * (S @ gbar + xbar)  # bam.py:112
    return mu, S


def bam_lowrank_update(samples, vs, mu0, S0, reg, dtype=np.float64):
    """Low-rank update without the SVD: Q = [sqrt(reg/B) (G - gbar)^T, sqrt(reg/(1+reg)) gbar] has Q Q^T = U exactly
    (bam.py:59) and S is invariant to the choice of factor, so bam.py:105-111 applies with K = B+1 columns."""
    X, G, mu0, S0 = (np.asarray(a, dtype=dtype) for a in (samples, vs, mu0, S0))
    reg = dtype(reg)
    xbar, gbar, U, V = bam_stats(X, G, mu0, S0, reg, dtype)
    B = X.shape[0]
    Q = np.concatenate([np.sqrt(reg / B) * (G - gbar).T, np.sqrt(reg / (1 + reg)) * gbar[:, None]], axis=1)
    K = Q.shape[1]
    I = np.identity(K, dtype=dtype)
    A = V.T @ Q
    T = A.T @ Q + 0.25 * I
    T = (T + T.T) / 2
    w, Z = np.linalg.eigh(T)
    R = (Z * np.sqrt(np.maximum(w, 0))) @ Z.T
    BB = 0.5 * I + R
    BB = BB @ BB
    S = V.T - A @ np.linalg.solve(BB, A.T)
    mu = 1 / (1 + reg) * mu0 + reg / (1 + reg) * (S @ gbar + xbar)
    return mu, S


class Regularizers:
    """bam.py:237-274.  The `iteration` argument is ignored; an internal counter advances on every call."""

    def __init__(self):
        self.counter = 0

    def reset(self):
        self.counter = 0

    def constant(self, reg0):
        def reg_iter(iteration):
            self.counter += 1
            return reg0
        return reg_iter

    def linear(self, reg0):
        def reg_iter(iteration):
            self.counter += 1
            return reg0 / self.counter
        return reg_iter

    def custom(self, func):
        def reg_iter(iteration):
            self.counter += 1
            return func(self.counter)
        return reg_iter


class BaM:
    """Follows gsmvi/bam.py:117-216 with an injectable sampler and update function."""

    def __init__(self, D, lp, lp_g, use_lowrank=False, jit_compile=True):
        self.D, self.lp, self.lp_g = D, lp, lp_g
        self.use_lowrank = use_lowrank
        self.jit_compile = jit_compile

    def fit(self, key, regf, mean=None, cov=None, batch_size=2, niter=5000, nprint=10, verbose=False,
            check_goodness_=True, monitor=None, retries=10, jitter=1e-6, sampler=None, update=None,
            dtype=np.float64, trace=None):
        if mean is None:
            mean = np.zeros(self.D, dtype=dtype)  # bam.py:163-164
        if cov is None:
            cov = np.identity(self.D, dtype=dtype)  # bam.py:165-166
        mean, cov = np.asarray(mean, dtype=dtype), np.asarray(cov, dtype=dtype)
        if update is None:
            update = bam_lowrank_update_literal if self.use_lowrank else bam_update_literal  # bam.py:170-173
        if sampler is None:
            sampler = numpy_mvn_sampler
        nevals = 1  # bam.py:168
        self.n_reverts = 0
        if nprint > niter:
            nprint = niter  # bam.py:177
        for i in range(niter + 1):  # bam.py:178
            if verbose and (i % (niter // nprint) == 0):
                print(f"Iteration {i} of {niter}")
            if monitor is not None and (i % monitor.checkpoint) == 0:  # bam.py:182-185
                monitor(i, [mean, cov], self.lp, key, nevals=nevals)
                nevals = 0
            j = 0
            while True:  # bam.py:188-206
                try:
                    samples = np.asarray(sampler(mean, cov, batch_size, i), dtype=dtype)  # bam.py:193
                    vs = np.asarray(self.lp_g(samples), dtype=dtype)  # bam.py:194
                    nevals += batch_size  # bam.py:195
                    reg = regf(i)  # bam.py:196
                    mean_new, cov_new = update(samples, vs, mean, cov, reg, dtype=dtype)  # bam.py:197
                    cov_new = cov_new + np.eye(self.D, dtype=dtype) * dtype(jitter)  # bam.py:198
                    cov_new = (cov_new + cov_new.T) / 2.0  # bam.py:199
                    break
                except Exception as e:  # bam.py:201-206
                    if j < retries:
                        j += 1
                    else:
                        raise e
            if check_goodness(cov_new):  # bam.py:208-212
                mean, cov = mean_new, cov_new
            else:
                self.n_reverts += 1
                if verbose:
                    print("Bad update for covariance matrix. Revert")
            if trace is not None:
                trace.append((mean.copy(), cov.copy()))
        if monitor is not None:  # bam.py:214-215
            monitor(i, [mean, cov], self.lp, key, nevals=nevals)
        return mean, cov


# --------------------------------------------------------------------------------------------------------------
# Monitor
# --------------------------------------------------------------------------------------------------------------


def gaussian_logprob(x, mu, cov):
    """log N(x | mu, cov) per row: what numpyro.distributions.MultivariateNormal.log_prob returns
    (monitors.py:107): -1/2 maha - sum log diag(L) - D/2 log 2 pi, L = chol(cov)."""
    x = np.atleast_2d(x)
    D = x.shape[1]
    L = np.linalg.cholesky(cov)
    y = scipy.linalg.solve_triangular(L, (x - mu).T, lower=True)
    return -0.5 * np.sum(y * y, axis=0) - np.sum(np.log(np.diag(L))) - 0.5 * D * np.log(2 * np.pi)


def reverse_kl(samples, lpq, lpp):
    """monitors.py:10-15."""
    logl = np.sum(lpp(samples))
    logq = np.sum(lpq(samples))
    return (logq - logl) / samples.shape[0]


def forward_kl(samples, lpq, lpp):
    """monitors.py:17-22."""
    logl = np.sum(lpp(samples))
    logq = np.sum(lpq(samples))
    return (logl - logq) / samples.shape[0]


class KLMonitor:
    """monitors.py:43-125 with an injectable sampler (default: the reference's seeded numpy sampler).  `key` handling:
    the reference splits a JAX key and seeds numpy with its first word (monitors.py:101-102); here `key_to_seed`
    maps the key to that seed so a stand-in split can be injected by the golden tests."""

    def __init__(self, batch_size_kl=8, checkpoint=20, offset_evals=0, ref_samples=None, sampler=None,
                 key_to_seed=None):
        self.batch_size_kl, self.checkpoint = batch_size_kl, checkpoint
        self.offset_evals, self.ref_samples = offset_evals, ref_samples
        self.sampler, self.key_to_seed = sampler, key_to_seed
        self.rkl, self.fkl, self.nevals = [], [], []

    def reset(self, batch_size_kl=None, checkpoint=None, offset_evals=None, ref_samples=None):  # monitors.py:69-81
        self.nevals, self.rkl, self.fkl = [], [], []
        if batch_size_kl is not None:
            self.batch_size_kl = batch_size_kl
        if checkpoint is not None:
            self.checkpoint = checkpoint
        if offset_evals is not None:
            self.offset_evals = offset_evals
        if ref_samples is not None:
            self.ref_samples = ref_samples

    def __call__(self, i, params, lp, key, nevals=1):
        mu, cov = params
        if self.key_to_seed is not None:
            np.random.seed(self.key_to_seed(key))  # monitors.py:101-102
        try:
            if self.sampler is not None:
                qsamples = self.sampler(mu, cov, self.batch_size_kl, i)
            else:
                qsamples = np.random.multivariate_normal(mean=mu, cov=cov, size=self.batch_size_kl)  # monitors.py:106
            lpq = lambda x: gaussian_logprob(x, mu, cov)  # monitors.py:107
            self.rkl.append(reverse_kl(qsamples, lpq, lp))  # monitors.py:108
            if self.ref_samples is not None:  # monitors.py:110-113
                idx = np.random.permutation(self.ref_samples.shape[0])[: self.batch_size_kl]
                self.fkl.append(forward_kl(self.ref_samples[idx], lpq, lp))
            else:
                self.fkl.append(np.nan)  # monitors.py:115
        except Exception:  # monitors.py:117-120
            self.rkl.append(np.nan)
            self.fkl.append(np.nan)
        self.nevals.append(self.offset_evals + nevals)  # monitors.py:122
        self.offset_evals = self.nevals[-1]  # monitors.py:123
        return key


# --------------------------------------------------------------------------------------------------------------
# ADVI (gsmvi/advi.py) - SURVEY.md section 8f-4.  PARITY UNPINNED against the reference itself: advi.py needs jax autodiff
# (jax.value_and_grad) and optax, neither importable here, and a numpy stand-in cannot differentiate.  The restatement
# below follows advi.py line by line with the gradient written out in closed form; tests/test_host_cpu.py checks that
# closed form against central finite differences of neg_elbo (the function the reference differentiates).
# --------------------------------------------------------------------------------------------------------------


def advi_neg_elbo(mu, Lm, Z, lp):
    """neg_elbo of advi.py:31-45 for fixed standard-normal draws Z [B, D]: x = mu + z L^T (numpyro's reparameterised
    sample), loss = -(sum_b lp(x_b) - sum_b log q(x_b)) with q = N(mu, L L^T), L = lower triangle of Lm."""
    Lt = np.tril(Lm)
    X = mu[None, :] + Z @ Lt.T
    D = mu.shape[0]
    logq = -0.5 * np.sum(Z * Z) - Z.shape[0] * (np.sum(np.log(np.abs(np.diag(Lt)))) + 0.5 * D * np.log(2 * np.pi))
    return -(lp(X) - logq), X


def advi_grad(mu, Lm, Z, G):
    """Gradient of advi_neg_elbo with respect to (mu, lower triangle of L) given the scores G = grad log p at the samples."""
    B = Z.shape[0]
    gL = -np.tril(G.T @ Z) - B * np.diag(1.0 / np.diag(Lm))
    return -G.sum(axis=0), gL


class ADVI:
    """Follows gsmvi/advi.py:8-112 with a draw tape and optax.adam written out (b1 = 0.9, b2 = 0.999, eps = 1e-8)."""

    def __init__(self, D, lp, lp_g):
        self.D, self.lp, self.lp_g = D, lp, lp_g

    def fit(self, key, lr, Z, mean=None, cov=None, batch_size=8, niter=1000, monitor=None, b1=0.9, b2=0.999, eps=1e-8,
            dtype=np.float64):
        D = self.D
        mean = np.zeros(D, dtype) if mean is None else np.asarray(mean, dtype)  # advi.py:77-78
        cov = np.identity(D, dtype=dtype) if cov is None else np.asarray(cov, dtype)  # advi.py:79-80
        Lm = np.linalg.cholesky(cov)  # advi.py:83
        mu = mean.copy()
        m_mu, v_mu, m_L, v_L = np.zeros(D, dtype), np.zeros(D, dtype), np.zeros((D, D), dtype), np.zeros((D, D), dtype)
        losses = []
        nevals = 1
        for i in range(niter + 1):  # advi.py:92
            if monitor is not None and (i % monitor.checkpoint) == 0:  # advi.py:95-100
                monitor(i, [mu, np.tril(Lm) @ np.tril(Lm).T], self.lp, key, nevals=nevals)
                nevals = 0
            z = np.asarray(Z[i], dtype)
            loss, X = advi_neg_elbo(mu, Lm, z, self.lp)
            g_mu, g_L = advi_grad(mu, Lm, z, np.asarray(self.lp_g(X), dtype))
            t = i + 1
            m_mu = b1 * m_mu + (1 - b1) * g_mu
            v_mu = b2 * v_mu + (1 - b2) * g_mu**2
            m_L = b1 * m_L + (1 - b1) * g_L
            v_L = b2 * v_L + (1 - b2) * g_L**2
            mu = mu - lr * (m_mu / (1 - b1**t)) / (np.sqrt(v_mu / (1 - b2**t)) + eps)
            Lm = Lm - np.tril(lr * (m_L / (1 - b1**t)) / (np.sqrt(v_L / (1 - b2**t)) + eps))
            losses.append(loss)
            nevals += batch_size
        cov = np.tril(Lm) @ np.tril(Lm).T  # advi.py:109
        if monitor is not None:
            monitor(i, [mu, cov], self.lp, key, nevals=nevals)
        return mu, cov, losses


# --------------------------------------------------------------------------------------------------------------
# Benchmark targets (examples/example_gsm_numpy.py:8-31 made reproducible)
# --------------------------------------------------------------------------------------------------------------


def example_target(D, seed):
    """The example's generator verbatim (example_gsm_numpy.py:11-14) on a seeded RandomState:
    mean ~ U(0,1)^D, L ~ N(0,1)^{DxD}, cov = L L^T + 1e-3 I."""
    rng = np.random.RandomState(seed)
    mean = rng.random_sample(D)
    L = rng.normal(size=D**2).reshape(D, D)
    cov = L @ L.T + np.eye(D) * 1e-3
    return mean, cov


def dense_gaussian_target(D, seed=0):
    """Benchmark family: as example_target but cov = A A^T / D + 1e-3 I so the spectrum is O(1) at any D."""
    rng = np.random.RandomState(seed)
    mean = rng.random_sample(D)
    A = rng.normal(size=(D, D))
    cov = A @ A.T / D + np.eye(D) * 1e-3
    return mean, cov


def illcond_gaussian_target(D, kappa=1e2, seed=0):
    """cov = Q diag(logspace(0, -log10 kappa, D)) Q^T with Q from the QR of a seeded Gaussian."""
    rng = np.random.RandomState(seed)
    mean = rng.random_sample(D)
    Q, _ = np.linalg.qr(rng.normal(size=(D, D)))
    ev = np.logspace(0, -np.log10(kappa), D)
    cov = (Q * ev) @ Q.T
    cov = (cov + cov.T) / 2
    return mean, cov


def gaussian_score_fns(mean, cov):
    """lp (sum over the batch, example_gsm.py:34 convention; constant dropped as example_gsm_numpy.py:17-22) and
    lp_g = -icov (x - mean) (example_gsm_numpy.py:24-29), vectorised."""
    icov = np.linalg.inv(cov)

    def lp(x):
        d = np.atleast_2d(x) - mean
        return -0.5 * np.sum((d @ icov) * d)

    def lp_g(x):
        return -(np.atleast_2d(x) - mean) @ icov.T

    return lp, lp_g, icov
