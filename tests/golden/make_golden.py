"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF (run in the build container, where
/root/reference exists; the GPU box only ever sees the committed .npz files).

  python tests/golden/make_golden.py

What is executed:
  * gsmvi/gsm_numpy.py is imported and run unmodified (pure numpy).
  * gsmvi/gsm.py, gsmvi/bam.py and gsmvi/monitors.py import `jax` / `numpyro`, which are not installable here (no
    network).  They are imported unmodified with a numpy-backed stand-in registered under those module names:
    jax.numpy -> numpy, jit -> identity, vmap -> Python loop, pure_callback -> direct call,
    jax.scipy.linalg.sqrtm -> scipy.linalg.sqrtm, xla_bridge platform -> 'gpu' (the branch that calls
    scipy.linalg.sqrtm on the host, bam.py:20-22), random.split -> a deterministic numpy split,
    numpyro MultivariateNormal.log_prob -> scipy.stats.multivariate_normal.logpdf.  The reference's *code* is what
    runs; only the array library underneath is numpy (fp64) instead of XLA.  `np.NaN` (removed in numpy 2, used at
    monitors.py:115) is aliased back to np.nan.
"""
import os
import sys
import types
import warnings

import numpy as np
import scipy.linalg
import scipy.stats

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GSMVI_REFERENCE", "/root/reference")


def install_shims():
    if not hasattr(np, "NaN"):
        np.NaN = np.nan
    jax = types.ModuleType("jax")
    jnp = np  # jax.numpy
    jax.numpy = jnp

    def jit(fn=None, **kw):
        if fn is None:
            return lambda f: f
        return fn

    def vmap(fn, in_axes=0):
        def mapped(*args):
            axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
            n = next(a.shape[0] for a, ax in zip(args, axes) if ax is not None)
            outs = [fn(*[a[i] if ax is not None else a for a, ax in zip(args, axes)]) for i in range(n)]
            if isinstance(outs[0], tuple):
                return tuple(np.stack([o[k] for o in outs]) for k in range(len(outs[0])))
            return np.stack(outs)
        return mapped

    class ShapeDtypeStruct:
        def __init__(self, shape, dtype):
            self.shape, self.dtype = shape, dtype

    def pure_callback(fn, result_shape, *args):
        return fn(*args)

    jax.jit, jax.vmap, jax.ShapeDtypeStruct, jax.pure_callback = jit, vmap, ShapeDtypeStruct, pure_callback
    jax.grad = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("grad is not used on the hot path"))

    random = types.ModuleType("jax.random")

    def PRNGKey(seed):
        return np.array([0, seed], dtype=np.uint32)

    def split(key, num=2):
        ss = np.random.SeedSequence([int(k) for k in np.asarray(key).ravel()])
        return [np.array(c.generate_state(2), dtype=np.uint32) for c in ss.spawn(num)]

    random.PRNGKey, random.split = PRNGKey, split
    jax.random = random

    jsp = types.ModuleType("jax.scipy")
    jspl = types.ModuleType("jax.scipy.linalg")
    jspl.sqrtm = scipy.linalg.sqrtm
    jsp.linalg = jspl
    jax.scipy = jsp

    lib = types.ModuleType("jax.lib")
    xb = types.ModuleType("jax.lib.xla_bridge")

    class _Backend:
        platform = "gpu"

    xb.get_backend = lambda: _Backend()
    lib.xla_bridge = xb
    jax.lib = lib

    numpyro = types.ModuleType("numpyro")
    dist = types.ModuleType("numpyro.distributions")

    class MultivariateNormal:
        def __init__(self, loc=None, covariance_matrix=None):
            self.loc, self.cov = np.asarray(loc), np.asarray(covariance_matrix)

        def log_prob(self, x):
            return scipy.stats.multivariate_normal(mean=self.loc, cov=self.cov, allow_singular=False).logpdf(x)

    dist.MultivariateNormal = MultivariateNormal
    numpyro.distributions = dist

    for name, mod in {"jax": jax, "jax.numpy": jnp, "jax.random": random, "jax.scipy": jsp, "jax.scipy.linalg": jspl,
                      "jax.lib": lib, "jax.lib.xla_bridge": xb, "numpyro": numpyro,
                      "numpyro.distributions": dist}.items():
        sys.modules[name] = mod
    return split


def example_target(D, seed):
    # examples/example_gsm_numpy.py:11-14 on a seeded RandomState
    rng = np.random.RandomState(seed)
    mean = rng.random_sample(D)
    L = rng.normal(size=D**2).reshape(D, D)
    cov = L @ L.T + np.eye(D) * 1e-3
    return mean, cov


def score_fns(mean, cov):
    icov = np.linalg.inv(cov)
    # examples/example_gsm_numpy.py:17-29 (loops as in the example)
    def lp(x):
        out = 0
        for i in range(x.shape[0]):
            out += -0.5 * np.dot(np.dot(mean - x[i], icov), mean - x[i])
        return out

    def lp_g(x):
        return np.array([-1.0 * np.dot(icov, x[i] - mean) for i in range(x.shape[0])])

    return lp, lp_g


def main():
    split = install_shims()
    sys.path.insert(0, REF)
    warnings.simplefilter("ignore")
    import contextlib
    import io

    from gsmvi import gsm_numpy  # the real, unmodified numpy reference
    from gsmvi import bam as ref_bam  # unmodified, on the numpy-backed jax stand-in
    from gsmvi import gsm as ref_gsm
    from gsmvi import monitors as ref_mon

    out = {}
    # ---- 1. single-call gsm_update vectors (gsm_numpy.py:27-55 and gsm.py:31-58)
    for D, B in [(5, 2), (10, 4), (64, 16), (12, 7)]:
        rng = np.random.RandomState(1000 + D)
        X = rng.normal(size=(B, D))
        G = rng.normal(size=(B, D))
        mu0 = rng.normal(size=D)
        A = rng.normal(size=(D, D))
        S0 = A @ A.T / D + 0.1 * np.eye(D)
        mu_n, S_n = gsm_numpy.gsm_update(X, G, mu0, S0)
        mu_j, S_j = ref_gsm.gsm_update(X, G, mu0, S0)
        k = f"gsm_update_D{D}_B{B}"
        out.update({k + "_X": X, k + "_G": G, k + "_mu0": mu0, k + "_S0": S0, k + "_mu_numpy": mu_n,
                    k + "_S_numpy": S_n, k + "_mu_jaxcode": np.asarray(mu_j), k + "_S_jaxcode": np.asarray(S_j)})

    # ---- 2. gsm_numpy.GSM.fit end state: BASELINE config 1 (D=10, 500 iters, key=99) and the example's D=5
    for D, niter, B in [(10, 500, 2), (5, 500, 2), (10, 60, 8)]:
        mean, cov = example_target(D, seed=D)
        lp, lp_g = score_fns(mean, cov)
        with contextlib.redirect_stdout(io.StringIO()):
            m_fit, c_fit = gsm_numpy.GSM(D=D, lp=lp, lp_g=lp_g).fit(99, niter=niter, batch_size=B, verbose=False)
        k = f"gsm_fit_D{D}_n{niter}_B{B}"
        out.update({k + "_target_mean": mean, k + "_target_cov": cov, k + "_mean": m_fit, k + "_cov": c_fit})

    # ---- 3. single-call bam_update / bam_lowrank_update vectors (bam.py:31-69, 72-114)
    for D, B, reg in [(5, 2, 100.0), (16, 4, 100.0), (16, 4, 1.0), (32, 8, 10.0), (24, 40, 5.0)]:
        rng = np.random.RandomState(2000 + D + B)
        mean, cov = example_target(D, seed=50 + D)
        cov = cov / D
        _, lp_g = score_fns(mean, cov)
        mu0 = rng.normal(size=D) * 0.1
        A = rng.normal(size=(D, D))
        S0 = A @ A.T / D + 0.5 * np.eye(D)
        X = np.random.RandomState(7).multivariate_normal(mu0, S0, size=B)
        G = lp_g(X)
        mu_f, S_f = ref_bam.bam_update(X, G, mu0, S0, reg)
        k = f"bam_update_D{D}_B{B}_reg{reg:g}"
        out.update({k + "_X": X, k + "_G": G, k + "_mu0": mu0, k + "_S0": S0, k + "_reg": np.float64(reg),
                    k + "_mu": np.asarray(mu_f), k + "_S": np.asarray(S_f)})
        if B < D:
            mu_l, S_l = ref_bam.bam_lowrank_update(X, G, mu0, S0, reg)
            out.update({k + "_mu_lowrank": np.asarray(mu_l), k + "_S_lowrank": np.asarray(S_l)})

    # ---- 4. BaM.fit end state with the stand-in key split (reference loop bam.py:140-216 incl. jitter/symmetrise)
    for D, B, niter, lowrank in [(5, 2, 100, True), (8, 4, 40, False)]:
        mean, cov = example_target(D, seed=70 + D)
        lp, lp_g = score_fns(mean, cov)
        regf = ref_bam.Regularizers().custom(lambda i: 100 / (1 + i))  # example_bam.py:58-59
        with contextlib.redirect_stdout(io.StringIO()):
            bam = ref_bam.BaM(D=D, lp=lp, lp_g=lp_g, use_lowrank=lowrank, jit_compile=True)
            m_fit, c_fit = bam.fit(np.array([0, 99], dtype=np.uint32), regf=regf, niter=niter, batch_size=B,
                                   verbose=False)
        k = f"bam_fit_D{D}_B{B}_n{niter}_lr{int(lowrank)}"
        out.update({k + "_target_mean": mean, k + "_target_cov": cov, k + "_mean": np.asarray(m_fit),
                    k + "_cov": np.asarray(c_fit)})
    # the per-iteration numpy seeds the stand-in split produces for key [0, 99] (so the oracle can replay them)
    key = np.array([0, 99], dtype=np.uint32)
    seeds = []
    for _ in range(128):
        key, ks = split(key, 2)
        seeds.append(int(ks[0]))
    out["standin_split_seeds_key99"] = np.array(seeds, dtype=np.uint64)

    # ---- 5. Regularizers (bam.py:237-274)
    r = ref_bam.Regularizers()
    f = r.linear(100.0)
    out["reg_linear_100"] = np.array([f(0) for _ in range(6)])
    r = ref_bam.Regularizers()
    f = r.custom(lambda i: 100 / (1 + i))
    out["reg_custom"] = np.array([f(123) for _ in range(6)])

    # ---- 6. KLMonitor (monitors.py:83-125): reverse and forward KL values for a fixed (mu, cov)
    D = 6
    mean, cov = example_target(D, seed=90)
    lp, lp_g = score_fns(mean, cov)
    rng = np.random.RandomState(3)
    ref_samples = rng.multivariate_normal(mean, cov, size=64)
    mon = ref_mon.KLMonitor(batch_size_kl=16, checkpoint=5, offset_evals=3, ref_samples=ref_samples)
    mu_q = mean + 0.1
    cov_q = cov * 1.3 + 0.05 * np.eye(D)
    key = np.array([0, 5], dtype=np.uint32)
    with contextlib.redirect_stdout(io.StringIO()):
        mon(0, (mu_q, cov_q), lp, key, nevals=1)
        mon(5, (mu_q, cov), lp, key, nevals=10)
    out.update({"mon_target_mean": mean, "mon_target_cov": cov, "mon_ref_samples": ref_samples, "mon_mu_q": mu_q,
                "mon_cov_q": cov_q, "mon_rkl": np.array(mon.rkl), "mon_fkl": np.array(mon.fkl),
                "mon_nevals": np.array(mon.nevals), "mon_seed": np.uint64(int(split(key)[1][0]))})

    path = os.path.join(HERE, "reference_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "with", len(out), "arrays,", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
