"""GPU parity tests of the monitor path: the device log-density reductions behind
KLMonitor (gsmvi/monitors.py:10-22, 83-125) against the oracle's Gaussian log-probability on the same points."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import gsmvi_oracle as orc


@pytest.fixture(scope="module")
def lib():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gsmvi_b200 import _lib as L
    L.lib()
    return L


@pytest.mark.parametrize("N,D", [(8, 10), (16, 200), (64, 1000)])
def test_gauss_logq_reduce_matches_oracle(lib, N, D):
    """sum_b log N(x_b | mu, L L^T) from the draws z_b (reverse KL: |z|^2 and log det only) and from arbitrary points x_b
    (forward KL: forward substitution per point), against orc.gaussian_logprob in fp64."""
    from gsmvi_b200._util import new_mat, new_vec
    rng = np.random.RandomState(N + D)
    mean_t, cov_t = orc.dense_gaussian_target(D, 1)
    Lc = np.linalg.cholesky(cov_t)
    mu = rng.normal(size=D)
    Z = rng.normal(size=(N, D)).astype(np.float32)
    X = mu + Z.astype(np.float64) @ Lc.T
    dev = "cuda"
    Lb, Lv = new_mat(D, D, dev); Lv.copy_(torch.as_tensor(Lc, dtype=torch.float32))
    Zb, Zv = new_mat(N, D, dev); Zv.copy_(torch.as_tensor(Z))
    Xb, Xv = new_mat(N, D, dev); Xv.copy_(torch.as_tensor(X, dtype=torch.float32))
    muv = new_vec(D, dev); muv[:D].copy_(torch.as_tensor(mu, dtype=torch.float32))
    out = torch.zeros(1, dtype=torch.float64, device=dev)
    # the factor the device holds is the fp32-rounded one: evaluate the oracle with exactly that covariance
    L32 = Lv.cpu().double().numpy()
    ref = float(np.sum(orc.gaussian_logprob(Xv.cpu().double().numpy(), muv[:D].cpu().double().numpy(), L32 @ L32.T)))
    lib.gauss_logq_reduce(Zb, N, D, muv, Lb, out, from_z=True)
    from_z = float(out.item())
    lib.gauss_logq_reduce(Xb, N, D, muv, Lb, out, from_z=False)
    from_x = float(out.item())
    # x was rounded to fp32 after mu + L z, so the two differ from each other by that rounding (~1e-6 relative)
    assert abs(from_z - ref) <= 2e-5 * abs(ref)
    assert abs(from_x - ref) <= 2e-5 * abs(ref)


def test_klmonitor_reverse_and_forward_kl_on_a_gaussian_target(lib):
    """KLMonitor through the public API: for q = N(mu, Sigma) and a Gaussian target p the Monte-Carlo estimates must agree
    with the closed-form KL(q || p) and KL(p || q) within sampling error, the bookkeeping (nevals, offset_evals, NaN fkl
    without reference samples) follows gsmvi/monitors.py:115, 122-123."""
    from gsmvi_b200.monitors import KLMonitor
    from gsmvi_b200.targets import DenseGaussianTarget
    D = 12
    mean_t, cov_t = orc.dense_gaussian_target(D, 3)
    tgt = DenseGaussianTarget(mean_t, cov_t)
    rng = np.random.RandomState(0)
    A = rng.normal(size=(D, D)) / np.sqrt(D)
    mu_q, cov_q = mean_t + 0.1 * rng.normal(size=D), cov_t + 0.2 * A @ A.T
    Pt = np.linalg.inv(cov_t)
    # lp: unnormalised log density summed over the batch (examples/example_gsm.py:34); add the constant for exact KL
    logZ = -0.5 * (D * np.log(2 * np.pi) + np.linalg.slogdet(cov_t)[1])
    lp = lambda x: tgt.lp(x) + x.shape[0] * logZ
    def kl(m0, S0, m1, S1):
        S1i = np.linalg.inv(S1)
        return 0.5 * (np.trace(S1i @ S0) + (m1 - m0) @ S1i @ (m1 - m0) - D + np.linalg.slogdet(S1)[1] - np.linalg.slogdet(S0)[1])
    ref_samples = mean_t + rng.normal(size=(40000, D)) @ np.linalg.cholesky(cov_t).T
    mon = KLMonitor(batch_size_kl=20000, checkpoint=5, offset_evals=3, ref_samples=ref_samples)
    mon(0, [mu_q, cov_q], lp, key=11, nevals=7)
    mon(5, [mu_q, cov_q], lp, key=11, nevals=2)
    rkl_true, fkl_true = kl(mu_q, cov_q, mean_t, cov_t), kl(mean_t, cov_t, mu_q, cov_q)
    for est in mon.rkl:
        assert abs(est - rkl_true) < 0.05 + 0.1 * rkl_true
    for est in mon.fkl:
        assert abs(est - fkl_true) < 0.05 + 0.1 * fkl_true
    assert mon.rkl[0] != mon.rkl[1]  # fresh draws per call
    assert mon.nevals == [10, 12] and mon.offset_evals == 12
    mon2 = KLMonitor(batch_size_kl=64, checkpoint=5)
    mon2(0, [mu_q, cov_q], lp, key=1)
    assert np.isnan(mon2.fkl[0]) and np.isfinite(mon2.rkl[0])
    mon2(1, [mu_q, -cov_q], lp, key=1)  # not positive definite: swallowed into NaN (monitors.py:117-120)
    assert np.isnan(mon2.rkl[1])


@pytest.mark.parametrize("D,B", [(24, 16), (640, 256)])
def test_klmonitor_inside_a_fit_uses_the_engine_factor(lib, D, B):
    """Inside GSM.fit the monitor receives the engine's Cholesky factor with [mean, cov] (monitors.MonitorParams.chol) and
    must not factor the covariance again; the reverse KL it records equals what a stand-alone monitor computes from the same
    (mean, cov) with the same key and call count (same Philox counters: the estimates differ only by the two factorisations'
    rounding), and the nevals bookkeeping of gsmvi/gsm.py:111-114, 131-132 is unchanged."""
    from gsmvi_b200 import monitors as mon_mod
    from gsmvi_b200.gsm import GSM
    from gsmvi_b200.monitors import KLMonitor
    from gsmvi_b200.targets import DenseGaussianTarget
    mean_t, cov_t = orc.dense_gaussian_target(D, 1)
    tgt = DenseGaussianTarget(mean_t, cov_t)
    seen = []

    class Spy(KLMonitor):
        def __call__(self, i, params, lp, key, nevals=1):
            seen.append((i, getattr(params, "chol", None) is not None,
                         params[0].detach().clone(), params[1].detach().clone()))
            return super().__call__(i, params, lp, key, nevals)

    calls = {"n": 0}
    real = mon_mod.L.potrf_check

    def counting(*a, **k):
        calls["n"] += 1
        return real(*a, **k)

    mon = Spy(batch_size_kl=512, checkpoint=2)
    mon_mod.L.potrf_check = counting
    try:
        GSM(D, tgt.lp, tgt.lp_g).fit(5, niter=4, batch_size=B, verbose=False, monitor=mon)
    finally:
        mon_mod.L.potrf_check = real
    assert [s[0] for s in seen] == [0, 2, 4, 4] and len(mon.rkl) == 4 and all(np.isfinite(mon.rkl))  # gsm.py:111-114, 131-132
    if D > 64:  # (the fp64 small-D engine keeps its factor in doubles and hands over none)
        assert all(s[1] for s in seen) and calls["n"] == 0
    # stand-alone monitor on the recorded states: same key, same call index -> same draws
    ref = KLMonitor(batch_size_kl=512, checkpoint=2)
    for (i, _, m, c) in seen:
        ref(i, [m, c], tgt.lp, 5, nevals=1)
    for a, b in zip(mon.rkl, ref.rkl):
        assert abs(a - b) <= 1e-4 * max(1.0, abs(b)), (mon.rkl, ref.rkl)
