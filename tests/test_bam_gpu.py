"""GPU parity tests of the BaM path through the C ABI, against the CPU oracle and the reference's golden vectors."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import gsmvi_oracle as orc
from test_gsm_gpu import record, relF


@pytest.fixture(scope="module")
def lib():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gsmvi_b200 import _lib as L
    L.lib()
    return L


def f32(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


@pytest.mark.parametrize("M,N,K,a_mn,b_mn", [(64, 64, 16, False, False), (300, 200, 100, False, False),
                                              (257, 257, 1024, True, True), (512, 130, 77, False, True),
                                              (130, 512, 200, True, False), (1024, 1024, 1024, False, False)])
def test_dgemm_matches_fp64(lib, M, N, K, a_mn, b_mn):
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g, dtype=torch.float64).cuda()
    B = torch.randn(N, K, generator=g, dtype=torch.float64).cuda()
    Cin = torch.randn(M, N, generator=g, dtype=torch.float64).cuda()
    C = torch.full((M, N), float("nan"), dtype=torch.float64, device="cuda")
    lib.dgemm(A.t().contiguous() if a_mn else A, B.t().contiguous() if b_mn else B, C, M, N, K, a_mn=a_mn, b_mn=b_mn,
              alpha=0.5, beta=-1.5, Cin=Cin, diag_add=2.0)
    ref = 0.5 * A @ B.t() - 1.5 * Cin + 2.0 * torch.eye(M, N, dtype=torch.float64, device="cuda")
    assert float((C - ref).norm() / ref.norm()) < 1e-14


@pytest.mark.parametrize("M,N,K,a_mn,b_mn", [(1024, 1024, 1024, False, True), (1100, 1024, 1000, True, True),
                                             (256, 384, 200, False, False)])
def test_dgemm_on_int8_tensor_cores_matches_fp64(lib, M, N, K, a_mn, b_mn):
    """Ozaki-split fp64 GEMM (tcgen05 kind::i8, exact int32 accumulation): 8 digits reproduce fp64 to ~1e-15 normwise."""
    g = torch.Generator().manual_seed(M + K)
    A = torch.randn(M, K, generator=g, dtype=torch.float64).cuda()
    B = torch.randn(N, K, generator=g, dtype=torch.float64).cuda()
    Aop = A.t().contiguous() if a_mn else A
    Bop = B.t().contiguous() if b_mn else B
    ref = A @ B.t()
    C = torch.full((M, N), float("nan"), dtype=torch.float64, device="cuda")
    lib.dgemm_oz(Aop, Bop, C, M, N, K, a_mn=a_mn, b_mn=b_mn, slices=8)
    assert float((C - ref).norm() / ref.norm()) < 5e-15
    Cin = torch.randn(M, N, generator=g, dtype=torch.float64).cuda()
    lib.dgemm_oz(Aop, Bop, C, M, N, K, a_mn=a_mn, b_mn=b_mn, alpha=-0.5, beta=2.0, Cin=Cin, slices=6)
    assert float((C - (-0.5 * ref + 2.0 * Cin)).norm() / ref.norm()) < 5e-11


@pytest.mark.parametrize("n", [1, 7, 64, 65, 200, 512, 1000, 2048])
def test_potrf64_matches_cholesky(lib, n):
    """The fp64 Cholesky of the BaM solve (register-blocked diagonal kernel, tile TRSM, look-ahead panels) against
    np.linalg.cholesky."""
    rng = np.random.RandomState(n)
    A = rng.normal(size=(n, n))
    S = A @ A.T / n + 0.05 * np.eye(n)
    ld = (n + 7) // 8 * 8
    buf = torch.zeros(n, ld, dtype=torch.float64, device="cuda")
    buf[:, :n] = torch.as_tensor(S)
    bad = torch.zeros(1, dtype=torch.int32, device="cuda")
    lib.potrf64(buf, n, bad)
    Lref = np.linalg.cholesky(S)
    Ld = buf[:, :n].cpu().numpy()
    assert int(bad.item()) == 0
    assert np.linalg.norm(Ld - Lref) / np.linalg.norm(Lref) < 1e-13
    assert np.abs(np.triu(Ld, 1)).max() == 0.0
    S[n // 2, n // 2] = -1.0
    buf[:, :n] = torch.as_tensor(S)
    lib.potrf64(buf, n, bad)
    assert int(bad.item()) == 1


def test_dgemm_tri_mirror_and_kranges(lib):
    n = 333
    g = torch.Generator().manual_seed(3)
    A = torch.randn(n, 200, generator=g, dtype=torch.float64).cuda()
    C = torch.zeros(n, n, dtype=torch.float64, device="cuda")
    lib.dgemm(A, A, C, n, n, 200, tri=True, mirror=True, alpha=2.0, diag_add=1.0)
    ref = 2.0 * A @ A.t() + torch.eye(n, dtype=torch.float64, device="cuda")
    assert float((C - ref).norm() / ref.norm()) < 1e-14 and torch.equal(C, C.t())
    Lm = torch.tril(torch.randn(n, n, generator=g, dtype=torch.float64)).cuda()
    U = torch.randn(n, n, generator=g, dtype=torch.float64).cuda()
    C2 = torch.empty(n, n, dtype=torch.float64, device="cuda")
    lib.dgemm(U, Lm, C2, n, n, n, b_mn=True, krange=lib.KR_B_UPPER)  # U L
    assert float((C2 - U @ Lm).norm() / (U @ Lm).norm()) < 1e-14
    C3 = torch.empty(n, n, dtype=torch.float64, device="cuda")
    lib.dgemm(Lm, C2, C3, n, n, n, a_mn=True, b_mn=True, krange=lib.KR_A_UPPER)  # L^T (U L)
    assert float((C3 - Lm.t() @ U @ Lm).norm() / (Lm.t() @ U @ Lm).norm()) < 1e-14


BAM_CASES = [(5, 2, 100.0), (16, 4, 100.0), (16, 4, 1.0), (32, 8, 10.0), (24, 40, 5.0)]


@pytest.mark.parametrize("D,B,reg", BAM_CASES)
def test_bam_update_matches_reference_golden(lib, golden, D, B, reg):
    from gsmvi_b200.bam import bam_lowrank_update, bam_update
    k = f"bam_update_D{D}_B{B}_reg{reg:g}"
    X, G, mu0, S0 = (golden[k + s] for s in ("_X", "_G", "_mu0", "_S0"))
    mu_ref, S_ref = golden[k + "_mu"], golden[k + "_S"]
    mu, S = bam_update(X, G, mu0, S0, reg)
    # oracle on the same fp32-rounded inputs the device sees
    mu_o, S_o = orc.bam_update(f32(X), f32(G), f32(mu0), f32(S0), reg)
    e_o = relF(S, S_o)
    asym = np.linalg.norm(S_ref - S_ref.T) / np.linalg.norm(S_ref)
    e_ref = relF(S, (S_ref + S_ref.T) / 2)
    record("bam_update_golden", dict(D=D, B=B, reg=reg, relF_vs_oracle=e_o, relF_vs_reference=e_ref, ref_asym=asym))
    assert e_o < 1e-4
    assert relF(mu, mu_o) < 1e-4
    # vs the reference's own (fp64, literal-formula) output: within its self-consistency and the fp32 input rounding
    assert e_ref < max(2e-4, 3 * asym)
    assert torch.equal(S, S.t())
    if B + 1 < D:
        mu_l, S_l = bam_lowrank_update(X, G, mu0, S0, reg)
        mu_lo, S_lo = orc.bam_lowrank_update(f32(X), f32(G), f32(mu0), f32(S0), reg)
        record("bam_lowrank_golden", dict(D=D, B=B, reg=reg, relF_vs_oracle=relF(S_l, S_lo), relF_vs_full=relF(S_l, S_o)))
        assert relF(S_l, S_lo) < 1e-4 and relF(mu_l, mu_lo) < 1e-4


@pytest.mark.parametrize("D,B,reg,kappa", [(256, 64, 10.0, 1e2), (512, 512, 100.0, 1e2), (1024, 256, 50.0, 1e2)])
def test_bam_update_matches_oracle_large(lib, D, B, reg, kappa):
    from gsmvi_b200.bam import bam_lowrank_update, bam_update
    rng = np.random.RandomState(D + B)
    mean_t, cov_t = orc.illcond_gaussian_target(D, kappa, 0)
    _, lp_g, _ = orc.gaussian_score_fns(mean_t, cov_t)
    mu0 = f32(rng.normal(size=D) * 0.1)
    A = rng.normal(size=(D, D))
    S0 = f32(A @ A.T / D * 0.5 + 0.5 * np.eye(D))
    S0 = (S0 + S0.T) / 2
    X = f32(mu0 + rng.normal(size=(B, D)) @ np.linalg.cholesky(S0).T)
    G = f32(lp_g(X))
    mu, S = bam_update(X, G, mu0, S0, reg)
    mu_o, S_o = orc.bam_update(X, G, mu0, S0, reg)
    xbar, gbar, U, V = orc.bam_stats(X, G, mu0, S0, reg)
    Sd = S.cpu().double().numpy()
    res = np.linalg.norm(Sd @ U @ Sd + Sd - V) / (np.linalg.norm(Sd) ** 2 * np.linalg.norm(U) + np.linalg.norm(V))
    record("bam_update_large", dict(D=D, B=B, reg=reg, kappa=kappa, relF_cov=relF(S, S_o), rel_mean=relF(mu, mu_o),
                                    qme_residual=res))
    assert relF(S, S_o) < 1e-4 and relF(mu, mu_o) < 1e-4
    assert res < 1e-6  # solves S U S + S = V
    if B + 1 < D:
        mu_l, S_l = bam_lowrank_update(X, G, mu0, S0, reg)
        record("bam_lowrank_large", dict(D=D, B=B, reg=reg, relF_vs_full_oracle=relF(S_l, S_o)))
        assert relF(S_l, S_o) < 1e-4 and relF(mu_l, mu_o) < 1e-4


def test_bam_update_full_size(lib):
    """BASELINE headline shape D = B = 4096 (configs[3]) on the ill-conditioned target, kappa = 1e2, reg = 100 (the first
    value of the example's schedule, example_bam.py:58-59): one device update vs the oracle's bam_update (a 4096 eigh on the
    host), plus the size-independent property that S solves S U S + S = V."""
    from gsmvi_b200.bam import bam_update
    D = B = 4096
    reg, kappa = 100.0, 1e2
    rng = np.random.RandomState(7)
    mean_t, cov_t = orc.illcond_gaussian_target(D, kappa, 0)
    P = np.linalg.inv(cov_t)
    mu0 = f32(rng.normal(size=D) * 0.1)
    S0 = np.eye(D)
    X = f32(mu0 + rng.normal(size=(B, D)))
    G = f32(-(X - mean_t) @ P)
    mu, S = bam_update(X, G, mu0, S0, reg)
    mu_o, S_o = orc.bam_update(X, G, mu0, S0, reg)
    xbar, gbar, U, V = orc.bam_stats(X, G, mu0, S0, reg)
    Sd = S.cpu().double().numpy()
    res = np.linalg.norm(Sd @ U @ Sd + Sd - V) / (np.linalg.norm(Sd) ** 2 * np.linalg.norm(U) + np.linalg.norm(V))
    record("bam_update_full_size", dict(D=D, B=B, reg=reg, kappa=kappa, relF_cov=relF(S, S_o), rel_mean=relF(mu, mu_o),
                                        qme_residual=res))
    assert relF(S, S_o) < 1e-4 and relF(mu, mu_o) < 1e-4
    assert res < 1e-6
    assert torch.equal(S, S.t())


FIT_CASES = [
    # D, B, niter, target, kappa, lowrank
    (16, 4, 60, "illcond", 1e1, False),
    (16, 4, 60, "illcond", 1e1, True),
    (64, 16, 30, "illcond", 1e1, False),
    (256, 64, 20, "illcond", 1e2, True),
    (1024, 256, 8, "illcond", 1e2, False),   # BASELINE config 3 shape
    (1024, 256, 8, "illcond", 1e2, True),
]


@pytest.mark.parametrize("D,B,niter,kind,kappa,lowrank", FIT_CASES)
def test_bam_fit_trajectory_parity(lib, D, B, niter, kind, kappa, lowrank):
    """Identical z-tape, target and regulariser schedule (example_bam.py:58-59: 100/(1+i)) fed to the device loop and
    the fp64 oracle loop; fitted (mu, Sigma) within 1e-4 relative (Frobenius)."""
    from gsmvi_b200.bam import BaM, Regularizers
    from gsmvi_b200.targets import DenseGaussianTarget
    mean_t, cov_t = orc.illcond_gaussian_target(D, kappa, 0)
    tgt = DenseGaussianTarget(mean_t, cov_t)
    P_dev = tgt.P.cpu().double().numpy()
    c_dev = tgt.c[:D].cpu().double().numpy()
    lp_g = lambda x: -(x @ P_dev.T) + c_dev
    Z = np.random.RandomState(1).normal(size=(niter + 1, B, D)).astype(np.float32)
    o = orc.BaM(D, None, lp_g, use_lowrank=lowrank)
    upd = orc.bam_lowrank_update if lowrank else orc.bam_update
    m_o, c_o = o.fit(99, orc.Regularizers().custom(lambda i: 100 / (1 + i)), niter=niter, batch_size=B,
                     sampler=orc.CholeskyTapeSampler(Z.astype(np.float64)), update=upd)
    b = BaM(D, tgt.lp, tgt.lp_g, use_lowrank=lowrank)
    m_d, c_d = b.fit(99, Regularizers().custom(lambda i: 100 / (1 + i)), niter=niter, batch_size=B, z_tape=Z,
                     verbose=False)
    e_c = relF(c_d, c_o)
    e_m = np.linalg.norm(m_d.cpu().double().numpy() - m_o) / np.linalg.norm(m_o)
    record("bam_fit_parity", dict(D=D, B=B, niter=niter, kappa=kappa, lowrank=lowrank, relF_cov=e_c, rel_mean=e_m,
                                  reverts_dev=b.n_reverts, reverts_oracle=o.n_reverts,
                                  ns_iters_max=max(b.ns_iters), ns_iters_mean=float(np.mean(b.ns_iters))))
    assert b.n_reverts == o.n_reverts
    assert e_c < 1e-4 and e_m < 1e-4
    assert torch.equal(c_d, c_d.t())


def test_regularizers_match_reference(golden):
    from gsmvi_b200.bam import Regularizers
    r = Regularizers()
    f = r.linear(100.0)
    assert np.array_equal(np.array([f(0) for _ in range(6)]), golden["reg_linear_100"])
    r = Regularizers()
    f = r.custom(lambda i: 100 / (1 + i))
    assert np.array_equal(np.array([f(123) for _ in range(6)]), golden["reg_custom"])
