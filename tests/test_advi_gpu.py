"""GPU parity tests of the ADVI baseline through the C ABI, against the CPU oracle's restatement of
gsmvi/advi.py on the same draw tape."""
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


@pytest.mark.parametrize("D,B,niter,score", [(16, 8, 200, "lp_g"), (16, 8, 200, "autograd"), (200, 64, 60, "lp_g")])
def test_advi_fit_trajectory_parity(lib, D, B, niter, score):
    """Identical z-tape, target and Adam hyper-parameters fed to the device loop and the fp64 oracle loop: (mean, cov) and
    the loss trace agree (fp32 state on the device against fp64 on the host; Adam's normalised steps make the comparison
    a trajectory one, tolerance 1e-3)."""
    from gsmvi_b200.advi import ADVI, adam
    mean_t, cov_t = orc.dense_gaussian_target(D, 4)
    lp_o, lp_g_o, icov = orc.gaussian_score_fns(mean_t, cov_t)
    Z = np.random.RandomState(2).normal(size=(niter + 1, B, D)).astype(np.float32)
    m_o, c_o, l_o = orc.ADVI(D, lp_o, lp_g_o).fit(0, 1e-2, Z.astype(np.float64), batch_size=B, niter=niter)
    P = torch.as_tensor(icov, dtype=torch.float32, device="cuda")
    m = torch.as_tensor(mean_t, dtype=torch.float32, device="cuda")
    lp = lambda x: -0.5 * torch.sum(((x - m) @ P) * (x - m))
    lp_g = (lambda x: -(x - m) @ P) if score == "lp_g" else None
    m_d, c_d, l_d = ADVI(D, lp, lp_g).fit(0, adam(1e-2), batch_size=B, niter=niter, z_tape=Z, verbose=False)
    e_c, e_m = relF(c_d, c_o), relF(m_d, m_o)
    e_l = np.max(np.abs(np.array(l_d) - np.array(l_o)) / np.maximum(np.abs(np.array(l_o)), 1.0))
    record("advi_fit_parity", dict(D=D, B=B, niter=niter, score=score, relF_cov=e_c, rel_mean=e_m, rel_loss=e_l))
    assert e_c < 1e-3 and e_m < 1e-3 and e_l < 1e-3
    assert len(l_d) == niter + 1 and l_d[-1] < l_d[0]
    assert torch.equal(c_d, c_d.t()) or relF(c_d, c_d.t().cpu().double().numpy()) < 1e-6


def test_advi_with_monitor_and_philox(lib):
    """examples/example_initializers.py:53-65: ADVI with a KLMonitor, warm-started from (mean, cov); Philox draws."""
    from gsmvi_b200.advi import ADVI
    from gsmvi_b200.monitors import KLMonitor
    from gsmvi_b200.targets import DenseGaussianTarget
    D, B = 16, 16
    mean_t, cov_t = orc.dense_gaussian_target(D, 9)
    tgt = DenseGaussianTarget(mean_t, cov_t)
    mon = KLMonitor(batch_size_kl=32, checkpoint=50)
    m, c, losses = ADVI(D, tgt.lp, tgt.lp_g).fit(3, 2e-2, mean=mean_t + 0.1, cov=np.eye(D) * 0.5, batch_size=B, niter=1500,
                                                 monitor=mon, verbose=False)
    assert len(losses) == 1501 and len(mon.rkl) == 1500 // 50 + 2
    assert np.mean(losses[-50:]) < np.mean(losses[:50])
    assert relF(c, cov_t) < 0.15 and np.max(np.abs(m.cpu().numpy() - mean_t)) < 0.1
    assert np.nanmean(mon.rkl[-5:]) < np.nanmean(mon.rkl[:2])
