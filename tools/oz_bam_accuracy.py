"""Single BaM update at D = 2048, B = 2048 (full solve, Ozaki path) vs the fp64 oracle, per digit count (GSMVI_OZ_SLICES)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gsm-vi_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import gsmvi_oracle as orc
from gsmvi_b200.bam import bam_update
D = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
B = D
kappa = float(sys.argv[2]) if len(sys.argv) > 2 else 1e2
reg = float(sys.argv[3]) if len(sys.argv) > 3 else 100.0
rng = np.random.RandomState(0)
mean_t, cov_t = orc.illcond_gaussian_target(D, kappa, 0)
P = np.linalg.inv(cov_t)
mu0 = rng.normal(size=D) * 0.1
A = rng.normal(size=(D, D)) / np.sqrt(D)
S0 = A @ A.T + 0.5 * np.eye(D)
X = (mu0 + rng.normal(size=(B, D)) @ np.linalg.cholesky(S0).T).astype(np.float32).astype(np.float64)
G = (-(X - mean_t) @ P).astype(np.float32).astype(np.float64)
mu_o, S_o = orc.bam_update_sym(X, G, mu0.astype(np.float32).astype(np.float64), S0.astype(np.float32).astype(np.float64), reg) if hasattr(orc, "bam_update_sym") else orc.bam_update(X, G, mu0.astype(np.float32).astype(np.float64), S0.astype(np.float32).astype(np.float64), reg)
mu, S = bam_update(X, G, mu0, S0, reg)
relF = lambda a, b: float(np.linalg.norm(a.double().cpu().numpy() - b) / np.linalg.norm(b))
print(json.dumps({"D": D, "kappa": kappa, "reg": reg, "slices": os.environ.get("GSMVI_OZ_SLICES", "8"), "relF_cov": relF(S, S_o), "rel_mean": relF(mu, mu_o)}))
