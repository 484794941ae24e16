import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gsm-vi_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import gsmvi_oracle as orc
from gsmvi_b200.bam import bam_update
g = np.load(os.path.join(ROOT, "tests/golden/reference_golden.npz"))
for (D, B, reg) in [(5, 2, 100.0), (16, 4, 1.0), (32, 8, 10.0)]:
    k = f"bam_update_D{D}_B{B}_reg{reg:g}"
    X, G, mu0, S0 = (g[k + s] for s in ("_X", "_G", "_mu0", "_S0"))
    xbar, gbar, U, V = orc.bam_stats(X, G, mu0, S0, reg)
    Lc = np.linalg.cholesky(V); M = np.eye(D) + 4 * Lc.T @ U @ Lc
    w, Q = np.linalg.eigh(M); N = (Q * np.sqrt(w)) @ Q.T
    print("oracle: |U|=%.6e |V|=%.6e |L|=%.6e |UL|=%.6e |M|=%.6e |N|=%.6e |R|=%.6e" % tuple(np.linalg.norm(a) for a in (U, V, Lc, U @ Lc, M, N, np.linalg.cholesky(np.eye(D) + N))))
    try:
        mu, S = bam_update(X, G, mu0, S0, reg)
        mo, So = orc.bam_update(X, G, mu0, S0, reg)
        print("relF", np.linalg.norm(S.cpu().double().numpy() - So) / np.linalg.norm(So))
    except Exception as e:
        print("EXC", e)
