"""Timing of the fp64 Cholesky (gsmvi_potrf64) by size: n = 64 is the diagonal-block kernel alone."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gsm-vi_b200"))
import numpy as np, torch
from gsmvi_b200 import _lib as L
for n in (64, 128, 256, 1024, 4096):
    rng = np.random.RandomState(0)
    A = rng.normal(size=(n, n)); S = A @ A.T / n + 0.05 * np.eye(n)
    ld = (n + 7) // 8 * 8
    src = torch.zeros(n, ld, dtype=torch.float64, device="cuda"); src[:, :n] = torch.as_tensor(S)
    buf = src.clone()
    bad = torch.zeros(1, dtype=torch.int32, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        buf.copy_(src); torch.cuda.synchronize()
        e0.record(); L.potrf64(buf, n, bad); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("potrf64 n=%d: %.3f ms (%.1f us per 64-column panel)" % (n, best, 1e3 * best / ((n + 63) // 64)), flush=True)
