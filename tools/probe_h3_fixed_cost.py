"""Per-tile fixed cost of gemm_h3_kernel: time of M = N = 4096 products as a function of K (T = a + b K per launch; a / waves
is what a tile pays for prologue + epilogue + CTA turnover, b K the tensor-pipe time)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gsm-vi_b200"))
import torch
from gsmvi_b200 import _lib as L
M = N = 4096
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = []
for K in (512, 1024, 2048, 4096, 8192):
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda")
    Ah = L.HOperand(M, K, "cuda").split_from(A); Bh = L.HOperand(N, K, "cuda").split_from(B)
    C = torch.empty(M, N, device="cuda")
    best = 1e9
    for _ in range(6):
        torch.cuda.synchronize(); e0.record(); L.gemm_h3(Ah, Bh, C, M, N, K); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res.append((K, best))
    print("K=%5d: %.4f ms  (%.0f TFLOP/s executed)" % (K, best, 3 * 2.0 * M * N * K / best / 1e9), flush=True)
(k1, t1), (k2, t2) = res[-2], res[-1]
b = (t2 - t1) / (k2 - k1)
a = t1 - b * k1
print("fit on the last two: T = %.4f ms + %.6f ms * K/1024 ; fixed share at K = 4096: %.1f%%" % (a, b * 1024, 100 * a / (a + b * 4096)))
