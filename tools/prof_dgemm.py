import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gsm-vi_b200"))
import torch
from gsmvi_b200 import _lib as L
n = 4096
g = torch.Generator().manual_seed(0)
A = torch.randn(n, n, generator=g, dtype=torch.float64).cuda(); B = torch.randn(n, n, generator=g, dtype=torch.float64).cuda()
C = torch.empty(n, n, dtype=torch.float64, device="cuda")
def t(fn, it=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
for (a_mn, b_mn) in [(False, False), (False, True), (True, True)]:
    ms = t(lambda: L.dgemm(A, B, C, n, n, n, a_mn=a_mn, b_mn=b_mn))
    print("dgemm a_mn=%d b_mn=%d: %.2f ms %.1f TF/s" % (a_mn, b_mn, ms, 2 * n**3 / ms / 1e9), flush=True)
ms = t(lambda: torch.matmul(A, B, out=C)); print("cublas dgemm: %.2f ms %.1f TF/s" % (ms, 2 * n**3 / ms / 1e9))
