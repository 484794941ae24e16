"""A/B of the two fp64 GEMM kernels (GSMVI_DGEMM=fma|mma in the environment): accuracy and throughput."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gsm-vi_b200"))
import torch
from gsmvi_b200 import _lib as L
def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters
g = torch.Generator().manual_seed(0)
for (M, N, K, a_mn, b_mn) in [(300, 200, 100, False, False), (257, 129, 77, True, True), (1024, 1024, 512, False, True)]:
    A = torch.randn(M, K, generator=g, dtype=torch.float64).cuda(); B = torch.randn(N, K, generator=g, dtype=torch.float64).cuda()
    C = torch.full((M, N), float("nan"), dtype=torch.float64, device="cuda")
    L.dgemm(A.t().contiguous() if a_mn else A, B.t().contiguous() if b_mn else B, C, M, N, K, a_mn=a_mn, b_mn=b_mn)
    print(os.environ.get("GSMVI_DGEMM", "mma"), M, N, K, a_mn, b_mn, "relerr", float((C - A @ B.t()).norm() / (A @ B.t()).norm()))
for n in (2048, 4096):
    A = torch.randn(n, n, dtype=torch.float64, device="cuda"); B = torch.randn(n, n, dtype=torch.float64, device="cuda")
    C = torch.empty(n, n, dtype=torch.float64, device="cuda")
    for (a_mn, b_mn) in [(False, False), (False, True), (True, True)]:
        ms = timeit(lambda: L.dgemm(A, B, C, n, n, n, a_mn=a_mn, b_mn=b_mn))
        print(os.environ.get("GSMVI_DGEMM", "mma"), "n=%d mn=%d%d: %.3f ms %.1f TF/s" % (n, a_mn, b_mn, ms, 2.0 * n**3 / ms / 1e9), flush=True)
