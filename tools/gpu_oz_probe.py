"""GPU probe of the int8-tensor-core fp64 GEMM (oz_gemm.cu): accuracy vs torch fp64 matmul, throughput vs the FP64 pipe."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gsm-vi_b200"))
import torch
from gsmvi_b200 import _lib as L
dev = "cuda"
out = {}
def rel(C, ref): return float((C - ref).norm() / ref.norm())
def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters
g = torch.Generator().manual_seed(0)
for (M, N, K, a_mn, b_mn) in [(1024, 1024, 1024, False, False), (1024, 1152, 640, False, True), (1100, 1024, 1000, True, True),
                               (1024, 1024, 2048, True, False)]:
    A = torch.randn(M, K, generator=g, dtype=torch.float64).to(dev)
    B = torch.randn(N, K, generator=g, dtype=torch.float64).to(dev)
    Aop = A.t().contiguous() if a_mn else A
    Bop = B.t().contiguous() if b_mn else B
    ref = A @ B.t()
    for s in (8, 7, 6):
        C = torch.full((M, N), float("nan"), dtype=torch.float64, device=dev)
        L.dgemm_oz(Aop, Bop, C, M, N, K, a_mn=a_mn, b_mn=b_mn, slices=s)
        torch.cuda.synchronize()
        k = "rand_%dx%dx%d_mn%d%d_s%d" % (M, N, K, a_mn, b_mn, s)
        out[k] = rel(C, ref)
        print(k, out[k], flush=True)
    Cn = torch.empty(M, N, dtype=torch.float64, device=dev)
    L.dgemm(Aop, Bop, Cn, M, N, K, a_mn=a_mn, b_mn=b_mn)
    out["native_%dx%dx%d" % (M, N, K)] = rel(Cn, ref)
# epilogue options + wide dynamic range + Gram (same operand)
M = N = 1024; K = 1536
A = (torch.randn(M, K, generator=g, dtype=torch.float64) * torch.logspace(-8, 3, M, dtype=torch.float64)[:, None]).to(dev)
A = A * torch.logspace(0, -6, K, dtype=torch.float64, device=dev)[None, :]
Cin = torch.randn(M, N, generator=g, dtype=torch.float64).to(dev)
C = torch.empty(M, N, dtype=torch.float64, device=dev)
L.dgemm_oz(A, A, C, M, N, K, alpha=0.5, beta=-2.0, Cin=Cin, diag_add=3.0, tri=True, mirror=True)
ref = 0.5 * (A @ A.t()) - 2.0 * torch.tril(Cin) - 2.0 * torch.tril(Cin, -1).t() + 3.0 * torch.eye(M, dtype=torch.float64, device=dev)
out["gram_dynrange_epilogue"] = rel(C, ref)
out["gram_rowwise_max_rel"] = float(((C - ref).norm(dim=1) / ref.norm(dim=1)).max())
print({k: v for k, v in out.items() if k.startswith("gram")}, flush=True)
# throughput
n = 4096
A = torch.randn(n, n, dtype=torch.float64, device=dev); B = torch.randn(n, n, dtype=torch.float64, device=dev)
C = torch.empty(n, n, dtype=torch.float64, device=dev)
ws = torch.empty(L.lib().gsmvi_dgemm_oz_workspace_bytes(n, n, n, 8) + 1024, dtype=torch.uint8, device=dev)
fl = 2.0 * n**3
for s in (8, 7, 6):
    ms = timeit(lambda: L.dgemm_oz(A, B, C, n, n, n, b_mn=True, slices=s, ws=ws))
    out["oz_s%d_4096" % s] = {"ms": ms, "tflops_fp64_equiv": fl / ms / 1e9, "int8_tops": (s * (s + 1) // 2) * fl / ms / 1e9}
    print("oz s=%d 4096^3: %.3f ms  %.1f TF/s fp64-equivalent, %.0f int8 TOP/s" % (s, ms, fl / ms / 1e9, (s * (s + 1) // 2) * fl / ms / 1e9), flush=True)
ms = timeit(lambda: L.dgemm(A, B, C, n, n, n, b_mn=True))
out["native_4096"] = {"ms": ms, "tflops": fl / ms / 1e9}
print("FP64-pipe dgemm 4096^3: %.3f ms %.1f TF/s" % (ms, fl / ms / 1e9), flush=True)
ms = timeit(lambda: torch.matmul(A, B, out=C))
out["cublas_4096"] = {"ms": ms, "tflops": fl / ms / 1e9}
print("cuBLAS dgemm 4096^3: %.3f ms %.1f TF/s" % (ms, fl / ms / 1e9), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "oz_probe.json"), "w"), indent=1)
