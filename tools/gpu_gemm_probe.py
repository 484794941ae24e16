"""GPU probe for the tcgen05 GEMM: correctness vs fp64 over layouts/modes + throughput. Run per group in its own
process so a trapped kernel cannot poison later groups:  python tools/gpu_gemm_probe.py <group>"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gsm-vi_b200"))
import torch
from gsmvi_b200 import _lib as L

dev = "cuda"
out = {}


def relerr(C, ref):
    return float((C.double() - ref).norm() / ref.norm().clamp_min(1e-300))


def run_case(name, M, N, K, a_mn=False, b_mn=False, npass=3, ld_pad=0, **kw):
    g = torch.Generator(device="cpu").manual_seed(hash(name) % (2**31))
    A = torch.randn(M, K, generator=g, dtype=torch.float32).to(dev)
    B = torch.randn(N, K, generator=g, dtype=torch.float32).to(dev)
    Aop = A.t().contiguous() if a_mn else A
    Bop = B.t().contiguous() if b_mn else B
    C = torch.full((M, N + ld_pad), float("nan"), device=dev)[:, :N]
    L.gemm_tf32(Aop, Bop, C, M, N, K, a_mn=a_mn, b_mn=b_mn, npass=npass, **kw)
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    e = relerr(C, ref)
    out[name] = e
    print("%-40s M=%d N=%d K=%d  relerr=%.3e" % (name, M, N, K, e), flush=True)
    return e


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


group = sys.argv[1]
if group == "basic1":
    run_case("kk_1pass_128x128x32", 128, 128, 32, npass=1)
    run_case("kk_1pass_128x128x256", 128, 128, 256, npass=1)
    run_case("kk_1pass_256x384x512", 256, 384, 512, npass=1)
    run_case("kk_1pass_ragged", 300, 200, 100, npass=1)
    run_case("kk_1pass_tiny", 10, 10, 12, npass=1)
elif group == "basic3":
    run_case("kk_3pass_128x128x32", 128, 128, 32)
    run_case("kk_3pass_128x128x256", 128, 128, 256)
    run_case("kk_3pass_256x384x512", 256, 384, 512)
    run_case("kk_3pass_ragged", 300, 200, 100)
    run_case("kk_3pass_tiny", 10, 10, 12)
    run_case("kk_3pass_1024", 1024, 1024, 1024)
    run_case("kk_3pass_K4096", 512, 512, 4096)
    run_case("kk_3pass_K8192", 512, 512, 8192)
    # same-sign sums (Gram-matrix diagonals): worst case for the truncating TMEM accumulate
    A = torch.rand(512, 8192, device=dev) + 0.5
    C = torch.empty(512, 512, device=dev)
    L.gemm_tf32(A, A, C, 512, 512, 8192)
    out["kk_3pass_positive_K8192"] = relerr(C, A.double() @ A.double().t())
    Cb = torch.matmul(A, A.t())
    out["cublas_fp32_positive_K8192"] = relerr(Cb, A.double() @ A.double().t())
    print({k: v for k, v in out.items() if "positive" in k}, flush=True)
elif group == "mn":
    run_case("mnA_1pass", 256, 256, 128, a_mn=True, npass=1)
    run_case("mnB_1pass", 256, 256, 128, b_mn=True, npass=1)
    run_case("mnAB_1pass", 256, 256, 128, a_mn=True, b_mn=True, npass=1)
    run_case("mnAB_3pass", 256, 384, 160, a_mn=True, b_mn=True, npass=3)
    run_case("mnAB_3pass_ragged", 300, 200, 100, a_mn=True, b_mn=True, npass=3)
elif group == "round":
    # does kind::tf32 truncate or round the 13 low mantissa bits?  a = 1 + 1.5*2^-11: trunc -> 1, rna -> 1+2^-10
    M = N = 128
    K = 32
    A = torch.zeros(M, K, device=dev)
    A[:, 0] = 1.0 + 1.5 * 2.0**-11
    B = torch.zeros(N, K, device=dev)
    B[:, 0] = 1.0
    C = torch.zeros(M, N, device=dev)
    L.gemm_tf32(A, B, C, M, N, K, npass=1)
    torch.cuda.synchronize()
    v = float(C[0, 0])
    out["tf32_input_handling"] = {"value": v, "mode": "truncate" if v == 1.0 else ("round" if v == 1.0 + 2.0**-10 else "other")}
    print(out, flush=True)
    C3 = torch.zeros(M, N, device=dev)
    L.gemm_tf32(A, B, C3, M, N, K, npass=3)
    torch.cuda.synchronize()
    out["split_exact"] = float(C3[0, 0]) - (1.0 + 1.5 * 2.0**-11)
    print(out, flush=True)
elif group == "epi":
    M, N, K = 384, 384, 256
    g = torch.Generator().manual_seed(5)
    A = torch.randn(M, K, generator=g).to(dev)
    Cin = torch.randn(M, N, generator=g).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    # alpha/beta/bias
    C = torch.empty(M, N, device=dev)
    L.gemm_tf32(A, A, C, M, N, K, alpha=0.5, beta=-2.0, Cin=Cin, bias_n=bias)
    ref = 0.5 * (A.double() @ A.double().t()) - 2.0 * Cin.double() + bias.double()
    out["alpha_beta_bias"] = relerr(C, ref)
    # tri + mirror, in place
    C2 = Cin.clone()
    C2 = (C2 + C2.t()) / 2
    sym = C2.clone()
    L.gemm_tf32(A, A, C2, M, N, K, alpha=1.0 / K, beta=1.0, Cin=C2, tri=True, mirror=True)
    ref = sym.double() + (A.double() @ A.double().t()) / K
    out["tri_mirror_inplace"] = relerr(C2, ref)
    out["tri_mirror_asym"] = float((C2 - C2.t()).abs().max())
    # signed K halves:  [D;E]^T [D;-E] with MN-major operand
    Bt = 96
    T = torch.randn(2 * Bt, M, generator=g).to(dev)
    C3 = torch.empty(M, M, device=dev)
    L.gemm_tf32(T, T, C3, M, M, 2 * Bt, a_mn=True, b_mn=True, tri=True, mirror=True, neg_from=Bt)
    Dm, Em = T[:Bt].double(), T[Bt:].double()
    ref = Dm.t() @ Dm - Em.t() @ Em
    out["signed_mn_tri"] = relerr(C3, ref)
    # krange: B lower triangular
    Lm = torch.tril(torch.randn(N, N, generator=g)).to(dev)
    Z = torch.randn(M, N, generator=g).to(dev)
    C4 = torch.empty(M, N, device=dev)
    L.gemm_tf32(Z, Lm, C4, M, N, N, krange=L.KR_B_LOWER, bias_n=bias)
    ref = Z.double() @ Lm.double().t() + bias.double()
    out["krange_b_lower"] = relerr(C4, ref)
    Um = Lm.t().contiguous()
    C5 = torch.empty(M, N, device=dev)
    L.gemm_tf32(Z, Um, C5, M, N, N, krange=L.KR_B_UPPER)
    out["krange_b_upper"] = relerr(C5, Z.double() @ Um.double().t())
    C6 = torch.empty(N, M, device=dev)
    L.gemm_tf32(Lm, Z, C6, N, M, N, krange=L.KR_A_LOWER)
    out["krange_a_lower"] = relerr(C6, Lm.double() @ Z.double().t())
    # sub-views (Cholesky panel style)
    big = torch.randn(640, 640, generator=g).to(dev)
    A21 = big[256:, 128:256]
    Cv = big[256:, 256:]
    before = Cv.clone()
    L.gemm_tf32(A21, A21, Cv, 384, 384, 128, alpha=-1.0, beta=1.0, Cin=Cv, tri=True)
    ref = before.double() - A21.double() @ A21.double().t()
    out["subview_syrk_lower"] = float((torch.tril(Cv.double() - ref)).norm() / ref.norm())
    out["subview_upper_untouched"] = float((torch.triu(Cv - before, 1)).abs().max())
    for k, v in out.items():
        print("%-28s %s" % (k, v), flush=True)
elif group == "presplit":
    from gsmvi_b200._util import new_mat
    M, N, K = 384, 512, 1024
    g = torch.Generator().manual_seed(7)
    A = torch.randn(M, K, generator=g).to(dev); B = torch.randn(N, K, generator=g).to(dev)
    ref = A.double() @ B.double().t()
    C = torch.empty(M, N, device=dev)
    Bhi = torch.empty_like(B); Blo = torch.empty_like(B); L.tf32_split(B, Bhi, Blo, N, K)
    L.gemm_tf32(A, Bhi, C, M, N, K, B_lo=Blo)
    out["presplit_B"] = relerr(C, ref)
    L.gemm_tf32(A, B, C, M, N, K)
    out["no_presplit"] = relerr(C, ref)
    # both pre-split with the round-to-nearest split (as the row pass writes T_hi / T_lo), MN-major
    At, Bt = A.t().contiguous(), B.t().contiguous()
    def rn_split(x):
        hi = (x.view(torch.int32) + 0x1000 & ~0x1FFF).view(torch.float32)
        lo = x - hi
        lo = ((lo.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
        return hi.contiguous(), lo.contiguous()
    Ahi, Alo = rn_split(At); Bhi, Blo2 = rn_split(Bt)
    L.gemm_tf32(Ahi, Bhi, C, M, N, K, a_mn=True, b_mn=True, A_lo=Alo, B_lo=Blo2)
    out["presplit_both_mn"] = relerr(C, ref)
    for k, v in out.items():
        print("%-20s %.3e" % (k, v), flush=True)
    n = 4096
    A = torch.randn(n, n, device=dev); B = torch.randn(n, n, device=dev); C = torch.empty(n, n, device=dev)
    Bhi = torch.empty_like(B); Blo = torch.empty_like(B); L.tf32_split(B, Bhi, Blo, n, n)
    Ahi = torch.empty_like(A); Alo = torch.empty_like(A); L.tf32_split(A, Ahi, Alo, n, n)
    fl = 2.0 * n**3
    for name, (a_, b_, kw) in [("none", (A, B, {})), ("B", (A, Bhi, {"B_lo": Blo})), ("both", (Ahi, Bhi, {"A_lo": Alo, "B_lo": Blo}))]:
        ms = timeit(lambda: L.gemm_tf32(a_, b_, C, n, n, n, **kw))
        out["perf_presplit_%s_4096" % name] = {"ms": ms, "tflops_alg": fl / ms / 1e9}
        print("presplit=%s 4096^3: %.3f ms %.1f TF/s algorithmic (%.1f executed)" % (name, ms, fl / ms / 1e9, 3 * fl / ms / 1e9), flush=True)
elif group == "perf":
    for n in (2048, 4096, 8192):
        A = torch.randn(n, n, device=dev)
        B = torch.randn(n, n, device=dev)
        C = torch.empty(n, n, device=dev)
        fl = 2.0 * n**3
        for npass in (1, 3):
            ms = timeit(lambda: L.gemm_tf32(A, B, C, n, n, n, npass=npass))
            out["ours_%dpass_%d" % (npass, n)] = {"ms": ms, "tflops_alg": fl / ms / 1e9, "tflops_exec": npass * fl / ms / 1e9}
            print("ours npass=%d n=%d: %.3f ms  %.1f TF/s algorithmic (%.1f executed)" % (npass, n, ms, fl / ms / 1e9, npass * fl / ms / 1e9), flush=True)
        ms = timeit(lambda: L.gemm_tf32(A, A, C, n, n, n, npass=3, tri=True, mirror=True, a_mn=True, b_mn=True))
        out["ours_3pass_syrk_mn_%d" % n] = {"ms": ms, "tflops_alg_dense": fl / ms / 1e9}
        print("ours syrk mn 3pass n=%d: %.3f ms  %.1f TF/s dense-counted" % (n, ms, fl / ms / 1e9), flush=True)
        torch.backends.cuda.matmul.allow_tf32 = True
        ms = timeit(lambda: torch.matmul(A, B.t(), out=C))
        out["cublas_tf32_%d" % n] = {"ms": ms, "tflops": fl / ms / 1e9}
        print("cublas tf32 n=%d: %.3f ms %.1f TF/s" % (n, ms, fl / ms / 1e9), flush=True)
        torch.backends.cuda.matmul.allow_tf32 = False
        ms = timeit(lambda: torch.matmul(A, B.t(), out=C))
        out["cublas_fp32_%d" % n] = {"ms": ms, "tflops": fl / ms / 1e9}
        print("cublas fp32 n=%d: %.3f ms %.1f TF/s" % (n, ms, fl / ms / 1e9), flush=True)
        if n <= 4096:
            Ad, Bd = A.double(), B.double()
            Cd = torch.empty(n, n, device=dev, dtype=torch.float64)
            ms = timeit(lambda: torch.matmul(Ad, Bd.t(), out=Cd), iters=5)
            out["cublas_fp64_%d" % n] = {"ms": ms, "tflops": fl / ms / 1e9}
            print("cublas fp64 n=%d: %.3f ms %.1f TF/s" % (n, ms, fl / ms / 1e9), flush=True)
            del Ad, Bd, Cd
else:
    raise SystemExit("unknown group")

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "gemm_probe_%s.json" % group), "w") as f:
    json.dump(out, f, indent=1)
