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
elif group in ("h3", "h3mn", "h3perf"):
    def hop(X):
        return L.HOperand(X.shape[0], X.shape[1], dev, ld=(X.shape[1] + 31) // 32 * 32).split_from(X)

    def h3_case(name, M, N, K, a_mn=False, b_mn=False, scale_a=1.0, scale_b=1.0, **kw):
        g = torch.Generator(device="cpu").manual_seed(abs(hash(name)) % (2**31))
        A = (torch.randn(M, K, generator=g, dtype=torch.float32) * scale_a).to(dev)
        B = (torch.randn(N, K, generator=g, dtype=torch.float32) * scale_b).to(dev)
        Aop = hop(A.t().contiguous() if a_mn else A)
        Bop = hop(B.t().contiguous() if b_mn else B)
        C = torch.full((M, N + 4), float("nan"), device=dev)[:, :N]
        L.gemm_h3(Aop, Bop, C, M, N, K, a_mn=a_mn, b_mn=b_mn, **kw)
        torch.cuda.synchronize()
        e = relerr(C, A.double() @ B.double().t())
        out[name] = e
        print("%-40s M=%d N=%d K=%d  relerr=%.3e" % (name, M, N, K, e), flush=True)
        return e

    if group == "h3":
        # the split itself: hi + lo/2^11 reproduces the fp32 value to 2^-22
        X = torch.randn(300, 200, device=dev) * 37.0
        o = hop(X)
        out["split_roundtrip"] = float(((o.dequant() - X).abs() / X.abs().clamp_min(1e-30)).max())
        out["split_scale"] = float(o.scale)
        print(out, flush=True)
        h3_case("kk_128x128x64", 128, 128, 64)
        h3_case("kk_128x128x256", 128, 128, 256)
        h3_case("kk_256x384x512", 256, 384, 512)
        h3_case("kk_ragged", 300, 200, 100)
        h3_case("kk_tiny", 10, 10, 12)
        h3_case("kk_1024", 1024, 1024, 1024)
        h3_case("kk_K4096", 512, 512, 4096)
        h3_case("kk_K8192", 512, 512, 8192)
        h3_case("kk_scales", 256, 256, 512, scale_a=3.0e-7, scale_b=8.0e5)
        # same-sign sums: worst case for the truncating TMEM accumulate
        A = torch.rand(512, 8192, device=dev) + 0.5
        C = torch.empty(512, 512, device=dev)
        Ao = hop(A)
        L.gemm_h3(Ao, Ao, C, 512, 512, 8192)
        out["positive_K8192"] = relerr(C, A.double() @ A.double().t())
        L.gemm_tf32(A, A, C, 512, 512, 8192)
        out["positive_K8192_tf32x3"] = relerr(C, A.double() @ A.double().t())
        # wide dynamic range inside one tensor (rows scaled by 2^-20 .. 1)
        g = torch.Generator().manual_seed(3)
        A = torch.randn(256, 512, generator=g).to(dev) * torch.logspace(-6, 0, 256, device=dev)[:, None]
        B = torch.randn(256, 512, generator=g).to(dev)
        C = torch.empty(256, 256, device=dev)
        L.gemm_h3(hop(A), hop(B), C, 256, 256, 512)
        ref = A.double() @ B.double().t()
        out["dynrange_rowwise_relerr_max"] = float(((C.double() - ref).norm(dim=1) / ref.norm(dim=1)).max())
        # epilogue: alpha/beta/bias, absmax, tri+mirror, krange, split-K
        M, N, K = 384, 384, 256
        A = torch.randn(M, K, generator=g).to(dev)
        Cin = torch.randn(M, N, generator=g).to(dev)
        bias = torch.randn(N, generator=g).to(dev)
        Ao = hop(A)
        C = torch.empty(M, N, device=dev)
        am = torch.zeros(1, dtype=torch.int32, device=dev)
        L.gemm_h3(Ao, Ao, C, M, N, K, alpha=0.5, beta=-2.0, Cin=Cin, bias_n=bias, absmax_out=am)
        ref = 0.5 * (A.double() @ A.double().t()) - 2.0 * Cin.double() + bias.double()
        out["alpha_beta_bias"] = relerr(C, ref)
        out["absmax_matches"] = float(am.view(torch.float32)) - float(C.abs().max())
        C2 = (Cin + Cin.t()) / 2
        sym = C2.clone()
        L.gemm_h3(Ao, Ao, C2, M, N, K, alpha=1.0 / K, beta=1.0, Cin=C2, tri=True, mirror=True)
        out["tri_mirror_inplace"] = relerr(C2, sym.double() + (A.double() @ A.double().t()) / K)
        out["tri_mirror_asym"] = float((C2 - C2.t()).abs().max())
        Lm = torch.tril(torch.randn(N, N, generator=g)).to(dev)
        Z = torch.randn(M, N, generator=g).to(dev)
        C4 = torch.empty(M, N, device=dev)
        L.gemm_h3(hop(Z), hop(Lm), C4, M, N, N, krange=L.KR_B_LOWER, bias_n=bias)
        out["krange_b_lower"] = relerr(C4, Z.double() @ Lm.double().t() + bias.double())
        S = 3
        Cs = torch.zeros(S, M, N, device=dev)
        L.gemm_h3(Ao, Ao, Cs.view(S * M, N), M, N, K, splits=S, split_stride=M * N)
        out["splitk3"] = relerr(Cs.sum(0), A.double() @ A.double().t())
        Cs = torch.zeros(7, M, N, device=dev)
        L.gemm_h3(Ao, Ao, Cs.view(7 * M, N), M, N, K, splits=7, split_stride=M * N)
        out["splitk7_more_than_blocks"] = relerr(Cs.sum(0), A.double() @ A.double().t())
        for k, v in out.items():
            print("%-32s %s" % (k, v), flush=True)
    elif group == "h3mn":
        h3_case("mnA", 256, 256, 128, a_mn=True)
        h3_case("mnB", 256, 256, 128, b_mn=True)
        h3_case("mnAB", 256, 256, 128, a_mn=True, b_mn=True)
        h3_case("mnAB_big", 256, 384, 1024, a_mn=True, b_mn=True)
        h3_case("mnAB_ragged", 300, 200, 100, a_mn=True, b_mn=True)
        h3_case("mnAB_tri_mirror", 384, 384, 192, a_mn=True, b_mn=True)
    else:
        for n in (2048, 4096, 8192):
            A = torch.randn(n, n, device=dev)
            B = torch.randn(n, n, device=dev)
            C = torch.empty(n, n, device=dev)
            Ao, Bo = hop(A), hop(B)
            fl = 2.0 * n**3
            ms = timeit(lambda: L.gemm_h3(Ao, Bo, C, n, n, n))
            out["h3_%d" % n] = {"ms": ms, "tflops_alg": fl / ms / 1e9, "tflops_exec": 3 * fl / ms / 1e9}
            print("h3 n=%d: %.3f ms  %.1f TF/s algorithmic (%.1f executed)" % (n, ms, fl / ms / 1e9, 3 * fl / ms / 1e9), flush=True)
            ms = timeit(lambda: L.gemm_h3(Ao, Ao, C, n, n, n, tri=True, mirror=True, a_mn=True, b_mn=True))
            out["h3_syrk_mn_%d" % n] = {"ms": ms, "tflops_alg_dense": fl / ms / 1e9}
            print("h3 syrk mn n=%d: %.3f ms  %.1f TF/s dense-counted" % (n, ms, fl / ms / 1e9), flush=True)
            ms = timeit(lambda: L.gemm_h3(Ao, Bo, C, n, n, n, krange=L.KR_B_LOWER))
            out["h3_trmm_%d" % n] = {"ms": ms}
            print("h3 B-lower n=%d: %.3f ms" % (n, ms), flush=True)
            ms = timeit(lambda: Ao.split_from(A))
            out["split_%d" % n] = {"ms": ms, "GBps": (12.0 * n * n) / ms / 1e6}
            print("absmax+split n=%d: %.3f ms" % (n, ms), flush=True)
            Ah, Bh = A.half(), B.half()
            Ch = torch.empty(n, n, device=dev, dtype=torch.float16)
            ms = timeit(lambda: torch.matmul(Ah, Bh.t(), out=Ch))
            out["cublas_f16_%d" % n] = {"ms": ms, "tflops": fl / ms / 1e9}
            print("cublas f16 n=%d: %.3f ms %.1f TF/s" % (n, ms, fl / ms / 1e9), flush=True)
        # left-looking Cholesky panel update shape: M x 128 output, long K, split-K
        n = 4096
        A = torch.randn(n, n, device=dev)
        Ao = hop(A)
        for (M, K, S) in [(2048, 2048, 9), (1024, 3072, 18), (3968, 128, 1), (3072, 1024, 6)]:
            Cs = torch.empty(S * M, 128, device=dev)
            ms = timeit(lambda: L.gemm_h3(Ao, Ao, Cs, M, 128, K, splits=S, split_stride=M * 128), iters=20)
            out["panel_update_M%d_K%d_S%d" % (M, K, S)] = {"ms": ms}
            print("panel update M=%d K=%d splits=%d: %.4f ms" % (M, K, S, ms), flush=True)
else:
    raise SystemExit("unknown group")

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "gemm_probe_%s.json" % group), "w") as f:
    json.dump(out, f, indent=1)
