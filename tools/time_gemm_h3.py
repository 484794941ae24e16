"""CUDA-event timing of the four batch-sized GEMM launches of a GSM step (D = B = 4096 by default) through both kernels
behind gsmvi_gemm_h3: the persistent 2-CTA kernel (h3x2_gemm.cuh) at several TMEM chunk lengths and the one-CTA kernel
(h3_gemm.cuh).  The variants are INTERLEAVED launch by launch (the part is power-limited on these GEMMs: a variant timed after
a long run of another one sees lower clocks), every launch bracketed by its own events."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gsm-vi_b200"))
import torch
from gsmvi_b200 import _lib as L
D = B = int(os.environ.get("GSMVI_PROF_D", "4096"))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
g = torch.Generator().manual_seed(0)
H = lambda X: L.HOperand(X.shape[0], X.shape[1], "cuda").split_from(X)
Z = H(torch.randn(B, D, generator=g).cuda())
Lm = H(torch.tril(torch.randn(D, D, generator=g)).cuda() / D**0.5)
P = torch.randn(D, D, generator=g).cuda(); P = H((P + P.t()) / 2)
T = H(torch.randn(3 * B, D, generator=g).cuda())
Ta = L.HOperand.from_tensors(T.hi[:2 * B], T.lo[:2 * B], T.scale, 2 * B, D)
Tb = L.HOperand.from_tensors(T.hi[B:], T.lo[B:], T.scale, 2 * B, D)
S32 = torch.randn(D, D, generator=g).cuda(); S32 = (S32 + S32.t()) / 2; S = H(S32)
Xh = H(torch.randn(B, D, generator=g).cuda()); Gh = H(torch.randn(B, D, generator=g).cuda())
X = torch.empty(B, D, device="cuda"); G = torch.empty(B, D, device="cuda"); W = torch.empty(B, D, device="cuda")
So = torch.empty(D, D, device="cuda"); bias = torch.zeros(D, device="cuda")
calls = {
    "sample": lambda: L.gemm_h3(Z, Lm, X, B, D, D, krange=L.KR_B_LOWER, bias_n=bias),
    "score": lambda: L.gemm_h3(Xh, P, G, B, D, D, alpha=-1.0, bias_n=bias),
    "w": lambda: L.gemm_h3(Gh, S, W, B, D, D),
    "cov_update": lambda: L.gemm_h3(Ta, Tb, So, D, D, 2 * B, a_mn=True, b_mn=True, alpha=-1.0 / B, beta=1.0, Cin=S32, tri=True, mirror=True),
}
flops = {"sample": B * D * D, "score": 2 * B * D * D, "w": 2 * B * D * D, "cov_update": 2 * D * D * 2 * B / 2 + D * 128 * 2 * B}
variants = [("pair_chunk4", 1, "4"), ("pair_chunk8", 1, "8"), ("pair_chunk16", 1, "16"), ("single", 0, "4")]
tot = {v[0]: {n: 0.0 for n in calls} for v in variants}
def select(v):
    L.h3_pair_kernel(v[1])
    os.environ["GSMVI_X2_CHUNK_KB"] = v[2]
for v in variants:          # warm-up
    select(v)
    for fn in calls.values():
        fn()
torch.cuda.synchronize()
evs = []
for r in range(reps):
    for name, fn in calls.items():
        for k in range(len(variants)):
            v = variants[(k + r) % len(variants)]   # rotate who goes first
            select(v)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            evs.append((v[0], name, e0, e1))
torch.cuda.synchronize()
for vn, name, e0, e1 in evs:
    tot[vn][name] += e0.elapsed_time(e1) / reps
out = {}
for vn in tot:
    out[vn] = {n: {"ms": round(ms, 4), "executed_tflops": round(3 * flops[n] / ms / 1e9, 1)} for n, ms in tot[vn].items()}
    out[vn]["total_ms"] = round(sum(tot[vn].values()), 4)
print(json.dumps(out, indent=1))
