"""The four batch-sized GEMM launches of one GSM step at D = B = 4096 on the scaled 3xFP16 engine, repeated; target of
the ncu --set full capture (profiles/r01_ncu_gemm_h3_summary.txt)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gsm-vi_b200"))
import torch
from gsmvi_b200 import _lib as L
D = B = int(os.environ.get("GSMVI_PROF_D", "4096"))
g = torch.Generator().manual_seed(0)
H = lambda X: L.HOperand(X.shape[0], X.shape[1], "cuda").split_from(X)
Z = H(torch.randn(B, D, generator=g).cuda())
Lm = H(torch.tril(torch.randn(D, D, generator=g)).cuda() / D**0.5)
P = torch.randn(D, D, generator=g).cuda(); P = H((P + P.t()) / 2)
T = H(torch.randn(3 * B, D, generator=g).cuda())
Ta = L.HOperand.from_tensors(T.hi[:2 * B], T.lo[:2 * B], T.scale, 2 * B, D)
Tb = L.HOperand.from_tensors(T.hi[B:], T.lo[B:], T.scale, 2 * B, D)
S32 = torch.eye(D).cuda(); S = H(S32)
Xh = H(torch.randn(B, D, generator=g).cuda()); Gh = H(torch.randn(B, D, generator=g).cuda())
X = torch.empty(B, D, device="cuda"); G = torch.empty(B, D, device="cuda"); W = torch.empty(B, D, device="cuda")
So = torch.empty(D, D, device="cuda"); bias = torch.zeros(D, device="cuda")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(reps):
    L.gemm_h3(Z, Lm, X, B, D, D, krange=L.KR_B_LOWER, bias_n=bias)                      # sample
    L.gemm_h3(Xh, P, G, B, D, D, alpha=-1.0, bias_n=bias)                               # score
    L.gemm_h3(Gh, S, W, B, D, D)                                                        # W = G Sigma
    L.gemm_h3(Ta, Tb, So, D, D, 2 * B, a_mn=True, b_mn=True, alpha=-1.0 / B, beta=1.0, Cin=S32, tri=True, mirror=True)
torch.cuda.synchronize()
print("done")
