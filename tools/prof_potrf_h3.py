import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gsm-vi_b200"))
import torch
from gsmvi_b200 import _lib as L
D = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
g = torch.Generator().manual_seed(0)
A = torch.randn(D, D, generator=g).cuda(); S = A @ A.t() / D + 0.1 * torch.eye(D, device="cuda")
bad = torch.zeros(1, dtype=torch.int32, device="cuda")
Lh = L.HOperand(D, D, "cuda")
ws3 = torch.empty(L.workspace_bytes(L.WS_POTRF_H3, 0, D) // 4, device="cuda")
Lo3 = torch.zeros(D, D, device="cuda")
for _ in range(2):
    L.potrf_h3(S, Lo3, Lh, D, bad, ws3, zero_upper=False)
torch.cuda.synchronize(); print("h3 ok", int(bad.item()))
if len(sys.argv) > 3:  # correctness against torch at this size
    ref = torch.linalg.cholesky(S.double())
    print("relF(L) vs fp64 cholesky: %.2e, split dequant max err %.2e" % (float((Lo3.double() - ref).norm() / ref.norm()),
          float((Lh.dequant() - Lo3).abs().max())))
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import time
s.record()
t0 = time.perf_counter()
for _ in range(reps):
    L.potrf_h3(S, Lo3, Lh, D, bad, ws3, zero_upper=False)
t1 = time.perf_counter()  # the host has enqueued everything (nothing above synchronises)
e.record(); torch.cuda.synchronize()
print("potrf_h3 D=%d: %.3f ms per factorisation (host: %.3f ms per call to enqueue its launches)" % (D, s.elapsed_time(e) / reps, 1e3 * (t1 - t0) / reps))

# stated baseline (SURVEY.md section 9 allows a library call beside ours): cuSOLVER fp32 potrf through torch.linalg.cholesky
torch.linalg.cholesky(S); torch.cuda.synchronize()
s.record()
for _ in range(reps):
    torch.linalg.cholesky(S)
e.record(); torch.cuda.synchronize()
print("cuSOLVER fp32 potrf (torch.linalg.cholesky) D=%d: %.3f ms per factorisation" % (D, s.elapsed_time(e) / reps))
