"""What bounds the pair kernel?  The score-shaped launch (4096^3) timed with parts of its work switched off through
GSMVI_X2_PROBE (results are garbage in those runs): bit 0 = no TMA loads (the MMAs run on whatever is in shared memory),
bit 1 = hi*hi MMAs only (a third of the tensor work on the same operand traffic)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gsm-vi_b200"))
import torch
from gsmvi_b200 import _lib as L
D = B = 4096
g = torch.Generator().manual_seed(0)
H = lambda X: L.HOperand(X.shape[0], X.shape[1], "cuda").split_from(X)
P = torch.randn(D, D, generator=g).cuda(); P = H((P + P.t()) / 2)
Xh = H(torch.randn(B, D, generator=g).cuda())
G = torch.empty(B, D, device="cuda"); bias = torch.zeros(D, device="cuda")
fn = lambda: L.gemm_h3(Xh, P, G, B, D, D, alpha=-1.0, bias_n=bias)
L.h3_pair_kernel(1)
variants = [("full", "0", "16"), ("no_tma", "1", "16"), ("hihi_only", "2", "16"), ("no_tma_hihi_only", "3", "16"),
            ("full_chunk4", "0", "4"), ("no_tma_chunk4", "1", "4"), ("no_tma_chunk64", "1", "64")]
reps = 20
evs = []
for r in range(reps + 2):
    for k in range(len(variants)):
        name, probe, chunk = variants[(k + r) % len(variants)]
        os.environ["GSMVI_X2_PROBE"] = probe
        os.environ["GSMVI_X2_CHUNK_KB"] = chunk
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        if r >= 2:
            evs.append((name, e0, e1))
torch.cuda.synchronize()
os.environ["GSMVI_X2_PROBE"] = "0"
tot = {}
for name, e0, e1 in evs:
    tot[name] = tot.get(name, 0.0) + e0.elapsed_time(e1) / reps
floor_cycles = 7 * 64 * 12 * 64   # 7 rounds of supertiles x 64 k-blocks x 12 MMAs x 64 cycles
out = {n: {"ms": round(ms, 4), "mma_floor_ms_at_1965MHz": round(floor_cycles / 1.965e6 * (1 / 3 if "hihi" in n else 1), 4)} for n, ms in tot.items()}
print(json.dumps(out, indent=1))
