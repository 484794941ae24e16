import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gsm-vi_b200"))
import torch
from gsmvi_b200 import _lib as L
n = 4096
A = torch.randn(n, n, dtype=torch.float64, device="cuda"); B = torch.randn(n, n, dtype=torch.float64, device="cuda")
C = torch.empty(n, n, dtype=torch.float64, device="cuda")
ws = torch.empty(L.lib().gsmvi_dgemm_oz_workspace_bytes(n, n, n, 8) + 1024, dtype=torch.uint8, device="cuda")
for _ in range(2):
    L.dgemm_oz(A, B, C, n, n, n, b_mn=True, slices=8, ws=ws)
torch.cuda.synchronize(); print("done")
