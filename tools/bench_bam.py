"""BaM iterations/s at D = B = 4096 (and the C3 shape) with the example schedule reg_i = 100 / (1 + i)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gsm-vi_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import gsmvi_oracle as orc
from gsmvi_b200.bam import BaMEngine
from gsmvi_b200.targets import DenseGaussianTarget
D = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
B = int(sys.argv[2]) if len(sys.argv) > 2 else D
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4
mean_t, cov_t = orc.dense_gaussian_target(D, 0)
tgt = DenseGaussianTarget(mean_t, cov_t)
eng = BaMEngine(D, B, tgt.lp_g, key=99, npass=3)
eng.step(0, 100.0)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for i in range(1, n + 1): eng.step(i, 100.0 / (1 + i))
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / n
err = float((eng.cov().double().cpu() - torch.as_tensor(cov_t)).norm() / torch.as_tensor(cov_t).norm())
print(json.dumps({"D": D, "B": B, "ms_per_iter": ms, "it_per_s": 1e3 / ms, "ns_iters": eng.ns_iters, "reverts": eng.n_reverts,
                  "relF_cov_vs_target_after_%d" % (n + 1): err, "oz_slices": os.environ.get("GSMVI_OZ_SLICES", "8")}))
