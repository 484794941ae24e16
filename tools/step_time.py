"""Steady-state GSM step time at the headline shape without phase stamps (CUDA events around N iterations of GSM.fit's
engine).  Usage: python tools/step_time.py [D] [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gsm-vi_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import gsmvi_oracle as orc
from gsmvi_b200.gsm import GSMEngine
from gsmvi_b200.targets import DenseGaussianTarget
D = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
mean_t, cov_t = orc.dense_gaussian_target(D, 0)
tgt = DenseGaussianTarget(mean_t, cov_t)
eng = GSMEngine(D, D, tgt.lp_g, 99, torch.zeros(D), torch.eye(D))
step = eng.step
for i in range(5): step(i)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for i in range(5, 5 + steps): step(i)
e.record(); torch.cuda.synchronize()
print("GSM step D=%d: %.4f ms (%d steps)" % (D, s.elapsed_time(e) / steps, steps))
