"""Where does GSM.fit spend its time beyond the iterations?  (bench.py's e2e figure)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gsm-vi_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import gsmvi_oracle as orc
from gsmvi_b200.gsm import GSM, GSMEngine
from gsmvi_b200.targets import DenseGaussianTarget
D = B = 4096
mean_t, cov_t = orc.dense_gaussian_target(D, 0)
tgt = DenseGaussianTarget(mean_t, cov_t)
mean_h = torch.zeros(D).pin_memory(); cov_h = torch.eye(D).pin_memory()
def T(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(2):
    t0 = T()
    eng = GSMEngine(D, B, tgt.lp_g, 99, mean_h, cov_h)
    t1 = T()
    for i in range(20): eng.step(i)
    t2 = T()
    m, c = eng.mean().clone(), eng.cov().clone()
    t3 = T()
    mh, ch = m.cpu(), c.cpu()
    t4 = T()
    print("rep %d: engine setup %.1f ms, 20 steps %.1f ms, clone %.1f ms, .cpu() %.1f ms" % (rep, 1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3)), flush=True)
    del eng
