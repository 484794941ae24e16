"""Per-phase CUDA-event timing of the GSM step at the headline shape (GSMVI_PHASE_TIMING=1), with and without the side
stream beside the Cholesky.  Usage: python tools/phase_timing.py [D] [steps]"""
import os, sys
os.environ["GSMVI_PHASE_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gsm-vi_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import gsmvi_oracle as orc
from gsmvi_b200.gsm import GSM
from gsmvi_b200 import gsm as gsm_mod
from gsmvi_b200.targets import DenseGaussianTarget
D = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
mean_t, cov_t = orc.dense_gaussian_target(D, 0)
tgt = DenseGaussianTarget(mean_t, cov_t)
g = GSM(D, tgt.lp, tgt.lp_g)
g.fit(99, niter=3, batch_size=D, verbose=False)
g.fit(99, niter=steps, batch_size=D, verbose=False)
gsm_mod.release_engines()
