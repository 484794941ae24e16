"""BaM iteration timing on N GPUs (torchrun): tensor-parallel solve vs the replicated one (GSMVI_BAM_TP=0), with the solve's
stage timing (GSMVI_BAM_TIMING=1).  python -m torch.distributed.run --nproc-per-node N tools/bench_bam_tp.py [D] [B] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gsm-vi_b200"))
import torch
import torch.distributed as dist

from gsmvi_b200.bam import BaMEngine
from gsmvi_b200.targets import DenseGaussianTarget, illcond_gaussian_target

D = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
group = None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    group = dist.group.WORLD
mean_t, cov_t = illcond_gaussian_target(D, 1e2, 0)
tgt = DenseGaussianTarget(mean_t, cov_t)
eng = BaMEngine(D, B, tgt.lp_g, key=99, npass=3, process_group=group)
eng.step(0, 100.0)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(1, steps + 1):
    eng.step(i, 100.0 / (1 + i))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
if world > 1:
    t = torch.tensor([ms], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
if rank == 0:
    print("BaM D=%d B=%d world=%d tp=%s: %.2f ms/step = %.2f it/s, ns_iters %s, reverts %d" % (
        D, B, world, os.environ.get("GSMVI_BAM_TP", "1"), ms, 1e3 / ms, eng.ns_iters, eng.n_reverts), flush=True)
eng.close()
if world > 1:
    dist.destroy_process_group()
