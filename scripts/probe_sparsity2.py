"""How sparse is the gradient that reaches the proposal fields (interlevel loss)?  Fraction of non-zero samples and of
warps (32 consecutive samples) with at least one non-zero, over a few eager training steps of a bench workload."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from soccernerfs_b200 import ops
from soccernerfs_b200.engine.trainer import TrainStep

work = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
dev = torch.device("cuda", 0)
model = bench.build_model(work, dev)
model.proposal_sampler.update_sched = lambda step: 0
trainer = TrainStep(model, use_cuda_graph=False, overlap_branches=False)
host = bench._make_batches(30, bench.RAYS_PER_RANK, seed=1000)
orig = ops._DensityField.backward
stats = []


def patched(ctx, grad_density):
    g = grad_density.reshape(-1)
    nz = g != 0
    m = g.numel() // 32 * 32
    warps = nz[:m].view(-1, 32).any(dim=1)
    stats.append((g.numel(), float(nz.float().mean()), float(warps.float().mean())))
    return orig(ctx, grad_density)


ops._DensityField.backward = staticmethod(patched)
for i in range(30):
    trainer(*bench._bundle(host[i].to(dev)))
    if i in (0, 1, 5, 10, 20, 29):
        torch.cuda.synchronize()
        print(f"step {i}:", [(n, round(a, 3), round(b, 3)) for n, a, b in stats[-2:]], flush=True)
