"""How sparse is the gradient that reaches the proposal fields?  Per level: fraction of samples with a zero upstream
gradient, fraction of 32-sample groups (one warp of kp_density_field_bwd) that are entirely zero, and what a perfect
compaction would leave."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from soccernerfs_b200 import ops
from soccernerfs_b200.engine.trainer import TrainStep

dev = torch.device("cuda", 0)
for name in sys.argv[1:] or ["cfg2"]:
    model = bench.build_model(name, dev)
    model.proposal_sampler.update_sched = lambda step: 0
    trainer = TrainStep(model, use_cuda_graph=False, overlap_branches=False)
    host = bench._make_batches(12, bench.RAYS_PER_RANK, seed=1000)
    stats = []
    orig = ops._DensityField.backward

    def spy(ctx, g):
        z = (g.reshape(-1) == 0)
        n = z.numel() // 32 * 32
        warps = z[:n].view(-1, 32)
        stats.append((z.numel(), float(z.float().mean()), float(warps.all(dim=1).float().mean()),
                      float((~warps).any(dim=1).float().mean())))
        return orig(ctx, g)

    ops._DensityField.backward = staticmethod(spy)
    for i in range(12):
        trainer(*bench._bundle(host[i].to(dev)))
    torch.cuda.synchronize()
    ops._DensityField.backward = staticmethod(orig)
    for s in stats[-4:]:
        print(name, "samples %d zero %.3f warps_all_zero %.3f warps_active %.3f dense-equivalent warps %.3f" % (s[0], s[1], s[2], s[3], 1 - s[1]))
