"""A few graph-replayed training steps between cudaProfilerStart/Stop:  ncu --graph-profiling node  lists the kernels of one replay."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2")
args = ap.parse_args()
from soccernerfs_b200.engine.trainer import TrainStep
dev = torch.device("cuda", 0)
model = bench.build_model(args.workload, dev)
model.proposal_sampler.update_sched = lambda step: 0
trainer = TrainStep(model, use_cuda_graph=True)
host = bench._make_batches(8, bench.RAYS_PER_RANK, seed=1000)
res = [h.to(dev) for h in host]
for i in range(6):
    trainer(*bench._bundle(res[i]))
torch.cuda.synchronize()
torch.cuda.profiler.start()
trainer(*bench._bundle(res[6]))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
