import os, sys
sys.path.insert(0, "/root/repo")
import torch
import bench
from soccernerfs_b200 import ops
from soccernerfs_b200.model_components.losses import regularizer_plan, _const, REG_NAMES
dev = torch.device("cuda", 0)
model = bench.build_model("cfg3", dev)
planes, terms, rows = regularizer_plan(model.field.grids, [p.grids for p in model.proposal_networks])
coef = torch.rand(len(planes), 4, device=dev) * 1e-3
targets = [torch.zeros_like(p, memory_format=torch.preserve_format) for p in planes]
pl = [p.detach() for p in planes]
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
full = torch.tensor([[0, p.numel() // 4] for p in planes], dtype=torch.int64, device=dev)
eighth = torch.tensor([[0, p.numel() // 32] for p in planes], dtype=torch.int64, device=dev)
none = torch.zeros(len(planes), 2, dtype=torch.int64, device=dev)
print("no range      ", t(lambda: ops.plane_reg_fused(pl, terms, coef, targets, False)))
print("range = full  ", t(lambda: ops.plane_reg_fused(pl, terms, coef, targets, False, write_range=full)))
print("range = 1/8   ", t(lambda: ops.plane_reg_fused(pl, terms, coef, targets, False, write_range=eighth)))
print("range = empty ", t(lambda: ops.plane_reg_fused(pl, terms, coef, targets, False, write_range=none)))
