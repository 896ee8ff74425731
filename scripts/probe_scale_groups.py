"""cfg3 scatter / gather launched per GROUP of scales (does splitting the L2 working set beat the repeated per-sample set-up?)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctypes import c_void_p
import torch
import bench
from soccernerfs_b200 import _lib, ops
from soccernerfs_b200.engine.trainer import TrainStep

dev = torch.device("cuda", 0)
model = bench.build_model("cfg3", dev)
model.proposal_sampler.update_sched = lambda step: 0
trainer = TrainStep(model, use_cuda_graph=False, overlap_branches=False)
host = bench._make_batches(4, bench.RAYS_PER_RANK, seed=1000)
for i in range(3):
    trainer(*bench._bundle(host[i].to(dev)))
field = model.field
pts = field._last_points
ms = [[ops.as_channel_last(p.detach()) for p in g] for g in field.grids]
flat = [p for g in ms for p in g]
K, NP = len(ms), len(ms[0])
c, m = flat[0].shape[1], pts.M
gout = torch.randn(m, K * c, device=dev)
grads = [torch.zeros_like(q) for q in flat]
flush = torch.empty(160 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
ps = pts.struct()

def scatter(groups):
    for grp in groups:
        tg = [g if (i // NP) in grp else None for i, g in enumerate(grads)]
        _lib.call("kp_hexplane_bwd", ops._plane_ptrs(flat), ops._plane_ptrs(tg), ops._plane_hw(flat), K, NP, c, ps, m, 1, 0x3F,
                  c_void_p(gout.data_ptr()), _lib.stream_ptr())

def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

for name, groups in (("all", [range(6)]), ("0-3|4-5", [range(4), (4, 5)]), ("0-3|4|5", [range(4), (4,), (5,)]),
                     ("0-2|3|4|5", [range(3), (3,), (4,), (5,)]), ("5|4|0-3", [(5,), (4,), range(4)])):
    print(f"scatter {name:12s} {timed(lambda: scatter(groups)):.3f} ms", flush=True)
