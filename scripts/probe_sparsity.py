"""How sparse is the gradient reaching the proposal density fields? (decides whether zero-skipping pays)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from soccernerfs_b200 import ops
from soccernerfs_b200.data.scene_box import SceneBox
from soccernerfs_b200.data.synthetic import perturb_time_planes, synthetic_rays
from soccernerfs_b200.engine.trainer import TrainStep
from soccernerfs_b200.models.kplanes import KPlanesModelConfig
from soccernerfs_b200.cameras.rays import RayBundle

dev = "cuda"
gen = torch.Generator().manual_seed(0)
o, d, t, aabb = synthetic_rays(4096, gen)
model = KPlanesModelConfig().setup(scene_box=SceneBox(aabb=aabb), num_train_data=1).to(dev)
perturb_time_planes(model)
model.proposal_sampler.update_sched = lambda s: 0
step = TrainStep(model)
orig = ops._DensityField.backward
stats = []
def spy(ctx, g):
    z = (g == 0)
    w = z.view(-1, 32).all(dim=1)
    stats.append((g.numel(), float(z.float().mean()), float(w.float().mean())))
    return orig(ctx, g)
ops._DensityField.backward = staticmethod(spy)
img = torch.rand(4096, 3, generator=gen).to(dev)
for i in range(40):
    o, d, t, _ = synthetic_rays(4096, gen)
    rb = RayBundle(origins=o.to(dev), directions=d.to(dev), pixel_area=torch.ones(4096, 1, device=dev), times=t.to(dev))
    out = step(rb, {"image": img})
    if i in (0, 1, 5, 10, 20, 39):
        print(i, float(out["loss"]), stats[-2:], flush=True)
