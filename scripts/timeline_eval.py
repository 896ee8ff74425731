"""Kernel timeline of full-frame inference (BASELINE config 5) through torch.profiler/CUPTI: per-kernel totals of one
frame -> stdout, so that the eval leg's time can be attributed (nsys is absent)."""
import argparse, collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from torch.profiler import profile, ProfilerActivity

ap = argparse.ArgumentParser()
ap.add_argument("--ray-tile", type=int, default=4)
args = ap.parse_args()
from soccernerfs_b200.cameras.cameras import Cameras
from soccernerfs_b200.engine.frame_renderer import FrameRenderer

dev = torch.device("cuda", 0)
model = bench.build_model("cfg3", dev)
model.eval()
h, w = 1080, 1920
pos = torch.tensor([[1.0, 0.0, 0.35]])
fwd = -pos / pos.norm(dim=-1, keepdim=True)
right = torch.cross(fwd, torch.tensor([[0.0, 0.0, 1.0]]), dim=-1)
right = right / right.norm(dim=-1, keepdim=True)
up = torch.cross(right, fwd, dim=-1)
c2w = torch.cat([torch.stack([right, up, -fwd], dim=-1), pos[..., None]], dim=-1)
cams = Cameras(c2w.to(dev), 1600.0, 1600.0, w / 2, h / 2, w, h, times=torch.tensor([0.5]).to(dev))
r = FrameRenderer(model, cams, ray_tile=args.ray_tile)
r.render(0)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    r.render(0)
    torch.cuda.synchronize()
tot = collections.Counter()
cnt = collections.Counter()
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        tot[e.name[:70]] += e.time_range.end - e.time_range.start
        cnt[e.name[:70]] += 1
s = sum(tot.values())
print(f"ray_tile {args.ray_tile}: kernels {sum(cnt.values())}, busy {s / 1000:.1f} ms")
for k, v in tot.most_common(14):
    print(f"{v / 1000:8.2f} ms {cnt[k]:5d}  {k}")
