"""Synthetic ray batches of the reference's scene shapes -- no dataset exists on the GPU box.

"broadcast": 19 train cameras on a ring around the origin looking inward, aabb [-1.5,1.5]^3
(NS/data/dataparsers/broadcaststyle_dataparser.py:196-232, :449-463), 100 frames, fps_downsample 4 ->
frame ids linspace(0,99,25).int(), times = id/99.  "stadium": 30 cameras, aabb [-1,1]^3
(stadiumwide_dataparser.py:83-112).  Same generator as oracle/kplanes_oracle.py::synthetic_rays.
"""
from __future__ import annotations

import torch


def synthetic_rays(n: int, gen: torch.Generator, scene: str = "broadcast", n_frames: int = 25):
    scale = 1.5 if scene == "broadcast" else 1.0
    n_cams = 19 if scene == "broadcast" else 30
    aabb = torch.tensor([[-scale] * 3, [scale] * 3])
    cam = torch.randint(0, n_cams, (n,), generator=gen)
    ang = cam.float() / n_cams * 6.283185307179586
    origins = torch.stack([torch.cos(ang), torch.sin(ang), torch.full_like(ang, 0.35)], dim=-1)
    target = (torch.rand(n, 3, generator=gen) - 0.5) * torch.tensor([1.6, 1.6, 0.6])
    d = target - origins
    directions = d / d.norm(dim=-1, keepdim=True)
    frame_ids = torch.linspace(0, 99, n_frames).to(torch.int32).float()
    times = (frame_ids[torch.randint(0, n_frames, (n,), generator=gen)] / 99.0)[:, None]
    return origins, directions, times, aabb


def perturb_time_planes(model, std: float = 0.05, seed: int = 0) -> None:
    """N(0,std) noise on the space-time planes so products / gradients are non-degenerate (they start at 1)."""
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for grids in list(model.field.grids) + [p.grids for p in model.proposal_networks]:
            if len(grids) == 6:
                for i in (2, 4, 5):
                    g = grids[i]
                    g.add_((std * torch.randn(g.shape, generator=gen)).to(g.device))
