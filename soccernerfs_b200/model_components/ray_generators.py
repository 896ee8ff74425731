"""RayGenerator (NS/model_components/ray_generators.py:25-59): (camera,row,col) pixel indices -> RayBundle, on the
device.  The pose optimiser of the reference is a caller-side component (camera_optimizers.py) and is accepted only
in its "off" mode (the K-Planes preset's default)."""
from __future__ import annotations

import torch
from torch import nn

from ..cameras.cameras import Cameras
from ..cameras.rays import RayBundle


class RayGenerator(nn.Module):
    def __init__(self, cameras: Cameras, pose_optimizer=None) -> None:
        super().__init__()
        self.cameras = cameras
        if pose_optimizer is not None and getattr(getattr(pose_optimizer, "config", None), "mode", "off") != "off":
            raise NotImplementedError("camera pose optimisation is not built (mode must be 'off')")
        self.pose_optimizer = pose_optimizer
        self.register_buffer("image_coords", cameras.get_image_coords(), persistent=False)

    def forward(self, ray_indices: torch.Tensor) -> RayBundle:
        """ray_indices [num_rays, 3] = camera, row, col."""
        return self.cameras.generate_rays_from_indices(ray_indices)
