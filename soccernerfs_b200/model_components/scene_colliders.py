"""Scene colliders: near/far per ray.  Mirror of NS/model_components/scene_colliders.py:28-110, 166-188.
``AABBBoxCollider`` runs the slab test in ``kp_aabb_intersect`` (same op order as the reference)."""
from __future__ import annotations

import torch
from torch import nn

from .. import ops
from ..cameras.rays import RayBundle
from ..data.scene_box import SceneBox


class SceneCollider(nn.Module):
    def __init__(self, **kwargs) -> None:
        self.kwargs = kwargs
        super().__init__()

    def set_nears_and_fars(self, ray_bundle) -> RayBundle:
        raise NotImplementedError

    def forward(self, ray_bundle: RayBundle) -> RayBundle:
        if ray_bundle.nears is not None and ray_bundle.fars is not None:
            return ray_bundle
        return self.set_nears_and_fars(ray_bundle)


class AABBBoxCollider(SceneCollider):
    def __init__(self, scene_box: SceneBox, near_plane: float = 0.0, **kwargs) -> None:
        super().__init__(**kwargs)
        self.scene_box = scene_box
        self.near_plane = near_plane
        self._aabb6 = tuple(float(v) for v in scene_box.aabb.detach().flatten().tolist())

    def _intersect_with_aabb(self, rays_o: torch.Tensor, rays_d: torch.Tensor, aabb: torch.Tensor = None):
        near_plane = self.near_plane if self.training else 0  # scene_colliders.py:88
        aabb6 = self._aabb6 if aabb is None else tuple(float(v) for v in aabb.detach().flatten().tolist())
        return ops.aabb_intersect(rays_o, rays_d, aabb6, near_plane)

    def set_nears_and_fars(self, ray_bundle: RayBundle) -> RayBundle:
        nears, fars = self._intersect_with_aabb(ray_bundle.origins, ray_bundle.directions)
        ray_bundle.nears = nears[..., None]
        ray_bundle.fars = fars[..., None]
        return ray_bundle


class NearFarCollider(SceneCollider):
    def __init__(self, near_plane: float, far_plane: float, **kwargs) -> None:
        self.near_plane = near_plane
        self.far_plane = far_plane
        super().__init__(**kwargs)

    def set_nears_and_fars(self, ray_bundle: RayBundle) -> RayBundle:
        ones = torch.ones_like(ray_bundle.origins[..., 0:1])
        near_plane = self.near_plane if self.training else 0
        ray_bundle.nears = ones * near_plane
        ray_bundle.fars = ones * self.far_plane
        return ray_bundle
