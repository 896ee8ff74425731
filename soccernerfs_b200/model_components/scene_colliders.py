"""Per-ray near / far planes.  Same classes and call surface as NS/model_components/scene_colliders.py:28-110, 166-188:
``collider(ray_bundle)`` fills ``ray_bundle.nears`` / ``.fars`` unless both are already present.
``AABBBoxCollider`` runs the slab test in ``kp_aabb_intersect`` (the reference's operation order, so the planes are
bit-identical to its torch path); ``NearFarCollider`` is two fills."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import nn

from .. import ops
from ..cameras.rays import RayBundle
from ..data.scene_box import SceneBox


def _as_six_floats(aabb: torch.Tensor) -> Tuple[float, ...]:
    return tuple(float(v) for v in aabb.detach().reshape(-1).tolist())


class SceneCollider(nn.Module):
    """Base class: subclasses implement ``set_nears_and_fars``."""

    def __init__(self, **kwargs) -> None:
        self.kwargs = kwargs
        super().__init__()

    def set_nears_and_fars(self, ray_bundle) -> RayBundle:
        raise NotImplementedError

    def forward(self, ray_bundle: RayBundle) -> RayBundle:
        already_set = ray_bundle.nears is not None and ray_bundle.fars is not None
        return ray_bundle if already_set else self.set_nears_and_fars(ray_bundle)


class AABBBoxCollider(SceneCollider):
    """Ray / scene-box intersection; the training-time ``near_plane`` is dropped to 0 in eval mode (:88)."""

    def __init__(self, scene_box: SceneBox, near_plane: float = 0.0, **kwargs) -> None:
        super().__init__(**kwargs)
        self.scene_box, self.near_plane = scene_box, near_plane
        self._box_key, self._box = None, None

    def _host_box(self):
        """Host copy of the scene box (a launch never waits on a device read), refreshed when ``scene_box.aabb`` is
        replaced or modified in place (e.g. a render crop)."""
        aabb = self.scene_box.aabb
        key = (aabb.data_ptr(), aabb._version)
        if key != self._box_key:
            self._box_key, self._box = key, _as_six_floats(aabb)
        return self._box

    def _intersect_with_aabb(self, rays_o: torch.Tensor, rays_d: torch.Tensor, aabb: Optional[torch.Tensor] = None):
        box = self._host_box() if aabb is None else _as_six_floats(aabb)
        return ops.aabb_intersect(rays_o, rays_d, box, self.near_plane if self.training else 0)

    def set_nears_and_fars(self, ray_bundle: RayBundle) -> RayBundle:
        t_near, t_far = self._intersect_with_aabb(ray_bundle.origins, ray_bundle.directions)
        ray_bundle.nears, ray_bundle.fars = t_near.unsqueeze(-1), t_far.unsqueeze(-1)
        return ray_bundle


class NearFarCollider(SceneCollider):
    """Constant planes for every ray (:166-188); like the box collider the near plane is 0 outside training."""

    def __init__(self, near_plane: float, far_plane: float, **kwargs) -> None:
        self.near_plane, self.far_plane = near_plane, far_plane
        super().__init__(**kwargs)

    def set_nears_and_fars(self, ray_bundle: RayBundle) -> RayBundle:
        like = ray_bundle.origins[..., :1]
        ray_bundle.nears = torch.full_like(like, float(self.near_plane if self.training else 0))
        ray_bundle.fars = torch.full_like(like, float(self.far_plane))
        return ray_bundle
