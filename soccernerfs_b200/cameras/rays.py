"""Ray data structures crossing the drop-in boundary: Frustums, RaySamples, RayBundle.

Mirror of NS/cameras/rays.py:31-277 (same field names, shapes and methods).  ``RaySamples.get_weights``
(rays.py:127-149) is the alpha-compositing weight computation and runs on the warp-per-ray CUDA kernel.
"""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import Callable, Dict, Optional

import torch

from ..utils.tensor_dataclass import TensorDataclass


@dataclass
class Frustums(TensorDataclass):
    origins: torch.Tensor  # [..., 3]
    directions: torch.Tensor  # [..., 3]
    starts: torch.Tensor  # [..., 1]
    ends: torch.Tensor  # [..., 1]
    pixel_area: torch.Tensor  # [..., 1]
    offsets: Optional[torch.Tensor] = None  # [..., 3]

    def get_positions(self) -> torch.Tensor:
        """Frustum centre: origins + directions * (starts + ends) / 2 (+ offsets).  rays.py:48-57."""
        pos = self.origins + self.directions * (self.starts + self.ends) / 2
        if self.offsets is not None:
            pos = pos + self.offsets
        return pos

    def get_start_positions(self) -> torch.Tensor:
        return self.origins + self.directions * self.starts

    def set_offsets(self, offsets):
        self.offsets = offsets

    @classmethod
    def get_mock_frustum(cls, device="cpu") -> "Frustums":
        one3, one1 = torch.ones((1, 3), device=device), torch.ones((1, 1), device=device)
        return Frustums(origins=one3, directions=one3.clone(), starts=one1, ends=one1.clone(), pixel_area=one1.clone())


@dataclass
class RaySamples(TensorDataclass):
    frustums: Frustums
    camera_indices: Optional[torch.Tensor] = None  # [..., 1]
    deltas: Optional[torch.Tensor] = None  # [..., 1]
    spacing_starts: Optional[torch.Tensor] = None  # [..., S, 1]
    spacing_ends: Optional[torch.Tensor] = None  # [..., S, 1]
    spacing_to_euclidean_fn: Optional[Callable] = None
    metadata: Optional[Dict[str, torch.Tensor]] = None
    times: Optional[torch.Tensor] = None  # [..., 1]

    def get_weights(self, densities: torch.Tensor) -> torch.Tensor:
        """Compositing weights alpha_i * T_i with T = exp(-cumsum(delta*sigma)); [..., S, 1] -> [..., S, 1]."""
        from .. import ops

        shape = densities.shape
        s = shape[-2]
        w = ops.get_weights(self.deltas.reshape(-1, s), densities.reshape(-1, s))
        return w.view(shape)

    @staticmethod
    def get_weights_and_transmittance_from_alphas(alphas: torch.Tensor, weights_only: bool = False):
        """rays.py:151-170 (SDF models; plain torch, not on the K-Planes path)."""
        transmittance = torch.cumprod(
            torch.cat([torch.ones((*alphas.shape[:1], 1, 1), device=alphas.device), 1.0 - alphas + 1e-7], 1), 1
        )
        weights = alphas * transmittance[:, :-1, :]
        return weights if weights_only else (weights, transmittance)


@dataclass
class RayBundle(TensorDataclass):
    origins: torch.Tensor  # [..., 3]
    directions: torch.Tensor  # [..., 3]
    pixel_area: torch.Tensor  # [..., 1]
    camera_indices: Optional[torch.Tensor] = None
    nears: Optional[torch.Tensor] = None
    fars: Optional[torch.Tensor] = None
    metadata: Optional[Dict[str, torch.Tensor]] = None
    times: Optional[torch.Tensor] = None

    def set_camera_indices(self, camera_index: int) -> None:
        self.camera_indices = torch.ones_like(self.origins[..., 0:1]).long() * camera_index

    def __len__(self) -> int:
        return torch.numel(self.origins) // self.origins.shape[-1]

    def sample(self, num_rays: int) -> "RayBundle":
        assert num_rays <= len(self)
        return self[random.sample(range(len(self)), k=num_rays)]

    def get_row_major_sliced_ray_bundle(self, start_idx: int, end_idx: int) -> "RayBundle":
        return self.flatten()[start_idx:end_idx]

    def get_ray_samples(self, bin_starts, bin_ends, spacing_starts=None, spacing_ends=None,
                        spacing_to_euclidean_fn: Optional[Callable] = None, deltas=None) -> RaySamples:
        """Frustums for bins [bin_starts, bin_ends] ([..., S, 1]) along every ray.  rays.py:233-277.
        ``deltas`` (extension): bin_ends - bin_starts when the sampler kernel already produced it."""
        shaped = self[..., None]
        frustums = Frustums(origins=shaped.origins, directions=shaped.directions, starts=bin_starts, ends=bin_ends,
                            pixel_area=shaped.pixel_area)
        return RaySamples(
            frustums=frustums,
            camera_indices=None if self.camera_indices is None else self.camera_indices[..., None],
            deltas=bin_ends - bin_starts if deltas is None else deltas,
            spacing_starts=spacing_starts,
            spacing_ends=spacing_ends,
            spacing_to_euclidean_fn=spacing_to_euclidean_fn,
            metadata=shaped.metadata,
            times=None if self.times is None else self.times[..., None],
        )
