"""Scene contraction for unbounded scenes (NS/field_components/spatial_distortions.py:28-88).

``KPlanesModel(bounded=False)`` contracts sample positions with ``SceneContraction(order=float("inf"))`` before the
planes are queried (NS/models/kplanes.py:203-206, NS/fields/kplanes_field.py:278-280).  On the kernel path the
contraction is evaluated inside the gather / scatter / proposal kernels (``KpPoints.norm_mode = 2``: the positions are
never materialised); ``forward`` is the plain tensor form for any other caller.  Only point inputs: the Gaussian
(mip-NeRF 360 covariance) branch of the reference is not on the K-Planes path.
"""
from __future__ import annotations

from typing import Optional, Union

import torch
from torch import nn


class SpatialDistortion(nn.Module):
    """Apply spatial distortions (spatial_distortions.py:28-39)."""

    def forward(self, positions: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError


class SceneContraction(SpatialDistortion):
    """f(x) = x if ||x|| < 1 else (2 - 1/||x||) (x / ||x||); ``order=float("inf")`` contracts to the cube [-2, 2]^3."""

    def __init__(self, order: Optional[Union[float, int]] = None) -> None:
        super().__init__()
        self.order = order

    def forward(self, positions: torch.Tensor) -> torch.Tensor:
        if not torch.is_tensor(positions):
            raise NotImplementedError("SceneContraction of Gaussians is not part of the K-Planes path")
        mag = torch.linalg.norm(positions, ord=self.order, dim=-1)[..., None]
        return torch.where(mag < 1, positions, (2 - (1 / mag)) * (positions / mag))
