"""Field base class.  Mirror of NS/fields/base_field.py:36-130 (the subset K-Planes uses)."""
from __future__ import annotations

from abc import abstractmethod
from enum import Enum
from typing import Dict, Optional, Tuple

import torch
from torch import nn

from ..cameras.rays import Frustums, RaySamples


class FieldHeadNames(Enum):
    """NS/field_components/field_heads.py:28-43."""

    RGB = "rgb"
    SH = "sh"
    DENSITY = "density"
    NORMALS = "normals"
    PRED_NORMALS = "pred_normals"
    UNCERTAINTY = "uncertainty"
    TRANSIENT_RGB = "transient_rgb"
    TRANSIENT_DENSITY = "transient_density"
    SEMANTICS = "semantics"
    SDF = "sdf"
    ALPHA = "alpha"
    GRADIENT = "gradient"
    PROBS = "probs"


class Field(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self._sample_locations = None
        self._density_before_activation = None

    def density_fn(self, positions: torch.Tensor) -> torch.Tensor:
        ray_samples = RaySamples(
            frustums=Frustums(
                origins=positions,
                directions=torch.ones_like(positions),
                starts=torch.zeros_like(positions[..., :1]),
                ends=torch.zeros_like(positions[..., :1]),
                pixel_area=torch.ones_like(positions[..., :1]),
            )
        )
        density, _ = self.get_density(ray_samples)
        return density

    @abstractmethod
    def get_density(self, ray_samples: RaySamples) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (density [..., 1], features)."""

    @abstractmethod
    def get_outputs(self, ray_samples: RaySamples, density_embedding: Optional[torch.Tensor] = None):
        """-> field outputs."""

    def forward(self, ray_samples: RaySamples, compute_normals: bool = False) -> Dict[FieldHeadNames, torch.Tensor]:
        density, density_embedding = self.get_density(ray_samples)
        field_outputs = self.get_outputs(ray_samples, density_embedding=density_embedding)
        field_outputs[FieldHeadNames.DENSITY] = density
        return field_outputs
