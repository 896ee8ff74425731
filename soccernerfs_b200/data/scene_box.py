"""SceneBox: the scene's axis-aligned bounding box.  The members of NS/data/scene_box.py:26-107 that the K-Planes path
and its callers use, same names and results."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Union

import torch


@dataclass
class SceneBox:
    aabb: torch.Tensor = None  # [2,3]: row 0 = minimum corner, row 1 = maximum corner

    def _extent(self) -> torch.Tensor:
        lo, hi = self.aabb
        return hi - lo

    def get_diagonal_length(self):
        return torch.sqrt(self._extent().pow(2).sum() + 1e-20)

    def get_center(self):
        return self.aabb[0] + self._extent() / 2.0

    def get_centered_and_scaled_scene_box(self, scale_factor: Union[float, torch.Tensor] = 1.0):
        centred = self.aabb - self.get_center()
        return SceneBox(aabb=centred * scale_factor)

    @staticmethod
    def get_normalized_positions(positions: torch.Tensor, aabb: torch.Tensor):
        """World positions -> [0,1]^3 inside ``aabb`` (scene_box.py:56-66): (p - min) / (max - min), in that order."""
        lo, hi = aabb[0], aabb[1]
        return (positions - lo) / (hi - lo)

    def to_json(self) -> Dict:
        lo, hi = self.aabb.tolist()
        return {"type": "aabb", "min_point": lo, "max_point": hi}

    @staticmethod
    def from_json(json_: Dict) -> "SceneBox":
        if json_["type"] != "aabb":
            raise AssertionError("only aabb scene boxes exist")
        return SceneBox(aabb=torch.tensor([json_["min_point"], json_["max_point"]]))

    @staticmethod
    def from_camera_poses(poses: torch.Tensor, scale_factor: float) -> "SceneBox":
        centres = poses[..., :3, -1]
        corners = torch.stack([centres.min(dim=0).values, centres.max(dim=0).values])
        return SceneBox(aabb=scale_factor * corners)
