"""SceneBox: axis-aligned bounding box of the scene.  Mirror of NS/data/scene_box.py:26-107 (subset used by K-Planes)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Union

import torch


@dataclass
class SceneBox:
    aabb: torch.Tensor = None  # [2,3]: min xyz, max xyz

    def get_diagonal_length(self):
        diff = self.aabb[1] - self.aabb[0]
        return torch.sqrt((diff**2).sum() + 1e-20)

    def get_center(self):
        return self.aabb[0] + (self.aabb[1] - self.aabb[0]) / 2.0

    def get_centered_and_scaled_scene_box(self, scale_factor: Union[float, torch.Tensor] = 1.0):
        return SceneBox(aabb=(self.aabb - self.get_center()) * scale_factor)

    @staticmethod
    def get_normalized_positions(positions: torch.Tensor, aabb: torch.Tensor):
        """[0,1] positions inside the aabb (scene_box.py:56-66)."""
        return (positions - aabb[0]) / (aabb[1] - aabb[0])

    def to_json(self) -> Dict:
        return {"type": "aabb", "min_point": self.aabb[0].tolist(), "max_point": self.aabb[1].tolist()}

    @staticmethod
    def from_json(json_: Dict) -> "SceneBox":
        assert json_["type"] == "aabb"
        return SceneBox(aabb=torch.tensor([json_["min_point"], json_["max_point"]]))

    @staticmethod
    def from_camera_poses(poses: torch.Tensor, scale_factor: float) -> "SceneBox":
        xyzs = poses[..., :3, -1]
        return SceneBox(aabb=scale_factor * torch.stack([torch.min(xyzs, dim=0)[0], torch.max(xyzs, dim=0)[0]]))
