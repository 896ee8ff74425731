"""Model base class and config plumbing.  Mirror of NS/models/base_model.py:38-217 and of the two config classes
K-Planes needs from NS/configs/base_config.py:33-58 (``PrintableConfig`` / ``InstantiateConfig``)."""
from __future__ import annotations

from abc import abstractmethod
from collections import defaultdict
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple, Type

import torch
from torch import nn
from torch.nn import Parameter

from ..cameras.rays import RayBundle
from ..data.scene_box import SceneBox
from ..model_components.scene_colliders import NearFarCollider


class PrintableConfig:
    def __str__(self):
        lines = [self.__class__.__name__ + ":"]
        for key, val in vars(self).items():
            if isinstance(val, tuple):
                val = "[" + "\n".join(str(v) for v in val) + "]"
            lines += f"{key}: {str(val)}".split("\n")
        return "\n    ".join(lines)


@dataclass
class InstantiateConfig(PrintableConfig):
    """``_target(self, **kwargs)`` instantiation contract (base_config.py:50-58)."""

    _target: Type

    def setup(self, **kwargs) -> Any:
        return self._target(self, **kwargs)


def to_immutable_dict(d: Dict[str, Any]):
    return field(default_factory=lambda: dict(d))


@dataclass
class ModelConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: Model)
    enable_collider: bool = True
    collider_params: Optional[Dict[str, float]] = to_immutable_dict({"near_plane": 2.0, "far_plane": 6.0})
    loss_coefficients: Dict[str, float] = to_immutable_dict({"rgb_loss_coarse": 1.0, "rgb_loss_fine": 1.0})
    eval_num_rays_per_chunk: int = 4096


class Model(nn.Module):
    config: ModelConfig

    def __init__(self, config: ModelConfig, scene_box: SceneBox, num_train_data: int, **kwargs) -> None:
        super().__init__()
        self.config = config
        self.scene_box = scene_box
        self.render_aabb = None
        self.num_train_data = num_train_data
        self.kwargs = kwargs
        self.collider = None
        self.populate_modules()
        self.callbacks = None
        self.device_indicator_param = nn.Parameter(torch.empty(0))

    @property
    def device(self):
        return self.device_indicator_param.device

    def get_training_callbacks(self, training_callback_attributes) -> List:
        return []

    def populate_modules(self):
        if self.config.enable_collider:
            self.collider = NearFarCollider(near_plane=self.config.collider_params["near_plane"],
                                            far_plane=self.config.collider_params["far_plane"])

    @abstractmethod
    def get_param_groups(self) -> Dict[str, List[Parameter]]:
        """parameter groups for the optimizers."""

    @abstractmethod
    def get_outputs(self, ray_bundle: RayBundle) -> Dict[str, torch.Tensor]:
        """ray bundle -> outputs."""

    def forward(self, ray_bundle: RayBundle) -> Dict[str, torch.Tensor]:
        if self.collider is not None:
            ray_bundle = self.collider(ray_bundle)
        return self.get_outputs(ray_bundle)

    def get_metrics_dict(self, outputs, batch) -> Dict[str, torch.Tensor]:
        return {}

    @abstractmethod
    def get_loss_dict(self, outputs, batch, metrics_dict=None) -> Dict[str, torch.Tensor]:
        """losses."""

    @torch.no_grad()
    def get_outputs_for_camera_ray_bundle(self, camera_ray_bundle: RayBundle) -> Dict[str, torch.Tensor]:
        """Full-image inference in ``eval_num_rays_per_chunk`` chunks (base_model.py:162-186)."""
        num_rays_per_chunk = self.config.eval_num_rays_per_chunk
        image_height, image_width = camera_ray_bundle.origins.shape[:2]
        num_rays = len(camera_ray_bundle)
        outputs_lists = defaultdict(list)
        for i in range(0, num_rays, num_rays_per_chunk):
            ray_bundle = camera_ray_bundle.get_row_major_sliced_ray_bundle(i, i + num_rays_per_chunk)
            outputs = self.forward(ray_bundle=ray_bundle)
            for output_name, output in outputs.items():
                outputs_lists[output_name].append(output)
        outputs = {}
        for output_name, outputs_list in outputs_lists.items():
            if not torch.is_tensor(outputs_list[0]):
                continue
            outputs[output_name] = torch.cat(outputs_list).view(image_height, image_width, -1)
        return outputs

    def get_image_metrics_and_images(self, outputs, batch) -> Tuple[Dict[str, float], Dict[str, torch.Tensor]]:
        raise NotImplementedError

    def load_model(self, loaded_state: Dict[str, Any]) -> None:
        state = {key.replace("module.", ""): value for key, value in loaded_state["model"].items()}
        self.load_state_dict(state)
