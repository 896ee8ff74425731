"""IST / ISG datamanager options (the surface `ns-train k-planes --pipeline.datamanager.*` exposes).

Field names, types and defaults of ``DynamicDataManagerConfig`` (NS/data/datamanagers/dynamic_datamanager.py:34-59;
preset overrides at NS/configs/method_configs.py:491-511).  The reference's own ``DynamicDataManager`` (image cache,
dataparsers, ray generator) is a caller of the hot path and is used unchanged when this package is plugged into
nerfstudio; ``importance_weights`` / ``make_pixel_sampler`` below give the same behaviour to users of
this package alone (synthetic data, tests).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Literal, Optional

import torch

from ..dynamic_dataset import compute_isg, compute_ist
from ..pixel_samplers import DynamicBasedPixelSampler, PixelSampler


@dataclass
class DynamicDataManagerConfig:
    """The dynamic-scene options of the reference; everything else of VanillaDataManagerConfig is out of scope."""

    train_num_rays_per_batch: int = 1024
    eval_num_rays_per_batch: int = 1024
    use_importance_sampling: bool = True
    """Whether to use importance sampling for the dynamic dataset."""
    is_pixel_ratio: float = 0.1
    """The ratio of pixels to sample using importance sampling per iteration."""
    ist_range: float = 0.25
    """The range of time differences to use for importance sampling."""
    isg: bool = False
    """Use ISG (true) or IST (False)."""
    isg_gamma: float = 5e-2
    """ISG gamma."""
    iters_to_start_is: int = 5000
    """Iterations before starting IST sampling."""
    pick_mode: Literal["normal", "randsteps", "lowfps"] = "normal"
    """Method for picking images in random loading."""


class ImportanceState:
    """What ``DynamicBasedPixelSampler`` reads from the dataset object (dynamic_dataset.py:60-68)."""

    def __init__(self, config: DynamicDataManagerConfig) -> None:
        self.use_importance_sampling = config.use_importance_sampling
        self.is_pixel_ratio = config.is_pixel_ratio
        self.ist_range = config.ist_range
        self.isg = config.isg
        self.isg_gamma = config.isg_gamma
        self.iters_to_start_ist = config.iters_to_start_is
        self.pick_mode = config.pick_mode


def importance_weights(config: DynamicDataManagerConfig, images: torch.Tensor, cam_ids: torch.Tensor,
                       cam_times: torch.Tensor, device="cpu") -> Optional[torch.Tensor]:
    """fp16 [B,H,W] weight maps for a cached image batch: ISG if ``config.isg`` else IST (dynamic_dataset.py:97-110)."""
    if not config.use_importance_sampling:
        return None
    if config.isg:
        return compute_isg(images, cam_ids, config.isg_gamma, device=device)
    return compute_ist(images, cam_ids, cam_times, config.ist_range, device=device)


def make_pixel_sampler(config: DynamicDataManagerConfig, num_rays_per_batch: int, **kwargs: Any) -> PixelSampler:
    """DynamicDataManager._get_pixel_sampler (dynamic_datamanager.py:97-113) without the patch / equirect cases."""
    if config.use_importance_sampling:
        return DynamicBasedPixelSampler(num_rays_per_batch, dataset=ImportanceState(config), **kwargs)
    return PixelSampler(num_rays_per_batch, **kwargs)
