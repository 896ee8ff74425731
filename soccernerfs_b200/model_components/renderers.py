"""Renderers on the B200 compositing kernels.

Same classes / arguments / train-vs-eval behaviour as NS/model_components/renderers.py:
``RGBRenderer`` :58-140, ``AccumulationRenderer`` :197-223, ``DepthRenderer`` :226-287, ``MedianRGBRenderer`` :290-362,
and the global ``background_color_override_context`` :43-55 (used by ns-render's crop).  Packed samples
(``ray_indices``; nerfacc's volumetric sampler) never occur on the K-Planes / nerfplayer-nerfacto path and raise.
"""
from __future__ import annotations

import contextlib
from typing import Generator, Optional, Union

import torch
from torch import nn

from .. import ops
from ..cameras.rays import RaySamples

BACKGROUND_COLOR_OVERRIDE: Optional[torch.Tensor] = None

COLORS_DICT = {  # NS/utils/colors.py
    "white": torch.tensor([1.0, 1.0, 1.0]),
    "black": torch.tensor([0.0, 0.0, 0.0]),
    "red": torch.tensor([1.0, 0.0, 0.0]),
    "green": torch.tensor([0.0, 1.0, 0.0]),
    "blue": torch.tensor([0.0, 0.0, 1.0]),
}
_DEVICE_COLORS = {}


@contextlib.contextmanager
def background_color_override_context(mode: torch.Tensor) -> Generator[None, None, None]:
    """Context manager for setting the background colour globally (renderers.py:43-55)."""
    global BACKGROUND_COLOR_OVERRIDE  # pylint: disable=global-statement
    old = BACKGROUND_COLOR_OVERRIDE
    try:
        BACKGROUND_COLOR_OVERRIDE = mode
        yield
    finally:
        BACKGROUND_COLOR_OVERRIDE = old


def _no_packed(ray_indices, num_rays):
    if ray_indices is not None and num_rays is not None:
        raise NotImplementedError("packed samples (nerfacc ray_indices) are not on the K-Planes path and are not built")


class RGBRenderer(nn.Module):
    """Standard volumetric rendering: sum_i w_i rgb_i + background * (1 - sum_i w_i)."""

    def __init__(self, background_color: Union[str, torch.Tensor] = "random") -> None:
        super().__init__()
        self.background_color = background_color

    @classmethod
    def combine_rgb(cls, rgb: torch.Tensor, weights: torch.Tensor, background_color="random", ray_indices=None,
                    num_rays=None, nan_to_num: bool = False) -> torch.Tensor:
        _no_packed(ray_indices, num_rays)
        s = rgb.shape[-2]
        batch = rgb.shape[:-2]
        w2, rgb3 = weights.reshape(-1, s), rgb.reshape(-1, s, 3)
        if BACKGROUND_COLOR_OVERRIDE is not None:
            background_color = BACKGROUND_COLOR_OVERRIDE
        if isinstance(background_color, str) and background_color == "last_sample":
            return ops.composite_rgb(w2, rgb3, "last_sample", nan_to_num).view(*batch, 3)
        if isinstance(background_color, str) and background_color == "random":
            background_color = torch.rand((w2.shape[0], 3), dtype=torch.float32, device=rgb.device)  # renderers.py:104-105
        if isinstance(background_color, str) and background_color in COLORS_DICT:
            key = (background_color, str(rgb.device))
            if key not in _DEVICE_COLORS:  # cached per device: no H2D copy in the steady state (CUDA-graph safe)
                _DEVICE_COLORS[key] = COLORS_DICT[background_color].to(rgb.device)
            background_color = _DEVICE_COLORS[key]
        assert isinstance(background_color, torch.Tensor)
        bg = background_color.to(rgb.device).reshape(-1, 3)
        return ops.composite_rgb(w2, rgb3, bg, nan_to_num).view(*batch, 3)

    def forward(self, rgb: torch.Tensor, weights: torch.Tensor, ray_indices=None, num_rays=None) -> torch.Tensor:
        # eval: nan_to_num(rgb) before and clamp(0,1) after (renderers.py:133-139); nan_to_num is fused in the kernel
        out = self.combine_rgb(rgb, weights, background_color=self.background_color, ray_indices=ray_indices,
                               num_rays=num_rays, nan_to_num=not self.training)
        if not self.training:
            torch.clamp_(out, min=0.0, max=1.0)
        return out


class AccumulationRenderer(nn.Module):
    """Accumulated opacity along a ray."""

    @classmethod
    def forward(cls, weights: torch.Tensor, ray_indices=None, num_rays=None) -> torch.Tensor:
        _no_packed(ray_indices, num_rays)
        s = weights.shape[-2]
        return ops.accumulate(weights.reshape(-1, s)).view(*weights.shape[:-2], 1)


class DepthRenderer(nn.Module):
    """Median (default) or expected depth along a ray (renderers.py:226-287)."""

    def __init__(self, method: str = "median") -> None:
        super().__init__()
        self.method = method

    def forward(self, weights: torch.Tensor, ray_samples: RaySamples, ray_indices=None, num_rays=None) -> torch.Tensor:
        _no_packed(ray_indices, num_rays)
        s = weights.shape[-2]
        batch = weights.shape[:-2]
        if self.method == "median":  # (starts + ends)/2 gathered at the median index, in one kernel
            fr = ray_samples.frustums
            return ops.median_depth(weights.reshape(-1, s), fr.starts.reshape(-1, s), fr.ends.reshape(-1, s)).view(*batch, 1)
        steps = (ray_samples.frustums.starts + ray_samples.frustums.ends) / 2
        if self.method == "expected":
            if weights.requires_grad and torch.is_grad_enabled():  # differentiable variant: plain torch (not on the k-planes path)
                depth = torch.sum(weights * steps, dim=-2) / (torch.sum(weights, -2) + 1e-10)
            else:
                depth = ops.expected_depth(weights.reshape(-1, s), steps.reshape(-1, s)).view(*batch, 1)
            return torch.clip(depth, steps.min(), steps.max())  # global min/max, as the reference (:283)
        raise NotImplementedError(f"Method {self.method} not implemented")


class MedianRGBRenderer(nn.Module):
    """RGB of the sample where the accumulated weight reaches 0.5 (renderers.py:290-362)."""

    def __init__(self, background_color: Union[str, torch.Tensor] = "random") -> None:
        super().__init__()
        self.background_color = background_color

    @classmethod
    def combine_rgb(cls, rgb: torch.Tensor, weights: torch.Tensor, background_color="random", ray_indices=None,
                    num_rays=None) -> torch.Tensor:
        _no_packed(ray_indices, num_rays)
        s = weights.shape[-2]
        idx = ops.median_index(weights.reshape(-1, s)).view(*weights.shape[:-2], 1)
        idx = idx.unsqueeze(dim=2).expand(-1, -1, 3)
        return torch.gather(rgb, dim=-2, index=idx)

    def forward(self, rgb: torch.Tensor, weights: torch.Tensor, ray_indices=None, num_rays=None) -> torch.Tensor:
        if not self.training:
            rgb = torch.nan_to_num(rgb)
        out = self.combine_rgb(rgb, weights, background_color=self.background_color, ray_indices=ray_indices, num_rays=num_rays)
        if not self.training:
            torch.clamp_(out, min=0.0, max=1.0)
        return out
