"""Pixel samplers: which (image, y, x) triples make up a ray batch.

``PixelSampler.sample_method`` / ``collate_image_dataset_batch`` (NS/data/pixel_samplers.py:51-128) and the
importance sampler ``DynamicBasedPixelSampler.sample_method`` (:340-426).  Bit-exactness with the reference comes from
making the SAME random calls in the SAME order (``random.shuffle``, ``torch.multinomial`` per image, ``torch.rand`` for
the uniform remainder), so a shared seed reproduces the reference's indices exactly.  Host-side by design (SURVEY.md
8a, row a18).

The importance sampler also has a DEVICE path (``kp_importance_pixels``), taken when the weight maps
live on the GPU: the host keeps the reference's control flow (shuffled image order, all-zero maps skipped, k per image) and
the per-image ``torch.multinomial`` calls -- ~200 ms of host time per 4096-ray step at 540x960 -- become six kernel
launches for the whole step without a host synchronisation.  It draws from the same distribution with its own
counter-based random stream, so it reproduces the reference's statistics, not its exact indices.
"""
from __future__ import annotations

import random
from math import floor
from typing import Dict, Optional, Union

import torch


_EXTENTS: Dict = {}


def _extent(values, device) -> torch.Tensor:
    """``torch.tensor([num_images, H, W], device=device)``, cached: building it from a Python list is a blocking
    host-to-device copy, i.e. a stream synchronisation per step once the sampler runs on the GPU."""
    key = (tuple(int(v) for v in values), str(device))
    t = _EXTENTS.get(key)
    if t is None:
        if len(_EXTENTS) > 64:
            _EXTENTS.clear()
        t = _EXTENTS[key] = torch.tensor(list(key[0]), device=device)
    return t


class PixelSampler:
    def __init__(self, num_rays_per_batch: int, keep_full_image: bool = False, **kwargs) -> None:
        self.kwargs = kwargs
        self.num_rays_per_batch = num_rays_per_batch
        self.keep_full_image = keep_full_image

    def set_num_rays_per_batch(self, num_rays_per_batch: int):
        self.num_rays_per_batch = num_rays_per_batch

    def sample_method(self, batch_size: int, num_images: int, image_height: int, image_width: int,
                      mask: Optional[torch.Tensor] = None, batch: Optional[Dict] = None,
                      device: Union[torch.device, str] = "cpu") -> torch.Tensor:
        """Uniform over all pixels of all images (or over the mask's non-zero pixels) -> int64 [batch_size, 3]."""
        if isinstance(mask, torch.Tensor):
            nonzero_indices = torch.nonzero(mask[..., 0], as_tuple=False)
            chosen = random.sample(range(len(nonzero_indices)), k=batch_size)
            return nonzero_indices[chosen]
        return torch.floor(
            torch.rand((batch_size, 3), device=device) * _extent((num_images, image_height, image_width), device)
        ).long()

    def collate_image_dataset_batch(self, batch: Dict, num_rays_per_batch: int, keep_full_image: bool = False):
        device = batch["image"].device
        num_images, image_height, image_width, _ = batch["image"].shape
        kwargs = dict(batch=batch, device=device)
        if "mask" in batch:
            kwargs["mask"] = batch["mask"]
        indices = self.sample_method(num_rays_per_batch, num_images, image_height, image_width, **kwargs)
        c, y, x = (i.flatten() for i in torch.split(indices, 1, dim=-1))
        collated = {k: v[c, y, x] for k, v in batch.items() if k not in ("image_idx", "iter_steps") and v is not None}
        assert collated["image"].shape == (num_rays_per_batch, 3), collated["image"].shape
        indices[:, 0] = batch["image_idx"][c]  # random image slots -> absolute camera indices
        collated["indices"] = indices
        if keep_full_image:
            collated["full_image"] = batch["image"]
        return collated

    def sample(self, image_batch: Dict):
        if isinstance(image_batch["image"], torch.Tensor):
            return self.collate_image_dataset_batch(image_batch, self.num_rays_per_batch, keep_full_image=self.keep_full_image)
        raise ValueError("image_batch['image'] must be a torch.Tensor")


class EquirectangularPixelSampler(PixelSampler):
    """Uniform over the SPHERE for equirectangular images (pixel_samplers.py:228-267): rows follow the density
    sin(phi) / 2 by inverse-transform sampling, columns and images are uniform.  Three ``torch.rand(batch_size)`` draws in
    the reference's order (image, row, column).  With a mask the reference falls back to the base sampler."""

    def sample_method(self, batch_size: int, num_images: int, image_height: int, image_width: int,
                      mask: Optional[torch.Tensor] = None, batch: Optional[Dict] = None,
                      device: Union[torch.device, str] = "cpu") -> torch.Tensor:
        if isinstance(mask, torch.Tensor):
            return super().sample_method(batch_size, num_images, image_height, image_width, mask=mask, device=device)
        image_u = torch.rand(batch_size, device=device)
        row_u = torch.acos(1 - 2 * torch.rand(batch_size, device=device)) / torch.pi  # polar angle / pi in [0, 1]
        col_u = torch.rand(batch_size, device=device)
        extent = _extent((num_images, image_height, image_width), device)
        return torch.floor(torch.stack((image_u, row_u, col_u), dim=-1) * extent).long()


class PatchPixelSampler(PixelSampler):
    """Square patches for patch-based losses (pixel_samplers.py:270-327): ``batch_size // patch_size**2`` top-left corners
    drawn uniformly (one ``torch.rand((n, 3))``), every patch expanded row-major.  The ray count is rounded down to whole
    patches."""

    def __init__(self, num_rays_per_batch: int, keep_full_image: bool = False, **kwargs) -> None:
        self.patch_size = int(kwargs["patch_size"])
        per_patch = self.patch_size**2
        super().__init__(num_rays_per_batch // per_patch * per_patch, keep_full_image, **kwargs)

    def set_num_rays_per_batch(self, num_rays_per_batch: int):
        per_patch = self.patch_size**2
        self.num_rays_per_batch = num_rays_per_batch // per_patch * per_patch

    def sample_method(self, batch_size: int, num_images: int, image_height: int, image_width: int,
                      mask: Optional[torch.Tensor] = None, batch: Optional[Dict] = None,
                      device: Union[torch.device, str] = "cpu") -> torch.Tensor:
        if mask is not None:  # (the reference tests ``if mask:``; a mask means the base sampler)
            return super().sample_method(batch_size, num_images, image_height, image_width, mask=mask, device=device)
        p = self.patch_size
        n = batch_size // (p * p)
        extent = _extent((num_images, image_height - p, image_width - p), device)
        corners = torch.rand((n, 3), device=device) * extent
        offsets = torch.arange(p, device=device)
        grid = corners.view(n, 1, 1, 3).repeat(1, p, p, 1)
        grid[..., 1] += offsets.view(1, p, 1)
        grid[..., 2] += offsets.view(1, 1, p)
        return torch.floor(grid).long().reshape(-1, 3)


class DynamicBasedPixelSampler(PixelSampler):
    """Samples ``is_pixel_ratio`` of the batch proportionally to the IST/ISG weight maps, the rest uniformly."""

    def __init__(self, num_rays_per_batch: int, keep_full_image: bool = False, device_sampler: Optional[bool] = None,
                 **kwargs) -> None:
        """``device_sampler``: None = the device path whenever the weight maps are CUDA tensors; False = always the
        reference's host loop (bit-exact indices for a shared seed; on CUDA maps it synchronises per image like the
        reference would); True = require the device path."""
        self.dataset = kwargs["dataset"]
        self.device_sampler = device_sampler
        self._nonzero_maps = None  # (identity of the weight tensor, host bool [B]: which maps have a non-zero pixel)
        super().__init__(num_rays_per_batch, keep_full_image, **kwargs)

    def _maps_with_mass(self, weights: torch.Tensor) -> list:
        """Which images have a non-zero pixel (the reference skips the others, :391-393).  The maps are dataset-level
        constants, so this one reduction + read-back is cached until the tensor changes."""
        key = (weights.data_ptr(), weights._version, tuple(weights.shape))
        if self._nonzero_maps is None or self._nonzero_maps[0] != key:
            flags = (weights.reshape(weights.shape[0], -1) != 0).any(dim=1).cpu().tolist()
            self._nonzero_maps = (key, flags)
        return self._nonzero_maps[1]

    def _sample_ist_device(self, weights: torch.Tensor, num_ist: int, per_image: int, num_images: int,
                           image_width: int) -> torch.Tensor:
        """The importance part of the batch on the device -> int64 [sampled, 3] (image, row, col)."""
        from .. import ops

        has_mass = self._maps_with_mass(weights)
        order = list(range(num_images))
        random.shuffle(order)
        sel, sampled = [], 0
        for i in order:  # the reference's walk (:381-404) without its sampling
            if sampled >= num_ist:
                break
            k = per_image if sampled + per_image <= num_ist else num_ist - sampled
            if not has_mass[i]:
                continue
            sel.append((i, k, sampled))
            sampled += k
        if not sel:
            return torch.zeros((0, 3), dtype=torch.int64, device=weights.device)
        seed = random.getrandbits(63)  # from the same (seedable) python generator that shuffled the images
        # (a non-contiguous map tensor is copied here, AFTER the cached non-zero test above was keyed on the caller's tensor)
        return ops.importance_pixels(weights.contiguous(), torch.tensor(sel, dtype=torch.int32), max(k for _, k, _ in sel),
                                     sampled, image_width, seed)

    def sample_method(self, batch_size: int, num_images: int, image_height: int, image_width: int,
                      mask: Optional[torch.Tensor] = None, batch: Optional[Dict] = None,
                      device: Union[torch.device, str] = "cpu") -> torch.Tensor:
        assert batch is not None, "Batch information must be provided for DynamicBasedPixelSampler"
        if "ist_weights" not in batch or batch["ist_weights"] is None:
            return super().sample_method(batch_size, num_images, image_height, image_width, mask=mask, device=device)
        sampled = 0
        use_ist = batch["iter_steps"] > self.dataset.iters_to_start_ist and batch["ist_weights"] is not None
        on_device = use_ist and (self.device_sampler or (self.device_sampler is None and batch["ist_weights"].is_cuda))
        if on_device:
            weights = batch["ist_weights"]
            if not weights.is_cuda or weights.dtype != torch.float16:
                raise RuntimeError("device_sampler=True needs fp16 CUDA weight maps (there is no CPU path)")
            num_ist = floor(self.dataset.is_pixel_ratio * batch_size)
            per_image = 10 * (-(-num_ist // num_images))
            indices = self._sample_ist_device(weights, num_ist, per_image, num_images, image_width)
            uniform = super().sample_method(batch_size - indices.shape[0], num_images, image_height, image_width, mask=mask,
                                            device=weights.device)
            return torch.cat((indices, uniform.to(weights.device)), dim=0)
        if use_ist:
            num_ist = floor(self.dataset.is_pixel_ratio * batch_size)
            per_image = 10 * (-(-num_ist // num_images))
            weights = batch["ist_weights"]
            indices = torch.zeros((num_ist, 3), device=device)
            order = list(range(num_images))
            random.shuffle(order)  # usually only part of the images is needed: shuffle so all maps get used over time
            for i in order:
                if sampled >= num_ist:
                    break
                wmap = weights[i]
                k = per_image if sampled + per_image <= num_ist else num_ist - sampled
                nnz = len(torch.nonzero(wmap))
                if nnz == 0:  # camera that sees no motion
                    continue
                samples = torch.multinomial(wmap.flatten(), k, replacement=(nnz < k))
                h, w = torch.div(samples, image_width, rounding_mode="floor"), samples % image_width
                indices[sampled: sampled + k, 0] = i
                indices[sampled: sampled + k, 1] = h
                indices[sampled: sampled + k, 2] = w
                sampled += k
            if sampled < num_ist:
                indices = indices[:sampled]
        uniform = super().sample_method(batch_size - sampled, num_images, image_height, image_width, mask=mask, device=device)
        if use_ist:
            return torch.cat((indices, uniform), dim=0).long()
        return uniform
