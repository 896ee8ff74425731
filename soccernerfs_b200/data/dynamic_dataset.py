"""Importance-sampling weight maps of the dynamic dataset: IST (temporal difference) and ISG (global median).

Arithmetic of ``DynamicDataset.compute_ist`` (NS/data/datasets/dynamic_dataset.py:328-470) and ``compute_isg``
(:215-326) on an image batch, without the file caching / tqdm / debug-map code around it.  Host-side torch ops like the reference's (it runs them once per image-cache reload, not per training step); both maps
also have a device path (``kp_ist_map`` / ``kp_isg_map``) used when the images are on the GPU.
"""
from __future__ import annotations

import torch

IST_ALPHA = 0.15  # dynamic_dataset.py:420: differences below this are camera shake / noise


def temporal_neighbours(cam_ids: torch.Tensor, cam_times: torch.Tensor, ist_range: float):
    """Per image the indices of the SAME camera's images whose time differs by (0.01, ist_range]
    (dynamic_dataset.py:398-407), as CSR lists (offsets int32 [B+1], neighbours int32) on the host."""
    cam_ids = cam_ids.reshape(-1).cpu()
    cam_times = cam_times.reshape(-1).cpu()
    offsets, nbrs = [0], []
    for i in range(cam_ids.shape[0]):
        same_cam = torch.where(cam_ids == cam_ids[i])[0]
        dt = torch.abs(cam_times[same_cam] - cam_times[i])
        close = same_cam[torch.where((dt <= ist_range) & (dt > 0.01))[0]]
        nbrs += close.tolist()
        offsets.append(len(nbrs))
    return torch.tensor(offsets, dtype=torch.int32), torch.tensor(nbrs, dtype=torch.int32)


def compute_ist_cuda(images: torch.Tensor, cam_ids: torch.Tensor, cam_times: torch.Tensor, ist_range: float) -> torch.Tensor:
    """``compute_ist`` with the per-pixel work in ``kp_ist_map`` (images on the GPU, [B,H,W,3] fp32) -> fp16 [B,H,W] on
    the GPU, bit-identical to the host version."""
    from .. import ops

    offsets, nbrs = temporal_neighbours(cam_ids, cam_times, ist_range)
    return ops.ist_map(images, offsets.to(images.device), nbrs.to(images.device), IST_ALPHA)


def compute_ist(images: torch.Tensor, cam_ids: torch.Tensor, cam_times: torch.Tensor, ist_range: float,
                device="cpu") -> torch.Tensor:
    """images [B,H,W,3], cam_ids [B] or [B,1], cam_times [B] or [B,1] -> fp16 [B,H,W].
    For every image: max abs difference to the images of the SAME camera whose time differs by (0.01, ist_range],
    mean over channels, values <= 0.15 zeroed; an image without such neighbours gets a uniform map."""
    if images.is_cuda:
        return compute_ist_cuda(images, cam_ids, cam_times, ist_range)
    b, h, w = images.shape[:3]
    cam_ids = cam_ids.reshape(-1)
    cam_times = cam_times.reshape(-1)
    out = torch.zeros(b, h, w, device=device)
    for i in range(b):
        same_cam = torch.where(cam_ids == cam_ids[i])[0]
        dt = torch.abs(cam_times[same_cam] - cam_times[i])
        close = same_cam[torch.where((dt <= ist_range) & (dt > 0.01))[0]]
        if len(close) == 0:
            out[i] = torch.ones(h, w, device=device)
            continue
        cur = images[i].to(device)
        max_diff = torch.zeros_like(cur)
        for j in close:
            max_diff = torch.maximum(max_diff, torch.abs(cur - images[j].to(device)))
        max_diff = max_diff.mean(dim=2)
        out[i] = torch.where(max_diff > IST_ALPHA, max_diff, torch.zeros_like(max_diff))
    return out.to(torch.float16)


def compute_isg(images: torch.Tensor, cam_ids: torch.Tensor, isg_gamma: float, device="cpu") -> torch.Tensor:
    """images [B,H,W,3], cam_ids [B] -> fp16 [B,H,W]: Geman-McClure residual to the per-camera median image.  Images on
    the GPU take the device path (``kp_isg_map``), bit-identical to this host version."""
    if images.is_cuda:
        from .. import ops

        return ops.isg_map(images, cam_ids, isg_gamma)
    b, h, w = images.shape[:3]
    cam_ids = cam_ids.reshape(-1)
    medians = {}
    for cid in torch.unique(cam_ids):
        sel = torch.where(cam_ids == cid)[0]
        medians[cid.item()] = torch.median(images[sel], dim=0).values.to(device)
    out = torch.zeros(b, h, w, device=device)
    for i in range(b):
        sq = torch.square(images[i].to(device) - medians[cam_ids[i].item()])
        psi = sq.div_(sq + isg_gamma**2)
        out[i] = (1.0 / 3) * torch.sum(psi, dim=-1)
    return out.to(torch.float16)
