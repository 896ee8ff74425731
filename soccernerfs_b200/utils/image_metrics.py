"""Image-quality metrics ``KPlanesModel.get_image_metrics_and_images`` reports (NS/models/kplanes.py:290-295, 454-515).

PSNR and SSIM are restated in torch with torchmetrics' defaults (the reference instantiates
``PeakSignalNoiseRatio(data_range=1.0)`` and ``structural_similarity_index_measure``): evaluation tooling, computed
once per eval image, not part of the hot path.  LPIPS (pretrained AlexNet) and the RetinaNet-based ``DynMetric`` need
downloaded weights; they are used through the reference's own objects when those import, and their keys are simply
absent otherwise -- never an exception in the middle of a training run.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F


def psnr(target: torch.Tensor, preds: torch.Tensor, data_range: float = 1.0) -> torch.Tensor:
    """10 log10(data_range^2 / mse) over the whole tensor (torchmetrics PeakSignalNoiseRatio, base 10, dim None)."""
    mse = torch.mean((preds - target) ** 2)
    return 10.0 * torch.log10(torch.as_tensor(data_range**2, device=mse.device, dtype=mse.dtype) / mse)


def ssim(preds: torch.Tensor, target: torch.Tensor, kernel_size: int = 11, sigma: float = 1.5, k1: float = 0.01,
         k2: float = 0.03, data_range: Optional[float] = None) -> torch.Tensor:
    """torchmetrics.functional.structural_similarity_index_measure defaults: [B,C,H,W] inputs, gaussian 11x11 window with
    sigma 1.5, reflect padding cropped from the result, data_range = max(range(preds), range(target)), mean over batch."""
    if data_range is None:
        data_range = float(max(preds.max() - preds.min(), target.max() - target.min()))
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    channels = preds.shape[1]
    dist = torch.arange((1 - kernel_size) / 2, (1 + kernel_size) / 2, 1, dtype=preds.dtype, device=preds.device)
    gauss = torch.exp(-((dist / sigma) ** 2) / 2)
    gauss = (gauss / gauss.sum())[None]
    kernel = (gauss.t() @ gauss).expand(channels, 1, kernel_size, kernel_size)
    pad = (kernel_size - 1) // 2
    preds = F.pad(preds, (pad, pad, pad, pad), mode="reflect")
    target = F.pad(target, (pad, pad, pad, pad), mode="reflect")
    stack = torch.cat((preds, target, preds * preds, target * target, preds * target))
    out = F.conv2d(stack, kernel, groups=channels)
    b = preds.shape[0]
    mu_p, mu_t, pp, tt, pt = (out[i * b:(i + 1) * b] for i in range(5))
    sig_p, sig_t, sig_pt = pp - mu_p**2, tt - mu_t**2, pt - mu_p * mu_t
    full = ((2 * mu_p * mu_t + c1) * (2 * sig_pt + c2)) / ((mu_p**2 + mu_t**2 + c1) * (sig_p + sig_t + c2))
    full = full[..., pad:-pad, pad:-pad]
    return full.reshape(b, -1).mean(-1).mean()


def optional_lpips(device):
    """torchmetrics' LearnedPerceptualImagePatchSimilarity if importable AND its weights are available, else None."""
    try:
        from torchmetrics.image.lpip import LearnedPerceptualImagePatchSimilarity

        return LearnedPerceptualImagePatchSimilarity(normalize=True).to(device)
    except Exception:
        return None


def optional_dynmetric(device):
    """The reference's RetinaNet-based dynamic-region metric (NS/utils/dynmetric.py) if nerfstudio is importable."""
    try:
        from nerfstudio.utils.dynmetric import DynMetric

        return DynMetric(device=device)
    except Exception:
        return None
