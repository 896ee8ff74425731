"""Visualisation colour maps for ``get_image_metrics_and_images`` (NS/utils/colormaps.py:26-82).

The reference indexes matplotlib's 256-entry ``viridis`` / ``turbo`` tables.  matplotlib is used when importable (then
the images are identical to the reference's); otherwise the tables are rebuilt from the published polynomial fits of
the two maps (turbo: A. Mikhailov's 5th-order fit from the Turbo release notes; viridis: M. Zucker's 6th-order fit),
which agree with the tables to ~1/255.  Evaluation tooling only -- nothing here is on the training path.
"""
from __future__ import annotations

from typing import Optional

import torch

_TABLES = {}

_TURBO = (
    (0.13572138, 4.61539260, -42.66032258, 132.13108234, -152.94239396, 59.28637943),
    (0.09140261, 2.19418839, 4.84296658, -14.18503333, 4.27729857, 2.82956604),
    (0.10667330, 12.64194608, -60.58204836, 110.36276771, -89.90310912, 27.34824973),
)
_VIRIDIS = (
    (0.2777273272234177, 0.1050930431085774, -0.3308618287255563, -4.634230498983486, 6.228269936347081, 4.776384997670288,
     -5.435455855934631),
    (0.005407344544966578, 1.404613529898575, 0.214847559468213, -5.799100973351585, 14.17993336680509, -13.74514537774601,
     4.645852612178535),
    (0.3340998053353061, 1.384590162594685, 0.09509516302823659, -19.33244095627987, 56.69055260068105, -65.35303263337234,
     26.3124352495832),
)


def _table(cmap: str) -> torch.Tensor:
    if cmap not in _TABLES:
        try:
            from matplotlib import cm  # the reference's own source of the tables

            _TABLES[cmap] = torch.tensor(cm.get_cmap(cmap).colors, dtype=torch.float32)
        except Exception:
            coeffs = {"turbo": _TURBO, "viridis": _VIRIDIS}.get(cmap)
            if coeffs is None:
                raise ValueError(f"colour map {cmap!r} needs matplotlib")
            x = torch.linspace(0.0, 1.0, 256, dtype=torch.float64)
            chans = [sum(c * x**i for i, c in enumerate(cs)) for cs in coeffs]
            _TABLES[cmap] = torch.stack(chans, dim=-1).clamp(0.0, 1.0).float()
    return _TABLES[cmap]


def apply_colormap(image: torch.Tensor, cmap: str = "viridis") -> torch.Tensor:
    """[..., 1] in [0,1] -> [..., 3] (colormaps.py:26-45)."""
    table = _table(cmap).to(image.device)
    image = torch.nan_to_num(image, 0)
    image_long = (image * 255).long()
    lo, hi = int(image_long.min()), int(image_long.max())
    assert lo >= 0, f"the min value is {lo}"
    assert hi <= 255, f"the max value is {hi}"
    return table[image_long[..., 0]]


def apply_depth_colormap(depth: torch.Tensor, accumulation: Optional[torch.Tensor] = None, near_plane: Optional[float] = None,
                         far_plane: Optional[float] = None, cmap: str = "turbo") -> torch.Tensor:
    """Depth -> colour, white where nothing accumulated (colormaps.py:48-82)."""
    near_plane = near_plane or float(torch.min(depth))
    far_plane = far_plane or float(torch.max(depth))
    depth = torch.clip((depth - near_plane) / (far_plane - near_plane + 1e-10), 0, 1)
    colored = apply_colormap(depth, cmap=cmap)
    if accumulation is not None:
        colored = colored * accumulation + (1 - accumulation)
    return colored
