"""Which rays of a batch sit on a DISCRETE decision of the reference's step?  (test infrastructure, like the rest of oracle/)

The training step is piecewise smooth in its inputs: ReLU units of the decoders and of the proposal networks, the
``searchsorted(side="right")`` bin lookups of the interlevel loss (NS/model_components/losses.py:46-75) and the median
index of the depth renderer (renderers.py:256-264) each pick a branch.  Two correct fp32 implementations that differ
in the last bits (summation order, FMA contraction, exp/expf) agree to ~1e-6 away from those decisions, but a ray one
of whose decisions lies within rounding of a tie can take the other branch -- an O(1) change of that ray's gradient
contribution that no precision fixes.  ``fragile_rays`` evaluates the oracle forward once, recomputes every
pre-activation in fp64 together with the magnitude of the dot product's terms, and flags the rays that have

  * a ReLU unit with |pre| < relu_window * sum_i |w_i x_i|   (field sigma / colour nets, both proposal nets),
  * a final-level bin edge within ``edge_window`` of a proposal-level bin edge (interlevel ``outer`` lookups),
  * a cumulative weight within ``median_window`` of 0.5 (median-depth index).

The parity tests drop the flagged rays from BOTH implementations' batch (rays are independent through the whole step;
only the loss means couple them) and then assert the stated fp32 bar on everything that is left, reporting the flagged
fraction.  Windows are multiples of the fp32 error bound of the respective quantity, not tuned to make a test pass:
relu_window 2e-6 is ~8x the unit round-off of the 4-term TF32 products (2^-22) on the terms' magnitude.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from . import kplanes_oracle as ko


def _relu_fragile(x: torch.Tensor, w: torch.Tensor, window: float) -> torch.Tensor:
    """[M] bool: some unit of relu(x @ w.T) has |pre| within window * sum|terms| of zero."""
    xd, wd = x.double(), w.double()
    pre = xd @ wd.t()
    mag = xd.abs() @ wd.abs().t()
    return (pre.abs() < window * mag).any(dim=-1)


@torch.no_grad()
def fragile_rays(mp: ko.ModelParams, origins, directions, times, rand, anneal: float = 1.0, near_plane: float = 0.0,
                 relu_window: float = 2e-6, edge_window: float = 2e-6, median_window: float = 2e-6
                 ) -> Tuple[torch.Tensor, Dict[str, float]]:
    """-> (bool [N]: ray is fragile, {criterion: fraction of rays it flags})."""
    nears, fars = ko.aabb_collider(origins, directions, mp.field.aabb, near_plane)
    out = ko.model_forward(mp, origins, directions, times, nears, fars, rand, anneal=anneal, training=True)
    n = origins.shape[0]
    flags: Dict[str, torch.Tensor] = {}
    # field decoders
    f = mp.field
    feats = out["features"]
    s = feats.shape[0] // n
    frag = _relu_fragile(feats, f.sigma_w[0], relu_window)
    h1 = torch.relu(feats @ f.sigma_w[0].t())
    o = h1 @ f.sigma_w[1].t()
    geo = o[:, : f.geo_feat_dim]
    if f.view_dependent:
        dirs = directions[:, None, :].expand(n, s, 3).reshape(-1, 3)
        cin = torch.cat([ko.sh4((dirs + 1.0) / 2.0), geo], dim=-1)
    else:
        cin = geo
    frag |= _relu_fragile(cin, f.color_w[0], relu_window)
    h2 = torch.relu(cin @ f.color_w[0].t())
    frag |= _relu_fragile(h2, f.color_w[1], relu_window)
    flags["field_relu"] = frag.view(n, s).any(-1)
    # proposal networks
    frag_p = torch.zeros(n, dtype=torch.bool)
    for lvl, p in enumerate(mp.proposals):
        smp = out["samples_list"][lvl]
        pos = smp.positions()
        sp = pos.shape[1]
        pts = ko._normalized(pos, p.aabb)
        if times is not None and len(p.grids) == 6:
            pts = torch.cat([pts, ((times * 2) - 1)[:, None, :].expand(n, sp, 1)], dim=-1)
        pf = ko.interpolate_kplanes(pts.reshape(-1, pts.shape[-1]), [p.grids], concat_features=False)
        frag_p |= _relu_fragile(pf, p.sigma_w[0], relu_window).view(n, sp).any(-1)
    flags["proposal_relu"] = frag_p
    # interlevel lookups: final-level edges vs every proposal level's edges
    c = out["samples_list"][-1].spacing_bins
    tie = torch.zeros(n, dtype=torch.bool)
    for lvl in range(len(mp.proposals)):
        cp = out["samples_list"][lvl].spacing_bins.contiguous()
        idx = torch.searchsorted(cp, c.contiguous(), side="right").clamp(1, cp.shape[-1] - 1)
        lo, hi = torch.take_along_dim(cp, idx - 1, -1), torch.take_along_dim(cp, idx, -1)
        dist = torch.minimum((c - lo).abs(), (hi - c).abs())
        tie |= (dist < edge_window).any(-1)
    flags["interlevel_edge_tie"] = tie
    # median index of the final level and of the proposal levels (prop_depth_i)
    med = torch.zeros(n, dtype=torch.bool)
    for w in out["weights_list"]:
        cw = torch.cumsum(w[..., 0].double(), dim=-1)
        med |= ((cw - 0.5).abs() < median_window).any(-1)
    flags["median_tie"] = med
    fragile = torch.zeros(n, dtype=torch.bool)
    for v in flags.values():
        fragile |= v
    stats = {k: float(v.float().mean()) for k, v in flags.items()}
    stats["any"] = float(fragile.float().mean())
    return fragile, stats
