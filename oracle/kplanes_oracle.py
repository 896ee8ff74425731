"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the K-Planes train/render hot path.

This is a plain-PyTorch (CPU, fp32, no autocast) *restatement* of the reference algorithm, written
functionally over explicit tensors so that every CUDA kernel can be checked against it on the same
seeded inputs.  It is NOT shipped and NOT on the product path: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may
import it (and there only as the checker / the CPU baseline being timed).

Parity pin: the reference's own tests hold no golden vectors for this path, so the
oracle is pinned against outputs of the reference's own modules run in the build container
(``oracle/ref_loader.py`` + ``oracle/make_golden.py`` -> ``tests/golden/*.npz``, checked by
``tests/test_oracle_golden.py`` everywhere and by ``tests/test_oracle_vs_reference.py`` where
/root/reference exists).  The tiny-cuda-nn decoders (un-vendored third party, pinned v1.6 in
``nerfstudio/Dockerfile:121``) are restated as bias-free fp32 dense stacks and the SH encoding as
``nerfstudio/utils/math.py:25-86`` of ``2x-1``; that boundary is **parity unpinned** by the
reference (no test, CUDA-only), see DESIGN.md.

NS = /root/reference/nerfstudio/nerfstudio.  All functions take/return torch tensors.
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

EPS_LOSS = 1.0e-7  # NS/model_components/losses.py:32


# --------------------------------------------------------------------------------------------
# Field: plane interpolation (NS/utils/interpolation.py:5-33, NS/fields/kplanes_field.py:77-126)
# --------------------------------------------------------------------------------------------
def grid_sample_plane(plane: torch.Tensor, coords: torch.Tensor) -> torch.Tensor:
    """plane [1,C,H,W], coords [M,2] in [-1,1] (coords[:,0] -> W, coords[:,1] -> H) -> [M,C].

    NS/utils/interpolation.py:24-32: bilinear, align_corners=True, padding_mode="border".
    """
    c = plane.shape[1]
    out = F.grid_sample(plane, coords.view(1, 1, -1, 2), align_corners=True, mode="bilinear", padding_mode="border")
    return out.view(c, -1).transpose(0, 1)


def bilinear_border_manual(plane: torch.Tensor, coords: torch.Tensor) -> torch.Tensor:
    """Hand restatement of ATen's grid_sampler_2d (bilinear / border / align_corners) used to document
    the exact arithmetic the CUDA kernel follows; checked against :func:`grid_sample_plane`."""
    _, c, h, w = plane.shape
    ix = ((coords[:, 0] + 1.0) / 2.0) * (w - 1)
    iy = ((coords[:, 1] + 1.0) / 2.0) * (h - 1)
    ix = ix.clamp(0.0, float(w - 1))
    iy = iy.clamp(0.0, float(h - 1))
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    x1 = x0 + 1.0
    y1 = y0 + 1.0
    nw = (x1 - ix) * (y1 - iy)
    ne = (ix - x0) * (y1 - iy)
    sw = (x1 - ix) * (iy - y0)
    se = (ix - x0) * (iy - y0)
    p = plane[0].permute(1, 2, 0)  # [H,W,C]

    def fetch(yy, xx):
        valid = ((xx >= 0) & (xx <= w - 1) & (yy >= 0) & (yy <= h - 1)).unsqueeze(-1)
        v = p[yy.long().clamp(0, h - 1), xx.long().clamp(0, w - 1)]
        return torch.where(valid, v, torch.zeros_like(v))

    return (
        fetch(y0, x0) * nw[:, None]
        + fetch(y0, x1) * ne[:, None]
        + fetch(y1, x0) * sw[:, None]
        + fetch(y1, x1) * se[:, None]
    )


def interpolate_kplanes(
    pts: torch.Tensor,
    ms_grids: Sequence[Sequence[torch.Tensor]],
    concat_features: bool,
    freeze_time_planes: bool = False,
) -> torch.Tensor:
    """pts [M,3|4] in [-1,1]; ms_grids[scale][plane] = [1,C,H,W].  NS/fields/kplanes_field.py:77-126.

    Plane order = combinations(range(D),2) = XY,XZ,XT,YZ,YT,ZT; plane (a,b) has H=reso[b], W=reso[a]
    (kplanes_field.py:61-67) and is sampled at pts[:,(a,b)] so coordinate a indexes W.
    """
    d = pts.shape[-1]
    combs = list(itertools.combinations(range(d), 2))
    per_scale = []
    for grids in ms_grids:
        interp = 1.0
        for ci, comb in enumerate(combs):
            if freeze_time_planes and d == 4 and 3 in comb:
                continue
            interp = interp * grid_sample_plane(grids[ci], pts[:, list(comb)])
        per_scale.append(interp)
    if concat_features:
        return torch.cat(per_scale, dim=-1)
    total = 0.0
    for t in per_scale:
        total = total + t
    return total


# --------------------------------------------------------------------------------------------
# Decoders (tcnn stand-ins) and activations
# --------------------------------------------------------------------------------------------
class _TruncExp(torch.autograd.Function):
    """NS/field_components/activations.py:25-41: fwd exp(x), bwd g*exp(clamp(x,-15,15))."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(-15, 15))


trunc_exp = _TruncExp.apply


def mlp(x: torch.Tensor, weights: Sequence[torch.Tensor], out_act: str = "none") -> torch.Tensor:
    """Bias-free dense stack, ReLU hidden activations (tcnn FullyFusedMLP semantics, fp32).
    weights[i] is [out_i, in_i] (torch.nn.Linear convention)."""
    for i, w in enumerate(weights):
        x = x @ w.t()
        if i + 1 < len(weights):
            x = torch.relu(x)
    if out_act == "sigmoid":
        x = torch.sigmoid(x)
    return x


def sh4(dirs01: torch.Tensor) -> torch.Tensor:
    """Degree-4 SH of (2x-1): NS/utils/math.py:25-86 applied as tcnn's SphericalHarmonics encoding
    would (inputs in [0,1], kplanes_field.py:39-44, :319-321)."""
    d = dirs01 * 2.0 - 1.0
    x, y, z = d[..., 0], d[..., 1], d[..., 2]
    xx, yy, zz = x**2, y**2, z**2
    comps = [
        torch.full_like(x, 0.28209479177387814),
        0.4886025119029199 * y,
        0.4886025119029199 * z,
        0.4886025119029199 * x,
        1.0925484305920792 * x * y,
        1.0925484305920792 * y * z,
        0.9461746957575601 * zz - 0.31539156525251999,
        1.0925484305920792 * x * z,
        0.5462742152960396 * (xx - yy),
        0.5900435899266435 * y * (3 * xx - yy),
        2.890611442640554 * x * y * z,
        0.4570457994644658 * y * (5 * zz - 1),
        0.3731763325901154 * z * (5 * zz - 3),
        0.4570457994644658 * x * (5 * zz - 1),
        1.445305721320277 * z * (xx - yy),
        0.5900435899266435 * x * (xx - 3 * yy),
    ]
    return torch.stack(comps, dim=-1)


# --------------------------------------------------------------------------------------------
# Parameter containers (reference layouts: planes NCHW [1,C,H,W])
# --------------------------------------------------------------------------------------------
@dataclass
class FieldParams:
    """KPlanesField parameters (NS/fields/kplanes_field.py:147-273)."""

    aabb: torch.Tensor  # [2,3]
    grids: List[List[torch.Tensor]]  # [scale][plane] -> [1,C,H,W]
    sigma_w: List[torch.Tensor]  # [hid,K*C], [16,hid]
    color_w: List[torch.Tensor]  # [64,in], [64,64], [3,64]
    concat: bool = True
    view_dependent: bool = True
    geo_feat_dim: int = 15

    def tensors(self) -> List[torch.Tensor]:
        return [p for g in self.grids for p in g] + list(self.sigma_w) + list(self.color_w)


@dataclass
class DensityFieldParams:
    """KPlanesDensityField parameters (NS/fields/kplanes_field.py:376-407)."""

    aabb: torch.Tensor
    grids: List[torch.Tensor]  # [plane] -> [1,C,H,W]
    sigma_w: List[torch.Tensor]  # [64,C], [1,64]

    def tensors(self) -> List[torch.Tensor]:
        return list(self.grids) + list(self.sigma_w)


def _band_limited(shape, coarse: int, draw) -> torch.Tensor:
    """A [1,C,H,W] field whose spatial axes (those longer than ``coarse``) are the bilinear upsampling of a
    ``coarse``-resolution random field: smooth at texel scale, like planes after TV-regularised training."""
    h, w = shape[2], shape[3]
    ch, cw = min(h, coarse), min(w, coarse)
    low = draw([shape[0], shape[1], ch, cw])
    if (ch, cw) == (h, w):
        return low
    return torch.nn.functional.interpolate(low, size=(h, w), mode="bilinear", align_corners=True).contiguous()


def init_planes(c: int, reso: Sequence[int], a: float, b: float, gen: torch.Generator, smooth_res: int = 0) -> List[torch.Tensor]:
    """NS/fields/kplanes_field.py:47-74 (time planes = 1, space planes U(a,b)).  ``smooth_res`` > 0 (test inputs only):
    the uniform noise is drawn at that resolution and bilinearly upsampled (see ``_band_limited``)."""
    planes = []
    for comb in itertools.combinations(range(len(reso)), 2):
        shape = [1, c] + [reso[cc] for cc in comb[::-1]]
        if len(reso) == 4 and 3 in comb:
            planes.append(torch.ones(shape))
        elif smooth_res > 0:
            planes.append(_band_limited(shape, smooth_res, lambda sh: torch.empty(sh).uniform_(a, b, generator=gen)))
        else:
            planes.append(torch.empty(shape).uniform_(a, b, generator=gen))
    return planes


def _time_noise(shape, time_noise: float, gen: torch.Generator, smooth_res: int = 0) -> torch.Tensor:
    """N(0, time_noise) on a space-time plane [1,C,T,R]; band-limited along the spatial axis when ``smooth_res`` > 0."""
    if smooth_res > 0 and shape[3] > smooth_res:
        low = time_noise * torch.randn([shape[0], shape[1], shape[2], smooth_res], generator=gen)
        return torch.nn.functional.interpolate(low, size=(shape[2], shape[3]), mode="bilinear", align_corners=True).contiguous()
    return time_noise * torch.randn(shape, generator=gen)


def xavier(out_d: int, in_d: int, gen: torch.Generator) -> torch.Tensor:
    bound = (6.0 / (in_d + out_d)) ** 0.5
    return torch.empty(out_d, in_d).uniform_(-bound, bound, generator=gen)


def make_field_params(
    aabb, spacetime_resolution, feat_dim, multiscale_res, gen, sigma_hidden=64, rgb_hidden=64,
    view_dependent=True, time_noise=0.05, concat=True, smooth_res=0,
) -> FieldParams:
    grids = []
    for m in multiscale_res:
        reso = [r * m for r in spacetime_resolution[:3]] + list(spacetime_resolution[3:])
        planes = init_planes(feat_dim, reso, 0.1, 0.5, gen, smooth_res)
        if len(reso) == 4 and time_noise > 0:
            for i in (2, 4, 5):  # N(0,0.05) noise on time planes so grads are non-degenerate
                planes[i] = planes[i] + _time_noise(planes[i].shape, time_noise, gen, smooth_res)
        grids.append(planes)
    k = feat_dim * len(multiscale_res) if concat else feat_dim
    in_color = 15 + (16 if view_dependent else 0)
    return FieldParams(
        aabb=aabb,
        grids=grids,
        sigma_w=[xavier(sigma_hidden, k, gen), xavier(16, sigma_hidden, gen)],
        color_w=[xavier(rgb_hidden, in_color, gen), xavier(rgb_hidden, rgb_hidden, gen), xavier(3, rgb_hidden, gen)],
        concat=concat,
        view_dependent=view_dependent,
    )


def make_density_params(aabb, resolution, feat_dim, gen, time_noise=0.05, smooth_res=0) -> DensityFieldParams:
    planes = init_planes(feat_dim, resolution, 0.1, 0.15, gen, smooth_res)
    if len(resolution) == 4 and time_noise > 0:
        for i in (2, 4, 5):
            planes[i] = planes[i] + _time_noise(planes[i].shape, time_noise, gen, smooth_res)
    return DensityFieldParams(aabb=aabb, grids=planes, sigma_w=[xavier(64, feat_dim, gen), xavier(1, 64, gen)])


# --------------------------------------------------------------------------------------------
# Ray samples (functional stand-in for RaySamples / Frustums, NS/cameras/rays.py:31-125, 233-277)
# --------------------------------------------------------------------------------------------
@dataclass
class Samples:
    origins: torch.Tensor  # [N,3]
    directions: torch.Tensor  # [N,3]
    starts: torch.Tensor  # [N,S]
    ends: torch.Tensor  # [N,S]
    spacing_bins: torch.Tensor  # [N,S+1] (spacing_starts ++ last spacing_end)
    nears: torch.Tensor  # [N,1]
    fars: torch.Tensor  # [N,1]
    times: Optional[torch.Tensor] = None  # [N,1]

    @property
    def deltas(self):
        return self.ends - self.starts  # rays.py:254

    def positions(self):
        """Frustums.get_positions, rays.py:54: origins + directions * (starts + ends) / 2."""
        return self.origins[:, None, :] + self.directions[:, None, :] * (self.starts + self.ends)[..., None] / 2

    def steps(self):
        return (self.starts + self.ends) / 2  # renderers.py:256


def aabb_collider(origins, directions, aabb, near_plane: float = 0.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """NS/model_components/scene_colliders.py:57-95.  Returns nears, fars as [N,1]."""
    dir_fraction = 1.0 / (directions + 1e-6)
    t1 = (aabb[0, 0] - origins[:, 0:1]) * dir_fraction[:, 0:1]
    t2 = (aabb[1, 0] - origins[:, 0:1]) * dir_fraction[:, 0:1]
    t3 = (aabb[0, 1] - origins[:, 1:2]) * dir_fraction[:, 1:2]
    t4 = (aabb[1, 1] - origins[:, 1:2]) * dir_fraction[:, 1:2]
    t5 = (aabb[0, 2] - origins[:, 2:3]) * dir_fraction[:, 2:3]
    t6 = (aabb[1, 2] - origins[:, 2:3]) * dir_fraction[:, 2:3]
    nears = torch.max(torch.cat([torch.minimum(t1, t2), torch.minimum(t3, t4), torch.minimum(t5, t6)], 1), 1).values
    fars = torch.min(torch.cat([torch.maximum(t1, t2), torch.maximum(t3, t4), torch.maximum(t5, t6)], 1), 1).values
    nears = torch.clamp(nears, min=near_plane)
    fars = torch.maximum(fars, nears + 1e-6)
    return nears[:, None], fars[:, None]


def spacing_fns(spacing: str):
    """(spacing_fn, spacing_fn_inv): "uniform" (UniformSampler, ray_samplers.py:129-150) or "piecewise"
    (UniformLinDispPiecewiseSampler, ray_samplers.py:221-246: uniform up to distance 1, linear in disparity beyond)."""
    if spacing == "uniform":
        return (lambda x: x), (lambda x: x)
    assert spacing == "piecewise"
    return (lambda x: torch.where(x < 1, x / 2, 1 - 1 / (2 * x))), (lambda x: torch.where(x < 0.5, 2 * x, 1 / (2 - 2 * x)))


def _make_samples(origins, directions, nears, fars, times, bins, spacing: str = "uniform") -> Samples:
    fn, fn_inv = spacing_fns(spacing)
    s_near, s_far = fn(nears), fn(fars)  # ray_samplers.py:114-116
    euclid = fn_inv(bins * s_far + (1 - bins) * s_near)
    smp = Samples(origins, directions, euclid[:, :-1], euclid[:, 1:], bins, nears, fars, times)
    smp.spacing = spacing
    return smp


def uniform_sampler(origins, directions, nears, fars, times, num_samples: int, t_rand: Optional[torch.Tensor],
                    spacing: str = "uniform"):
    """UniformSampler / SpacedSampler.generate_ray_samples, NS/model_components/ray_samplers.py:79-126.
    t_rand [N,S+1] (or [N,1] for single_jitter) is the reference's ``torch.rand`` draw; None = eval."""
    bins = torch.linspace(0.0, 1.0, num_samples + 1)[None, :]
    if t_rand is not None:
        centers = (bins[..., 1:] + bins[..., :-1]) / 2.0
        upper = torch.cat([centers, bins[..., -1:]], -1)
        lower = torch.cat([bins[..., :1], centers], -1)
        bins = lower + (upper - lower) * t_rand
    else:
        bins = bins.expand(origins.shape[0], -1)
    return _make_samples(origins, directions, nears, fars, times, bins, spacing)


def pdf_cdf(weights: torch.Tensor, histogram_padding: float = 0.01, eps: float = 1e-5) -> torch.Tensor:
    """weights [N,S] -> cdf [N,S+1].  ray_samplers.py:302-312."""
    w = weights + histogram_padding
    w_sum = torch.sum(w, dim=-1, keepdim=True)
    padding = torch.relu(eps - w_sum)
    w = w + padding / w.shape[-1]
    w_sum = w_sum + padding
    pdf = w / w_sum
    cdf = torch.min(torch.ones_like(pdf), torch.cumsum(pdf, dim=-1))
    return torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)


def pdf_u(n_rays: int, num_samples: int, rand: Optional[torch.Tensor]) -> torch.Tensor:
    """ray_samplers.py:314-328.  rand [N,S_out+1] (or [N,1]) in training, None in eval."""
    num_bins = num_samples + 1
    u = torch.linspace(0.0, 1.0 - (1.0 / num_bins), steps=num_bins)
    if rand is not None:
        u = u.expand(n_rays, num_bins) + rand / num_bins
    else:
        u = (u + 1.0 / (2 * num_bins)).expand(n_rays, num_bins)
    return u.contiguous()


def pdf_sampler(prev: Samples, weights: torch.Tensor, num_samples: int, rand: Optional[torch.Tensor]):
    """PDFSampler.generate_ray_samples (include_original=False), ray_samplers.py:274-369.
    Returns (Samples, inds int64 [N,S_out+1])."""
    cdf = pdf_cdf(weights)
    u = pdf_u(weights.shape[0], num_samples, rand)
    existing = prev.spacing_bins
    inds = torch.searchsorted(cdf, u, side="right")
    below = torch.clamp(inds - 1, 0, existing.shape[-1] - 1)
    above = torch.clamp(inds, 0, existing.shape[-1] - 1)
    cdf_g0 = torch.gather(cdf, -1, below)
    bins_g0 = torch.gather(existing, -1, below)
    cdf_g1 = torch.gather(cdf, -1, above)
    bins_g1 = torch.gather(existing, -1, above)
    t = torch.clip(torch.nan_to_num((u - cdf_g0) / (cdf_g1 - cdf_g0), 0), 0, 1)
    bins = (bins_g0 + t * (bins_g1 - bins_g0)).detach()
    # the euclidean positions come from the level-0 sampler's closure (ray_samplers.py:359): same spacing function
    return _make_samples(prev.origins, prev.directions, prev.nears, prev.fars, prev.times, bins,
                         getattr(prev, "spacing", "uniform")), inds


# --------------------------------------------------------------------------------------------
# Field evaluation
# --------------------------------------------------------------------------------------------
def _normalized(positions, aabb):
    return (positions - aabb[0]) / (aabb[1] - aabb[0])  # NS/data/scene_box.py:56-66


def density_field(p: DensityFieldParams, positions: torch.Tensor, times: Optional[torch.Tensor]) -> torch.Tensor:
    """KPlanesDensityField.density_fn/get_density, kplanes_field.py:410-460.  positions [N,S,3] world,
    times [N,1].  NOTE (SURVEY finding 3): positions are normalised to [0,1] only, NOT to [-1,1]."""
    n, s = positions.shape[:2]
    pts = _normalized(positions, p.aabb)
    if times is not None and len(p.grids) == 6:
        t = (times * 2) - 1
        pts = torch.cat([pts, t[:, None, :].expand(n, s, 1)], dim=-1)
    feats = interpolate_kplanes(pts.reshape(-1, pts.shape[-1]), [p.grids], concat_features=False)
    return trunc_exp(mlp(feats, p.sigma_w)).view(n, s, 1)


def field_forward(p: FieldParams, smp: Samples) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """KPlanesField.forward = get_density + get_outputs, kplanes_field.py:275-370 (bounded, MLP decoder,
    no appearance embedding).  Returns density [N,S,1], rgb [N,S,3], features [M,K*C]."""
    positions = smp.positions()
    n, s = positions.shape[:2]
    pts = _normalized(positions, p.aabb) * 2.0 - 1.0
    if smp.times is not None and len(p.grids[0]) == 6:
        t = (smp.times * 2) - 1
        pts = torch.cat([pts, t[:, None, :].expand(n, s, 1)], dim=-1)
    feats = interpolate_kplanes(pts.reshape(-1, pts.shape[-1]), p.grids, concat_features=p.concat)
    o = mlp(feats, p.sigma_w)
    geo, sigma_raw = torch.split(o, [p.geo_feat_dim, 1], dim=-1)
    density = trunc_exp(sigma_raw).view(n, s, 1)
    if p.view_dependent:
        dirs = smp.directions[:, None, :].expand(n, s, 3).reshape(-1, 3)
        cin = torch.cat([sh4((dirs + 1.0) / 2.0), geo], dim=-1)
    else:
        cin = geo
    rgb = mlp(cin, p.color_w, out_act="sigmoid").view(n, s, 3)
    return density, rgb, feats


# --------------------------------------------------------------------------------------------
# Compositing (NS/cameras/rays.py:127-149, NS/model_components/renderers.py)
# --------------------------------------------------------------------------------------------
def get_weights(deltas: torch.Tensor, densities: torch.Tensor) -> torch.Tensor:
    """deltas, densities [N,S,1] -> weights [N,S,1].  rays.py:137-147."""
    delta_density = deltas * densities
    alphas = 1 - torch.exp(-delta_density)
    transmittance = torch.cumsum(delta_density[..., :-1, :], dim=-2)
    transmittance = torch.cat([torch.zeros((*transmittance.shape[:1], 1, 1)), transmittance], dim=-2)
    transmittance = torch.exp(-transmittance)
    return torch.nan_to_num(alphas * transmittance)


def render_rgb(rgb, weights, background, training: bool = True) -> torch.Tensor:
    """RGBRenderer.forward/combine_rgb, renderers.py:71-140.  background: "last_sample" | [N,3]|[3] tensor
    (the "random" mode's ``rand_like`` draw is passed in explicitly)."""
    if not training:
        rgb = torch.nan_to_num(rgb)
    comp = torch.sum(weights * rgb, dim=-2)
    acc = torch.sum(weights, dim=-2)
    if isinstance(background, str):
        assert background == "last_sample"
        background = rgb[..., -1, :]
    comp = comp + background * (1.0 - acc)
    if not training:
        comp = comp.clamp(0.0, 1.0)
    return comp


def render_accumulation(weights) -> torch.Tensor:
    return torch.sum(weights, dim=-2)  # renderers.py:222


def median_index(weights) -> torch.Tensor:
    """renderers.py:260-263 / :324-327.  -> int64 [N,1]."""
    cum = torch.cumsum(weights[..., 0], dim=-1)
    split = torch.ones((*weights.shape[:-2], 1)) * 0.5
    idx = torch.searchsorted(cum, split, side="left")
    return torch.clamp(idx, 0, weights.shape[-2] - 1)


def render_depth_median(weights, steps) -> torch.Tensor:
    return torch.gather(steps, dim=-1, index=median_index(weights))  # renderers.py:264


def render_depth_expected(weights, steps) -> torch.Tensor:
    """renderers.py:266-283 (note the *global* steps.min()/max() clip)."""
    depth = torch.sum(weights * steps[..., None], dim=-2) / (torch.sum(weights, -2) + 1e-10)
    return torch.clip(depth, steps.min(), steps.max())


def render_median_rgb(rgb, weights, training: bool = True) -> torch.Tensor:
    """MedianRGBRenderer, renderers.py:320-362."""
    if not training:
        rgb = torch.nan_to_num(rgb)
    idx = median_index(weights).unsqueeze(2).expand(-1, -1, 3)
    out = torch.gather(rgb, dim=-2, index=idx)
    if not training:
        out = out.clamp(0.0, 1.0)
    return out


# --------------------------------------------------------------------------------------------
# Losses (NS/model_components/losses.py)
# --------------------------------------------------------------------------------------------
def outer(t0_starts, t0_ends, t1_starts, t1_ends, y1):
    """losses.py:46-75."""
    cy1 = torch.cat([torch.zeros_like(y1[..., :1]), torch.cumsum(y1, dim=-1)], dim=-1)
    idx_lo = torch.searchsorted(t1_starts.contiguous(), t0_starts.contiguous(), side="right") - 1
    idx_lo = torch.clamp(idx_lo, min=0, max=y1.shape[-1] - 1)
    idx_hi = torch.searchsorted(t1_ends.contiguous(), t0_ends.contiguous(), side="right")
    idx_hi = torch.clamp(idx_hi, min=0, max=y1.shape[-1] - 1)
    cy1_lo = torch.take_along_dim(cy1[..., :-1], idx_lo, dim=-1)
    cy1_hi = torch.take_along_dim(cy1[..., 1:], idx_hi, dim=-1)
    return cy1_hi - cy1_lo


def lossfun_outer(t, w, t_env, w_env):
    """losses.py:78-95."""
    w_outer = outer(t[..., :-1], t[..., 1:], t_env[..., :-1], t_env[..., 1:], w_env)
    return torch.clip(w - w_outer, min=0) ** 2 / (w + EPS_LOSS)


def interlevel_loss(weights_list: List[torch.Tensor], sdist_list: List[torch.Tensor]):
    """losses.py:106-121.  weights_list[i] [N,S_i,1]; sdist_list[i] [N,S_i+1] (spacing bins)."""
    c = sdist_list[-1].detach()
    w = weights_list[-1][..., 0].detach()
    total = 0.0
    for sdist, weights in zip(sdist_list[:-1], weights_list[:-1]):
        total = total + torch.mean(lossfun_outer(c, w, sdist, weights[..., 0]))
    return total


def lossfun_distortion(t, w):
    """losses.py:125-136."""
    ut = (t[..., 1:] + t[..., :-1]) / 2
    dut = torch.abs(ut[..., :, None] - ut[..., None, :])
    loss_inter = torch.sum(w * torch.sum(w[..., None, :] * dut, dim=-1), dim=-1)
    loss_intra = torch.sum(w**2 * (t[..., 1:] - t[..., :-1]), dim=-1) / 3
    return loss_inter + loss_intra


def distortion_loss(weights_list, sdist_list):
    """losses.py:139-144."""
    return torch.mean(lossfun_distortion(sdist_list[-1], weights_list[-1][..., 0]))


def compute_plane_tv(t, only_w=False):
    """losses.py:356-366."""
    _, _, h, w = t.shape
    h_tv = torch.square(t[..., 1:, :] - t[..., : h - 1, :]).mean()
    w_tv = torch.square(t[..., :, 1:] - t[..., :, : w - 1]).mean()
    return h_tv + w_tv if not only_w else w_tv


def compute_plane_smoothness(t):
    """losses.py:369-380."""
    _, _, h, _ = t.shape
    first = t[..., 1:, :] - t[..., : h - 1, :]
    second = first[..., 1:, :] - first[..., : h - 2, :]
    return torch.square(second).mean()


def space_tv_loss(multi_res_grids):
    """losses.py:383-406."""
    total = 0.0
    for grids in multi_res_grids:
        spatial = [0, 1, 2] if len(grids) == 3 else [0, 1, 3]
        for gid, grid in enumerate(grids):
            total = total + compute_plane_tv(grid, only_w=gid not in spatial)
    return total


def time_smoothness_loss(multi_res_grids):
    """losses.py:409-428."""
    total = 0.0
    for grids in multi_res_grids:
        for gid in [] if len(grids) == 3 else [2, 4, 5]:
            total = total + compute_plane_smoothness(grids[gid])
    return torch.as_tensor(total)


def sparse_transients_loss(multi_res_grids):
    """losses.py:431-452."""
    total = 0.0
    for grids in multi_res_grids:
        if len(grids) == 3:
            continue
        for gid in [2, 4, 5]:
            total = total + torch.abs(1 - grids[gid]).mean()
    return torch.as_tensor(total)


def ds_nerf_depth_loss(weights, termination_depth, steps, lengths, sigma):
    """losses.py:213-235.  weights/steps/lengths [N,S,1], termination_depth [N,1]."""
    depth_mask = termination_depth > 0
    loss = -torch.log(weights + EPS_LOSS) * torch.exp(-((steps - termination_depth[:, None]) ** 2) / (2 * sigma)) * lengths
    loss = loss.sum(-2) * depth_mask
    return torch.mean(loss)


# K-Planes default loss coefficients, NS/models/kplanes.py:148-161
DEFAULT_LOSS_COEFFICIENTS = {
    "rgb_loss": 1.0,
    "interlevel_loss": 1.0,
    "distortion_loss": 0.001,
    "space_tv_loss": 0.0002,
    "time_smoothness_loss": 0.001,
    "sparse_transients_loss": 0.0001,
    "space_tv_proposal_loss": 0.0002,
    "time_smoothness_proposal_loss": 0.00001,
    "sparse_transients_proposal_loss": 0.0001,
    "depth_loss": 0.05,
}


# --------------------------------------------------------------------------------------------
# Whole model forward + loss (NS/models/kplanes.py:349-388, 414-452; ray_samplers.py:559-600)
# --------------------------------------------------------------------------------------------
@dataclass
class ModelParams:
    field: FieldParams
    proposals: List[DensityFieldParams]
    num_proposal_samples: Tuple[int, ...] = (256, 128)
    num_nerf_samples: int = 48
    loss_coefficients: Dict[str, float] = field(default_factory=lambda: dict(DEFAULT_LOSS_COEFFICIENTS))

    def tensors(self) -> List[torch.Tensor]:
        out = []
        for p in self.proposals:
            out += p.tensors()
        return out + self.field.tensors()


def model_forward(
    mp: ModelParams,
    origins, directions, times, nears, fars,
    rand: Optional[Dict[str, torch.Tensor]],
    anneal: float = 1.0,
    training: bool = True,
    background="random",
) -> Dict[str, torch.Tensor]:
    """KPlanesModel.get_outputs.  ``rand`` carries the reference's torch.rand draws in call order:
    "t_rand" [N,S0+1] (initial sampler), "u1".."uL" [N,S_l+1] (PDF levels), "bg" [N,3] (random
    background).  rand=None -> eval (deterministic samplers, last_sample background)."""
    weights_list, samples_list, inds_list = [], [], []
    n_prop = len(mp.proposals)
    smp, weights = None, None
    for lvl in range(n_prop + 1):
        is_prop = lvl < n_prop
        s = mp.num_proposal_samples[lvl] if is_prop else mp.num_nerf_samples
        if lvl == 0:
            smp = uniform_sampler(origins, directions, nears, fars, times, s, None if rand is None else rand["t_rand"])
        else:
            annealed = torch.pow(weights, anneal)  # ray_samplers.py:584
            smp, inds = pdf_sampler(smp, annealed[..., 0], s, None if rand is None else rand[f"u{lvl}"])
            inds_list.append(inds)
        if is_prop:
            density = density_field(mp.proposals[lvl], smp.positions(), times)
            weights = get_weights(smp.deltas[..., None], density)
            weights_list.append(weights)
            samples_list.append(smp)
    density, rgb, feats = field_forward(mp.field, smp)
    weights = get_weights(smp.deltas[..., None], density)
    weights_list.append(weights)
    samples_list.append(smp)
    if training:
        bg = rand["bg"] if background == "random" else background
    else:
        bg = "last_sample" if background == "random" else background
    out = {
        "rgb": render_rgb(rgb, weights, bg, training),
        "accumulation": render_accumulation(weights),
        "depth": render_depth_median(weights, smp.steps()),
        "median_rgb": render_median_rgb(rgb, weights, training),
        "weights_list": weights_list,
        "samples_list": samples_list,
        "inds_list": inds_list,
        "density": density,
        "rgb_samples": rgb,
        "features": feats,
    }
    for i in range(n_prop):
        out[f"prop_depth_{i}"] = render_depth_median(weights_list[i], samples_list[i].steps())
    return out


def model_loss_dict(mp: ModelParams, out: Dict, image: torch.Tensor, training: bool = True) -> Dict[str, torch.Tensor]:
    """KPlanesModel.get_loss_dict, NS/models/kplanes.py:414-452 (+ misc.scale_dict, NS/utils/misc.py:116-129)."""
    ld = {"rgb_loss": F.mse_loss(image, out["rgb"])}
    if training:
        sdists = [s.spacing_bins for s in out["samples_list"]]
        ld["distortion_loss"] = distortion_loss(out["weights_list"], sdists)
        ld["interlevel_loss"] = interlevel_loss(out["weights_list"], sdists)
        nerf = mp.field.grids
        prop = [p.grids for p in mp.proposals]
        ld["space_tv_loss"] = space_tv_loss(nerf)
        ld["space_tv_proposal_loss"] = space_tv_loss(prop)
        if len(nerf[0]) == 6:
            ld["sparse_transients_loss"] = sparse_transients_loss(nerf)
            ld["sparse_transients_proposal_loss"] = sparse_transients_loss(prop)
            ld["time_smoothness_loss"] = time_smoothness_loss(nerf)
            ld["time_smoothness_proposal_loss"] = time_smoothness_loss(prop)
    return {k: v * mp.loss_coefficients[k] if k in mp.loss_coefficients else v for k, v in ld.items()}


# --------------------------------------------------------------------------------------------
# Synthetic scene shapes -- shared by tests, smoke and bench so inputs are identical
# --------------------------------------------------------------------------------------------
def undistort(coords: torch.Tensor, k: torch.Tensor, eps: float = 1e-3, max_iterations: int = 10) -> torch.Tensor:
    """radial_and_tangential_undistort (NS/cameras/camera_utils.py:363-401): Newton's method on the OpenCV forward model
    (residual and Jacobian of _compute_residual_and_jacobian, :298-360), started at the distorted point, a step only
    where |det J| > eps.  coords [..., 2], k [..., 6] = (k1, k2, k3, k4, p1, p2), broadcastable."""
    k1, k2, k3, k4, p1, p2 = (k[..., i] for i in range(6))
    xd, yd = coords[..., 0], coords[..., 1]
    x, y = xd, yd
    for _ in range(max_iterations):
        r = x * x + y * y
        d = 1.0 + r * (k1 + r * (k2 + r * (k3 + r * k4)))
        fx = d * x + 2 * p1 * x * y + p2 * (r + 2 * x * x) - xd
        fy = d * y + 2 * p2 * x * y + p1 * (r + 2 * y * y) - yd
        d_r = k1 + r * (2.0 * k2 + r * (3.0 * k3 + r * 4.0 * k4))
        d_x, d_y = 2.0 * x * d_r, 2.0 * y * d_r
        fx_x = d + d_x * x + 2.0 * p1 * y + 6.0 * p2 * x
        fx_y = d_y * x + 2.0 * p1 * x + 2.0 * p2 * y
        fy_x = d_x * y + 2.0 * p2 * y + 2.0 * p1 * x
        fy_y = d + d_y * y + 2.0 * p2 * x + 6.0 * p1 * y
        den = fy_x * fx_y - fx_x * fy_y
        ok = torch.abs(den) > eps
        x = x + torch.where(ok, (fx * fy_y - fy * fx_y) / den, torch.zeros_like(den))
        y = y + torch.where(ok, (fy * fx_x - fx * fy_x) / den, torch.zeros_like(den))
    return torch.stack([x, y], dim=-1)


def intersect_aabb(origins: torch.Tensor, directions: torch.Tensor, aabb6: torch.Tensor):
    """nerfstudio.utils.math._intersect_aabb (NS/utils/math.py:201-238) with intersect_aabb's constants (:267-270,
    max_bound = invalid_value = 1e10): what Cameras.generate_rays(aabb_box=...) stores as nears / fars
    (cameras.py:478-497).  origins / directions [N,3], aabb6 [6] = min then max -> t_min [N], t_max [N]."""
    tx_min = (aabb6[:3] - origins) / directions
    tx_max = (aabb6[3:] - origins) / directions
    t_min = torch.max(torch.min(tx_min, tx_max), dim=-1).values
    t_max = torch.min(torch.max(tx_min, tx_max), dim=-1).values
    t_min = torch.clamp(t_min, min=0, max=1e10)
    t_max = torch.clamp(t_max, min=0, max=1e10)
    miss = t_max <= t_min
    return torch.where(miss, 1e10, t_min), torch.where(miss, 1e10, t_max)


CAMERA_PERSPECTIVE, CAMERA_FISHEYE, CAMERA_EQUIRECTANGULAR = 1, 2, 3  # CameraType, NS/cameras/cameras.py:42-47


def generate_rays(c2w, fx, fy, cx, cy, cam_times, cam_idx, y_idx, x_idx, pixel_offset: float = 0.5,
                  distortion_params=None, camera_type=None):
    """Cameras._generate_rays_from_coords (NS/cameras/cameras.py:505-741), for rays given as (camera, row, col) like
    RayGenerator.forward (NS/model_components/ray_generators.py:43-59).
    c2w [C,3,4]; fx,fy,cx,cy [C,1]; cam_times [C,1] | None; indices int64 [N]; distortion_params [C,6] | None (OpenCV
    k1,k2,k3,k4,p1,p2, undone for every non-equirectangular camera, :635-654); camera_type int [C,1] | None (all
    perspective): per-type direction models :665-697.
    -> origins [N,3], directions [N,3], pixel_area [N,1], directions_norm [N,1], times [N,1] | None."""
    y = y_idx.float() + pixel_offset  # get_image_coords(pixel_offset=0.5), cameras.py:299-326
    x = x_idx.float() + pixel_offset
    fx_, fy_, cx_, cy_ = (t[cam_idx, 0] for t in (fx, fy, cx, cy))
    coord = torch.stack([(x - cx_) / fx_, -(y - cy_) / fy_], -1)  # cameras.py:624-627
    coord_x = torch.stack([(x - cx_ + 1) / fx_, -(y - cy_) / fy_], -1)
    coord_y = torch.stack([(x - cx_) / fx_, -(y - cy_ + 1) / fy_], -1)
    coord_stack = torch.stack([coord, coord_x, coord_y], dim=0)  # [3,N,2]
    ctype = None if camera_type is None else camera_type.reshape(-1)[cam_idx]  # [N]
    if distortion_params is not None:
        und = undistort(coord_stack, distortion_params[cam_idx][None])
        keep = None if ctype is None else (ctype == CAMERA_EQUIRECTANGULAR)
        coord_stack = und if keep is None else torch.where(keep[None, :, None], coord_stack, und)
    dirs = torch.cat([coord_stack, -torch.ones_like(coord_stack[..., :1])], dim=-1)  # perspective: (u, v, -1), :665-670
    if ctype is not None:
        theta = torch.clip(torch.sqrt(torch.sum(coord_stack**2, dim=-1)), 0.0, math.pi)  # fisheye, :672-683
        st = torch.sin(theta)
        fish = torch.stack([coord_stack[..., 0] * st / theta, coord_stack[..., 1] * st / theta, -torch.cos(theta)], dim=-1)
        lon = -torch.pi * coord_stack[..., 0]  # equirectangular, :685-697
        lat = torch.pi * (0.5 - coord_stack[..., 1])
        equi = torch.stack([-torch.sin(lon) * torch.sin(lat), torch.cos(lat), -torch.cos(lon) * torch.sin(lat)], dim=-1)
        sel = ctype[None, :, None]
        dirs = torch.where(sel == CAMERA_FISHEYE, fish, torch.where(sel == CAMERA_EQUIRECTANGULAR, equi, dirs))
    c2w_r = c2w[cam_idx]  # [N,3,4]
    rotation = c2w_r[..., :3, :3]
    dirs = torch.sum(dirs[..., None, :] * rotation, dim=-1)  # :708-710
    eps = torch.tensor([np.finfo(float).eps * 4.0]).to(dirs)  # camera_utils._EPS, camera_utils.py:28
    norm = torch.maximum(torch.linalg.vector_norm(dirs, dim=-1, keepdims=True), eps)  # normalize_with_norm :240-252
    dirs = dirs / norm
    origins = c2w_r[..., :3, 3]
    directions = dirs[0]
    dx = torch.sqrt(torch.sum((directions - dirs[1]) ** 2, dim=-1))  # :722-723
    dy = torch.sqrt(torch.sum((directions - dirs[2]) ** 2, dim=-1))
    pixel_area = (dx * dy)[..., None]
    times = cam_times[cam_idx, 0][..., None] if cam_times is not None else None
    return origins, directions, pixel_area, norm[0], times


def synthetic_rays(n: int, gen: torch.Generator, scene: str = "broadcast", n_frames: int = 25):
    """Broadcast-style scene: cameras on a ring of radius ~1 around the origin looking inward, aabb
    +-1.5 (broadcaststyle_dataparser.py:449-463); stadium: aabb +-1.  Returns origins, directions
    (unit), times [N,1] drawn from the fps-downsampled frame ids / 99, and the aabb."""
    scale = 1.5 if scene == "broadcast" else 1.0
    n_cams = 19 if scene == "broadcast" else 30
    aabb = torch.tensor([[-scale] * 3, [scale] * 3])
    cam = torch.randint(0, n_cams, (n,), generator=gen)
    ang = cam.float() / n_cams * 6.283185307179586
    origins = torch.stack([torch.cos(ang), torch.sin(ang), torch.full_like(ang, 0.35)], dim=-1)
    target = (torch.rand(n, 3, generator=gen) - 0.5) * torch.tensor([1.6, 1.6, 0.6])
    d = target - origins
    directions = d / d.norm(dim=-1, keepdim=True)
    frame_ids = torch.linspace(0, 99, n_frames).to(torch.int32).float()
    times = (frame_ids[torch.randint(0, n_frames, (n,), generator=gen)] / 99.0)[:, None]
    return origins, directions, times, aabb


def make_model_params(cfg: str, gen: torch.Generator, aabb: torch.Tensor, smooth_res: int = 0) -> ModelParams:
    """BASELINE.json configs resolved as in SURVEY.md 8 table.  ``smooth_res`` > 0: band-limited planes (noise drawn
    at that spatial resolution and bilinearly upsampled) instead of per-texel white noise -- the conditioning of a
    TV-regularised, trained field rather than of the random initialisation."""
    if cfg == "cfg1":
        res, ms, hid, vd, nerf_s, prop_t = (64, 64, 64, 16), (1, 2, 4), 64, True, 48, 16
    elif cfg == "cfg2":
        res, ms, hid, vd, nerf_s, prop_t = (64, 64, 64, 50), (1, 2, 4, 8), 64, True, 48, 150
    elif cfg == "cfg3":
        res, ms, hid, vd, nerf_s, prop_t = (64, 64, 64, 100), (1, 2, 4, 8, 16, 32), 128, False, 64, 100
    elif cfg == "tiny":
        res, ms, hid, vd, nerf_s, prop_t = (16, 16, 16, 6), (1, 2), 64, True, 16, 6
    else:
        raise ValueError(cfg)
    fieldp = make_field_params(aabb, res, 32, ms, gen, sigma_hidden=hid, view_dependent=vd, smooth_res=smooth_res)
    if cfg == "tiny":
        props = [make_density_params(aabb, [24, 24, 24, prop_t], 8, gen), make_density_params(aabb, [32, 32, 32, prop_t], 8, gen)]
        nprop = (32, 24)
    else:
        props = [make_density_params(aabb, [128, 128, 128, prop_t], 8, gen, smooth_res=smooth_res),
                 make_density_params(aabb, [256, 256, 256, prop_t], 8, gen, smooth_res=smooth_res)]
        nprop = (256, 128)
    return ModelParams(field=fieldp, proposals=props, num_proposal_samples=nprop, num_nerf_samples=nerf_s)


def make_rand(n: int, mp: ModelParams, gen: torch.Generator) -> Dict[str, torch.Tensor]:
    """The reference's training-mode torch.rand draws, in the order it makes them
    (ray_samplers.py:106, :319 per PDF level, renderers.py:104-105)."""
    rand = {"t_rand": torch.rand(n, mp.num_proposal_samples[0] + 1, generator=gen)}
    sizes = list(mp.num_proposal_samples[1:]) + [mp.num_nerf_samples]
    for lvl, s in enumerate(sizes, start=1):
        rand[f"u{lvl}"] = torch.rand(n, s + 1, generator=gen)
    rand["bg"] = torch.rand(n, 3, generator=gen)
    return rand


def train_step(mp: ModelParams, origins, directions, times, image, rand, anneal=1.0, near_plane=0.0):
    """One fwd+bwd of the reference training step on CPU (Model.forward = collider + get_outputs,
    then get_loss_dict and loss.backward()).  Returns (outputs, loss_dict, grads aligned with
    mp.tensors())."""
    params = mp.tensors()
    for p in params:
        p.requires_grad_(True)
        p.grad = None
    nears, fars = aabb_collider(origins, directions, mp.field.aabb, near_plane)
    out = model_forward(mp, origins, directions, times, nears, fars, rand, anneal=anneal, training=True)
    ld = model_loss_dict(mp, out, image)
    loss = sum(ld.values())
    loss.backward()
    grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
    return out, ld, grads
