"""Ray samplers on the B200 kernels: SpacedSampler / UniformSampler / UniformLinDispPiecewiseSampler,
PDFSampler and ProposalNetworkSampler.

Same classes, constructor arguments, state (`_anneal`, `_steps_since_update`, `_step`) and call surface as
NS/model_components/ray_samplers.py:31-150, 221-369, 510-600.  Random numbers are still drawn with
``torch.rand`` in exactly the reference's order and shapes (so a shared seed reproduces the reference's
sample positions on the same device); the arithmetic runs in ``kp_uniform_bins`` / ``kp_pdf_resample``.
The other spacing functions of the reference (LinearDisparity, Sqrt, Log) are not used by K-Planes /
nerfplayer-nerfacto and are not built.
"""
from __future__ import annotations

import contextlib
import functools
from abc import abstractmethod
from typing import Callable, List, Optional, Tuple

import torch
from torch import nn

from .. import ops
from ..cameras.rays import RayBundle, RaySamples

_UNIFORM, _PIECEWISE = 0, 1


def _spacing_fns(mode: int):
    if mode == _UNIFORM:
        return (lambda x: x), (lambda x: x)
    return (lambda x: torch.where(x < 1, x / 2, 1 - 1 / (2 * x))), (lambda x: torch.where(x < 0.5, 2 * x, 1 / (2 - 2 * x)))


class Sampler(nn.Module):
    """Generate samples (ray_samplers.py:31-51)."""

    def __init__(self, num_samples: Optional[int] = None) -> None:
        super().__init__()
        self.num_samples = num_samples

    @abstractmethod
    def generate_ray_samples(self) -> RaySamples:
        """Generate ray samples."""

    def forward(self, *args, **kwargs) -> RaySamples:
        return self.generate_ray_samples(*args, **kwargs)


def _to_ray_samples(ray_bundle: RayBundle, spacing_bins, euclid_bins, spacing_to_euclidean_fn, frustums=None) -> RaySamples:
    """``frustums`` = contiguous (starts, ends, deltas) [N,S] written by the sampler kernel; without it they are
    strided views of ``euclid_bins`` like in the reference."""
    shape = ray_bundle.origins.shape[:-1]
    sb = spacing_bins.view(*shape, -1)
    if frustums is None:
        eb = euclid_bins.view(*shape, -1)
        starts, ends, deltas = eb[..., :-1, None], eb[..., 1:, None], None
    else:
        starts, ends, deltas = (x.view(*shape, -1, 1) for x in frustums)
    rs = ray_bundle.get_ray_samples(
        bin_starts=starts,
        bin_ends=ends,
        spacing_starts=sb[..., :-1, None],
        spacing_ends=sb[..., 1:, None],
        spacing_to_euclidean_fn=spacing_to_euclidean_fn,
        deltas=deltas,
    )
    rs._kp_sdist = sb  # the [.., S+1] edges spacing_starts/ends are views of (saves re-concatenating them)
    return rs


def spacing_edges(ray_samples: RaySamples) -> torch.Tensor:
    """[..., S+1] spacing-domain bin edges: cat(spacing_starts, spacing_ends[-1]) (ray_samplers.py:331-338,
    losses.py:98-103), taken from the sampler's own edge tensor when the samples came from one of ours."""
    sb = getattr(ray_samples, "_kp_sdist", None)
    if sb is not None and sb.shape[:-1] == ray_samples.spacing_starts.shape[:-2]:
        return sb
    return torch.cat([ray_samples.spacing_starts[..., 0], ray_samples.spacing_ends[..., -1:, 0]], dim=-1)


class SpacedSampler(Sampler):
    """Stratified bins in a spacing domain (ray_samplers.py:54-126).  ``spacing_mode``: 0 uniform, 1 piecewise
    uniform / linear-in-disparity -- the two spacings the K-Planes and nerfplayer-nerfacto models use."""

    def __init__(self, spacing_mode: int = _UNIFORM, num_samples: Optional[int] = None, train_stratified=True,
                 single_jitter=False) -> None:
        super().__init__(num_samples=num_samples)
        self.train_stratified = train_stratified
        self.single_jitter = single_jitter
        self.spacing_mode = spacing_mode
        self.spacing_fn, self.spacing_fn_inv = _spacing_fns(spacing_mode)

    def generate_ray_samples(self, ray_bundle: Optional[RayBundle] = None, num_samples: Optional[int] = None) -> RaySamples:
        assert ray_bundle is not None
        assert ray_bundle.nears is not None
        assert ray_bundle.fars is not None
        num_samples = num_samples or self.num_samples
        assert num_samples is not None
        num_rays = ray_bundle.origins.shape[0]
        device = ray_bundle.origins.device
        t_rand = None
        if self.train_stratified and self.training:  # same draw as ray_samplers.py:104-108
            if self.single_jitter:
                t_rand = torch.rand((num_rays, 1), dtype=torch.float32, device=device)
            else:
                t_rand = torch.rand((num_rays, num_samples + 1), dtype=torch.float32, device=device)
        sb, eb, *fr = ops.uniform_bins(ray_bundle.nears, ray_bundle.fars, num_samples, t_rand, self.spacing_mode,
                                       want_frustums=True)
        s_near, s_far = (self.spacing_fn(x) for x in (ray_bundle.nears, ray_bundle.fars))
        spacing_to_euclidean_fn = lambda x: self.spacing_fn_inv(x * s_far + (1 - x) * s_near)  # noqa: E731
        # tag the closure so PDFSampler can evaluate the same function inside its kernel instead of calling it
        spacing_to_euclidean_fn._kp_spacing = (self.spacing_mode, ray_bundle.nears, ray_bundle.fars)
        return _to_ray_samples(ray_bundle, sb, eb, spacing_to_euclidean_fn, fr)


class UniformSampler(SpacedSampler):
    """ray_samplers.py:129-150."""

    def __init__(self, num_samples: Optional[int] = None, train_stratified=True, single_jitter=False) -> None:
        super().__init__(_UNIFORM, num_samples, train_stratified, single_jitter)


class UniformLinDispPiecewiseSampler(SpacedSampler):
    """ray_samplers.py:221-246."""

    def __init__(self, num_samples: Optional[int] = None, train_stratified=True, single_jitter=False) -> None:
        super().__init__(_PIECEWISE, num_samples, train_stratified, single_jitter)


class PDFSampler(Sampler):
    """Inverse-CDF resampling of a weight histogram (ray_samplers.py:249-369): one warp per ray."""

    def __init__(self, num_samples: Optional[int] = None, train_stratified: bool = True, single_jitter: bool = False,
                 include_original: bool = True, histogram_padding: float = 0.01) -> None:
        super().__init__(num_samples=num_samples)
        self.train_stratified = train_stratified
        self.include_original = include_original
        self.histogram_padding = histogram_padding
        self.single_jitter = single_jitter
        self.last_inds: Optional[torch.Tensor] = None  # searchsorted indices of the last call when record_inds
        self.record_inds = False

    def generate_ray_samples(self, ray_bundle: Optional[RayBundle] = None, ray_samples: Optional[RaySamples] = None,
                             weights: torch.Tensor = None, num_samples: Optional[int] = None, eps: float = 1e-5,
                             anneal=1.0) -> RaySamples:
        """``anneal`` (extension): exponent applied to ``weights`` inside the kernel -- the proposal sampler's
        torch.pow(weights, anneal) of ray_samplers.py:584 without a separate pass."""
        if ray_samples is None or ray_bundle is None:
            raise ValueError("ray_samples and ray_bundle must be provided")
        num_samples = num_samples or self.num_samples
        assert num_samples is not None
        assert ray_samples.spacing_starts is not None and ray_samples.spacing_ends is not None, \
            "ray_sample spacing_starts and spacing_ends must be provided"
        assert ray_samples.spacing_to_euclidean_fn is not None, "ray_samples.spacing_to_euclidean_fn must be provided"
        w = weights[..., 0]
        s_in = w.shape[-1]
        w = w.reshape(-1, s_in)
        n = w.shape[0]
        rand = None
        if self.train_stratified and self.training:  # same draw as ray_samplers.py:318-321
            if self.single_jitter:
                rand = torch.rand((n, 1), device=w.device)
            else:
                rand = torch.rand((n, num_samples + 1), device=w.device)
        existing_bins = spacing_edges(ray_samples)
        tag = getattr(ray_samples.spacing_to_euclidean_fn, "_kp_spacing", None)
        kernel_euclid = tag is not None and not self.include_original
        mode, nears, fars = tag if kernel_euclid else (0, torch.zeros(n, device=w.device), torch.ones(n, device=w.device))
        sb, eb, inds, _, *fr = ops.pdf_resample(w, existing_bins.reshape(n, s_in + 1), nears, fars, num_samples, rand,
                                                self.histogram_padding, eps, spacing=mode, want_inds=self.record_inds,
                                                anneal=anneal, want_frustums=kernel_euclid)
        self.last_inds = inds
        if self.include_original:
            sb, _ = torch.sort(torch.cat([existing_bins.reshape(n, -1), sb], -1), -1)
        if not kernel_euclid:  # foreign spacing function: evaluate its closure like the reference (:359)
            shape = ray_bundle.origins.shape[:-1]
            eb = ray_samples.spacing_to_euclidean_fn(sb.view(*shape, -1))
            fr = None
        return _to_ray_samples(ray_bundle, sb, eb, ray_samples.spacing_to_euclidean_fn, fr)


def _density_field_of(fn: Callable):
    """If ``fn`` is (a functools.partial of) KPlanesDensityField.density_fn return (field, times) else None."""
    from ..fields.kplanes_field import KPlanesDensityField

    times, bound = None, fn
    if isinstance(fn, functools.partial):
        if fn.args or set(fn.keywords) - {"times"}:
            return None
        times, bound = fn.keywords.get("times"), fn.func
    field = getattr(bound, "__self__", None)
    if isinstance(field, KPlanesDensityField) and getattr(bound, "__func__", None) is KPlanesDensityField.density_fn:
        return field, times
    return None


class ProposalNetworkSampler(Sampler):
    """Proposal-network sampling loop (ray_samplers.py:510-600)."""

    def __init__(self, num_proposal_samples_per_ray: Tuple[int] = (64,), num_nerf_samples_per_ray: int = 32,
                 num_proposal_network_iterations: int = 2, single_jitter: bool = False,
                 update_sched: Callable = lambda x: 1, initial_sampler: Optional[Sampler] = None) -> None:
        super().__init__()
        self.num_proposal_samples_per_ray = num_proposal_samples_per_ray
        self.num_nerf_samples_per_ray = num_nerf_samples_per_ray
        self.num_proposal_network_iterations = num_proposal_network_iterations
        self.update_sched = update_sched
        if self.num_proposal_network_iterations < 1:
            raise ValueError("num_proposal_network_iterations must be >= 1")
        if initial_sampler is None:
            self.initial_sampler = UniformLinDispPiecewiseSampler(single_jitter=single_jitter)
        else:
            self.initial_sampler = initial_sampler
        self.pdf_sampler = PDFSampler(include_original=False, single_jitter=single_jitter)
        self._anneal = 1.0
        self._steps_since_update = 0
        self._step = 0
        # extension: a CUDA stream the proposal fields' density + weights are evaluated on.  Autograd replays a node's
        # backward on the stream its forward ran on, so the back-propagation through the proposal networks (which
        # depends only on the interlevel loss) then runs concurrently with the main field's backward.
        self.side_stream: Optional["torch.cuda.Stream"] = None

    def set_anneal(self, anneal: float) -> None:
        self._anneal = anneal

    def step_cb(self, step):
        self._step = step
        self._steps_since_update += 1

    def generate_ray_samples(self, ray_bundle: Optional[RayBundle] = None,
                             density_fns: Optional[List[Callable]] = None) -> Tuple[RaySamples, List, List]:
        assert ray_bundle is not None
        assert density_fns is not None
        weights_list, ray_samples_list = [], []
        n = self.num_proposal_network_iterations
        weights, ray_samples = None, None
        updated = self._steps_since_update > self.update_sched(self._step) or self._step < 10
        self.side_stream_used = False  # whether this call put work on ``side_stream`` (the trainer joins it only then)
        for i_level in range(n + 1):
            is_prop = i_level < n
            num_samples = self.num_proposal_samples_per_ray[i_level] if is_prop else self.num_nerf_samples_per_ray
            if i_level == 0:
                ray_samples = self.initial_sampler(ray_bundle, num_samples=num_samples)
            else:
                assert weights is not None
                # annealed_weights = pow(weights, _anneal) (ray_samplers.py:584) is applied inside the PDF kernel;
                # _anneal may be a device scalar tensor (CUDA-graph training step) or the reference's python float
                ray_samples = self.pdf_sampler(ray_bundle, ray_samples, weights, num_samples=num_samples, anneal=self._anneal)
            if is_prop:
                side = self.side_stream if (updated and torch.is_grad_enabled() and ray_bundle.origins.is_cuda) else None
                if side is not None:
                    main = torch.cuda.current_stream()
                    side.wait_stream(main)
                    self.side_stream_used = True
                with torch.cuda.stream(side) if side is not None else contextlib.nullcontext():
                    with torch.set_grad_enabled(updated and torch.is_grad_enabled()):
                        fast = _density_field_of(density_fns[i_level])
                        if fast is not None and (fast[1] is None or fast[1] is ray_bundle.times):
                            # same numbers as density_fn(get_positions()), without materialising positions
                            density, _ = fast[0].get_density(ray_samples)
                        else:
                            density = density_fns[i_level](ray_samples.frustums.get_positions())
                    weights = ray_samples.get_weights(density)
                if side is not None:
                    main.wait_stream(side)
                    weights.record_stream(main)
                weights_list.append(weights)
                ray_samples_list.append(ray_samples)
        if updated:
            self._steps_since_update = 0
        assert ray_samples is not None
        return ray_samples, weights_list, ray_samples_list
