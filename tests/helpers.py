"""Shared test helpers: build the CUDA-path model from oracle parameters, feed both the same random draws."""
from __future__ import annotations

import contextlib
from typing import Dict, List

import torch

from oracle import kplanes_oracle as ko
from soccernerfs_b200.cameras.rays import RayBundle
from soccernerfs_b200.data.scene_box import SceneBox
from soccernerfs_b200.models.kplanes import KPlanesModel, KPlanesModelConfig

CFG = {
    # name: (spacetime_resolution, multiscale, sigma_hidden, view_dependent, nerf_samples, prop resolutions, prop samples)
    "tiny": ((16, 16, 16, 6), (1, 2), 64, True, 16, ([24, 24, 24, 6], [32, 32, 32, 6]), (32, 24)),
    "cfg1": ((64, 64, 64, 16), (1, 2, 4), 64, True, 48, ([128, 128, 128, 16], [256, 256, 256, 16]), (256, 128)),
    "cfg2": ((64, 64, 64, 50), (1, 2, 4, 8), 64, True, 48, ([128, 128, 128, 150], [256, 256, 256, 150]), (256, 128)),
    "cfg3": ((64, 64, 64, 100), (1, 2, 4, 8, 16, 32), 128, False, 64, ([128, 128, 128, 100], [256, 256, 256, 100]), (256, 128)),
}


def model_config(cfg: str) -> KPlanesModelConfig:
    res, ms, hid, vd, nerf_s, prop_res, prop_s = CFG[cfg]
    return KPlanesModelConfig(
        spacetime_resolution=res, multiscale_res=ms, sigma_net_hidden_dim=hid, disable_viewing_dependent=not vd,
        num_nerf_samples_per_ray=nerf_s, num_proposal_samples_per_ray=prop_s,
        proposal_net_args_list=[{"feature_dim": 8, "resolution": list(r)} for r in prop_res],
    )


def model_params_in_oracle_order(model: KPlanesModel) -> List[torch.nn.Parameter]:
    """Parameters ordered like oracle ModelParams.tensors(): proposals (planes, w1, w2)..., field planes, sigma, color."""
    out = []
    for p in model.proposal_networks:
        out += list(p.grids) + list(p.sigma_net.weights)
    out += [q for gs in model.field.grids for q in gs] + list(model.field.sigma_net.weights) + list(model.field.color_net.weights)
    return out


def build_model(cfg: str, mp: ko.ModelParams, aabb: torch.Tensor, device) -> KPlanesModel:
    model = model_config(cfg).setup(scene_box=SceneBox(aabb=aabb.clone()), num_train_data=1).to(device)
    with torch.no_grad():
        for dst, src in zip(model_params_in_oracle_order(model), mp.tensors()):
            assert dst.shape == src.shape, (dst.shape, src.shape)
            dst.copy_(src.detach().to(device))
    return model


def ray_bundle(origins, directions, times, device) -> RayBundle:
    n = origins.shape[0]
    return RayBundle(origins=origins.to(device), directions=directions.to(device), pixel_area=torch.ones(n, 1, device=device),
                     times=None if times is None else times.to(device))


@contextlib.contextmanager
def rand_queue(queue, device):
    """Serve torch.rand calls of the CUDA-path samplers / renderer from ``queue`` (same draws as the oracle)."""
    real = torch.rand
    q = list(queue)

    def fake(*size, **kw):
        if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)):
            size = tuple(size[0])
        t = q.pop(0)
        assert tuple(t.shape) == tuple(size), (t.shape, size)
        return t.to(device)

    torch.rand = fake
    try:
        yield q
    finally:
        torch.rand = real


def rand_list(rand: Dict[str, torch.Tensor]) -> List[torch.Tensor]:
    keys = ["t_rand"] + sorted(k for k in rand if k.startswith("u")) + ["bg"]
    return [rand[k] for k in keys]


def train_step_cuda(model: KPlanesModel, origins, directions, times, image, rand, anneal: float, device, forced_bins=None):
    """collider + get_outputs + get_loss_dict + backward on the CUDA path.  Returns (outputs, loss_dict, grads).
    ``forced_bins``: per PDF level (spacing_bins, euclid_bins) [N,S+1] that REPLACE the bins the PDF kernel resampled
    ("given the reference's samples": the rest of the step then sees bit-identical sample positions)."""
    model.train()
    model.proposal_sampler.set_anneal(anneal)
    model.proposal_sampler.pdf_sampler.record_inds = True
    for p in model.parameters():
        p.grad = None
    rb = ray_bundle(origins, directions, times, device)
    inds = []
    orig = model.proposal_sampler.pdf_sampler.generate_ray_samples

    def rec(*a, **k):
        r = orig(*a, **k)
        inds.append(model.proposal_sampler.pdf_sampler.last_inds)
        if forced_bins is not None:
            from soccernerfs_b200.model_components.ray_samplers import _to_ray_samples

            sb, eb = (t.to(device).contiguous() for t in forced_bins[len(inds) - 1])
            r = _to_ray_samples(a[0], sb, eb, a[1].spacing_to_euclidean_fn, None)
        return r

    model.proposal_sampler.pdf_sampler.generate_ray_samples = rec
    try:
        with rand_queue(rand_list(rand), device):
            out = model(rb)
    finally:
        model.proposal_sampler.pdf_sampler.generate_ray_samples = orig
    ld = model.get_loss_dict(out, {"image": image.to(device)}, model.get_metrics_dict(out, {"image": image.to(device)}))
    loss = sum(ld.values())
    loss.backward()
    grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in model_params_in_oracle_order(model)]
    out["inds_list"] = inds
    return out, ld, grads
