"""The reference's own unit tests for this path, re-stated against our mirror classes with the same calls and scenarios
(REF/nerfstudio/tests/model_components/test_ray_sampler.py, test_renderers.py, tests/cameras/test_rays.py,
tests/utils/test_tensor_dataclass.py) -- a user of the reference should be able to run the code they already have.
The reference only checks shapes / loose bounds there; where it is cheap we also check the values against the oracle.
Samplers and renderers launch CUDA kernels (gpu marker); the ray containers are host logic."""
import pytest
import torch

from oracle import kplanes_oracle as ko
from soccernerfs_b200.cameras.rays import Frustums, RayBundle, RaySamples

DEV = "cuda"


def _bundle(n=10, device="cpu"):
    origins = torch.zeros((n, 3), device=device)
    return RayBundle(origins=origins, directions=torch.ones_like(origins), pixel_area=torch.ones((n, 1), device=device))


# ---- tests/cameras/test_rays.py ------------------------------------------------------------------------
def test_frustum_get_position():
    frustum = Frustums(origins=torch.tensor([[0.0, 1.0, 2.0]]), directions=torch.tensor([[0.0, 1.0, 0.0]]),
                       starts=torch.tensor([[2.0]]), ends=torch.tensor([[3.0]]), pixel_area=torch.ones((1, 1)))
    assert frustum.get_positions() == pytest.approx(torch.tensor([[0.0, 3.5, 2.0]]), abs=1e-6)


def test_frustum_apply_masks_and_mock():
    frustum = Frustums(origins=torch.ones((5, 3)), directions=torch.ones((5, 3)), starts=torch.ones((5, 1)),
                       ends=torch.ones((5, 1)), pixel_area=torch.ones((5, 1)))
    kept = frustum[torch.tensor([False, True, False, True, True])]
    assert kept.origins.shape == (3, 3) and kept.directions.shape == (3, 3)
    assert kept.starts.shape == (3, 1) and kept.ends.shape == (3, 1) and kept.pixel_area.shape == (3, 1)
    Frustums.get_mock_frustum()


# ---- tests/model_components/test_ray_sampler.py -----------------------------------------------------------
@pytest.mark.gpu
def test_uniform_sampler():
    from soccernerfs_b200.model_components.ray_samplers import UniformSampler
    from soccernerfs_b200.model_components.scene_colliders import NearFarCollider

    num_samples = 15
    sampler = UniformSampler(num_samples=num_samples)
    ray_bundle = NearFarCollider(near_plane=2, far_plane=4)(_bundle(device=DEV))
    ray_samples = sampler(ray_bundle)
    positions = ray_samples.frustums.get_positions()
    assert positions.shape[-2] == num_samples
    # every sample lies between the collider's planes along the ray, in order
    mid = (ray_samples.frustums.starts + ray_samples.frustums.ends) / 2
    assert bool((mid >= 2).all()) and bool((mid <= 4).all()) and bool((mid[:, 1:] >= mid[:, :-1]).all())
    # eval mode: the deterministic bins of the oracle, bit for bit
    sampler.eval()
    ev = sampler(ray_bundle)
    ref = ko.uniform_sampler(torch.zeros(10, 3), torch.ones(10, 3), torch.full((10, 1), 2.0), torch.full((10, 1), 4.0),
                             None, num_samples, None)
    assert torch.equal(ev.frustums.starts[..., 0].cpu(), ref.starts) and torch.equal(ev.frustums.ends[..., 0].cpu(), ref.ends)


@pytest.mark.gpu
def test_pdf_sampler():
    from soccernerfs_b200.model_components.ray_samplers import PDFSampler, UniformSampler
    from soccernerfs_b200.model_components.scene_colliders import NearFarCollider

    num_samples = 15
    ray_bundle = NearFarCollider(near_plane=2, far_plane=4)(_bundle(device=DEV))
    coarse = UniformSampler(num_samples=num_samples)(ray_bundle)
    weights = torch.ones((10, num_samples, 1), device=DEV)
    pdf_sampler = PDFSampler(num_samples)  # include_original=True, the reference's default
    fine = pdf_sampler(ray_bundle, coarse, weights, num_samples)
    assert fine.frustums.starts.shape == (10, 2 * num_samples + 1, 1)  # existing 16 edges + 16 new ones -> 31 bins
    edges = torch.cat([fine.frustums.starts[..., 0], fine.frustums.ends[:, -1:, 0]], -1)
    assert bool((edges[:, 1:] >= edges[:, :-1]).all()) and bool((edges >= 2 - 1e-5).all()) and bool((edges <= 4 + 1e-5).all())
    # without the original bins: exactly num_samples bins
    fine2 = PDFSampler(num_samples, include_original=False)(ray_bundle, coarse, weights, num_samples)
    assert fine2.frustums.starts.shape == (10, num_samples, 1)


# ---- tests/model_components/test_renderers.py -------------------------------------------------------------
@pytest.mark.gpu
def test_rgb_renderer():
    from soccernerfs_b200.model_components import renderers

    num_samples = 10
    rgb_samples = torch.ones((3, num_samples, 3), device=DEV)
    weights = torch.ones((3, num_samples, 1), device=DEV)
    weights /= torch.sum(weights, dim=-2, keepdim=True)
    rgb_renderer = renderers.RGBRenderer()
    rgb = rgb_renderer(rgb=rgb_samples, weights=weights)
    assert torch.max(rgb) > 0.9
    rgb = rgb_renderer(rgb=rgb_samples * 0, weights=weights)
    assert float(torch.max(rgb)) == pytest.approx(0, abs=1e-6)


@pytest.mark.gpu
def test_acc_renderer():
    from soccernerfs_b200.model_components import renderers

    weights = torch.ones((3, 10, 1), device=DEV)
    weights /= torch.sum(weights, dim=-2, keepdim=True)
    accumulation = renderers.AccumulationRenderer()(weights=weights)
    assert accumulation.shape == (3, 1) and torch.max(accumulation) > 0.9


@pytest.mark.gpu
def test_depth_renderer():
    from soccernerfs_b200.model_components import renderers

    num_samples = 10
    weights = torch.ones((num_samples, 1), device=DEV)  # a single unbatched ray, like the reference's test
    weights /= torch.sum(weights, dim=-2, keepdim=True)
    frustums = Frustums.get_mock_frustum(device=DEV)
    frustums.starts = torch.linspace(0, 100, num_samples, device=DEV)[..., None]
    frustums.ends = torch.linspace(1, 101, num_samples, device=DEV)[..., None]
    ray_samples = RaySamples(frustums=frustums, camera_indices=torch.ones((num_samples, 1), device=DEV),
                             deltas=torch.ones((num_samples, 1), device=DEV))
    steps = (frustums.starts + frustums.ends) / 2
    depth = renderers.DepthRenderer(method="median")(weights=weights, ray_samples=ray_samples)
    assert torch.min(depth) > 0
    assert torch.equal(depth.cpu(), ko.render_depth_median(weights.cpu()[None], steps.cpu()[None, :, 0])[0])
    depth = renderers.DepthRenderer(method="expected")(weights=weights, ray_samples=ray_samples)
    assert torch.min(depth) > 0
    assert float(depth) == pytest.approx(float(ko.render_depth_expected(weights.cpu()[None], steps.cpu()[None, :, 0])), rel=1e-6)


# ---- tests/utils/test_tensor_dataclass.py ------------------------------------------------------------------
def test_tensor_dataclass_broadcast_reshape_index():
    """The container semantics RayBundle / RaySamples / Frustums rely on (NS/utils/tensor_dataclass.py): batch-shape
    broadcasting at construction (nested dataclasses and dict fields included), reshape / flatten / indexing."""
    from dataclasses import dataclass
    from typing import Dict

    from soccernerfs_b200.utils.tensor_dataclass import TensorDataclass

    @dataclass
    class Nested(TensorDataclass):
        x: torch.Tensor

    @dataclass
    class Holder(TensorDataclass):
        a: torch.Tensor
        b: torch.Tensor
        c: Nested = None
        d: Dict = None

    @dataclass
    class OnlyOptional(TensorDataclass):
        vals: torch.Tensor = None

    OnlyOptional(vals=torch.ones(1))
    with pytest.raises(ValueError):
        OnlyOptional()
    assert Holder(a=torch.ones((4, 6, 3)), b=torch.ones((6, 2))).b.shape == (4, 6, 2)
    assert Holder(a=torch.ones((4, 6, 3)), b=torch.ones(2)).b.shape == (4, 6, 2)
    with pytest.raises(RuntimeError):
        Holder(a=torch.ones((4, 6, 3)), b=torch.ones((3, 2)))
    t = Holder(a=torch.ones((4, 6, 3)), b=torch.ones((6, 2)), c=Nested(x=torch.ones((6, 5))),
               d={"t1": torch.ones((4, 6, 3)), "t2": {"t3": torch.ones((6, 4))}})
    assert t.shape == (4, 6) and t.size == 24 and t.ndim == 2 and len(t) == 4
    assert t.c.x.shape == (4, 6, 5) and t.d["t2"]["t3"].shape == (4, 6, 4)
    r = t.reshape((2, 12))
    assert r.shape == (2, 12) and r.a.shape == (2, 12, 3) and r.d["t2"]["t3"].shape == (2, 12, 4)
    f = t.flatten()
    assert f.shape == (24,) and f.b.shape == (24, 2) and f[0:4].shape == (4,)
    assert t[:, 1].shape == (4,) and t[:, 1].a.shape == (4, 3) and t[:, 1].d["t1"].shape == (4, 3)
    assert t[:, 0:2].shape == (4, 2) and t[:, 0:2].d["t2"]["t3"].shape == (4, 2, 4)
    assert t[..., 1].a.shape == (4, 3) and t[0].shape == (6,) and t[0, ...].a.shape == (6, 3)
    u = Holder(a=torch.ones((2, 3, 4, 5)), b=torch.ones((4, 5)), d={"t1": torch.ones((2, 3, 4, 5))})
    assert u[0, ...].shape == (3, 4) and u[0, ...].a.shape == (3, 4, 5)


# ---- tests/cameras/test_cameras.py:124-159 (the oracle's ray generation; the kernels are pinned to the same reference
#      rays numerically by tests/test_gpu_parity.py::test_lens_ray_generation_vs_reference_fixture) ------------------------
def test_equirectangular_camera():
    height = 100  # width is twice the height
    c2w = torch.eye(4)[None, :3, :]
    one = torch.ones(1, 1)
    yy, xx = torch.meshgrid(torch.arange(height), torch.arange(2 * height), indexing="ij")
    cam = torch.zeros(height * 2 * height, dtype=torch.int64)
    origins, directions, pixel_area, _, _ = ko.generate_rays(
        c2w, one * height, one * height, one * height, one * 0.5 * height, None, cam, yy.reshape(-1), xx.reshape(-1),
        camera_type=torch.tensor([[ko.CAMERA_EQUIRECTANGULAR]]))
    assert torch.allclose(origins[0], torch.tensor([0.0, 0.0, 0.0]))
    directions = directions.view(height, 2 * height, 3)
    threshold = 0.9
    x, y, z = torch.tensor([1.0, 0.0, 0.0]), torch.tensor([0.0, 1.0, 0.0]), torch.tensor([0.0, 0.0, 1.0])
    # top pixels point up
    assert directions[0, 0] @ y > threshold and directions[0, height] @ y > threshold and directions[0, -1] @ y > threshold
    # middle pixels point horizontally; the middle of the image is camera forwards
    assert directions[height // 2, 0] @ z > threshold and directions[height // 2, height // 2] @ -x > threshold
    assert directions[height // 2, height] @ -z > threshold and directions[height // 2, 3 * height // 2] @ x > threshold
    assert directions[height // 2, -1] @ z > threshold
    # bottom pixels point down
    assert directions[-1, 0] @ -y > threshold and directions[-1, height] @ -y > threshold and directions[-1, -1] @ -y > threshold
    assert bool((pixel_area > 0).all())
