"""a18: IST/ISG weight maps and importance pixel sampling vs the fixture produced by the reference's own code
(oracle/make_golden.py::gen_importance).  Host-side logic: runs on CPU.  Indices are bit-exact given the seeds."""
import random

import torch

from soccernerfs_b200.data.datamanagers.dynamic_datamanager import (
    DynamicDataManagerConfig,
    importance_weights,
    make_pixel_sampler,
)
from soccernerfs_b200.data.pixel_samplers import DynamicBasedPixelSampler, PixelSampler
from tests.conftest import load_golden


def test_datamanager_option_defaults_match_reference():
    """NS/data/datamanagers/dynamic_datamanager.py:34-59."""
    c = DynamicDataManagerConfig()
    assert (c.use_importance_sampling, c.is_pixel_ratio, c.ist_range, c.isg, c.isg_gamma, c.iters_to_start_is, c.pick_mode) == (
        True, 0.1, 0.25, False, 5e-2, 5000, "normal")


def test_ist_and_isg_weight_maps_match_reference():
    g = load_golden("importance")
    cfg = DynamicDataManagerConfig(ist_range=0.25)
    ist = importance_weights(cfg, g["images"], g["cam_ids"], g["cam_times"])
    assert ist.dtype == torch.float16 and torch.equal(ist.float(), g["ist"])
    isg = importance_weights(DynamicDataManagerConfig(isg=True), g["images"], g["cam_ids"], g["cam_times"])
    assert torch.equal(isg.float(), g["isg"])
    assert importance_weights(DynamicDataManagerConfig(use_importance_sampling=False), g["images"], g["cam_ids"], g["cam_times"]) is None
    # image 11 is the only frame of camera 3 -> uniform map; frames closer than 0.01 are not compared
    assert bool((g["ist"][11] == 1).all())


def test_importance_pixel_sampler_bit_exact():
    g = load_golden("importance")
    cfg = DynamicDataManagerConfig(is_pixel_ratio=0.3, iters_to_start_is=100)
    sampler = make_pixel_sampler(cfg, 64)
    assert isinstance(sampler, DynamicBasedPixelSampler)
    b = g["images"].shape[0]
    for name, steps, weights in (("ist_on", 500, g["ist"]), ("ist_off", 50, g["ist"]), ("no_weights", 500, None)):
        torch.manual_seed(1234)
        random.seed(99)
        batch = {"image": g["images"], "image_idx": torch.arange(b) + 100, "iter_steps": steps, "ist_weights": weights}
        col = sampler.collate_image_dataset_batch(batch, 64)
        assert torch.equal(col["indices"], g[f"{name}_indices"]), name
        assert torch.equal(col["image"], g[f"{name}_image"]), name
    # with IST on, the importance share of the batch lands on moving pixels only
    idx = g["ist_on_indices"]
    n_ist = int(0.3 * 64)
    w = g["ist"][idx[:n_ist, 0] - 100, idx[:n_ist, 1], idx[:n_ist, 2]]
    assert bool((w > 0).all())
    torch.manual_seed(4321)
    h, wd = g["images"].shape[1:3]
    assert torch.equal(PixelSampler(32).sample_method(32, b, h, wd), g["uniform"])
    assert isinstance(make_pixel_sampler(DynamicDataManagerConfig(use_importance_sampling=False), 8), PixelSampler)


def test_temporal_neighbour_lists():
    """The CSR neighbour lists handed to kp_ist_map: same camera, 0.01 < |dt| <= ist_range (host logic)."""
    from soccernerfs_b200.data.dynamic_dataset import temporal_neighbours

    cam_ids = torch.tensor([0, 0, 0, 1, 0])
    cam_times = torch.tensor([0.0, 0.1, 0.105, 0.1, 0.5])
    off, nb = temporal_neighbours(cam_ids, cam_times, 0.25)
    assert off.tolist() == [0, 2, 3, 4, 4, 4]  # image 1 and 2 are only 0.005 apart; image 3 is another camera; 4 too far
    assert nb.tolist() == [1, 2, 0, 0]
    assert off.dtype == torch.int32 and nb.dtype == torch.int32


import pytest  # noqa: E402


@pytest.mark.gpu
def test_ist_map_kernel_bit_exact_vs_reference_fixture():
    """(f4) kp_ist_map == the IST map the reference's own compute_ist produced (fp16, every pixel identical)."""
    from soccernerfs_b200.data.dynamic_dataset import compute_ist

    g = load_golden("importance")
    ist = compute_ist(g["images"].cuda(), g["cam_ids"], g["cam_times"], 0.25)
    assert ist.is_cuda and ist.dtype == torch.float16 and ist.shape == g["ist"].shape
    assert torch.equal(ist.float().cpu(), g["ist"])
    assert bool((ist[11] == 1).all())  # the only frame of its camera: uniform map


@pytest.mark.gpu
def test_isg_map_kernel_bit_exact_vs_reference_fixture():
    """(f4) kp_isg_map == the ISG map the reference's own compute_isg produced (per-camera median + Geman-McClure, fp16)."""
    from soccernerfs_b200.data.dynamic_dataset import compute_isg

    g = load_golden("importance")
    isg = compute_isg(g["images"].cuda(), g["cam_ids"], 5e-2)
    assert isg.is_cuda and isg.dtype == torch.float16 and isg.shape == g["isg"].shape
    assert torch.equal(isg.float().cpu(), g["isg"])
    # a larger random case (even and odd frame counts per camera) against the host implementation
    gen = torch.Generator().manual_seed(7)
    images = torch.rand(23, 20, 31, 3, generator=gen)
    cam_ids = torch.tensor([0] * 8 + [1] * 9 + [2] * 5 + [3])
    assert torch.equal(compute_isg(images.cuda(), cam_ids, 5e-2).cpu(), compute_isg(images, cam_ids, 5e-2))


def test_equirectangular_and_patch_samplers_match_reference():
    """EquirectangularPixelSampler / PatchPixelSampler (NS/data/pixel_samplers.py:228-327) vs the reference's own classes
    under the same torch seed (fixture pixel_samplers): identical indices; and the sampler choice of
    DynamicDataManager._get_pixel_sampler (dynamic_datamanager.py:97-113)."""
    import types

    from soccernerfs_b200.data.datamanagers.dynamic_datamanager import DynamicDataManagerConfig, make_pixel_sampler
    from soccernerfs_b200.data.pixel_samplers import (DynamicBasedPixelSampler, EquirectangularPixelSampler, PatchPixelSampler,
                                                      PixelSampler)

    g = load_golden("pixel_samplers")
    torch.manual_seed(2468)
    eq = EquirectangularPixelSampler(96).sample_method(96, 7, 40, 80)
    assert eq.dtype == torch.int64 and torch.equal(eq, g["equirect"])
    assert int(eq[:, 1].max()) < 40 and int(eq[:, 2].max()) < 80 and int(eq[:, 0].max()) < 7
    torch.manual_seed(1357)
    patch = PatchPixelSampler(100, patch_size=4)
    assert patch.num_rays_per_batch == int(g["patch_rays"]) == 96  # whole 4x4 patches only
    idx = patch.sample_method(patch.num_rays_per_batch, 5, 30, 50)
    assert torch.equal(idx, g["patch"])
    blocks = idx.view(-1, 4, 4, 3)
    assert bool((blocks[..., 0] == blocks[:, :1, :1, 0]).all())  # one image per patch, rows / columns contiguous
    assert torch.equal(blocks[:, :, 0, 1] - blocks[:, :1, 0, 1], torch.arange(4).expand(blocks.shape[0], 4))
    assert torch.equal(blocks[:, 0, :, 2] - blocks[:, 0, :1, 2], torch.arange(4).expand(blocks.shape[0], 4))
    patch.set_num_rays_per_batch(50)
    assert patch.num_rays_per_batch == int(g["patch_rays_after_set"]) == 48
    cams = lambda t: types.SimpleNamespace(camera_type=torch.tensor(t)[:, None])  # noqa: E731
    cfg = DynamicDataManagerConfig()
    assert type(make_pixel_sampler(cfg, 64, cameras=cams([3, 3]))) is EquirectangularPixelSampler
    assert type(make_pixel_sampler(cfg, 64, cameras=cams([1, 3]))) is DynamicBasedPixelSampler
    assert type(make_pixel_sampler(DynamicDataManagerConfig(patch_size=2), 64, cameras=cams([3, 3]))) is PatchPixelSampler
    assert type(make_pixel_sampler(DynamicDataManagerConfig(use_importance_sampling=False), 64)) is PixelSampler
