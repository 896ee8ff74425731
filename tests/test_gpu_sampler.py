"""GPU tests of the on-device importance pixel sampler (kp_importance_pixels): its random stream vs
the numpy restatement (oracle/device_sampler.py), its distribution vs the multinomial it replaces
(NS/data/pixel_samplers.py:396-398), and DynamicBasedPixelSampler's device path vs its host (reference) path."""
import random
import types

import numpy as np
import pytest
import torch

from oracle import device_sampler as ds

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _maps(h, w, gen):
    """Eight weight maps: dense, 5 % non-zero, three non-zero pixels, exactly k = 10 non-zero, one pixel, fp16 subnormals
    next to large weights, a copy of the dense one (another image index = other keys), all-zero."""
    hw = h * w
    m = torch.zeros(8, hw)
    m[0] = torch.rand(hw, generator=gen) + 0.01
    idx = torch.randperm(hw, generator=gen)
    m[1, idx[: hw // 20]] = torch.rand(hw // 20, generator=gen) + 0.05
    m[2, idx[:3]] = torch.tensor([0.5, 1.0, 2.0])
    m[3, idx[:10]] = torch.rand(10, generator=gen) + 0.1
    m[4, idx[7]] = 0.3
    m[5, idx[:200]] = 6e-8
    m[5, idx[200:220]] = 1000.0
    m[6] = m[0]
    return m.half().view(8, h, w)


@pytest.mark.parametrize("h,w", [(40, 64), (41, 63)])  # 16-byte rows (vector loads) and ragged rows (scalar loads)
def test_device_sampler_follows_the_restatement(h, w):
    from soccernerfs_b200 import ops

    gen = torch.Generator().manual_seed(11)
    maps = _maps(h, w, gen)
    k = 10
    sel = torch.tensor([[0, k, 0], [1, k, 10], [2, k, 20], [3, k, 30], [4, 4, 40], [5, k, 44], [6, k, 54], [7, 3, 64], [1, 1, 67]],
                       dtype=torch.int32)
    seed = 0x1234_5678_9ABC_DEF
    out = ops.importance_pixels(maps.to(DEV), sel, k, 68, w, seed).cpu()
    again = ops.importance_pixels(maps.to(DEV), sel.to(DEV), k, 68, w, seed).cpu()
    assert torch.equal(out, again)  # deterministic, whatever order the atomics took
    flat = maps.view(8, -1).numpy()
    mismatched = 0
    for img, kk, first in sel.tolist():
        rows = out[first: first + kk]
        if img == 7:  # all-zero map
            assert bool((rows == -1).all())
            continue
        assert bool((rows[:, 0] == img).all())
        pix = (rows[:, 1] * w + rows[:, 2]).numpy()
        assert bool((rows[:, 1] >= 0).all() and (rows[:, 1] < h).all() and (rows[:, 2] >= 0).all() and (rows[:, 2] < w).all())
        assert bool((flat[img][pix] > 0).all())
        nnz = int(np.count_nonzero(flat[img]))
        if nnz >= kk:
            assert len(set(pix.tolist())) == kk  # without replacement
        ref = ds.sample_image(flat[img], img, kk, seed)
        if nnz >= kk:
            mismatched += 0 if np.array_equal(pix, ref) else 1
        else:  # with replacement: the same draws unless a cumulative sum sits within rounding of a target
            mismatched += 0 if np.array_equal(pix, ref) else 1
    # logf of the device and numpy's log differ in the last place on ~20 % of the keys; that moves a winner only when two
    # keys are within an ulp of each other
    assert mismatched <= 1, mismatched
    # the subnormal / huge map: the 10 draws are all from the 20 heavy pixels
    heavy = set(np.nonzero(flat[5] > 1)[0].tolist())
    rows = out[44:54]
    assert set((rows[:, 1] * w + rows[:, 2]).tolist()) <= heavy


def test_device_sampler_distribution():
    """Many images with the same 16-pixel map: single draws follow w / sum(w) (chi-square vs the exact probabilities), the
    first of 3 draws without replacement does too, the 3 are distinct, and maps with fewer non-zero pixels than draws are
    sampled with replacement in proportion to the weights."""
    from soccernerfs_b200 import ops

    w = torch.tensor([0.0, 0.5, 0.25, 0.0, 1.0, 0.125, 2.0, 0.0, 0.75, 0.375, 0.0, 1.5, 0.0, 0.0, 0.0, 0.0], dtype=torch.float16)
    p = (w.double() / w.double().sum()).numpy()
    n = 16384
    maps = w[None].expand(n, 16).contiguous().to(DEV)
    nz = p > 0
    for k in (1, 3):
        sel = torch.stack([torch.arange(n), torch.full((n,), k), torch.arange(n) * k], dim=-1).to(torch.int32)
        out = ops.importance_pixels(maps, sel, k, n * k, 4, seed=2024 + k).cpu().view(n, k, 3)
        pix = (out[..., 1] * 4 + out[..., 2]).numpy()
        assert bool((out[..., 0] == torch.arange(n)[:, None]).all())
        if k == 3:
            s = np.sort(pix, axis=1)
            assert bool((s[:, 0] < s[:, 1]).all() and (s[:, 1] < s[:, 2]).all())
        counts = np.bincount(pix[:, 0], minlength=16).astype(np.float64)
        assert counts[~nz].sum() == 0
        chi2 = float((((counts - n * p) ** 2)[nz] / (n * p[nz])).sum())
        assert chi2 < 40.0, (k, chi2)  # 7 degrees of freedom: P(chi2 > 40) ~ 1e-6
    w2 = torch.zeros(n, 16, dtype=torch.float16)
    w2[:, 3], w2[:, 9] = 1.0, 3.0
    sel = torch.stack([torch.arange(n), torch.full((n,), 5), torch.arange(n) * 5], dim=-1).to(torch.int32)
    out = ops.importance_pixels(w2.to(DEV), sel, 5, n * 5, 4, seed=5).cpu()
    pix = (out[:, 1] * 4 + out[:, 2]).numpy()
    assert set(np.unique(pix).tolist()) == {3, 9}
    assert abs(float((pix == 9).mean()) - 0.75) < 0.01


def test_dynamic_pixel_sampler_device_path_follows_the_host_walk():
    """DynamicBasedPixelSampler with CUDA weight maps: same image order / per-image counts / skipped empty maps as the host
    (reference) path under the same python seed, every importance pixel has a non-zero weight, the remainder is uniform
    in range, nothing leaves the device, and the collated batch is consistent."""
    from soccernerfs_b200.data.pixel_samplers import DynamicBasedPixelSampler

    gen = torch.Generator().manual_seed(3)
    b, h, w = 12, 36, 64
    maps = torch.zeros(b, h, w)
    for i in range(b):
        if i in (2, 7):
            continue  # cameras that see no motion
        y0, x0 = int(torch.randint(0, h - 8, (1,), generator=gen)), int(torch.randint(0, w - 8, (1,), generator=gen))
        maps[i, y0: y0 + 8, x0: x0 + 8] = torch.rand(8, 8, generator=gen) + 0.2
    maps = maps.half()
    state = types.SimpleNamespace(iters_to_start_ist=10, is_pixel_ratio=0.25)
    n = 256
    host = DynamicBasedPixelSampler(n, dataset=state, device_sampler=False)
    dev = DynamicBasedPixelSampler(n, dataset=state)
    random.seed(77)
    torch.manual_seed(77)
    ref = host.sample_method(n, b, h, w, batch={"ist_weights": maps, "iter_steps": 11})
    random.seed(77)
    out = dev.sample_method(n, b, h, w, batch={"ist_weights": maps.to(DEV), "iter_steps": 11})
    assert out.is_cuda and out.dtype == torch.int64 and out.shape == (n, 3)
    num_ist = 64
    o = out.cpu()
    assert torch.equal(o[:num_ist, 0], ref[:num_ist, 0])  # same images, same counts, in the same order
    assert not bool(((o[:num_ist, 0] == 2) | (o[:num_ist, 0] == 7)).any())
    assert bool((maps[o[:num_ist, 0], o[:num_ist, 1], o[:num_ist, 2]] > 0).all())
    for img in torch.unique(o[:num_ist, 0]).tolist():
        rows = o[:num_ist][o[:num_ist, 0] == img]
        assert len({(int(r[1]), int(r[2])) for r in rows}) == rows.shape[0]  # 64 non-zero pixels >= k: no repeats
    rest = o[num_ist:]
    assert bool((rest >= 0).all() and (rest[:, 0] < b).all() and (rest[:, 1] < h).all() and (rest[:, 2] < w).all())
    # before iters_to_start_ist: uniform only; without weight maps: the base sampler
    early = dev.sample_method(n, b, h, w, batch={"ist_weights": maps.to(DEV), "iter_steps": 5}, device=DEV)
    assert early.shape == (n, 3) and early.is_cuda
    # the collated batch of a device-resident image cache
    images = torch.rand(b, h, w, 3, generator=gen).to(DEV)
    batch = {"image": images, "image_idx": torch.arange(100, 100 + b, device=DEV), "ist_weights": maps.to(DEV), "iter_steps": 11}
    random.seed(78)
    col = dev.collate_image_dataset_batch(batch, n)
    idx = col["indices"]
    assert idx.is_cuda and torch.equal(col["image"], images[idx[:, 0] - 100, idx[:, 1], idx[:, 2]])
    assert bool((col["ist_weights"][:num_ist] > 0).all())
    with pytest.raises(RuntimeError):
        DynamicBasedPixelSampler(n, dataset=state, device_sampler=True).sample_method(
            n, b, h, w, batch={"ist_weights": maps, "iter_steps": 11})


def test_device_sampler_full_resolution_maps():
    """The preset's shape: 4096-ray batch, 10 % importance pixels, 540 x 960 maps (sparse IST-like and dense ISG-like):
    every pixel valid, distinct per image; device time of the preset's 41 images x 10 pixels reported."""
    from soccernerfs_b200 import ops
    from soccernerfs_b200.data.pixel_samplers import DynamicBasedPixelSampler

    gen = torch.Generator().manual_seed(4)
    b, h, w = 60, 540, 960
    maps = torch.zeros(b, h * w, dtype=torch.float16)
    idx = torch.randint(0, h * w, (b, 5000), generator=gen)
    maps.scatter_(1, idx, (torch.rand(b, 5000, generator=gen) * 0.8 + 0.15).half())
    maps[b // 2:] = (torch.rand(b - b // 2, h * w, generator=gen) * 0.3 + 1e-3).half()  # dense (ISG-like)
    maps = maps.view(b, h, w).to(DEV)
    state = types.SimpleNamespace(iters_to_start_ist=0, is_pixel_ratio=0.1)
    sampler = DynamicBasedPixelSampler(4096, dataset=state)
    random.seed(5)
    batch = {"ist_weights": maps, "iter_steps": 1}
    out = sampler.sample_method(4096, b, h, w, batch=batch)
    num_ist = 409  # floor(0.1 * 4096); 10 * ceil(409 / 60) = 70 per image: five images of 70 and one of 59
    o = out[:num_ist]
    assert out.shape == (4096, 3) and bool((maps[o[:, 0], o[:, 1], o[:, 2]] > 0).all())
    key = o[:, 0] * (h * w) + o[:, 1] * w + o[:, 2]
    assert torch.unique(key).numel() == num_ist
    counts = torch.bincount(o[:, 0], minlength=b)
    assert int(counts.max()) == 70 and int((counts > 0).sum()) == 6 and int(counts.sum()) == num_ist
    # the preset's own walk (>= 409 images in the cache: 41 images x 10 pixels), sparse and dense maps alternating
    imgs = torch.arange(41) + 10
    sel = torch.stack([imgs, torch.full((41,), 10), torch.arange(41) * 10], dim=-1).to(torch.int32).to(DEV)
    for _ in range(2):
        res = ops.importance_pixels(maps, sel, 10, 410, w, seed=9)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10):
        res = ops.importance_pixels(maps, sel, 10, 410, w, seed=9)
    t1.record()
    torch.cuda.synchronize()
    print(f"\ndevice importance sampler: {t0.elapsed_time(t1) / 10:.3f} ms for 41 maps of {h}x{w} x 10 pixels")
    assert bool((maps[res[:, 0], res[:, 1], res[:, 2]] > 0).all()) and torch.equal(res[:, 0].cpu(), imgs.repeat_interleave(10))
    assert t0.elapsed_time(t1) / 10 < 20.0  # the host loop it replaces takes ~200 ms


def _ring_cameras(n_cams, n_frames, h, w):
    """One camera per image (nerfstudio's convention): n_cams cameras on a ring looking at the origin, n_frames each."""
    from soccernerfs_b200.cameras.cameras import Cameras

    c2w, times, ids = [], [], []
    for c in range(n_cams):
        ang = 2 * np.pi * c / n_cams
        pos = torch.tensor([np.cos(ang), np.sin(ang), 0.35], dtype=torch.float32) * 1.2
        z = pos / pos.norm()  # the camera looks along -z
        x = torch.linalg.cross(torch.tensor([0.0, 0.0, 1.0]), z)
        x = x / x.norm()
        y = torch.linalg.cross(z, x)
        m = torch.stack([x, y, z, pos], dim=-1)
        for f in range(n_frames):
            c2w.append(m)
            times.append(f / (n_frames - 1))
            ids.append(float(c))
    return Cameras(torch.stack(c2w), 50.0, 50.0, w / 2, h / 2, w, h, times=torch.tensor(times)[:, None], ids=torch.tensor(ids)[:, None])


def test_device_datamanager_feeds_the_graphed_train_step():
    """DynamicDataManager.next_train on a device-resident image cache (IST maps by kp_ist_map, importance + uniform pixel
    sampling, pixel gather, ray generation: all CUDA tensors) and its batches through the CUDA-graph train step: the keys a
    datamanager adds (indices, ist_weights) and the bundle's directions_norm do not push the step off the graph."""
    from soccernerfs_b200.data.datamanagers.dynamic_datamanager import DynamicDataManager, DynamicDataManagerConfig
    from soccernerfs_b200.data.dynamic_dataset import compute_ist
    from soccernerfs_b200.data.scene_box import SceneBox
    from soccernerfs_b200.engine.trainer import TrainStep
    from soccernerfs_b200.models.kplanes import KPlanesModelConfig

    n_cams, n_frames, h, w = 4, 5, 36, 64
    cams = _ring_cameras(n_cams, n_frames, h, w)
    images = torch.full((n_cams * n_frames, h, w, 3), 0.2)
    for c in range(n_cams):
        for f in range(n_frames):  # a bright square that moves with time
            x0 = 4 + 9 * f + c
            images[c * n_frames + f, 10:22, x0: x0 + 12] = torch.tensor([0.9, 0.8, 0.1])
    cfg = DynamicDataManagerConfig(train_num_rays_per_batch=512, is_pixel_ratio=0.25, ist_range=0.3, iters_to_start_is=3)
    random.seed(1)
    torch.manual_seed(1)
    dm = DynamicDataManager(cfg, cams, images, device=DEV)
    cache = dm.image_cache.batch
    assert cache["image"].is_cuda and cache["ist_weights"].is_cuda and cache["ist_weights"].dtype == torch.float16
    host_maps = compute_ist(images, cams.ids, cams.times, cfg.ist_range)
    assert torch.equal(cache["ist_weights"].cpu(), host_maps) and int((host_maps > 0).sum()) > 1000
    for step in range(6):
        rb, batch = dm.next_train(step)
        idx = batch["indices"]
        assert idx.is_cuda and rb.origins.is_cuda and batch["image"].is_cuda and idx.shape == (512, 3)
        assert torch.equal(batch["image"], cache["image"][idx[:, 0], idx[:, 1], idx[:, 2]])
        assert torch.equal(rb.times, dm.cameras.times[idx[:, 0]]) and torch.equal(rb.camera_indices, idx[:, 0:1])
        if step >= 3:  # iter_steps = step + 1 > iters_to_start_is: a quarter of the batch follows the maps
            assert bool((batch["ist_weights"][:128] > 0).all())
            # on the moving square (bright) or where it is in a neighbouring frame (background): 20-50 % bright expected,
            # against 2 % of bright pixels under uniform sampling
            assert float((batch["image"][:128].max(dim=-1).values > 0.5).float().mean()) > 0.06
    # prefetch on a side stream: the same batches in the same order
    seqs = []
    for prefetch in (False, True):
        random.seed(2)
        torch.manual_seed(2)
        d2 = DynamicDataManager(cfg, cams, images, device=DEV, prefetch=prefetch)
        seqs.append([d2.next_train(i) for i in range(6)])
    torch.cuda.synchronize()
    for (rb_a, b_a), (rb_b, b_b) in zip(*seqs):
        assert torch.equal(b_a["indices"], b_b["indices"]) and torch.equal(b_a["image"], b_b["image"])
        assert torch.equal(rb_a.origins, rb_b.origins) and torch.equal(rb_a.directions, rb_b.directions)
    dm = d2  # train from the prefetching manager
    model_cfg = KPlanesModelConfig(spacetime_resolution=(16, 16, 16, 5), multiscale_res=(1, 2), num_nerf_samples_per_ray=16,
                                   num_proposal_samples_per_ray=(32, 24),
                                   proposal_net_args_list=[{"feature_dim": 8, "resolution": [24, 24, 24, 5]},
                                                           {"feature_dim": 8, "resolution": [32, 32, 32, 5]}])
    model = model_cfg.setup(scene_box=SceneBox(aabb=torch.tensor([[-1.5, -1.5, -1.5], [1.5, 1.5, 1.5]])), num_train_data=len(cams)).to(DEV)
    trainer = TrainStep(model, max_steps=100, warm_up_end=2, use_cuda_graph=True)
    losses = []
    try:
        for step in range(16):
            rb, batch = dm.next_train(6 + step)
            assert trainer._graphable(rb, batch)
            losses.append(float(trainer(rb, batch)["loss"]))
        assert trainer._graphs, "the step never reached its CUDA-graph replay"
    finally:
        trainer.close()
    assert all(l == l and l < 1e3 for l in losses) and min(losses[8:]) < losses[0], losses
