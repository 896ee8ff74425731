"""Pins oracle/kplanes_oracle.py to the fixtures produced by the REAL reference (oracle/make_golden.py).

CPU-only; runs everywhere.  Tolerances: integer outputs bit-exact; fp32 outputs identical up to
reassociation (<= 1e-6 rel).
"""
import torch

from oracle import kplanes_oracle as ko
from tests.conftest import load_golden, rel_err

TOL = 2e-6


def _grids(g, prefix, n_scales, n_planes=6):
    return [[g[f"{prefix}_{i}_{j}"] for j in range(n_planes)] for i in range(n_scales)]


def test_interpolate_kplanes_matches_reference():
    g = load_golden("interp")
    grids = _grids(g, "grid", 2)
    for gs in grids:
        for p in gs:
            p.requires_grad_(True)
    out = ko.interpolate_kplanes(g["pts"], grids, True)
    assert rel_err(out, g["out_cat"]) < TOL
    (out * g["grad_out"]).sum().backward()
    for i in range(2):
        for j in range(6):
            assert rel_err(grids[i][j].grad, g[f"ggrid_{i}_{j}"]) < TOL
    assert rel_err(ko.interpolate_kplanes(g["pts"], grids, False), g["out_sum"]) < TOL
    g3 = [[g[f"grid3_{j}"] for j in range(3)]]
    assert rel_err(ko.interpolate_kplanes(g["pts"][:, :3], g3, True), g["out_static"]) < TOL


def test_manual_bilinear_equals_grid_sample():
    g = load_golden("interp")
    for j, comb in enumerate([(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]):
        plane = g[f"grid_1_{j}"]
        a = ko.grid_sample_plane(plane, g["pts"][:, list(comb)])
        b = ko.bilinear_border_manual(plane, g["pts"][:, list(comb)])
        assert rel_err(b, a) < TOL


def test_collider_and_samplers_match_reference():
    g = load_golden("samplers")
    o, d, t, aabb = g["origins"], g["directions"], g["times"], g["aabb"]
    nt, ft = ko.aabb_collider(o, d, aabb, near_plane=0.05)
    assert torch.equal(nt, g["nears_train"]) and torch.equal(ft, g["fars_train"])
    ne, fe = ko.aabb_collider(o, d, aabb, near_plane=0.0)
    assert torch.equal(ne, g["nears_eval"]) and torch.equal(fe, g["fars_eval"])
    for mode in ("train", "eval"):
        tr = g[f"{mode}_t_rand"] if mode == "train" else None
        ur = g[f"{mode}_u_rand"] if mode == "train" else None
        s0 = ko.uniform_sampler(o, d, nt, ft, t, 40, tr)
        assert torch.equal(s0.spacing_bins, g[f"{mode}_bins0"])
        assert torch.equal(s0.starts, g[f"{mode}_starts0"]) and torch.equal(s0.ends, g[f"{mode}_ends0"])
        s1, inds = ko.pdf_sampler(s0, g[f"{mode}_weights"][..., 0], 24, ur)
        assert torch.equal(inds, g[f"{mode}_inds1"])  # bit-exact indices
        assert torch.equal(s1.spacing_bins, g[f"{mode}_bins1"])
        assert torch.equal(s1.starts, g[f"{mode}_starts1"]) and torch.equal(s1.ends, g[f"{mode}_ends1"])
        assert rel_err(s1.positions(), g[f"{mode}_positions1"]) < TOL


def test_compositing_matches_reference():
    g = load_golden("render")
    starts = g["starts"]
    deltas = (starts[:, 1:] - starts[:, :-1])[..., None]
    steps = (starts[:, :-1] + starts[:, 1:]) / 2
    density = g["density"].clone().requires_grad_(True)
    rgb = g["rgb"].clone().requires_grad_(True)
    w = ko.get_weights(deltas, density)
    assert rel_err(w, g["weights"]) < TOL
    comp = ko.render_rgb(rgb, w, g["bg"], training=True)
    acc = ko.render_accumulation(w)
    assert rel_err(comp, g["comp"]) < TOL and rel_err(acc, g["acc"]) < TOL
    assert torch.equal(ko.render_depth_median(w, steps), g["depth_median"])
    assert rel_err(ko.render_depth_expected(w, steps), g["depth_expected"]) < TOL
    assert torch.equal(ko.render_median_rgb(rgb, w), g["median_rgb"])
    ((comp * g["go_rgb"]).sum() + (acc * g["go_acc"]).sum() + (w * g["go_w"]).sum()).backward()
    assert rel_err(density.grad, g["g_density"]) < TOL
    assert rel_err(rgb.grad, g["g_rgb"]) < TOL
    assert rel_err(ko.render_rgb(g["rgb"], g["weights"], "last_sample", training=False), g["comp_eval"]) < TOL


def test_losses_match_reference():
    g = load_golden("losses")
    ws = [g[f"w{i}"].clone().requires_grad_(True) for i in range(3)]
    bs = [g[f"b{i}"] for i in range(3)]
    il = ko.interlevel_loss(ws, bs)
    dl = ko.distortion_loss(ws, bs)
    assert rel_err(il, g["interlevel"]) < TOL and rel_err(dl, g["distortion"]) < TOL
    (il + dl).backward()
    for i in range(3):
        assert rel_err(ws[i].grad, g[f"g_w{i}"]) < TOL
    grids = _grids(g, "grid", 2)
    for gs in grids:
        for p in gs:
            p.requires_grad_(True)
    tv, ts, st = ko.space_tv_loss(grids), ko.time_smoothness_loss(grids), ko.sparse_transients_loss(grids)
    assert rel_err(tv, g["space_tv"]) < TOL and rel_err(ts, g["time_smoothness"]) < TOL and rel_err(st, g["sparse_transients"]) < TOL
    (0.7 * tv + 1.3 * ts + 0.4 * st).backward()
    for i in range(2):
        for j in range(6):
            assert rel_err(grids[i][j].grad, g[f"ggrid_{i}_{j}"]) < TOL
    g3 = [[g[f"grid3_{j}"] for j in range(3)]]
    assert rel_err(ko.space_tv_loss(g3), g["space_tv_static"]) < TOL
    assert float(ko.time_smoothness_loss(g3)) == float(g["time_smoothness_static"]) == 0.0
    assert float(ko.sparse_transients_loss(g3)) == float(g["sparse_transients_static"]) == 0.0
    st_ = g["ds_starts"]
    steps = ((st_[:, :-1] + st_[:, 1:]) / 2)[..., None]
    lengths = (st_[:, 1:] - st_[:, :-1])[..., None]
    assert rel_err(ko.ds_nerf_depth_loss(g["w2"], g["ds_term"], steps, lengths, torch.tensor([0.01])), g["ds_loss"]) < TOL


def load_tiny_model(g):
    """Rebuild oracle ModelParams from the model_tiny fixture (parameters stored in tensors() order)."""
    aabb = g["aabb"]
    gen = torch.Generator().manual_seed(0)
    mp = ko.make_model_params("tiny", gen, aabb)
    for i, p in enumerate(mp.tensors()):
        p.data.copy_(g[f"param_{i}"])
    return mp


def test_model_step_matches_reference():
    g = load_golden("model_tiny")
    mp = load_tiny_model(g)
    rand = {k[5:]: v for k, v in g.items() if k.startswith("rand_")}
    out, ld, grads = ko.train_step(mp, g["origins"], g["directions"], g["times"], g["image"], rand, anneal=float(g["anneal"]))
    assert torch.equal(out["inds_list"][0], g["inds1"]) and torch.equal(out["inds_list"][1], g["inds2"])
    for i in range(3):
        assert torch.equal(out["samples_list"][i].spacing_bins, g[f"bins_{i}"])
        assert rel_err(out["weights_list"][i], g[f"weights_{i}"]) < TOL
    for k in ("rgb", "accumulation", "depth", "median_rgb", "prop_depth_0", "prop_depth_1", "density"):
        assert rel_err(out[k], g[k]) < TOL, k
    for k, v in ld.items():
        assert rel_err(v, g["loss_" + k]) < TOL, k
    for i, gr in enumerate(grads):
        assert rel_err(gr, g[f"grad_{i}"]) < 1e-5, i


def test_ray_generation_matches_reference():
    """(f2) the oracle's pixel -> ray restatement vs rays produced by the reference's own Cameras (explicit
    (camera,row,col) triplets as RayGenerator issues them, and one whole frame)."""
    g = load_golden("raygen")
    cam = (g["c2w"], g["fx"], g["fy"], g["cx"], g["cy"], g["times"])
    ri = g["ray_indices"]
    o, d, pa, nrm, t = ko.generate_rays(*cam, ri[:, 0], ri[:, 1], ri[:, 2])
    assert torch.equal(o, g["origins"]) and torch.equal(t, g["ray_times"])
    assert rel_err(d, g["directions"]) < 1e-6 and rel_err(pa, g["pixel_area"]) < 1e-5 and rel_err(nrm, g["directions_norm"]) < 1e-6
    h, w = (int(v) for v in g["hw"])
    yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    c = torch.full((h * w,), int(g["frame_cam"]), dtype=torch.int64)
    o, d, pa, nrm, t = ko.generate_rays(*cam, c, yy.reshape(-1), xx.reshape(-1))
    assert torch.equal(o.view(h, w, 3), g["frame_origins"]) and torch.equal(t.view(h, w, 1), g["frame_times"])
    assert rel_err(d.view(h, w, 3), g["frame_directions"]) < 1e-6
    assert rel_err(pa.view(h, w, 1), g["frame_pixel_area"]) < 1e-5


def test_lens_models_match_reference():
    """(f2) OpenCV undistortion (camera_utils.py:298-401) and the fisheye / equirectangular direction models
    (cameras.py:665-697), mixed in one camera batch, vs rays of the reference's own Cameras (fixture raygen_lens);
    disable_distortion and the dataparsers' shared-distortion perspective batch included.  The restatement evaluates
    the reference's expressions in the reference's order: bit-identical on the CPU."""
    g = load_golden("raygen_lens")
    cam = (g["c2w"], g["fx"], g["fy"], g["cx"], g["cy"], g["times"])
    ri = g["ray_indices"]
    idx = (ri[:, 0], ri[:, 1], ri[:, 2])
    assert set(g["types"].reshape(-1).tolist()) == {1, 2, 3}
    o, d, pa, nrm, t = ko.generate_rays(*cam, *idx, distortion_params=g["dist"], camera_type=g["types"])
    assert torch.equal(o, g["origins"]) and torch.equal(t, g["ray_times"])
    assert torch.equal(d, g["directions"]) and torch.equal(pa, g["pixel_area"]) and torch.equal(nrm, g["directions_norm"])
    _, d, pa, _, _ = ko.generate_rays(*cam, *idx, camera_type=g["types"])  # disable_distortion=True, cameras.py:636
    assert torch.equal(d, g["nodist_directions"]) and torch.equal(pa, g["nodist_pixel_area"])
    assert not torch.equal(g["nodist_directions"], g["directions"])
    _, d, pa, _, _ = ko.generate_rays(*cam, *idx, distortion_params=g["dist"][0:1].expand(6, 6))
    assert torch.equal(d, g["persp_directions"]) and torch.equal(pa, g["persp_pixel_area"])
    h, w = (int(v) for v in g["hw"])
    yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    for c in (0, 3, 4, 5):
        ci = torch.full((h * w,), c, dtype=torch.int64)
        _, d, pa, nrm, _ = ko.generate_rays(*cam, ci, yy.reshape(-1), xx.reshape(-1), distortion_params=g["dist"],
                                            camera_type=g["types"])
        assert torch.equal(d.view(h, w, 3), g[f"frame{c}_directions"]), c
        assert torch.equal(pa.view(h, w, 1), g[f"frame{c}_pixel_area"]), c
        assert torch.equal(nrm.view(h, w, 1), g[f"frame{c}_directions_norm"]), c
    # zero distortion: the Newton step is exactly zero (the reason Cameras drops an all-zero table)
    pts = torch.randn(64, 2, generator=torch.Generator().manual_seed(3))
    assert torch.equal(ko.undistort(pts, torch.zeros(1, 6)), pts)
    # and undistortion inverts the OpenCV forward model where it converges (mild distortion, points inside the frame)
    k = g["dist"][0]
    und = ko.undistort(0.4 * pts.clamp(-1, 1), k[None])
    x, y = und[:, 0], und[:, 1]
    r = x * x + y * y
    dd = 1.0 + r * (k[0] + r * (k[1] + r * (k[2] + r * k[3])))
    fwd = torch.stack([dd * x + 2 * k[4] * x * y + k[5] * (r + 2 * x * x), dd * y + 2 * k[5] * x * y + k[4] * (r + 2 * y * y)], -1)
    assert (fwd - 0.4 * pts.clamp(-1, 1)).abs().max() < 1e-6


def test_decoder_depths_match_reference():
    """sigma_net_layers / rgb_net_layers / hidden widths other than the presets' (kplanes.py:96-103): the oracle's field
    with such weight stacks vs the reference's own KPlanesField (fixture field_depths), outputs and all gradients."""
    g = load_golden("field_depths")
    n, s = g["bins"].shape[0], g["bins"].shape[1] - 1
    for tag, n_sigma, n_color, view in (("a", 3, 2, True), ("b", 1, 4, False)):
        grids = [[g[f"{tag}_grid_{i}_{j}"].clone().requires_grad_(True) for j in range(6)] for i in range(2)]
        sw = [g[f"{tag}_sigma_w{i}"].clone().requires_grad_(True) for i in range(n_sigma)]
        cw = [g[f"{tag}_color_w{i}"].clone().requires_grad_(True) for i in range(n_color)]
        p = ko.FieldParams(aabb=g["aabb"], grids=grids, sigma_w=sw, color_w=cw, concat=True, view_dependent=view)
        smp = ko.Samples(origins=g["origins"], directions=g["directions"], starts=g["bins"][:, :-1], ends=g["bins"][:, 1:],
                         spacing_bins=g["bins"], nears=torch.zeros(n, 1), fars=torch.full((n, 1), 1.5), times=g["times"])
        dens, rgb, _ = ko.field_forward(p, smp)
        assert dens.shape == (n, s, 1) and rel_err(dens, g[f"{tag}_density"]) < 2e-6 and rel_err(rgb, g[f"{tag}_rgb"]) < 2e-6
        ((dens * g[f"{tag}_gd"]).sum() + (rgb * g[f"{tag}_gr"]).sum()).backward()
        for i in range(2):
            for j in range(6):
                assert rel_err(grids[i][j].grad, g[f"{tag}_ggrid_{i}_{j}"]) < 2e-6, (tag, i, j)
        for name, ws in (("sigma", sw), ("color", cw)):
            for i, w in enumerate(ws):
                assert rel_err(w.grad, g[f"{tag}_{name}_gw{i}"]) < 2e-6, (tag, name, i)


def _same_with_nans(a: torch.Tensor, b: torch.Tensor) -> bool:
    return torch.equal(torch.isnan(a), torch.isnan(b)) and torch.equal(torch.nan_to_num(a, nan=0.0), torch.nan_to_num(b, nan=0.0))


def test_crop_box_bounds_match_reference():
    """(f2) the slab test behind Cameras.generate_rays(aabb_box=...) (math.py:201-272, cameras.py:478-497) vs the
    reference's own function: hits, misses (1e10 twice), axis-parallel rays (infinite slabs) and the 0/0 slab (NaN)."""
    g = load_golden("raygen_crop")
    t_min, t_max = ko.intersect_aabb(g["origins"], g["directions"], g["aabb"])
    assert _same_with_nans(t_min, g["t_min"]) and _same_with_nans(t_max, g["t_max"])
    assert int(torch.isnan(g["t_min"]).sum()) == 8 and 0 < int((g["t_min"] == 1e10).sum()) < t_min.numel() - 8
    r = load_golden("raygen")
    h, w = (int(v) for v in r["hw"])
    t_min, t_max = ko.intersect_aabb(r["frame_origins"].reshape(-1, 3), r["frame_directions"].reshape(-1, 3), g["box"].reshape(-1))
    assert torch.equal(t_min.view(h, w, 1), g["frame_nears"]) and torch.equal(t_max.view(h, w, 1), g["frame_fars"])


def test_cfg4_piecewise_single_jitter_samplers_match_reference():
    """BASELINE config 4: UniformLinDispPiecewiseSampler + PDFSampler with single_jitter=True and the expected-depth
    renderer, vs the reference's own classes (fixture samplers_cfg4)."""
    g = load_golden("samplers_cfg4")
    o, d, t, nears, fars = g["origins"], g["directions"], g["times"], g["nears"], g["fars"]
    for mode in ("train", "eval"):
        tr = g[f"{mode}_t_rand"] if mode == "train" else None  # [N,1]: one jitter per ray
        ur = g[f"{mode}_u_rand"] if mode == "train" else None
        s0 = ko.uniform_sampler(o, d, nears, fars, t, 64, tr, spacing="piecewise")
        assert torch.equal(s0.spacing_bins, g[f"{mode}_bins0"])
        assert torch.equal(s0.starts, g[f"{mode}_starts0"]) and torch.equal(s0.ends, g[f"{mode}_ends0"])
        s1, inds = ko.pdf_sampler(s0, g[f"{mode}_weights"][..., 0], 24, ur)
        assert torch.equal(inds, g[f"{mode}_inds1"])
        assert torch.equal(s1.spacing_bins, g[f"{mode}_bins1"])
        assert torch.equal(s1.starts, g[f"{mode}_starts1"]) and torch.equal(s1.ends, g[f"{mode}_ends1"])
        depth = ko.render_depth_expected(g[f"{mode}_w1"], s1.steps())
        assert rel_err(depth, g[f"{mode}_depth_expected"]) < 1e-6
