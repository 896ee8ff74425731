"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the REAL reference modules.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
                                                        python -m oracle.make_golden --check   (regenerate into a scratch
                                                        directory and compare with the committed fixtures, array by array)

Every fixture stores the seeded inputs and the outputs the reference's own code produced for them
(``oracle/ref_loader.py`` explains the five stub modules).  ``tests/test_oracle_golden.py`` then pins
``oracle/kplanes_oracle.py`` to these vectors on any machine; the ``-m gpu`` tests pin the CUDA path
to the same vectors.  Random draws the reference makes internally (``torch.rand`` in the samplers and
the random background) are fed from a queue so that they can be stored with the fixture.
"""
from __future__ import annotations

import contextlib
import functools
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import kplanes_oracle as ko  # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


@contextlib.contextmanager
def rand_queue(queue):
    """Serve torch.rand / torch.rand_like calls made by the reference from ``queue`` (in order)."""
    real_rand, real_like = torch.rand, torch.rand_like
    q = list(queue)

    def fake_rand(*size, **kw):
        if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)):
            size = tuple(size[0])
        t = q.pop(0)
        assert tuple(t.shape) == tuple(size), (t.shape, size)
        return t.clone()

    def fake_like(x, **kw):
        t = q.pop(0)
        assert t.shape == x.shape, (t.shape, x.shape)
        return t.clone()

    torch.rand, torch.rand_like = fake_rand, fake_like
    try:
        yield q
    finally:
        torch.rand, torch.rand_like = real_rand, real_like


@contextlib.contextmanager
def record_searchsorted(store):
    real = torch.searchsorted

    def rec(*a, **k):
        r = real(*a, **k)
        store.append(r.clone())
        return r

    torch.searchsorted = rec
    try:
        yield
    finally:
        torch.searchsorted = real


def save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    conv = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        conv[k] = np.asarray(v)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **conv)
    print(f"wrote {path}: {os.path.getsize(path)/1024:.1f} KiB")


def load_ref_density_field(R, p: ko.DensityFieldParams, resolution):
    f = R.kplanes_field.KPlanesDensityField(p.aabb, resolution=resolution, feature_dim=p.grids[0].shape[1], linear_decoder=False)
    with torch.no_grad():
        for dst, src in zip(f.grids, p.grids):
            dst.copy_(src)
        for lin, w in zip(f.sigma_net.layers, p.sigma_w):
            lin.weight.copy_(w)
    return f


def load_ref_field(R, p: ko.FieldParams, res, ms, sigma_hidden):
    f = R.kplanes_field.KPlanesField(
        p.aabb, spacetime_resolution=res, feat_dim=p.grids[0][0].shape[1], multiscale_res=ms,
        concat_features_across_scales=p.concat, linear_decoder=False,
        disable_viewing_dependent=not p.view_dependent, sigma_net_hidden_dim=sigma_hidden,
    )
    with torch.no_grad():
        for gs, ps in zip(f.grids, p.grids):
            for dst, src in zip(gs, ps):
                dst.copy_(src)
        for lin, w in zip(f.sigma_net.layers, p.sigma_w):
            lin.weight.copy_(w)
        for lin, w in zip(f.color_net.layers, p.color_w):
            lin.weight.copy_(w)
    return f


def gen_interp(R):
    g = torch.Generator().manual_seed(101)
    res, ms, c = (12, 10, 14, 5), (1, 2), 8
    grids = []
    for m in ms:
        reso = [r * m for r in res[:3]] + [res[3]]
        grids.append([pl + 0.3 * torch.randn(pl.shape, generator=g) for pl in ko.init_planes(c, reso, 0.1, 0.5, g)])
    pts = torch.rand(257, 4, generator=g) * 2.4 - 1.2  # includes out-of-range -> border clamp
    pts[0] = torch.tensor([-1.0, -1.0, -1.0, -1.0])
    pts[1] = torch.tensor([1.0, 1.0, 1.0, 1.0])
    pts[2] = torch.tensor([0.0, 0.0, 0.0, 0.0])
    plist = [torch.nn.ParameterList([torch.nn.Parameter(p.clone()) for p in gs]) for gs in grids]
    out_cat = R.kplanes_field.interpolate_kplanes(pts, plist, True, False, False)
    go = torch.randn(out_cat.shape, generator=g)
    (out_cat * go).sum().backward()
    grads = [p.grad for gs in plist for p in gs]
    out_sum = R.kplanes_field.interpolate_kplanes(pts, plist, False, False, False)
    # static (3-plane) case
    grids3 = [[pl + 0.3 * torch.randn(pl.shape, generator=g) for pl in ko.init_planes(c, list(res[:3]), 0.1, 0.5, g)]]
    out3 = R.kplanes_field.interpolate_kplanes(pts[:, :3], grids3, True, False, False)
    arrs = dict(pts=pts, out_cat=out_cat, out_sum=out_sum, grad_out=go, out_static=out3)
    for i, gs in enumerate(grids):
        for j, p in enumerate(gs):
            arrs[f"grid_{i}_{j}"] = p
            arrs[f"ggrid_{i}_{j}"] = grads[i * 6 + j]
    for j, p in enumerate(grids3[0]):
        arrs[f"grid3_{j}"] = p
    save("interp", **arrs)


def gen_samplers(R):
    g = torch.Generator().manual_seed(202)
    n = 64
    origins, directions, times, aabb = ko.synthetic_rays(n, g)
    RS = R.ray_samplers
    box = R.SceneBox(aabb=aabb)
    col = R.scene_colliders.AABBBoxCollider(box, near_plane=0.05)
    col.train()
    rb = R.rays.RayBundle(origins=origins, directions=directions, pixel_area=torch.ones(n, 1), times=times)
    rb = col(rb)
    nears_t, fars_t = rb.nears.clone(), rb.fars.clone()
    col.eval()
    rb_e = R.rays.RayBundle(origins=origins, directions=directions, pixel_area=torch.ones(n, 1), times=times)
    rb_e = col(rb_e)
    arrs = dict(origins=origins, directions=directions, times=times, aabb=aabb, nears_train=nears_t, fars_train=fars_t,
                nears_eval=rb_e.nears, fars_eval=rb_e.fars)
    for mode in ("train", "eval"):
        us = RS.UniformSampler()
        pdf = RS.PDFSampler(include_original=False)
        us.train(mode == "train")
        pdf.train(mode == "train")
        s0, s1 = 40, 24
        t_rand = torch.rand(n, s0 + 1, generator=g)
        u_rand = torch.rand(n, s1 + 1, generator=g)
        weights = torch.rand(n, s0, 1, generator=g) ** 4
        weights[3] = 0.0  # all-zero weights row (histogram padding only)
        weights[4, :, 0] = torch.nn.functional.one_hot(torch.tensor(7), s0).float()  # a spike
        ss = []
        with rand_queue([t_rand, u_rand] if mode == "train" else []), record_searchsorted(ss):
            rs0 = us(rb, num_samples=s0)
            rs1 = pdf(rb, rs0, weights, num_samples=s1)
        arrs.update({
            f"{mode}_t_rand": t_rand, f"{mode}_u_rand": u_rand, f"{mode}_weights": weights,
            f"{mode}_bins0": torch.cat([rs0.spacing_starts[..., 0], rs0.spacing_ends[..., -1:, 0]], -1),
            f"{mode}_starts0": rs0.frustums.starts[..., 0], f"{mode}_ends0": rs0.frustums.ends[..., 0],
            f"{mode}_bins1": torch.cat([rs1.spacing_starts[..., 0], rs1.spacing_ends[..., -1:, 0]], -1),
            f"{mode}_starts1": rs1.frustums.starts[..., 0], f"{mode}_ends1": rs1.frustums.ends[..., 0],
            f"{mode}_inds1": ss[0], f"{mode}_positions1": rs1.frustums.get_positions(),
        })
    save("samplers", **arrs)


def gen_render(R):
    g = torch.Generator().manual_seed(303)
    n, s = 48, 37
    starts = torch.sort(torch.rand(n, s + 1, generator=g) * 4 + 0.1, dim=-1).values
    fr = R.rays.Frustums(origins=torch.zeros(n, s, 3), directions=torch.ones(n, s, 3), starts=starts[:, :-1, None],
                         ends=starts[:, 1:, None], pixel_area=torch.ones(n, s, 1))
    rs = R.rays.RaySamples(frustums=fr, deltas=(starts[:, 1:] - starts[:, :-1])[..., None])
    density = (torch.rand(n, s, 1, generator=g) ** 3 * 12).requires_grad_(True)
    with torch.no_grad():
        density[5] = 0.0
        density[6, :10] = 1e4  # saturating ray
    rgb = torch.rand(n, s, 3, generator=g).requires_grad_(True)
    bg = torch.rand(n, 3, generator=g)
    w = rs.get_weights(density)
    rr = R.renderers
    r_rgb = rr.RGBRenderer(background_color=bg)
    r_rgb.train()
    comp = r_rgb(rgb, w)
    acc = rr.AccumulationRenderer()(w)
    dmed = rr.DepthRenderer("median")(w, rs)
    dexp = rr.DepthRenderer("expected")(w, rs)
    mr = rr.MedianRGBRenderer()
    mr.train()
    med_rgb = mr(rgb, w)
    go_rgb = torch.randn(n, 3, generator=g)
    go_acc = torch.randn(n, 1, generator=g)
    go_w = torch.randn(n, s, 1, generator=g) * 0.1
    ((comp * go_rgb).sum() + (acc * go_acc).sum() + (w * go_w).sum()).backward()
    r_eval = rr.RGBRenderer(background_color="last_sample")
    r_eval.eval()
    comp_eval = r_eval(rgb.detach(), w.detach())
    save("render", starts=starts, density=density, rgb=rgb, bg=bg, weights=w, comp=comp, acc=acc, depth_median=dmed,
         depth_expected=dexp, median_rgb=med_rgb, go_rgb=go_rgb, go_acc=go_acc, go_w=go_w, g_density=density.grad,
         g_rgb=rgb.grad, comp_eval=comp_eval)


def gen_losses(R):
    g = torch.Generator().manual_seed(404)
    n = 40
    L = R.losses

    def mk(s):
        b = torch.sort(torch.rand(n, s + 1, generator=g), dim=-1).values
        b[:, 0], b[:, -1] = 0.0, 1.0
        w = torch.rand(n, s, 1, generator=g) ** 2
        w = w / w.sum(1, keepdim=True) * torch.rand(n, 1, 1, generator=g)
        return b, w

    class RSamp:  # minimal object with spacing_starts/ends like RaySamples
        def __init__(self, b):
            self.spacing_starts = b[:, :-1, None]
            self.spacing_ends = b[:, 1:, None]

    (b0, w0), (b1, w1), (b2, w2) = mk(32), mk(20), mk(12)
    w0.requires_grad_(True), w1.requires_grad_(True), w2.requires_grad_(True)
    il = L.interlevel_loss([w0, w1, w2], [RSamp(b0), RSamp(b1), RSamp(b2)])
    dl = L.distortion_loss([w0, w1, w2], [RSamp(b0), RSamp(b1), RSamp(b2)])
    (il + dl).backward()
    arrs = dict(b0=b0, b1=b1, b2=b2, w0=w0, w1=w1, w2=w2, interlevel=il, distortion=dl, g_w0=w0.grad, g_w1=w1.grad, g_w2=w2.grad)
    # plane regularisers on a 2-scale dynamic field + static field
    grids = []
    for m in (1, 2):
        reso = [6 * m, 5 * m, 7 * m, 4]
        grids.append([torch.nn.Parameter(pl + 0.3 * torch.randn(pl.shape, generator=g)) for pl in ko.init_planes(4, reso, 0.1, 0.5, g)])
    tv, ts, st = L.space_tv_loss(grids), L.time_smoothness_loss(grids), L.sparse_transients_loss(grids)
    (0.7 * tv + 1.3 * ts + 0.4 * st).backward()
    arrs.update(space_tv=tv, time_smoothness=ts, sparse_transients=st)
    for i, gs in enumerate(grids):
        for j, p in enumerate(gs):
            arrs[f"grid_{i}_{j}"] = p
            arrs[f"ggrid_{i}_{j}"] = p.grad
    grids3 = [[torch.nn.Parameter(pl + 0.3 * torch.randn(pl.shape, generator=g)) for pl in ko.init_planes(4, [6, 5, 7], 0.1, 0.5, g)]]
    arrs.update(space_tv_static=L.space_tv_loss(grids3), time_smoothness_static=L.time_smoothness_loss(grids3),
                sparse_transients_static=L.sparse_transients_loss(grids3))
    for j, p in enumerate(grids3[0]):
        arrs[f"grid3_{j}"] = p
    # DS-NeRF depth loss
    s = 12
    starts = torch.sort(torch.rand(n, s + 1, generator=g) * 4 + 0.1, dim=-1).values
    steps = ((starts[:, :-1] + starts[:, 1:]) / 2)[..., None]
    lengths = (starts[:, 1:] - starts[:, :-1])[..., None]
    term = torch.rand(n, 1, generator=g) * 4
    term[:5] = 0.0
    dsl = L.ds_nerf_depth_loss(w2.detach(), term, steps, lengths, torch.tensor([0.01]))
    arrs.update(ds_starts=starts, ds_term=term, ds_loss=dsl)
    save("losses", **arrs)


def gen_model(R):
    """Whole get_outputs + get_loss_dict + backward, composed from the reference's own classes exactly as
    NS/models/kplanes.py:188-309, 349-388, 414-452 does (the Model class itself drags in torchmetrics /
    RetinaNet downloads)."""
    g = torch.Generator().manual_seed(505)
    n = 96
    origins, directions, times, aabb = ko.synthetic_rays(n, g)
    mp = ko.make_model_params("tiny", g, aabb)
    image = torch.rand(n, 3, generator=g)
    rand = ko.make_rand(n, mp, g)
    anneal = 0.37

    field = load_ref_field(R, mp.field, (16, 16, 16, 6), (1, 2), 64)
    props = [load_ref_density_field(R, mp.proposals[0], [24, 24, 24, 6]), load_ref_density_field(R, mp.proposals[1], [32, 32, 32, 6])]
    RS = R.ray_samplers
    sampler = RS.ProposalNetworkSampler(
        num_nerf_samples_per_ray=mp.num_nerf_samples, num_proposal_samples_per_ray=mp.num_proposal_samples,
        num_proposal_network_iterations=2, single_jitter=False, update_sched=lambda step: 1,
        initial_sampler=RS.UniformSampler(single_jitter=False),
    )
    sampler.set_anneal(anneal)
    col = R.scene_colliders.AABBBoxCollider(R.SceneBox(aabb=aabb), near_plane=0.0)
    rr = R.renderers
    r_rgb, r_acc, r_depth, r_med = rr.RGBRenderer("random"), rr.AccumulationRenderer(), rr.DepthRenderer(), rr.MedianRGBRenderer()
    for m in (field, *props, sampler, col, r_rgb, r_med):
        m.train()

    rb = R.rays.RayBundle(origins=origins, directions=directions, pixel_area=torch.ones(n, 1), times=times)
    rb = col(rb)
    ss = []
    queue = [rand["t_rand"], rand["u1"], rand["u2"], rand["bg"]]
    with rand_queue(queue), record_searchsorted(ss):
        density_fns = [functools.partial(p.density_fn, times=rb.times) for p in props]
        ray_samples, weights_list, ray_samples_list = sampler(rb, density_fns=density_fns)
        fo = field(ray_samples)
        FH = R.kplanes_field.FieldHeadNames
        weights = ray_samples.get_weights(fo[FH.DENSITY])
        weights_list.append(weights)
        ray_samples_list.append(ray_samples)
        rgb = r_rgb(rgb=fo[FH.RGB], weights=weights)
    inds = ss[:2]
    acc = r_acc(weights)
    depth = r_depth(weights, ray_samples)
    med = r_med(rgb=fo[FH.RGB], weights=weights)
    pd0 = r_depth(weights=weights_list[0], ray_samples=ray_samples_list[0])
    pd1 = r_depth(weights=weights_list[1], ray_samples=ray_samples_list[1])

    L = R.losses
    coef = mp.loss_coefficients
    nerf_g, prop_g = field.grids, [p.grids for p in props]
    ld = {
        "rgb_loss": L.MSELoss()(image, rgb),
        "distortion_loss": L.distortion_loss(weights_list, ray_samples_list),
        "interlevel_loss": L.interlevel_loss(weights_list, ray_samples_list),
        "space_tv_loss": L.space_tv_loss(nerf_g),
        "space_tv_proposal_loss": L.space_tv_loss(prop_g),
        "sparse_transients_loss": L.sparse_transients_loss(nerf_g),
        "sparse_transients_proposal_loss": L.sparse_transients_loss(prop_g),
        "time_smoothness_loss": L.time_smoothness_loss(nerf_g),
        "time_smoothness_proposal_loss": L.time_smoothness_loss(prop_g),
    }
    ld = {k: v * coef[k] for k, v in ld.items()}
    loss = sum(ld.values())
    loss.backward()

    arrs = dict(origins=origins, directions=directions, times=times, aabb=aabb, image=image, anneal=np.float32(anneal),
                nears=rb.nears, fars=rb.fars, rgb=rgb, accumulation=acc, depth=depth, median_rgb=med, prop_depth_0=pd0,
                prop_depth_1=pd1, inds1=inds[0], inds2=inds[1], loss=loss, density=fo[FH.DENSITY], rgb_samples=fo[FH.RGB])
    for k, v in rand.items():
        arrs["rand_" + k] = v
    for k, v in ld.items():
        arrs["loss_" + k] = v
    for i, (w, rs_) in enumerate(zip(weights_list, ray_samples_list)):
        arrs[f"weights_{i}"] = w
        arrs[f"bins_{i}"] = torch.cat([rs_.spacing_starts[..., 0], rs_.spacing_ends[..., -1:, 0]], -1)
    # parameters in oracle order (ModelParams.tensors()) and their reference gradients
    ref_params = []
    for p in props:
        ref_params += list(p.grids) + [l.weight for l in p.sigma_net.layers]
    ref_params += [q for gs in field.grids for q in gs] + [l.weight for l in field.sigma_net.layers] + [l.weight for l in field.color_net.layers]
    for i, (mine, ref) in enumerate(zip(mp.tensors(), ref_params)):
        assert mine.shape == ref.shape
        arrs[f"param_{i}"] = mine
        arrs[f"grad_{i}"] = ref.grad
    save("model_tiny", **arrs)


def gen_field_variants(R):
    """Non-default field branches pinned to the reference's own modules: (a) unbounded scene = SceneContraction(order=inf)
    in KPlanesField and KPlanesDensityField (kplanes_field.py:278-280, 436-438); (b) linear_decoder=True = learned
    colour basis + linear density (kplanes_field.py:219-246, 303-304, 349-354).  Outputs and plane / weight gradients."""
    from nerfstudio.field_components.spatial_distortions import SceneContraction

    g = torch.Generator().manual_seed(321)
    res, ms, c = (12, 10, 14, 5), (1, 2), 8
    n, s = 96, 6
    aabb = torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]])
    origins = (torch.rand(n, 3, generator=g) - 0.5) * 1.6
    d = torch.randn(n, 3, generator=g)
    directions = d / d.norm(dim=-1, keepdim=True)
    bins = torch.sort(torch.rand(n, s + 1, generator=g), -1).values * 5.0  # up to 5 units away: well outside the unit cube
    times = torch.rand(n, 1, generator=g)
    rb = R.rays.RayBundle(origins=origins, directions=directions, pixel_area=torch.ones(n, 1), times=times,
                          nears=torch.zeros(n, 1), fars=torch.full((n, 1), 5.0))
    rs = rb.get_ray_samples(bin_starts=bins[:, :-1, None], bin_ends=bins[:, 1:, None])
    arrs = dict(aabb=aabb, origins=origins, directions=directions, bins=bins, times=times)

    def noisy_planes(field_grids):
        with torch.no_grad():
            for gs in field_grids:
                for p_ in gs:
                    p_.add_(0.3 * torch.randn(p_.shape, generator=g))

    # (a) contraction
    f = R.kplanes_field.KPlanesField(aabb, spacetime_resolution=res, feat_dim=c, multiscale_res=ms, concat_features_across_scales=True,
                                     linear_decoder=False, spatial_distortion=SceneContraction(order=float("inf")))
    noisy_planes(f.grids)
    out = f(rs)
    dens, rgb = out[R.kplanes_field.FieldHeadNames.DENSITY], out[R.kplanes_field.FieldHeadNames.RGB]
    gd, gr = torch.randn(dens.shape, generator=g), torch.randn(rgb.shape, generator=g)
    ((dens * gd).sum() + (rgb * gr).sum()).backward()
    arrs.update(con_density=dens, con_rgb=rgb, con_gd=gd, con_gr=gr)
    for i, gs in enumerate(f.grids):
        for j, p_ in enumerate(gs):
            arrs[f"con_grid_{i}_{j}"], arrs[f"con_ggrid_{i}_{j}"] = p_.detach(), p_.grad
    for name, net in (("sigma", f.sigma_net), ("color", f.color_net)):
        for i, lin in enumerate(net.layers):
            arrs[f"con_{name}_w{i}"], arrs[f"con_{name}_gw{i}"] = lin.weight.detach(), lin.weight.grad
    df = R.kplanes_field.KPlanesDensityField(aabb, resolution=[16, 14, 18, 5], feature_dim=c, linear_decoder=False,
                                             spatial_distortion=SceneContraction(order=float("inf")))
    noisy_planes([df.grids])
    dd, _ = df.get_density(rs)
    gdd = torch.randn(dd.shape, generator=g)
    (dd * gdd).sum().backward()
    arrs.update(pcon_density=dd, pcon_gd=gdd)
    for j, p_ in enumerate(df.grids):
        arrs[f"pcon_grid_{j}"], arrs[f"pcon_ggrid_{j}"] = p_.detach(), p_.grad
    for i, lin in enumerate(df.sigma_net.layers):
        arrs[f"pcon_w{i}"], arrs[f"pcon_gw{i}"] = lin.weight.detach(), lin.weight.grad

    # (b) linear decoder (bounded)
    lf = R.kplanes_field.KPlanesField(aabb * 3.0, spacetime_resolution=res, feat_dim=c, multiscale_res=ms,
                                      concat_features_across_scales=True, linear_decoder=True, linear_decoder_layers=2)
    noisy_planes(lf.grids)
    out = lf(rs)
    dens, rgb = out[R.kplanes_field.FieldHeadNames.DENSITY], out[R.kplanes_field.FieldHeadNames.RGB]
    gd, gr = torch.randn(dens.shape, generator=g), torch.randn(rgb.shape, generator=g)
    ((dens * gd).sum() + (rgb * gr).sum()).backward()
    arrs.update(lin_aabb=aabb * 3.0, lin_density=dens, lin_rgb=rgb, lin_gd=gd, lin_gr=gr)
    for i, gs in enumerate(lf.grids):
        for j, p_ in enumerate(gs):
            arrs[f"lin_grid_{i}_{j}"], arrs[f"lin_ggrid_{i}_{j}"] = p_.detach(), p_.grad
    for name, net in (("sigma", lf.sigma_net), ("basis", lf.color_basis)):
        for i, lin in enumerate(net.layers):
            arrs[f"lin_{name}_w{i}"], arrs[f"lin_{name}_gw{i}"] = lin.weight.detach(), lin.weight.grad
    save("field_variants", **arrs)


def gen_field_depths(R):
    """Decoder depths / widths other than the presets' (KPlanesModelConfig.sigma_net_layers / rgb_net_layers /
    *_hidden_dim, kplanes.py:96-103 -> tcnn n_hidden_layers / n_neurons, kplanes_field.py:249-273), by the reference's own
    KPlanesField: (a) two hidden sigma layers of 32 and one hidden colour layer of 48, view-dependent; (b) no hidden
    sigma layer at all and three hidden colour layers, disable_viewing_dependent.  Outputs and all gradients."""
    torch.manual_seed(97531)  # the reference fields' initial weights come from the global generator
    g = torch.Generator().manual_seed(654)
    res, ms, c = (12, 10, 14, 5), (1, 2), 8
    n, s = 80, 5
    aabb = torch.tensor([[-1.5, -1.5, -1.5], [1.5, 1.5, 1.5]])
    origins = (torch.rand(n, 3, generator=g) - 0.5) * 1.6
    d = torch.randn(n, 3, generator=g)
    directions = d / d.norm(dim=-1, keepdim=True)
    bins = torch.sort(torch.rand(n, s + 1, generator=g), -1).values * 1.5
    times = torch.rand(n, 1, generator=g)
    rb = R.rays.RayBundle(origins=origins, directions=directions, pixel_area=torch.ones(n, 1), times=times,
                          nears=torch.zeros(n, 1), fars=torch.full((n, 1), 1.5))
    rs = rb.get_ray_samples(bin_starts=bins[:, :-1, None], bin_ends=bins[:, 1:, None])
    arrs = dict(aabb=aabb, origins=origins, directions=directions, bins=bins, times=times)
    variants = {"a": dict(sigma_net_layers=2, sigma_net_hidden_dim=32, rgb_net_layers=1, rgb_net_hidden_dim=48),
                "b": dict(sigma_net_layers=0, rgb_net_layers=3, rgb_net_hidden_dim=64, disable_viewing_dependent=True)}
    for tag, kw in variants.items():
        f = R.kplanes_field.KPlanesField(aabb, spacetime_resolution=res, feat_dim=c, multiscale_res=ms,
                                         concat_features_across_scales=True, linear_decoder=False, **kw)
        with torch.no_grad():
            for gs in f.grids:
                for p_ in gs:
                    p_.add_(0.3 * torch.randn(p_.shape, generator=g))
        out = f(rs)
        dens, rgb = out[R.kplanes_field.FieldHeadNames.DENSITY], out[R.kplanes_field.FieldHeadNames.RGB]
        gd, gr = torch.randn(dens.shape, generator=g), torch.randn(rgb.shape, generator=g)
        ((dens * gd).sum() + (rgb * gr).sum()).backward()
        arrs.update({f"{tag}_density": dens, f"{tag}_rgb": rgb, f"{tag}_gd": gd, f"{tag}_gr": gr})
        for i, gs in enumerate(f.grids):
            for j, p_ in enumerate(gs):
                arrs[f"{tag}_grid_{i}_{j}"], arrs[f"{tag}_ggrid_{i}_{j}"] = p_.detach(), p_.grad
        for name, net in (("sigma", f.sigma_net), ("color", f.color_net)):
            assert len(net.layers) == kw[f"{'sigma' if name == 'sigma' else 'rgb'}_net_layers"] + 1
            for i, lin in enumerate(net.layers):
                arrs[f"{tag}_{name}_w{i}"], arrs[f"{tag}_{name}_gw{i}"] = lin.weight.detach(), lin.weight.grad
    save("field_depths", **arrs)


def main():
    torch.set_num_threads(1)
    R = load_reference()
    gen_interp(R)
    gen_samplers(R)
    gen_render(R)
    gen_losses(R)
    gen_model(R)
    gen_importance(R)
    gen_raygen(R)
    gen_samplers_cfg4(R)
    gen_field_variants(R)  # (its reference modules draw their initial weights from torch's GLOBAL generator, whose state
    # here is what the generators above left: later additions go below, and seed the global generator themselves)
    gen_raygen_lens(R)
    gen_raygen_crop(R)
    gen_pixel_samplers(R)
    gen_field_depths(R)


def _reference_method(path, class_name, method_name, extra_globals):
    """Compile ONE method of a reference class straight from its source file (the module itself cannot be imported on
    Python 3.12: its dataparser imports hit the mutable-dataclass-default error), so the fixture is still produced by
    the reference's own code."""
    import ast

    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == class_name:
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and fn.name == method_name:
                    fn.returns = None
                    for a in fn.args.args:
                        a.annotation = None
                    mod = ast.Module(body=[fn], type_ignores=[])
                    ast.fix_missing_locations(mod)
                    ns = dict(extra_globals)
                    exec(compile(mod, path, "exec"), ns)
                    return ns[method_name]
    raise KeyError(method_name)


def gen_importance(R):
    """a18: IST / ISG weight maps and the importance pixel sampler (NS/data/datasets/dynamic_dataset.py:215-470,
    NS/data/pixel_samplers.py:51-128, 340-426)."""
    import random
    import time
    import types

    from nerfstudio.data import pixel_samplers as ps

    g = torch.Generator().manual_seed(606)
    b, h, w = 12, 18, 24
    cam_ids = torch.tensor([0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 3])
    cam_times = torch.tensor([0.0, 0.1, 0.2, 0.6, 0.0, 0.005, 0.3, 0.31, 0.5, 0.5, 0.9, 0.4])[:, None]
    images = torch.rand(b, h, w, 3, generator=g) * 0.2
    for i in range(b):  # a moving blob per image so that temporal differences are sparse and non-trivial
        y, x = (3 + i) % (h - 4), (5 + 2 * i) % (w - 4)
        images[i, y:y + 4, x:x + 4] += 0.7
    path = os.path.join(os.environ.get("KPLANES_REFERENCE_ROOT", "/root/reference"), "nerfstudio", "nerfstudio", "data",
                        "datasets", "dynamic_dataset.py")
    glb = dict(torch=torch, time=time, DEBUG_IST_MAPS=False, tqdm=lambda x: x, str=str, print=lambda *a, **k: None)
    fake = types.SimpleNamespace(eval_dataset=False, ist_range=0.25, isg_gamma=5e-2,
                                 cameras=types.SimpleNamespace(times=cam_times, ids=cam_ids[:, None]))
    batch = {"image": images, "image_idx": torch.arange(b)}
    ist = _reference_method(path, "DynamicDataset", "compute_ist", glb)(fake, batch, "cpu")
    fake_isg = types.SimpleNamespace(eval_dataset=False, ist_range=0.25, isg_gamma=5e-2,
                                     cameras=types.SimpleNamespace(times=cam_times, ids=cam_ids))
    isg = _reference_method(path, "DynamicDataset", "compute_isg", glb)(fake_isg, batch, "cpu")

    ds = types.SimpleNamespace(iters_to_start_ist=100, is_pixel_ratio=0.3)
    sampler = ps.DynamicBasedPixelSampler(64, dataset=ds)
    out = {}
    for name, steps, weights in (("ist_on", 500, ist.float()), ("ist_off", 50, ist.float()), ("no_weights", 500, None)):
        torch.manual_seed(1234)
        random.seed(99)
        bt = {"image": images, "image_idx": torch.arange(b) + 100, "iter_steps": steps, "ist_weights": weights}
        col = sampler.collate_image_dataset_batch(bt, 64)
        out[f"{name}_indices"] = col["indices"]
        out[f"{name}_image"] = col["image"]
    torch.manual_seed(4321)
    uni = ps.PixelSampler(32).sample_method(32, b, h, w)
    save("importance", images=images, cam_ids=cam_ids, cam_times=cam_times, ist=ist.float(), isg=isg.float(), uniform=uni, **out)


def gen_pixel_samplers(R):
    """The two other pixel samplers DynamicDataManager._get_pixel_sampler can return (dynamic_datamanager.py:97-113):
    EquirectangularPixelSampler (pixel_samplers.py:228-267) and PatchPixelSampler (:270-327), by the reference's own
    classes under a fixed torch seed (the outputs are functions of torch's CPU random stream only)."""
    from nerfstudio.data import pixel_samplers as ps

    out = {}
    torch.manual_seed(2468)
    out["equirect"] = ps.EquirectangularPixelSampler(96).sample_method(96, 7, 40, 80)
    torch.manual_seed(1357)
    patch = ps.PatchPixelSampler(100, patch_size=4)
    out["patch_rays"] = torch.tensor(patch.num_rays_per_batch)
    out["patch"] = patch.sample_method(patch.num_rays_per_batch, 5, 30, 50)
    patch.set_num_rays_per_batch(50)
    out["patch_rays_after_set"] = torch.tensor(patch.num_rays_per_batch)
    save("pixel_samplers", **out)


def gen_raygen(R):
    """(f2) pixel -> ray generation by the reference's own Cameras (NS/cameras/cameras.py:327-741), driven the two ways
    the hot path's callers do: RayGenerator.forward (ray_generators.py:43-59: coords = image_coords[y, x],
    generate_rays(camera_indices=c[:, None], coords=coords)) and generate_rays(camera_indices=i, keep_shape=True) for a
    whole frame (scripts/render.py, base_model.py:162-186)."""
    from nerfstudio.cameras.cameras import Cameras

    g = torch.Generator().manual_seed(606)
    n_cams, h, w = 5, 36, 64
    # random rigid poses: QR of a random matrix, det forced to +1
    rot = torch.linalg.qr(torch.randn(n_cams, 3, 3, generator=g)).Q
    rot = rot * torch.sign(torch.linalg.det(rot))[:, None, None]
    pos = torch.randn(n_cams, 3, 1, generator=g) * 2.0
    c2w = torch.cat([rot, pos], dim=-1).float()
    fx = 50.0 + 20.0 * torch.rand(n_cams, 1, generator=g)
    fy = 50.0 + 20.0 * torch.rand(n_cams, 1, generator=g)
    cx = w / 2 + torch.randn(n_cams, 1, generator=g)
    cy = h / 2 + torch.randn(n_cams, 1, generator=g)
    times = torch.rand(n_cams, 1, generator=g)
    cams = Cameras(camera_to_worlds=c2w, fx=fx, fy=fy, cx=cx, cy=cy, width=w, height=h, times=times)
    n = 257
    ray_indices = torch.stack([torch.randint(0, n_cams, (n,), generator=g), torch.randint(0, h, (n,), generator=g),
                               torch.randint(0, w, (n,), generator=g)], dim=-1)
    image_coords = cams.get_image_coords()
    coords = image_coords[ray_indices[:, 1], ray_indices[:, 2]]
    rb = cams.generate_rays(camera_indices=ray_indices[:, 0].unsqueeze(-1), coords=coords)
    frame = cams.generate_rays(camera_indices=3, keep_shape=True)
    save("raygen", c2w=c2w, fx=fx, fy=fy, cx=cx, cy=cy, times=times, hw=torch.tensor([h, w]), ray_indices=ray_indices,
         origins=rb.origins, directions=rb.directions, pixel_area=rb.pixel_area, ray_times=rb.times,
         directions_norm=rb.metadata["directions_norm"], frame_cam=torch.tensor(3),
         frame_origins=frame.origins, frame_directions=frame.directions, frame_pixel_area=frame.pixel_area,
         frame_times=frame.times, frame_directions_norm=frame.metadata["directions_norm"])


def gen_raygen_lens(R):
    """(f2) the lens models of the reference's Cameras (cameras.py:635-697): OpenCV radial + tangential distortion undone
    by camera_utils.radial_and_tangential_undistort, fisheye and equirectangular direction models, mixed in ONE camera
    batch the way the reference's per-ray masks allow; also disable_distortion=True and a batch that is perspective with
    distortion only (what the dataparsers build from k1..p2 of transforms.json, broadcaststyle_dataparser.py:481-509)."""
    from nerfstudio.cameras.cameras import Cameras, CameraType

    g = torch.Generator().manual_seed(808)
    n_cams, h, w = 6, 40, 72
    rot = torch.linalg.qr(torch.randn(n_cams, 3, 3, generator=g)).Q
    rot = rot * torch.sign(torch.linalg.det(rot))[:, None, None]
    pos = torch.randn(n_cams, 3, 1, generator=g) * 2.0
    c2w = torch.cat([rot, pos], dim=-1).float()
    fx = 50.0 + 20.0 * torch.rand(n_cams, 1, generator=g)
    fy = 50.0 + 20.0 * torch.rand(n_cams, 1, generator=g)
    cx = w / 2 + torch.randn(n_cams, 1, generator=g)
    cy = h / 2 + torch.randn(n_cams, 1, generator=g)
    # camera 5 is equirectangular: fx = fy = height = width / 2 (cameras.py:689)
    fx[5], fy[5], cx[5], cy[5] = h, h, w / 2, h / 2
    times = torch.rand(n_cams, 1, generator=g)
    types = torch.tensor([[CameraType.PERSPECTIVE.value], [CameraType.FISHEYE.value], [CameraType.PERSPECTIVE.value],
                          [CameraType.FISHEYE.value], [CameraType.PERSPECTIVE.value], [CameraType.EQUIRECTANGULAR.value]])
    # k1, k2, k3, k4, p1, p2: broadcast-camera magnitudes, one camera with zeros, one strong enough that |det| crosses eps
    dist = torch.tensor([[-0.12, 0.03, -0.004, 0.0005, 0.002, -0.001],
                         [0.05, -0.01, 0.002, 0.0, 0.0, 0.0],
                         [0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
                         [-0.2, 0.05, 0.0, 0.0, -0.003, 0.004],
                         [-0.9, 0.3, 0.0, 0.0, 0.01, 0.01],
                         [0.1, 0.1, 0.1, 0.1, 0.1, 0.1]])
    cams = Cameras(camera_to_worlds=c2w, fx=fx, fy=fy, cx=cx, cy=cy, width=w, height=h, times=times,
                   distortion_params=dist, camera_type=types)
    n = 311
    ray_indices = torch.stack([torch.randint(0, n_cams, (n,), generator=g), torch.randint(0, h, (n,), generator=g),
                               torch.randint(0, w, (n,), generator=g)], dim=-1)
    coords = cams.get_image_coords()[ray_indices[:, 1], ray_indices[:, 2]]
    rb = cams.generate_rays(camera_indices=ray_indices[:, 0].unsqueeze(-1), coords=coords)
    rb_off = cams.generate_rays(camera_indices=ray_indices[:, 0].unsqueeze(-1), coords=coords, disable_distortion=True)
    out = dict(c2w=c2w, fx=fx, fy=fy, cx=cx, cy=cy, times=times, hw=torch.tensor([h, w]), types=types, dist=dist,
               ray_indices=ray_indices, origins=rb.origins, directions=rb.directions, pixel_area=rb.pixel_area,
               ray_times=rb.times, directions_norm=rb.metadata["directions_norm"],
               nodist_directions=rb_off.directions, nodist_pixel_area=rb_off.pixel_area)
    for cam in (0, 3, 4, 5):  # whole frames: distorted perspective, distorted fisheye, strong distortion, equirectangular
        fr = cams.generate_rays(camera_indices=cam, keep_shape=True)
        out[f"frame{cam}_directions"], out[f"frame{cam}_pixel_area"] = fr.directions, fr.pixel_area
        out[f"frame{cam}_directions_norm"] = fr.metadata["directions_norm"]
    # what the dataparsers build: perspective cameras sharing ONE distortion row
    persp = Cameras(camera_to_worlds=c2w, fx=fx, fy=fy, cx=cx, cy=cy, width=w, height=h, times=times,
                    distortion_params=dist[0], camera_type=CameraType.PERSPECTIVE)
    rbp = persp.generate_rays(camera_indices=ray_indices[:, 0].unsqueeze(-1), coords=coords)
    out["persp_directions"], out["persp_pixel_area"] = rbp.directions, rbp.pixel_area
    save("raygen_lens", **out)


def gen_raygen_crop(R):
    """(f2) crop-box rendering (scripts/render.py:101-106): Cameras.generate_rays(camera_indices=i, aabb_box=SceneBox)
    stores nears / fars from nerfstudio.utils.math.intersect_aabb (cameras.py:478-497; nerfacc is absent, so its
    _intersect_aabb branch runs, math.py:260-270) -- one whole frame of the raygen fixture's cameras against a box most
    rays miss, and the function alone on rays built to hit every branch (axis-parallel directions with 0 components,
    origins inside the box, on a face, rays pointing away, a 0/0 slab)."""
    from nerfstudio.cameras.cameras import Cameras
    from nerfstudio.data.scene_box import SceneBox
    from nerfstudio.utils import math as ns_math

    g0 = np.load(os.path.join(OUT, "raygen.npz"))
    t = {k: torch.from_numpy(g0[k]) for k in ("c2w", "fx", "fy", "cx", "cy", "times")}
    h, w = (int(v) for v in g0["hw"])
    cams = Cameras(camera_to_worlds=t["c2w"], fx=t["fx"], fy=t["fy"], cx=t["cx"], cy=t["cy"], width=w, height=h, times=t["times"])
    cam = int(g0["frame_cam"])
    o, d = torch.from_numpy(g0["frame_origins"])[0, 0], torch.from_numpy(g0["frame_directions"])
    centre = o + 3.0 * d[h // 2, w // 3]  # a box in front of the camera, off the optical axis
    box = torch.stack([centre - torch.tensor([0.9, 0.7, 0.8]), centre + torch.tensor([0.9, 0.7, 0.8])])
    frame = cams.generate_rays(camera_indices=cam, aabb_box=SceneBox(aabb=box))
    assert frame.nears.shape == (h, w, 1)
    hit = (frame.nears < 1e10).float().mean()
    assert 0.05 < float(hit) < 0.95, float(hit)
    g = torch.Generator().manual_seed(909)
    n = 512
    aabb = torch.tensor([-1.0, -0.5, -2.0, 1.5, 0.75, 0.25])
    origins = (torch.rand(n, 3, generator=g) - 0.5) * 6.0
    directions = torch.randn(n, 3, generator=g)
    directions = directions / directions.norm(dim=-1, keepdim=True)
    origins[:64] = (torch.rand(64, 3, generator=g) - 0.5) * torch.tensor([2.0, 1.0, 2.0]) + torch.tensor([0.25, 0.125, -0.875])
    directions[64:96] = torch.eye(3)[torch.randint(0, 3, (32,), generator=g)] * torch.tensor([1.0, -1.0, 1.0])  # zeros: +-inf slabs
    origins[96:104, 0] = -1.0  # on the x-min face ...
    directions[96:104] = torch.tensor([0.0, 0.6, 0.8])  # ... moving inside it: 0/0 = NaN in the x slab
    origins[104:112] = torch.tensor([3.0, 0.0, -1.0])
    directions[104:112] = torch.tensor([1.0, 0.0, 0.0])  # pointing away: both crossings negative -> clamped to 0 -> a miss
    t_min, t_max = ns_math.intersect_aabb(origins, directions, aabb)
    assert bool(torch.isnan(t_min).any()) and bool((t_min == 1e10).any()) and bool((t_min < 1e10).any())
    save("raygen_crop", frame_cam=torch.tensor(cam), box=box, frame_nears=frame.nears, frame_fars=frame.fars,
         aabb=aabb, origins=origins, directions=directions, t_min=t_min, t_max=t_max)


def gen_samplers_cfg4(R):
    """BASELINE config 4 (nerfplayer-nerfacto: only the sampler + compositing are shared with the path): the piecewise
    uniform / linear-in-disparity initial sampler and PDF resampling with single_jitter=True (nerfacto.py:125,
    nerfplayer_nerfacto.py:173-179), and DepthRenderer("expected") (:190), all by the reference's own classes."""
    g = torch.Generator().manual_seed(707)
    n = 48
    origins, directions, times, aabb = ko.synthetic_rays(n, g)
    RS = R.ray_samplers
    nears = 0.05 + 0.1 * torch.rand(n, 1, generator=g)
    fars = 2.0 + 4.0 * torch.rand(n, 1, generator=g)  # beyond distance 1: both branches of the piecewise spacing
    nears[5], fars[5] = 1.2, 1.9  # a ray entirely in the disparity branch
    nears[6], fars[6] = 0.1, 0.8  # and one entirely in the uniform branch
    arrs = dict(origins=origins, directions=directions, times=times, nears=nears, fars=fars)
    for mode in ("train", "eval"):
        ini = RS.UniformLinDispPiecewiseSampler(single_jitter=True)
        pdf = RS.PDFSampler(include_original=False, single_jitter=True)
        ini.train(mode == "train")
        pdf.train(mode == "train")
        s0, s1 = 64, 24
        t_rand = torch.rand(n, 1, generator=g)
        u_rand = torch.rand(n, 1, generator=g)
        weights = torch.rand(n, s0, 1, generator=g) ** 3
        rb = R.rays.RayBundle(origins=origins, directions=directions, pixel_area=torch.ones(n, 1), times=times,
                              nears=nears.clone(), fars=fars.clone())
        ss = []
        with rand_queue([t_rand, u_rand] if mode == "train" else []), record_searchsorted(ss):
            rs0 = ini(rb, num_samples=s0)
            rs1 = pdf(rb, rs0, weights, num_samples=s1)
        w1 = torch.rand(n, s1, 1, generator=g)
        w1 = w1 / w1.sum(-2, keepdim=True)
        depth = R.renderers.DepthRenderer(method="expected")(weights=w1, ray_samples=rs1)
        arrs.update({
            f"{mode}_t_rand": t_rand, f"{mode}_u_rand": u_rand, f"{mode}_weights": weights,
            f"{mode}_bins0": torch.cat([rs0.spacing_starts[..., 0], rs0.spacing_ends[..., -1:, 0]], -1),
            f"{mode}_starts0": rs0.frustums.starts[..., 0], f"{mode}_ends0": rs0.frustums.ends[..., 0],
            f"{mode}_bins1": torch.cat([rs1.spacing_starts[..., 0], rs1.spacing_ends[..., -1:, 0]], -1),
            f"{mode}_starts1": rs1.frustums.starts[..., 0], f"{mode}_ends1": rs1.frustums.ends[..., 0],
            f"{mode}_inds1": ss[0], f"{mode}_w1": w1, f"{mode}_depth_expected": depth,
        })
    save("samplers_cfg4", **arrs)


def check() -> list:
    """Regenerate every fixture into a scratch directory and compare it with the committed one (same keys, arrays equal
    bit for bit, NaNs in the same places).  -> names of the fixtures that differ or are missing on either side."""
    import tempfile

    global OUT
    committed = OUT
    with tempfile.TemporaryDirectory() as tmp:
        # gen_raygen_crop reads raygen.npz from OUT: the scratch copy it finds there was written by this very run
        OUT = tmp
        try:
            main()
        finally:
            OUT = committed
        names = sorted({f for d in (committed, tmp) for f in os.listdir(d) if f.endswith(".npz")})
        bad = []
        for f in names:
            pa, pb = os.path.join(committed, f), os.path.join(tmp, f)
            if not (os.path.exists(pa) and os.path.exists(pb)):
                bad.append(f)
                continue
            a, b = np.load(pa), np.load(pb)
            if set(a.files) != set(b.files) or not all(
                    a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k], equal_nan=a[k].dtype.kind == "f") for k in a.files):
                bad.append(f)
    return bad


if __name__ == "__main__":
    if "--check" in sys.argv[1:]:
        differing = check()
        print("fixtures that do not reproduce:", differing if differing else "none")
        sys.exit(1 if differing else 0)
    main()
