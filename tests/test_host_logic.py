"""CPU tests of the host-side mirror of the reference interface (no kernels involved)."""
import numpy as np
import pytest
import torch

from soccernerfs_b200.cameras.rays import Frustums, RayBundle, RaySamples
from soccernerfs_b200.data.scene_box import SceneBox
from soccernerfs_b200.models.kplanes import KPlanesModelConfig


def test_frustum_positions_known_answer():
    """Reference KAT: REF/nerfstudio/tests/cameras/test_rays.py:11-30."""
    fr = Frustums(origins=torch.ones((5, 3)), directions=torch.tensor([[0.0, 1.0, 0.0]] * 5) * torch.tensor([1.0, 2.5, 1.0]),
                  starts=torch.ones((5, 1)), ends=torch.ones((5, 1)), pixel_area=torch.ones((5, 1)))
    fr = Frustums(origins=torch.tensor([0.0, 1.0, 2.0]).expand(5, 3), directions=torch.tensor([0.0, 1.0, 0.0]).expand(5, 3),
                  starts=torch.full((5, 1), 2.0), ends=torch.full((5, 1), 3.0), pixel_area=torch.ones((5, 1)))
    assert torch.allclose(fr.get_positions(), torch.tensor([0.0, 3.5, 2.0]).expand(5, 3))


def test_tensor_dataclass_broadcast_and_index():
    rb = RayBundle(origins=torch.zeros(4, 3), directions=torch.ones(4, 3), pixel_area=torch.ones(1, 1), times=torch.rand(4, 1))
    assert rb.shape == (4,) and rb.pixel_area.shape == (4, 1) and len(rb) == 4
    rs = rb.get_ray_samples(bin_starts=torch.zeros(4, 7, 1), bin_ends=torch.ones(4, 7, 1))
    assert rs.shape == (4, 7) and rs.frustums.origins.shape == (4, 7, 3) and rs.times.shape == (4, 7, 1)
    assert rs.frustums.origins.stride(-2) == 0  # per-ray broadcast is a view (the kernels rely on this)
    assert rs[1:3].shape == (2, 7) and rs[..., 0].shape == (4,)
    assert rb.flatten()[1:3].origins.shape == (2, 3)
    img = RayBundle(origins=torch.zeros(2, 3, 3), directions=torch.ones(2, 3, 3), pixel_area=torch.ones(2, 3, 1))
    assert len(img) == 6 and img.get_row_major_sliced_ray_bundle(1, 5).shape == (4,)
    with pytest.raises(RuntimeError):
        rb[0] = rb[1]


def test_config_defaults_match_reference():
    """NS/models/kplanes.py:67-177."""
    c = KPlanesModelConfig()
    assert c.spacetime_resolution == (64, 64, 64, 50) and c.feature_dim == 32 and c.multiscale_res == (1, 2, 4, 8)
    assert c.num_proposal_samples_per_ray == (256, 128) and c.num_nerf_samples_per_ray == 48
    assert c.proposal_net_args_list[0] == {"feature_dim": 8, "resolution": [128, 128, 128, 150]}
    assert c.loss_coefficients["distortion_loss"] == 0.001 and c.loss_coefficients["time_smoothness_proposal_loss"] == 0.00001
    assert c.background_color_train == "random" and c.background_color_eval == "last_sample" and c.bounded


def test_model_surface_and_state_dict_layout():
    m = KPlanesModelConfig(spacetime_resolution=(8, 8, 8, 4), multiscale_res=(1, 2),
                           proposal_net_args_list=[{"feature_dim": 8, "resolution": [8, 8, 8, 4]}]).setup(
        scene_box=SceneBox(aabb=torch.tensor([[-1.0] * 3, [1.0] * 3])), num_train_data=3)
    groups = m.get_param_groups()
    assert set(groups) == {"proposal_networks", "fields"}
    sd = m.state_dict()
    # reference checkpoint layout: field.grids.<scale>.<plane> of logical shape [1,C,H,W]
    assert sd["field.grids.1.2"].shape == (1, 32, 4, 16) and sd["proposal_networks.0.grids.5"].shape == (1, 8, 4, 8)
    assert sd["field.grids.0.0"].permute(0, 2, 3, 1).is_contiguous()  # channel-last memory
    assert float(sd["field.grids.0.2"].min()) == 1.0  # time planes initialised to 1
    assert 0.1 <= float(sd["field.grids.0.0"].min()) and float(sd["field.grids.0.0"].max()) <= 0.5
    assert 0.1 <= float(sd["proposal_networks.0.grids.0"].min()) and float(sd["proposal_networks.0.grids.0"].max()) <= 0.15
    # loading a reference-layout (NCHW-contiguous) checkpoint keeps channel-last storage
    ref_like = {k: v.contiguous() for k, v in sd.items()}
    m.load_state_dict(ref_like)
    assert m.field.grids[0][0].permute(0, 2, 3, 1).is_contiguous()


def test_proposal_schedule_and_anneal_callbacks():
    """update schedule kplanes.py:254-259; anneal kplanes.py:326-331; sampler state ray_samplers.py:546-557, :573."""
    m = KPlanesModelConfig(spacetime_resolution=(8, 8, 8, 4), multiscale_res=(1,),
                           proposal_net_args_list=[{"feature_dim": 8, "resolution": [8, 8, 8, 4]}]).setup(
        scene_box=SceneBox(aabb=torch.tensor([[-1.0] * 3, [1.0] * 3])), num_train_data=3)
    cbs = m.get_training_callbacks(None)
    assert len(cbs) == 2
    cbs[0].run_callback_at_location(step=100, location=1)
    frac = 100 / 1000
    assert m.proposal_sampler._anneal == pytest.approx(10 * frac / (9 * frac + 1))
    cbs[1].run_callback_at_location(step=100, location=2)
    assert m.proposal_sampler._step == 100 and m.proposal_sampler._steps_since_update == 1
    sched = m.proposal_sampler.update_sched
    assert sched(0) == 1 and sched(5000) == 5 and sched(2500) == pytest.approx(2.5) and sched(10**6) == 5


def test_cosine_schedule():
    from soccernerfs_b200.engine.optimizers import cosine_decay_factor

    assert cosine_decay_factor(0, 512, 30000) == 0.0 and cosine_decay_factor(256, 512, 30000) == 0.5
    assert cosine_decay_factor(512, 512, 30000) == pytest.approx(1.0)
    assert cosine_decay_factor(30000, 512, 30000) == pytest.approx(0.0, abs=1e-12)
    mid = (30000 + 512) // 2
    assert cosine_decay_factor(mid, 512, 30000) == pytest.approx(0.5, abs=1e-3)


def test_non_default_branches_construct_with_reference_shapes():
    """linear decoder / appearance embedding / scene contraction (kplanes_field.py:196-246, 278-280): same module names and
    parameter shapes as the reference; anything the kernels do not evaluate (a non-L-inf contraction) fails loudly."""
    from soccernerfs_b200.field_components.spatial_distortions import SceneContraction
    from soccernerfs_b200.fields.kplanes_field import KPlanesDensityField, KPlanesField

    aabb = torch.tensor([[-1.0] * 3, [1.0] * 3])
    kw = dict(spacetime_resolution=(8, 8, 8, 4), feat_dim=8, multiscale_res=(1, 2), concat_features_across_scales=True)
    lin = KPlanesField(aabb, linear_decoder=True, linear_decoder_layers=2, use_appearance_embedding=True, appearance_dim=5,
                       num_images=7, **kw)
    assert [tuple(w.shape) for w in lin.color_basis.weights] == [(128, 8), (128, 128), (48, 128)]  # 3 + 5 -> 128 -> 128 -> 3 * 16
    assert [tuple(w.shape) for w in lin.sigma_net.weights] == [(1, 16)] and not hasattr(lin, "color_net")
    assert tuple(lin.appearance_embedding.embedding.weight.shape) == (7, 5)
    app = KPlanesField(aabb, linear_decoder=False, use_appearance_embedding=True, appearance_dim=5, num_images=7, **kw)
    assert app.color_net.weights[0].shape == (64, 16 + 15 + 5)
    con = KPlanesField(aabb, linear_decoder=False, spatial_distortion=SceneContraction(order=float("inf")), **kw)
    assert con._contract and KPlanesDensityField(aabb, [8, 8, 8, 4], 8, spatial_distortion=SceneContraction(order=float("inf")))._contract
    with pytest.raises(NotImplementedError):
        KPlanesField(aabb, linear_decoder=False, spatial_distortion=SceneContraction(order=None), **kw)
    x = torch.tensor([[0.5, 0.2, 0.1], [3.0, -1.0, 0.5]])
    assert torch.allclose(SceneContraction(order=float("inf"))(x), torch.tensor([[0.5, 0.2, 0.1], [5 / 3, -5 / 9, 5 / 18]]))


def test_grad_bucket_span_of():
    """GradBucket.span_of: element range of adjacent parameters, None when they are not adjacent."""
    import torch

    from soccernerfs_b200.distributed import GradBucket

    ps = [torch.nn.Parameter(torch.zeros(n)) for n in (3, 5, 7, 2)]
    b = GradBucket(ps)
    assert b.span_of([ps[1], ps[2]]) == (3, 15)
    assert b.span_of([ps[3]]) == (15, 17)
    assert b.span_of([ps[0], ps[2]]) is None
    assert b.span_of([torch.nn.Parameter(torch.zeros(1))]) is None


def test_tcnn_param_layout_round_trip():
    """utils/checkpoint.py: flat tcnn params <-> our weight matrices (first-layer input and last-layer output padded
    to 16, row-major [out, in], layer order)."""
    import torch

    from soccernerfs_b200.utils.checkpoint import tcnn_layer_shapes, tcnn_params_to_weights, weights_to_tcnn_params

    assert tcnn_layer_shapes([31, 64, 64, 3]) == [(64, 32), (64, 64), (16, 64)]
    assert tcnn_layer_shapes([8, 64, 1]) == [(64, 16), (16, 64)]
    gen = torch.Generator().manual_seed(0)
    ws = [torch.randn(64, 31, generator=gen), torch.randn(64, 64, generator=gen), torch.randn(3, 64, generator=gen)]
    flat = weights_to_tcnn_params(ws)
    assert flat.numel() == 64 * 32 + 64 * 64 + 16 * 64
    assert float(flat[:32][31]) == 0.0 and torch.equal(flat[:31], ws[0][0])  # row 0 of layer 0, then its padding column
    back = tcnn_params_to_weights(flat, [31, 64, 64, 3])
    assert all(torch.equal(a, b) for a, b in zip(ws, back))
    import pytest

    with pytest.raises(ValueError):
        tcnn_params_to_weights(flat[:-1], [31, 64, 64, 3])


def test_reference_checkpoint_round_trip():
    """A pipeline state dict in the reference's naming / layouts (planes NCHW-contiguous under ``_model.``, MLPs as flat
    tcnn ``params``) loads into the channel-last model and back without loss."""
    import torch

    from soccernerfs_b200 import ops
    from soccernerfs_b200.utils.checkpoint import load_reference_state_dict, to_reference_state_dict
    from tests.helpers import model_config
    from soccernerfs_b200.data.scene_box import SceneBox

    aabb = torch.tensor([[-1.5] * 3, [1.5] * 3])
    torch.manual_seed(1)
    a = model_config("tiny").setup(scene_box=SceneBox(aabb=aabb), num_train_data=1)
    torch.manual_seed(2)
    b = model_config("tiny").setup(scene_box=SceneBox(aabb=aabb), num_train_data=1)
    sd = to_reference_state_dict(a)
    assert "_model.field.grids.0.0" in sd and sd["_model.field.grids.0.0"].is_contiguous()
    assert "_model.field.sigma_net.params" in sd and "_model.proposal_networks.1.sigma_net.params" in sd
    sd["_datamanager.train_camera_optimizer.pose_adjustment"] = torch.zeros(3, 6)
    unused = load_reference_state_dict(b, {"step": 7, "pipeline": sd})
    # the parameter-less tcnn SH encoding and foreign (datamanager) keys are reported, not consumed
    assert unused == ["_model.field.direction_encoder.params", "_datamanager.train_camera_optimizer.pose_adjustment"]
    assert sd["_model.field.direction_encoder.params"].numel() == 0
    for (na, pa), (nb, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert na == nb and torch.equal(pa, pb), na
    assert ops.is_channel_last(b.field.grids[0][0])  # storage layout untouched by the load


def test_cameras_reject_unsupported_and_out_of_range():
    """cameras/cameras.py host-side contract: camera types / distortion parsed like cameras.py:178-218 (all-zero
    distortion and all-perspective batches select the plain kernel), camera-optimizer deltas rejected, host indices
    range-checked."""
    import pytest
    import torch

    from soccernerfs_b200.cameras.cameras import Cameras, CameraType

    c2w = torch.eye(4)[:3].repeat(2, 1, 1)
    cams = Cameras(c2w, 50.0, 50.0, 32.0, 18.0, 64, 36, times=torch.tensor([0.1, 0.2]))
    assert cams.shape == (2,) and len(cams) == 2 and cams.get_image_coords().shape == (36, 64, 2)
    assert float(cams.get_image_coords()[3, 5, 0]) == 3.5 and float(cams.get_image_coords()[3, 5, 1]) == 5.5
    assert cams._cam_types is None and cams._distortion is None  # the plain perspective kernel
    assert torch.equal(cams.camera_type, torch.ones(2, 1, dtype=torch.int64))
    fish = Cameras(c2w, 50.0, 50.0, 32.0, 18.0, 64, 36, camera_type=CameraType.FISHEYE)
    assert fish._cam_types.dtype == torch.int32 and fish._cam_types.tolist() == [2, 2] and fish._distortion is None
    mixed = Cameras(c2w, 50.0, 50.0, 32.0, 18.0, 64, 36, camera_type=[CameraType.PERSPECTIVE, CameraType.EQUIRECTANGULAR],
                    distortion_params=torch.tensor([0.1, 0, 0, 0, 0, 0.01]))  # one row shared by all cameras
    assert mixed._cam_types.tolist() == [1, 3] and mixed.distortion_params.shape == (2, 6)
    assert mixed._lens(disable_distortion=True) is None and mixed._lens(False) is mixed._distortion
    for ct in (torch.tensor([1, 2]), torch.tensor([[1], [2]]), 2):
        assert Cameras(c2w, 50.0, 50.0, 32.0, 18.0, 64, 36, camera_type=ct).camera_type.shape == (2, 1)
    zero = Cameras(c2w, 50.0, 50.0, 32.0, 18.0, 64, 36, distortion_params=torch.zeros(2, 6))
    assert zero.distortion_params is None  # Newton on an all-zero model is the identity: dropped
    with pytest.raises(ValueError):
        Cameras(c2w, 50.0, 50.0, 32.0, 18.0, 64, 36, camera_type=torch.tensor([1, 7]))  # cameras.py:699-701
    with pytest.raises(AssertionError):
        Cameras(c2w, 50.0, 50.0, 32.0, 18.0, 64, 36, camera_type=torch.tensor([1.0, 2.0]))  # cameras.py:203-205
    with pytest.raises(ValueError):
        Cameras(c2w, 50.0, 50.0, 32.0, 18.0, 64, 36, camera_type="fisheye")
    with pytest.raises(ValueError):
        Cameras(c2w, 50.0, 50.0, 32.0, 18.0, 64, 36, distortion_params=torch.ones(2, 4))
    with pytest.raises(ValueError):
        Cameras(c2w, 50.0, 50.0, 32.0, 18.0, 64, 36, distortion_params=torch.ones(3, 6))
    with pytest.raises(NotImplementedError):
        cams.generate_rays(camera_indices=0, coords=torch.tensor([[0.5, 0.5]]), distortion_params_delta=torch.zeros(1, 6))
    with pytest.raises(IndexError):
        cams.generate_rays_from_indices(torch.tensor([[2, 0, 0]]))
    with pytest.raises(IndexError):
        cams.generate_rays_from_indices(torch.tensor([[0, 36, 0]]))
    with pytest.raises(NotImplementedError):
        cams.generate_rays(camera_indices=0, coords=torch.tensor([[0.25, 0.5]]))  # two different sub-pixel offsets
    # one shared sub-pixel offset goes to the kernel as its pixel_offset with the integer parts as indices: pixel centres
    # (0.5) and the integer coordinates of the reference's own test (tests/cameras/test_cameras.py:119-122, offset 0)
    from soccernerfs_b200 import ops

    seen = {}

    def fake_generate_rays(c2w_, intr, times, ray_indices=None, pixel_offset=0.5, **kw):
        seen.update(ray_indices=ray_indices.clone(), pixel_offset=pixel_offset, kw=kw)
        n = ray_indices.shape[0]
        return torch.zeros(n, 3), torch.zeros(n, 3), torch.ones(n), torch.ones(n), torch.zeros(n)

    real, ops.generate_rays = ops.generate_rays, fake_generate_rays
    try:
        rb = cams.generate_rays(camera_indices=1, coords=torch.ones(10, 2))
        assert seen["pixel_offset"] == 0.0 and seen["ray_indices"].tolist() == [[1, 1, 1]] * 10 and rb.origins.shape == (10, 3)
        assert seen["kw"]["distortion"] is None and seen["kw"]["cam_types"] is None
        cams.generate_rays(camera_indices=torch.tensor([[0], [1]]), coords=torch.tensor([[3.5, 5.5], [0.5, 63.5]]))
        assert seen["pixel_offset"] == 0.5 and seen["ray_indices"].tolist() == [[0, 3, 5], [1, 0, 63]]
        mixed.generate_rays(camera_indices=0, coords=torch.tensor([[2.25, 7.25]]), disable_distortion=True)
        assert seen["pixel_offset"] == 0.25 and seen["kw"]["distortion"] is None and seen["kw"]["cam_types"] is not None
    finally:
        ops.generate_rays = real


def test_reference_optimizer_state_round_trip():
    """utils/checkpoint.py: Adam moments in the reference's numbering / layouts (planes NCHW, flat tcnn vectors with the
    SH sign flips) <-> our per-tensor FusedAdam state."""
    from soccernerfs_b200.data.scene_box import SceneBox
    from soccernerfs_b200.engine.optimizers import Optimizers
    from soccernerfs_b200.models.kplanes import KPlanesModelConfig
    from soccernerfs_b200.utils.checkpoint import (TCNN_SH_SIGNS, _reference_group_layout, load_reference_optimizer_state,
                                                   to_reference_optimizer_state, tcnn_layer_shapes)

    cfg = KPlanesModelConfig(spacetime_resolution=(8, 8, 8, 4), multiscale_res=(1, 2), num_nerf_samples_per_ray=8,
                             proposal_net_args_list=[{"feature_dim": 8, "resolution": [8, 8, 8, 4]}] * 2)
    aabb = torch.tensor([[-1.0, -1, -1], [1, 1, 1]])
    torch.manual_seed(0)
    models = [cfg.setup(scene_box=SceneBox(aabb=aabb), num_train_data=1) for _ in range(2)]
    opts = [Optimizers(m.get_param_groups()) for m in models]
    for group, opt in opts[0].optimizers.items():  # a populated state on the first model
        for p in models[0].get_param_groups()[group]:
            if p.requires_grad:
                opt.state[p] = {"step": 7, "exp_avg": torch.randn_like(p, memory_format=torch.preserve_format),
                                "exp_avg_sq": torch.rand_like(p, memory_format=torch.preserve_format)}
    ref = to_reference_optimizer_state(opts[0], models[0])
    layout = _reference_group_layout(models[0], "fields")
    kinds = [k for k, _ in layout]
    assert kinds == ["tensor"] * 13 + ["empty", "tcnn", "tcnn"]  # aabb, 12 planes, SH encoding, sigma_net, color_net
    assert set(ref["fields"]["state"]) == set(range(1, 13)) | {14, 15}  # aabb (no grad) and the SH encoding carry no state
    color = ref["fields"]["state"][15]
    assert color["exp_avg"].numel() == sum(o * i for o, i in tcnn_layer_shapes([31, 64, 64, 3]))
    w0 = opts[0].optimizers["fields"].state[models[0].field.color_net.weights[0]]["exp_avg"]
    assert torch.equal(color["exp_avg"][:32 * 64].view(64, 32)[:, :16], w0[:, :16] * TCNN_SH_SIGNS)
    assert ref["fields"]["state"][1]["exp_avg"].is_contiguous()  # planes in NCHW order
    load_reference_optimizer_state(opts[1], models[1], ref)
    for group in ("fields", "proposal_networks"):
        for p, q in zip(models[0].get_param_groups()[group], models[1].get_param_groups()[group]):
            if p.requires_grad:
                a, b = opts[0].optimizers[group].state[p], opts[1].optimizers[group].state[q]
                assert b["step"] == 7 and torch.equal(a["exp_avg"], b["exp_avg"]) and torch.equal(a["exp_avg_sq"], b["exp_avg_sq"])
                assert b["exp_avg"].stride() == q.stride()


def test_graph_capture_eligibility_of_a_batch():
    """TrainStep._graphable: the captured step has static buffers for origins / directions / times / image only, so a
    batch or bundle carrying something the model would USE (depth supervision, preset near / far bounds, unknown per-ray
    metadata, an appearance embedding's camera indices, no times) takes the eager iteration; what a datamanager adds and
    the model never reads (indices, ist_weights, mask, directions_norm without depth supervision) does not."""
    import types

    import torch

    from soccernerfs_b200.cameras.rays import RayBundle
    from soccernerfs_b200.engine.trainer import TrainStep

    n = 4
    me = types.SimpleNamespace(model=types.SimpleNamespace(field=types.SimpleNamespace(use_appearance_embedding=False)))
    ok = lambda rb, batch: TrainStep._graphable(me, rb, batch)  # noqa: E731

    def bundle(**kw):
        base = dict(origins=torch.zeros(n, 3), directions=torch.ones(n, 3), pixel_area=torch.ones(n, 1), times=torch.zeros(n, 1))
        base.update(kw)
        return RayBundle(**base)

    image = {"image": torch.zeros(n, 3)}
    assert ok(bundle(), image)
    assert ok(bundle(metadata={"directions_norm": torch.ones(n, 1)}, camera_indices=torch.zeros(n, 1, dtype=torch.long)),
              {**image, "indices": torch.zeros(n, 3), "ist_weights": torch.zeros(n), "mask": torch.ones(n, 1)})
    assert not ok(bundle(), {**image, "depth_image": torch.ones(n, 1)})
    assert not ok(bundle(), {"indices": torch.zeros(n, 3)})
    assert not ok(bundle(times=None), image)
    assert not ok(bundle(nears=torch.zeros(n, 1), fars=torch.ones(n, 1)), image)
    assert not ok(bundle(metadata={"directions_norm": torch.ones(n, 1), "something_else": torch.ones(n, 1)}), image)
    me.model.field.use_appearance_embedding = True
    assert not ok(bundle(), image)
