"""Drop-in boundary against the REAL reference objects: our ``KPlanesModel`` must be constructible from
the reference's own ``KPlanesModelConfig`` + ``SceneBox`` through the ``_target`` mechanism the plugin entry point uses
(NS/configs/base_config.py:50-58, soccernerfs_b200/configs/method_configs.py), expose the parameter groups the preset's
optimizers are keyed by, share the plane parameters' names and shapes with the reference model, and its host logic must
accept the reference's ``RayBundle`` / ``RaySamples`` dataclasses.  Runs only where /root/reference exists (the build
container); the reference is imported through oracle/ref_loader.py (test infrastructure)."""
import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference checkout not present")


def _preset_kwargs():
    # the model node of method_configs["k-planes"] (NS/configs/method_configs.py:513-545), small planes for speed
    return dict(
        eval_num_rays_per_chunk=1 << 15, multiscale_res=(1, 2), spacetime_resolution=(16, 16, 16, 10), feature_dim=32,
        concat_features_across_scales=True, disable_viewing_dependent=True,
        proposal_net_args_list=[{"feature_dim": 8, "resolution": (24, 24, 24, 10)}, {"feature_dim": 8, "resolution": (32, 32, 32, 10)}],
        sigma_net_layers=1, sigma_net_hidden_dim=128, rgb_net_layers=2, rgb_net_hidden_dim=64,
        num_proposal_samples_per_ray=(256, 128), num_nerf_samples_per_ray=64, bounded=True,
    )


def test_our_model_is_built_from_the_reference_config_via_target():
    from soccernerfs_b200.models.kplanes import KPlanesModel, KPlanesModelConfig

    ref = ref_loader.load_reference_model_module()
    rk = ref.models_kplanes
    aabb = torch.tensor([[-1.5, -1.5, -1.5], [1.5, 1.5, 1.5]])
    ref_cfg = rk.KPlanesModelConfig(**_preset_kwargs())
    ref_model = ref_cfg.setup(scene_box=ref.SceneBox(aabb=aabb), num_train_data=7)  # the reference's own model
    ours_cfg = rk.KPlanesModelConfig(**_preset_kwargs())
    ours_cfg._target = KPlanesModel  # what the plugin entry point does to the reference's TrainerConfig
    ours = ours_cfg.setup(scene_box=ref.SceneBox(aabb=aabb), num_train_data=7)
    assert isinstance(ours, KPlanesModel) and ours.config is ours_cfg
    # every config field the reference declares exists, with the same default, on our config class
    import dataclasses

    ref_fields = {f.name: f for f in dataclasses.fields(rk.KPlanesModelConfig)}
    own_fields = {f.name: f for f in dataclasses.fields(KPlanesModelConfig)}
    assert set(ref_fields) <= set(own_fields), set(ref_fields) - set(own_fields)
    ref_default, own_default = rk.KPlanesModelConfig(), KPlanesModelConfig()
    for name in ref_fields:
        if name == "_target":
            continue
        a, b = getattr(ref_default, name), getattr(own_default, name)
        a = a.value if hasattr(a, "value") else a
        b = b.value if hasattr(b, "value") else b
        assert (list(a) == list(b)) if isinstance(a, (tuple, list)) else (a == b), name
    # optimizer groups of the preset (method_configs.py:546-557)
    assert set(ours.get_param_groups()) == set(ref_model.get_param_groups()) == {"proposal_networks", "fields"}
    # planes and aabb: same state-dict names and shapes as the reference model (checkpoints interchange)
    ref_sd, own_sd = ref_model.state_dict(), ours.state_dict()
    plane_keys = [k for k in ref_sd if ".grids." in k or k.endswith("aabb")]
    assert len(plane_keys) == 2 * 6 + 2 * 6 + 3
    for k in plane_keys:
        assert k in own_sd and own_sd[k].shape == ref_sd[k].shape, k
    # decoders: same number of weights per network (tcnn keeps them flat; utils/checkpoint.py converts)
    def count(sd, prefix):
        return sum(v.numel() for k, v in sd.items() if k.startswith(prefix) and "grids" not in k and "aabb" not in k)

    for prefix in ("field.sigma_net", "field.color_net", "proposal_networks.0.sigma_net", "proposal_networks.1.sigma_net"):
        assert count(own_sd, prefix) == count(ref_sd, prefix) > 0, prefix
    # callbacks: the two the reference registers (anneal before, sampler step after each iteration, kplanes.py:318-347)
    assert len(ours.get_training_callbacks(None)) == len(ref_model.get_training_callbacks(None)) == 2
    assert ours.temporal_distortion == ref_model.temporal_distortion


def test_host_logic_accepts_reference_ray_dataclasses():
    """The reference's RayBundle -> get_ray_samples -> RaySamples objects feed our field's ray-form extraction and the
    samplers' spacing-edge helper unchanged (duck typing on the TensorDataclass fields, NS/cameras/rays.py:105-193)."""
    from soccernerfs_b200.fields.kplanes_field import _ray_form
    from soccernerfs_b200.model_components.ray_samplers import spacing_edges

    ref = ref_loader.load_reference()
    n, s = 5, 7
    gen = torch.Generator().manual_seed(0)
    d = torch.randn(n, 3, generator=gen)
    rb = ref.rays.RayBundle(origins=torch.rand(n, 3, generator=gen), directions=d / d.norm(dim=-1, keepdim=True),
                            pixel_area=torch.ones(n, 1), nears=torch.zeros(n, 1), fars=torch.full((n, 1), 3.0),
                            times=torch.rand(n, 1, generator=gen))
    bins = torch.sort(torch.rand(n, s + 1, generator=gen), -1).values
    rs = rb.get_ray_samples(bin_starts=bins[:, :-1, None] * 3, bin_ends=bins[:, 1:, None] * 3, spacing_starts=bins[:, :-1, None],
                            spacing_ends=bins[:, 1:, None], spacing_to_euclidean_fn=lambda x: x * 3)
    o, dd, st, en, t = _ray_form(rs)
    assert o.shape == (n, 3) and dd.shape == (n, 3) and st.shape == (n, s) and en.shape == (n, s) and t.shape == (n,)
    assert torch.equal(o, rb.origins) and torch.equal(t, rb.times[:, 0]) and torch.equal(st, bins[:, :-1] * 3)
    assert torch.equal(spacing_edges(rs), bins)


def test_method_specification_retargets_the_reference_trainer_config():
    """configs/method_configs.py::make_method_specification on a stand-in TrainerConfig tree: only the model node's
    _target changes; the datamanager / optimizer nodes (IST/ISG options) are the reference's own objects, untouched."""
    import sys
    import types
    from dataclasses import dataclass, field

    from soccernerfs_b200.models.kplanes import KPlanesModel

    ref = ref_loader.load_reference_model_module()
    rk = ref.models_kplanes

    @dataclass
    class MethodSpecification:
        config: object
        description: str

    @dataclass
    class Node:
        model: object = None
        datamanager: object = None

    @dataclass
    class Trainer:
        method_name: str = "k-planes"
        pipeline: Node = field(default_factory=Node)

    mod = types.ModuleType("nerfstudio.plugins.types")
    mod.MethodSpecification = MethodSpecification
    had = sys.modules.get("nerfstudio.plugins.types")
    sys.modules["nerfstudio.plugins.types"] = mod
    try:
        from soccernerfs_b200.configs.method_configs import make_method_specification

        dm = {"ist_range": 1.0, "is_pixel_ratio": 0.15}
        base = Trainer(pipeline=Node(model=rk.KPlanesModelConfig(**_preset_kwargs()), datamanager=dm))
        spec = make_method_specification(base)
    finally:
        if had is None:
            del sys.modules["nerfstudio.plugins.types"]
        else:
            sys.modules["nerfstudio.plugins.types"] = had
    assert spec.config.method_name == "k-planes" and spec.config.pipeline.model._target is KPlanesModel
    assert base.pipeline.model._target is rk.KPlanesModel  # the reference's preset object itself is not modified
    assert spec.config.pipeline.datamanager == dm


def test_cameras_from_reference_cameras():
    """Cameras.from_reference reads the reference's own Cameras (NS/cameras/cameras.py:56-146) by its field names: poses,
    intrinsics, sizes, per-camera types, OpenCV distortion, times and ids arrive unchanged (host logic only)."""
    from soccernerfs_b200.cameras.cameras import Cameras

    ref_loader.load_reference()
    from nerfstudio.cameras.cameras import Cameras as RefCameras
    from nerfstudio.cameras.cameras import CameraType as RefCameraType

    g = torch.Generator().manual_seed(3)
    n = 4
    c2w = torch.cat([torch.linalg.qr(torch.randn(n, 3, 3, generator=g)).Q, torch.randn(n, 3, 1, generator=g)], dim=-1)
    dist = torch.tensor([[0.1, -0.02, 0.0, 0.0, 0.001, 0.002]]).repeat(n, 1)
    ref = RefCameras(camera_to_worlds=c2w, fx=torch.rand(n, 1, generator=g) * 100 + 500, fy=600.0, cx=480.0, cy=270.0, width=960,
                     height=540, distortion_params=dist, times=torch.rand(n, 1, generator=g), ids=torch.arange(n).float()[:, None],
                     camera_type=[RefCameraType.PERSPECTIVE, RefCameraType.FISHEYE, RefCameraType.PERSPECTIVE, RefCameraType.EQUIRECTANGULAR])
    ours = Cameras.from_reference(ref)
    assert len(ours) == n and torch.equal(ours.camera_to_worlds, ref.camera_to_worlds)
    for name in ("fx", "fy", "cx", "cy", "times"):
        assert torch.equal(getattr(ours, name), getattr(ref, name)), name
    assert torch.equal(ours.width, ref.width.long()) and torch.equal(ours.height, ref.height.long())
    assert torch.equal(ours.camera_type, ref.camera_type.long()) and ours._cam_types.tolist() == [1, 2, 1, 3]
    assert torch.equal(ours.distortion_params, ref.distortion_params) and torch.equal(ours.ids, ref.ids)
    assert ours._height_host == [540] * n and ours._width_host == [960] * n
    plain = Cameras.from_reference(RefCameras(camera_to_worlds=c2w, fx=500.0, fy=500.0, cx=480.0, cy=270.0,
                                              distortion_params=torch.zeros(6)))  # what the dataparsers build without k1..p2
    assert plain._distortion is None and plain._cam_types is None and plain._width_host == [960] * n


def test_committed_fixtures_reproduce_from_the_reference():
    """oracle/make_golden.py run against the reference checkout regenerates every committed tests/golden/*.npz bit for bit
    (what pins the oracle, and through it the CUDA path, to the reference's own outputs)."""
    from oracle import make_golden

    threads = torch.get_num_threads()
    try:
        assert make_golden.check() == []
    finally:
        torch.set_num_threads(threads)
