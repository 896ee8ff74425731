"""The ``ns-train k-planes`` preset, with the model on the B200 kernels.

Two forms of the same configuration (NS/configs/method_configs.py:481-560; README.md:39-45 for the 32x variant):

* ``KPLANES_MODEL`` / ``KPLANES_DATAMANAGER`` / ``KPLANES_OPTIMIZERS`` -- this package's own config objects with the
  preset's values, usable without nerfstudio (bench.py, tests, ``engine.trainer.TrainStep``).
* ``kplanes_b200`` -- a ``nerfstudio.plugins.types.MethodSpecification`` for the entry-point group
  ``nerfstudio.method_configs`` (NS/plugins/registry.py:32-51; declared in this repo's ``pyproject.toml``).  It is the
  reference's OWN ``method_configs["k-planes"]`` TrainerConfig -- datamanager, IST/ISG options, optimizers, schedulers,
  viewer untouched -- with the model node's ``_target`` pointed at ``soccernerfs_b200.models.kplanes.KPlanesModel``
  (every config node instantiates ``self._target(self, **kwargs)``, NS/configs/base_config.py:50-58).  Discovered
  methods are merged after the built-ins (method_configs.py:700-702), so under the same name ``k-planes`` the new path
  replaces the reference's and every CLI spelling keeps working, e.g.
  ``ns-train k-planes --pipeline.datamanager.ist-range 0.75 --pipeline.model.multiscale-res 1 2 4 8 16 32
  broadcaststyle-data --fps-downsample 4``.  It only exists when nerfstudio itself is importable.
"""
from __future__ import annotations

import copy
from typing import Dict, Optional

from ..data.datamanagers.dynamic_datamanager import DynamicDataManagerConfig
from ..models.kplanes import KPlanesModel, KPlanesModelConfig

PRESET_LOSS_COEFFICIENTS = {
    "rgb_loss": 1.0, "interlevel_loss": 1.0, "distortion_loss": 0.001, "space_tv_loss": 0.02, "time_smoothness_loss": 1.0,
    "sparse_transients_loss": 0.001, "space_tv_proposal_loss": 0.02, "time_smoothness_proposal_loss": 1.0,
    "sparse_transients_proposal_loss": 0.001, "depth_loss": 0.05,
}


def kplanes_model_config(multiscale_res=(1, 2, 4, 8, 16)) -> KPlanesModelConfig:
    """``pipeline.model`` of the preset (method_configs.py:513-545); ``multiscale_res=(1, 2, 4, 8, 16, 32)`` is the
    "32x" run of README.md:39-45."""
    return KPlanesModelConfig(
        eval_num_rays_per_chunk=1 << 15, multiscale_res=tuple(multiscale_res), spacetime_resolution=(64, 64, 64, 100),
        feature_dim=32, concat_features_across_scales=True, disable_viewing_dependent=True,
        proposal_net_args_list=[{"feature_dim": 8, "resolution": (128, 128, 128, 100)},
                                {"feature_dim": 8, "resolution": (256, 256, 256, 100)}],
        sigma_net_layers=1, sigma_net_hidden_dim=128, rgb_net_layers=2, rgb_net_hidden_dim=64,
        num_proposal_samples_per_ray=(256, 128), num_nerf_samples_per_ray=64, bounded=True,
        loss_coefficients=dict(PRESET_LOSS_COEFFICIENTS), depth_sigma=0.01, is_euclidean_depth=False,
    )


KPLANES_MODEL = kplanes_model_config()
KPLANES_DATAMANAGER = DynamicDataManagerConfig(  # method_configs.py:491-511
    train_num_rays_per_batch=4096, eval_num_rays_per_batch=512, use_importance_sampling=True, is_pixel_ratio=0.15, isg=False,
    ist_range=1.0, isg_gamma=5e-2, iters_to_start_is=2000,
)
KPLANES_OPTIMIZERS: Dict[str, Dict[str, float]] = {  # method_configs.py:546-557: Adam + cosine decay for both groups
    name: {"lr": 1e-2, "eps": 1e-12, "warm_up_end": 512, "max_steps": 30000, "learning_rate_alpha": 0.0}
    for name in ("proposal_networks", "fields")
}
KPLANES_TRAINER = {"method_name": "k-planes", "max_num_iterations": 30000, "steps_per_eval_image": 500,
                   "steps_per_eval_batch": 1000, "steps_per_save": 10000, "mixed_precision": True}


def make_method_specification(base_config=None):
    """MethodSpecification for the entry point.  ``base_config``: the reference's TrainerConfig to retarget (default:
    its own ``method_configs["k-planes"]``)."""
    from nerfstudio.plugins.types import MethodSpecification

    if base_config is None:
        from nerfstudio.configs.method_configs import method_configs

        base_config = method_configs["k-planes"]
    config = copy.deepcopy(base_config)
    config.pipeline.model._target = KPlanesModel  # pylint: disable=protected-access
    return MethodSpecification(
        config=config,
        description="K-Planes (multiscale hexplane field, proposal sampling) on hand-written sm_100a kernels (soccernerfs_b200).",
    )


kplanes_b200: Optional[object] = None
try:  # only when the reference package itself is installed next to this one
    import nerfstudio.plugins.types  # noqa: F401

    kplanes_b200 = make_method_specification()
except Exception:  # nerfstudio absent (GPU box, tests): the plain preset objects above are what is used
    kplanes_b200 = None
