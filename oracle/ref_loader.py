"""TEST INFRASTRUCTURE ONLY -- imports the *real* reference hot path from /root/reference.

This module is used in the build container (where /root/reference exists) by
``oracle/make_golden.py`` to generate the committed fixtures under ``tests/golden/`` and by
``tests/test_oracle_vs_reference.py`` to validate ``oracle/kplanes_oracle.py`` against the
reference's own code.  It never runs on the GPU box (the reference is not there) and nothing in
``soccernerfs_b200/`` may import it.

The reference (nerfstudio 0.1.19 fork) cannot be imported as-is on Python 3.12 / CPU: five
third-party modules are missing and ``nerfstudio/configs/base_config.py:118`` uses a mutable
dataclass default.  We pre-register stub modules for exactly those, as described in SURVEY.md
section 8(c):

* ``torchtyping``                -- annotations only.
* ``nerfacc``                    -- imported by ray_samplers.py:22-24 / renderers.py:33, never executed
                                    on the K-Planes path (ray_indices is always None).
* ``matplotlib``                 -- colour maps, not on the path.
* ``nerfstudio.configs(.base_config)`` -- PrintableConfig / InstantiateConfig only.
* ``tinycudann``                 -- CUDA-only, un-vendored third party (Dockerfile:121 pins v1.6).
  Its published semantics are restated as: ``Network`` = bias-free dense stack with the requested
  hidden width / hidden-layer count / activations, fp32 here (the fp16 of FullyFusedMLP is *not*
  the parity target); ``Encoding`` = degree-4 spherical harmonics of ``2x-1``
  evaluated with the reference's own ``nerfstudio/utils/math.py:25-86``.

It also injects ``kplanes_field.Frustums`` (the reference forgets that import,
``fields/kplanes_field.py:421`` vs ``:28``).
"""
from __future__ import annotations

import os
import sys
import types

import torch
from torch import nn

REF_ROOT = os.environ.get("KPLANES_REFERENCE_ROOT", "/root/reference")
REF_PKG = os.path.join(REF_ROOT, "nerfstudio")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_PKG, "nerfstudio", "fields"))


class _TensorTypeMeta(type):
    def __getitem__(cls, item):
        return cls


class _TensorType(metaclass=_TensorTypeMeta):
    pass


class _StubNetwork(nn.Module):
    """Bias-free fp32 stand-in for ``tcnn.Network`` (FullyFusedMLP / CutlassMLP)."""

    def __init__(self, n_input_dims, n_output_dims, network_config, seed=None):
        super().__init__()
        width = int(network_config["n_neurons"])
        hidden = int(network_config["n_hidden_layers"])
        dims = [n_input_dims] + [width] * hidden + [n_output_dims]
        self.layers = nn.ModuleList([nn.Linear(a, b, bias=False) for a, b in zip(dims[:-1], dims[1:])])
        for lin in self.layers:
            nn.init.xavier_uniform_(lin.weight)
        self.activation = network_config.get("activation", "ReLU")
        self.output_activation = network_config.get("output_activation", "None")
        self.n_input_dims = n_input_dims
        self.n_output_dims = n_output_dims

    @staticmethod
    def _act(name, x):
        if name == "ReLU":
            return torch.relu(x)
        if name == "Sigmoid":
            return torch.sigmoid(x)
        if name == "None":
            return x
        raise NotImplementedError(name)

    def forward(self, x):
        for i, lin in enumerate(self.layers):
            x = lin(x)
            x = self._act(self.activation if i + 1 < len(self.layers) else self.output_activation, x)
        return x


class _StubEncoding(nn.Module):
    """``tcnn.Encoding`` SphericalHarmonics degree 4: inputs in [0,1] are mapped to 2x-1."""

    def __init__(self, n_input_dims, encoding_config):
        super().__init__()
        assert encoding_config["otype"] == "SphericalHarmonics"
        self.degree = int(encoding_config["degree"])
        self.n_input_dims = n_input_dims
        self.n_output_dims = self.degree**2

    def forward(self, x):
        from nerfstudio.utils.math import components_from_spherical_harmonics

        return components_from_spherical_harmonics(self.degree, x * 2.0 - 1.0)


def _install_stubs() -> None:
    if "torchtyping" not in sys.modules:
        m = types.ModuleType("torchtyping")
        m.TensorType = _TensorType
        m.patch_typeguard = lambda: None
        sys.modules["torchtyping"] = m
    if "nerfacc" not in sys.modules:
        m = types.ModuleType("nerfacc")

        class OccupancyGrid:  # noqa: D401 - import-only placeholder
            pass

        class ContractionType:
            AABB = 0
            UN_BOUNDED_SPHERE = 2

        m.OccupancyGrid = OccupancyGrid
        m.ContractionType = ContractionType
        sys.modules["nerfacc"] = m
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if "tinycudann" not in sys.modules:
        m = types.ModuleType("tinycudann")
        m.Network = _StubNetwork
        m.Encoding = _StubEncoding
        sys.modules["tinycudann"] = m

    # nerfstudio.configs.base_config is unimportable on py>=3.11; only two tiny classes are needed.
    import importlib.machinery
    from dataclasses import dataclass
    from typing import Any, Type

    if "nerfstudio" not in sys.modules:
        pkg = types.ModuleType("nerfstudio")
        pkg.__path__ = [os.path.join(REF_PKG, "nerfstudio")]
        pkg.__spec__ = importlib.machinery.ModuleSpec("nerfstudio", None, is_package=True)
        sys.modules["nerfstudio"] = pkg
    if "nerfstudio.configs" not in sys.modules:
        cfg = types.ModuleType("nerfstudio.configs")
        cfg.__path__ = [os.path.join(REF_PKG, "nerfstudio", "configs")]  # other submodules (config_utils) are the real ones
        sys.modules["nerfstudio.configs"] = cfg
        base = types.ModuleType("nerfstudio.configs.base_config")

        class PrintableConfig:
            pass

        @dataclass
        class InstantiateConfig(PrintableConfig):
            _target: Type

            def setup(self, **kwargs) -> Any:
                return self._target(self, **kwargs)

        base.PrintableConfig = PrintableConfig
        base.InstantiateConfig = InstantiateConfig
        sys.modules["nerfstudio.configs.base_config"] = base
        cfg.base_config = base


_REF = None


def load_reference():
    """Return a namespace with the reference's hot-path symbols (imported from /root/reference)."""
    global _REF
    if _REF is not None:
        return _REF
    if not reference_available():
        raise RuntimeError(f"reference checkout not found under {REF_ROOT}")
    _install_stubs()

    from nerfstudio.cameras import rays
    from nerfstudio.data.scene_box import SceneBox
    from nerfstudio.field_components import activations
    from nerfstudio.fields import kplanes_field
    from nerfstudio.model_components import losses, ray_samplers, renderers, scene_colliders
    from nerfstudio.utils import math as ns_math

    kplanes_field.Frustums = rays.Frustums  # missing import in the reference (SURVEY finding 2)

    ns = types.SimpleNamespace(
        rays=rays,
        SceneBox=SceneBox,
        activations=activations,
        kplanes_field=kplanes_field,
        losses=losses,
        ray_samplers=ray_samplers,
        renderers=renderers,
        scene_colliders=scene_colliders,
        math=ns_math,
        StubNetwork=_StubNetwork,
    )
    _REF = ns
    return ns


def load_reference_model_module():
    """``nerfstudio.models.kplanes`` of the reference (KPlanesModelConfig / KPlanesModel), for the drop-in boundary test.
    On top of ``load_reference``'s stubs it needs import-only stand-ins for the evaluation metrics the module pulls in at
    import time (torchmetrics PSNR / SSIM / LPIPS and the RetinaNet-based DynMetric, which downloads weights)."""
    ns = load_reference()

    class _Metric(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    def _stub(name, **attrs):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules[name] = m

    _stub("torchmetrics", PeakSignalNoiseRatio=_Metric)
    _stub("torchmetrics.functional", structural_similarity_index_measure=lambda *a, **k: None)
    _stub("torchmetrics.image")
    _stub("torchmetrics.image.lpip", LearnedPerceptualImagePatchSimilarity=_Metric)
    _stub("nerfstudio.utils.dynmetric", DynMetric=_Metric)
    from nerfstudio.models import kplanes as ref_kplanes

    ns.models_kplanes = ref_kplanes
    return ns
