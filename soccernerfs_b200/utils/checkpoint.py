"""Checkpoint compatibility with the reference.

The reference saves ``{"step", "pipeline": pipeline.state_dict(), "optimizers", "scalers"}``
(NS/engine/trainer.py:352-380); the model's tensors sit under the ``_model.`` prefix:

* ``_model.field.grids.<k>.<p>`` / ``_model.proposal_networks.<i>.grids.<p>``: planes ``[1,C,H,W]`` (NCHW,
  ``REF/scripts/plot_kplane.py:34-47``).  Ours have the same names and logical shapes and are stored channel-last, so
  they load with a plain strided copy.
* ``_model.field.sigma_net.params`` / ``color_net.params`` / ``proposal_networks.<i>.sigma_net.params``: tiny-cuda-nn
  keeps ALL weight matrices of a ``FullyFusedMLP`` in one flat fp32 ``params`` tensor: per layer a ROW-MAJOR
  ``[out, in]`` matrix, the first layer's ``in`` padded up to a multiple of 16 and the last layer's ``out`` padded up to
  a multiple of 16, matrices back to back in layer order (tiny-cuda-nn v1.6 ``fully_fused_mlp.cu``: ``m_weight_matrices``
  are ``GPUMatrix<T, RM>`` carved out of ``params`` in order).  tinycudann is not installed here, so this layout is
  restated from its published source and is NOT pinned by a fixture ("parity unpinned", like the decoder arithmetic
  itself); the conversion is exercised by a round-trip test only.
* ``optimizers``: ``torch.optim.Adam.state_dict()`` per group, state numbered by the reference's parameter order;
  ``load_reference_optimizer_state`` / ``to_reference_optimizer_state`` convert the moments both ways.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch

_PAD = 16


def _pad16(n: int) -> int:
    return (n + _PAD - 1) // _PAD * _PAD


def tcnn_layer_shapes(dims: Sequence[int]) -> List[Tuple[int, int]]:
    """Padded (out, in) of every weight matrix of a FullyFusedMLP with layer widths ``dims`` = [in, hidden..., out]."""
    shapes = []
    for i in range(len(dims) - 1):
        fan_in = _pad16(dims[i]) if i == 0 else dims[i]
        fan_out = _pad16(dims[i + 1]) if i == len(dims) - 2 else dims[i + 1]
        shapes.append((fan_out, fan_in))
    return shapes


def tcnn_params_to_weights(flat: torch.Tensor, dims: Sequence[int]) -> List[torch.Tensor]:
    """Flat tcnn ``params`` -> our unpadded ``[out_i, in_i]`` weight matrices (padding rows / columns dropped)."""
    shapes = tcnn_layer_shapes(dims)
    need = sum(o * i for o, i in shapes)
    if flat.numel() != need:
        raise ValueError(f"tcnn params have {flat.numel()} elements, layer widths {list(dims)} need {need}")
    out, off = [], 0
    for k, (o, i) in enumerate(shapes):
        w = flat[off: off + o * i].view(o, i)
        out.append(w[: dims[k + 1], : dims[k]].float().clone())
        off += o * i
    return out


def weights_to_tcnn_params(weights: Sequence[torch.Tensor]) -> torch.Tensor:
    """Inverse of ``tcnn_params_to_weights`` (padding filled with zeros)."""
    dims = [weights[0].shape[1]] + [w.shape[0] for w in weights]
    chunks = []
    for w, (o, i) in zip(weights, tcnn_layer_shapes(dims)):
        m = torch.zeros(o, i, dtype=torch.float32, device=w.device)
        m[: w.shape[0], : w.shape[1]] = w.detach().float()
        chunks.append(m.reshape(-1))
    return torch.cat(chunks)


def _mlps(model) -> Dict[str, torch.nn.Module]:
    out = {"field.sigma_net": model.field.sigma_net, "field.color_net": model.field.color_net}
    for i, p in enumerate(model.proposal_networks):
        out[f"proposal_networks.{i}.sigma_net"] = p.sigma_net
    return out


def _dims(mlp) -> List[int]:
    return [mlp.weights[0].shape[1]] + [w.shape[0] for w in mlp.weights]


# tiny-cuda-nn's SphericalHarmonics (degree 4) differs from NS/utils/math.py:25-86 -- the basis this package and the
# oracle use -- by the sign of every odd-index component (tcnn: -y, -x, -yz, -xz, y(-3x2+y2), y(1-5z2), x(1-5z2),
# x(-x2+3y2)).  A colour net trained behind tcnn's encoding therefore has the negated first-layer columns 1,3,...,15.
# Restated from tiny-cuda-nn's published source (include/tiny-cuda-nn/common_device.h, v1.6); tcnn is absent here, so
# like the flat layout this is unpinned, and it can be switched off.
TCNN_SH_SIGNS = torch.tensor([1.0, -1.0] * 8)


def _sh_fix(mlp_name: str, model, weights: List[torch.Tensor], tcnn_sh_convention: bool) -> List[torch.Tensor]:
    if not (tcnn_sh_convention and mlp_name == "field.color_net" and not model.field.disable_viewing_dependent):
        return weights
    w0 = weights[0].clone()
    w0[:, :16] = w0[:, :16] * TCNN_SH_SIGNS.to(w0.device)
    return [w0] + list(weights[1:])


def _repack_planes(dst: torch.Tensor, src: torch.Tensor) -> None:
    """NCHW-contiguous [1,C,H,W] <-> channel-last parameter; on the GPU through kp_repack_* (one coalesced transpose
    kernel per plane -- a 32x checkpoint is 2.3 GB of planes), on the CPU a strided copy."""
    from .. import ops

    if dst.is_cuda and src.is_cuda and dst.dim() == 4 and dst.shape[0] == 1 and dst.dtype == src.dtype == torch.float32:
        if ops.is_channel_last(dst) and src.is_contiguous():
            ops.repack(src, dst, to_channel_last=True)
            return
        if dst.is_contiguous() and ops.is_channel_last(src):
            ops.repack(src, dst, to_channel_last=False)
            return
    dst.copy_(src.to(dst.device))


def load_reference_state_dict(model, state: Dict[str, torch.Tensor], prefix: str = "_model.",
                              tcnn_sh_convention: bool = True) -> List[str]:
    """Copy a reference pipeline ``state_dict`` (or a checkpoint's ``["pipeline"]``) into ``model``.
    Returns the keys that were not consumed (datamanager / camera-optimizer state, empty tcnn encodings...)."""
    if "pipeline" in state and isinstance(state["pipeline"], dict):
        state = state["pipeline"]
    own = dict(model.named_parameters())
    mlps = _mlps(model)
    unused = []
    with torch.no_grad():
        for key, value in state.items():
            name = key[len(prefix):] if key.startswith(prefix) else key
            name = name.replace("module.", "")
            if name.endswith(".params") and name[: -len(".params")] in mlps:
                mlp_name = name[: -len(".params")]
                mlp = mlps[mlp_name]
                ws = _sh_fix(mlp_name, model, tcnn_params_to_weights(value.reshape(-1), _dims(mlp)), tcnn_sh_convention)
                for dst, src in zip(mlp.weights, ws):
                    dst.copy_(src.to(dst.device))
            elif name in own:
                if own[name].shape != value.shape:
                    raise ValueError(f"{key}: shape {tuple(value.shape)} does not match {tuple(own[name].shape)}")
                _repack_planes(own[name].data, value.to(own[name].device))  # NCHW source -> channel-last parameter
            else:
                unused.append(key)
    return unused


def to_reference_state_dict(model, prefix: str = "_model.", tcnn_sh_convention: bool = True) -> Dict[str, torch.Tensor]:
    """The model's tensors under the reference's names and layouts (planes NCHW-contiguous, MLPs as flat tcnn params,
    the parameter-less tcnn SH encoding as an empty ``direction_encoder.params``).  The reference's strict
    ``load_pipeline`` additionally expects keys this path does not own (lpips weights, datamanager / camera optimizer):
    load the result with ``strict=False`` or merge it into a reference checkpoint's ``pipeline`` dict."""
    out: Dict[str, torch.Tensor] = {}
    mlp_params = {f"{k}.weights.{i}" for k, m in _mlps(model).items() for i in range(len(m.weights))}
    for name, p in model.named_parameters():
        if name in mlp_params or p.numel() == 0:
            continue
        if p.dim() == 4:
            dst = torch.empty(p.shape, dtype=torch.float32, device=p.device)
            _repack_planes(dst, p.detach())
            out[prefix + name] = dst
        else:
            out[prefix + name] = p.detach().contiguous().clone()
    if not model.field.disable_viewing_dependent:
        out[prefix + "field.direction_encoder.params"] = torch.zeros(0, dtype=torch.float32)
    for k, m in _mlps(model).items():
        ws = _sh_fix(k, model, [w.detach() for w in m.weights], tcnn_sh_convention)
        out[prefix + k + ".params"] = weights_to_tcnn_params(ws)
    return out


# ---- optimizer state (NS/engine/trainer.py:362-373: {"optimizers": {group: torch.optim.Adam.state_dict()}}) ----------
def _reference_group_layout(model, group: str):
    """The reference's parameter order inside an optimizer group, as (kind, object) entries:
    ``Model.get_param_groups`` lists ``module.parameters()`` (NS/models/kplanes.py:311-316), i.e. registration order --
    aabb, the planes, [tcnn SH encoding's empty params], sigma_net.params, color_net.params (kplanes_field.py:170-273;
    proposal networks: aabb, planes, sigma_net.params, :386-407).  torch.optim.Adam numbers its state by that order."""
    def field_entries(f, name, has_color):
        ent = [("tensor", f.aabb)] + [("tensor", p) for g in (f.grids if has_color else [f.grids]) for p in g]
        if has_color and not f.disable_viewing_dependent:
            ent.append(("empty", None))
        ent.append(("tcnn", (name + ".sigma_net", f.sigma_net)))
        if has_color:
            ent.append(("tcnn", (name + ".color_net", f.color_net)))
        return ent

    if group == "fields":
        return field_entries(model.field, "field", True)
    if group == "proposal_networks":
        out = []
        for i, net in enumerate(model.proposal_networks):
            out += field_entries(net, f"proposal_networks.{i}", False)
        return out
    raise KeyError(group)


def load_reference_optimizer_state(optimizers, model, ref_state: Dict[str, Dict], tcnn_sh_convention: bool = True) -> None:
    """Copy the Adam moments of a reference checkpoint's ``["optimizers"]`` into ``engine.optimizers.Optimizers``:
    plane moments by strided copy, the flat tcnn moment vectors split like the weights (``exp_avg`` takes the SH sign
    flips of the weights; ``exp_avg_sq`` is sign-free)."""
    for group, opt in optimizers.optimizers.items():
        if group not in ref_state:
            continue
        ref = ref_state[group]["state"]
        for idx, (kind, obj) in enumerate(_reference_group_layout(model, group)):
            st = ref.get(idx, ref.get(str(idx)))
            if st is None or kind == "empty":
                continue
            step = int(st["step"]) if not torch.is_tensor(st["step"]) else int(st["step"].item())
            if kind == "tensor":
                targets = [(obj, st["exp_avg"], st["exp_avg_sq"])]
            else:
                name, mlp = obj
                dims = _dims(mlp)
                m = _sh_fix(name, model, tcnn_params_to_weights(st["exp_avg"].reshape(-1), dims), tcnn_sh_convention)
                v = tcnn_params_to_weights(st["exp_avg_sq"].reshape(-1), dims)
                targets = list(zip(mlp.weights, m, v))
            for p, m_src, v_src in targets:
                if not p.requires_grad:
                    continue
                own = opt.state[p]
                own["step"] = step
                for key, src in (("exp_avg", m_src), ("exp_avg_sq", v_src)):
                    if key not in own:
                        own[key] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    own[key].copy_(src.to(p.device).reshape(p.shape))


def to_reference_optimizer_state(optimizers, model, tcnn_sh_convention: bool = True) -> Dict[str, Dict]:
    """Inverse of ``load_reference_optimizer_state``: ``{group: {"state": {index: {...}}, "param_groups": [...]}}`` in the
    reference's parameter numbering and layouts."""
    out: Dict[str, Dict] = {}
    for group, opt in optimizers.optimizers.items():
        layout = _reference_group_layout(model, group)
        state: Dict[int, Dict] = {}
        for idx, (kind, obj) in enumerate(layout):
            if kind == "empty":
                continue
            if kind == "tensor":
                st = opt.state.get(obj)
                if st:
                    state[idx] = {"step": torch.tensor(float(st["step"])), "exp_avg": st["exp_avg"].detach().contiguous().clone(),
                                  "exp_avg_sq": st["exp_avg_sq"].detach().contiguous().clone()}
                continue
            name, mlp = obj
            sts = [opt.state.get(w) for w in mlp.weights]
            if not all(sts):
                continue
            m = _sh_fix(name, model, [s_["exp_avg"].detach() for s_ in sts], tcnn_sh_convention)
            state[idx] = {"step": torch.tensor(float(sts[0]["step"])), "exp_avg": weights_to_tcnn_params(m),
                          "exp_avg_sq": weights_to_tcnn_params([s_["exp_avg_sq"].detach() for s_ in sts])}
        pg = {k: v for k, v in opt.param_groups[0].items() if k not in ("params", "hyper_dev")}
        pg["params"] = list(range(len(layout)))
        out[group] = {"state": state, "param_groups": [pg]}
    return out
