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


def load_reference_state_dict(model, state: Dict[str, torch.Tensor], prefix: str = "_model.") -> List[str]:
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
                mlp = mlps[name[: -len(".params")]]
                for dst, src in zip(mlp.weights, tcnn_params_to_weights(value.reshape(-1), _dims(mlp))):
                    dst.copy_(src.to(dst.device))
            elif name in own:
                if own[name].shape != value.shape:
                    raise ValueError(f"{key}: shape {tuple(value.shape)} does not match {tuple(own[name].shape)}")
                own[name].copy_(value.to(own[name].device))  # NCHW source -> channel-last parameter: strided copy
            else:
                unused.append(key)
    return unused


def to_reference_state_dict(model, prefix: str = "_model.") -> Dict[str, torch.Tensor]:
    """The model's tensors under the reference's names and layouts (planes NCHW-contiguous, MLPs as flat tcnn params)."""
    out: Dict[str, torch.Tensor] = {}
    mlp_params = {f"{k}.weights.{i}" for k, m in _mlps(model).items() for i in range(len(m.weights))}
    for name, p in model.named_parameters():
        if name in mlp_params:
            continue
        out[prefix + name] = p.detach().contiguous().clone()
    for k, m in _mlps(model).items():
        out[prefix + k + ".params"] = weights_to_tcnn_params(list(m.weights))
    return out
