"""Multi-GPU data parallelism for the K-Planes step.

Training: planes and MLPs are replicated, the ray batch is sharded (each rank owns a contiguous slice of the
global batch), every loss mean is a local mean, and ONE NCCL all-reduce (sum) over a flat gradient buffer per
step -- plane + MLP gradients together -- followed by a 1/world_size scale folded into the Adam kernel
(``grad_scale``) reproduces the global-batch gradient.  The reference instead wraps the model in DDP with every
rank drawing its own full batch (NS/pipelines/base_pipeline.py:244-246, scripts/train.py:84); with
``global_batch = world * 4096`` the two coincide.
Evaluation: image tiles / ray chunks are assigned round-robin to ranks, no collective on the data path.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


class GradBucket:
    """All gradients of a parameter list as views into one flat fp32 buffer (one memset, one all-reduce)."""

    def __init__(self, params: Iterable[torch.nn.Parameter]) -> None:
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad and p.numel() > 0]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.params:
            v = torch.as_strided(self.flat, p.shape, p.stride(), storage_offset=off)  # same dense layout as the param
            self.views.append(v)
            off += p.numel()

    def attach_zeroed(self, sink: bool = False) -> None:
        """Zero the bucket and install the views as ``.grad`` so autograd accumulates in place.  ``sink=True``
        additionally publishes every view as the parameter's gradient sink (ops.grad_sink): the backward kernels
        then accumulate straight into the bucket."""
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            p.grad = v
            if sink:
                p._kp_grad_sink = v

    def detach_sinks(self) -> None:
        for p in self.params:
            if hasattr(p, "_kp_grad_sink"):
                del p._kp_grad_sink

    def all_reduce(self, group=None, async_op: bool = False):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def shard_slice(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of ``n_items`` owned by ``rank`` (remainder spread over the first ranks)."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def round_robin_chunks(n_rays: int, chunk: int, rank: int, world: int) -> List[Tuple[int, int]]:
    """Eval sharding: ray chunks [i*chunk, (i+1)*chunk) with i % world == rank (no collective needed)."""
    out = []
    for i, start in enumerate(range(0, n_rays, chunk)):
        if i % world == rank:
            out.append((start, min(start + chunk, n_rays)))
    return out


def world_info() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


@torch.no_grad()
def render_frame_sharded(model, camera_ray_bundle, rank: int, world: int, chunk: Optional[int] = None):
    """Evaluation sharding (BASELINE config 5): this rank renders the ray chunks ``i % world == rank`` of a
    full [H,W] camera ray bundle -- the chunking of ``Model.get_outputs_for_camera_ray_bundle``
    (NS/models/base_model.py:162-186) distributed round-robin, with NO collective on the data path.
    Returns ``{(start, end): outputs_dict}`` for the owned chunks; ``assemble_frame`` stitches the pieces of all ranks
    (after e.g. ``dist.gather_object`` or peer copies to the writer rank)."""
    chunk = chunk or model.config.eval_num_rays_per_chunk
    flat = camera_ray_bundle.flatten()
    pieces = {}
    for start, end in round_robin_chunks(len(flat), chunk, rank, world):
        pieces[(start, end)] = {k: v for k, v in model.forward(ray_bundle=flat[start:end]).items() if torch.is_tensor(v)}
    return pieces


def assemble_frame(pieces_per_rank, image_height: int, image_width: int):
    """Concatenate the chunk outputs of all ranks in ray order -> {name: [H,W,C]}."""
    merged = {}
    for pieces in pieces_per_rank:
        merged.update(pieces)
    keys = sorted(merged)
    assert keys[0][0] == 0 and all(a[1] == b[0] for a, b in zip(keys, keys[1:])) and keys[-1][1] == image_height * image_width
    names = merged[keys[0]].keys()
    return {n: torch.cat([merged[k][n] for k in keys]).view(image_height, image_width, -1) for n in names}
