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

    def __init__(self, params: Iterable[torch.nn.Parameter], arena: Optional["PeerArena"] = None) -> None:
        """``arena``: keep the bucket in NVLink peer memory (PeerArena) so that ``all_reduce`` runs our in-place
        peer-memory kernel instead of NCCL."""
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad and p.numel() > 0]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.arena, self.arena_offset = arena, 0
        if arena is not None:
            self.flat, self.arena_offset = arena.take(total)
        else:
            self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.params:
            # same dense layout as the param (as_strided offsets are absolute in the storage: the flat buffer may itself
            # be a slice of a peer-memory arena)
            v = torch.as_strided(self.flat, p.shape, p.stride(), storage_offset=self.flat.storage_offset() + off)
            self.views.append(v)
            off += p.numel()

    def enable_touched_marks(self, line_planes) -> bool:
        """Sparse gradient exchange (csrc/peer_allreduce.cu::peer_sharded_adam_sparse_kernel): one byte per 128-byte line of
        the bucket, kept in the peer arena next to it.  ``line_planes``: the parameters whose scatter kernel marks the lines
        it reduces into (channel-last planes with 32 features: one texel = one line); every other parameter of the bucket
        is marked "always dense".  Returns False (nothing changed) when the bucket's layout does not line up with 128-byte
        lines."""
        if self.arena is None or self.flat.numel() == 0:
            return False
        off, offsets = 0, []
        for p in self.params:
            offsets.append(off)
            off += p.numel()
        wanted = {id(p) for p in line_planes}
        for p, o in zip(self.params, offsets):
            if o % 32 or p.numel() % 32:
                return False
            if id(p) in wanted and not (p.dim() == 4 and p.shape[1] == 32 and p.stride(1) == 1):
                return False
        n_lines = (self.flat.numel() + 63) // 64 * 64 // 32
        raw, self.touched_offset = self.arena.take((n_lines + 3) // 4)
        self.touched = raw.view(torch.uint8)[:n_lines]
        self.touched.zero_()
        for p, o in zip(self.params, offsets):
            lines = self.touched[o // 32: (o + p.numel()) // 32]
            if id(p) in wanted:
                p._kp_touched = lines
            else:
                lines.fill_(2)
        self.param_offsets = dict((id(p), o) for p, o in zip(self.params, offsets))
        return True

    def _zero_spans(self, skip) -> List[Tuple[int, int]]:
        """Element ranges of the bucket NOT covered by the parameters in ``skip`` (ids), merged."""
        key = frozenset(skip)
        cache = self.__dict__.setdefault("_span_cache", {})
        if key not in cache:
            spans, off = [], 0
            for p in self.params:
                if id(p) not in key:
                    if spans and spans[-1][1] == off:
                        spans[-1] = (spans[-1][0], off + p.numel())
                    else:
                        spans.append((off, off + p.numel()))
                off += p.numel()
            cache[key] = spans
        return cache[key]

    def attach_zeroed(self, sink: bool = False, skip=None) -> None:
        """Zero the bucket and install the views as ``.grad`` so autograd accumulates in place.  ``sink=True``
        additionally publishes every view as the parameter's gradient sink (ops.grad_sink): the backward kernels
        then accumulate straight into the bucket.  ``skip``: ids of parameters whose gradient is about to be WRITTEN in
        full by a kernel (the fused regulariser sweep) -- their part of the bucket is not memset."""
        if skip:
            for a, b in self._zero_spans(skip):
                self.flat[a:b].zero_()
        else:
            self.flat.zero_()
        for p, v in zip(self.params, self.views):
            p.grad = v
            if sink:
                p._kp_grad_sink = v

    def detach_sinks(self) -> None:
        for p in self.params:
            if hasattr(p, "_kp_grad_sink"):
                del p._kp_grad_sink
            if hasattr(p, "_kp_touched"):
                del p._kp_touched

    def all_reduce(self, group=None, async_op: bool = False, span: Optional[Tuple[int, int]] = None):
        """Sum the bucket (or its [begin, end) element range ``span``) over the process group."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        begin, end = (0, self.flat.numel()) if span is None else span
        if end <= begin:
            return None
        if self.arena is not None:
            self.arena.all_reduce(self.arena_offset + begin, end - begin)
            return None
        return dist.all_reduce(self.flat[begin:end], op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    def span_of(self, params: Iterable[torch.nn.Parameter]) -> Optional[Tuple[int, int]]:
        """[begin, end) element range covered by ``params`` if they are adjacent in the bucket, else None."""
        wanted = {id(p) for p in params}
        begin, end, covered, off = None, None, 0, 0
        for p in self.params:
            if id(p) in wanted:
                begin = off if begin is None else begin
                end = off + p.numel()
                covered += p.numel()
            off += p.numel()
        if begin is None or covered != end - begin or len(wanted) != sum(1 for p in self.params if id(p) in wanted):
            return None
        return begin, end


class ShardedAdamGroup:
    """One parameter group under the sharded optimizer (csrc/peer_allreduce.cu::peer_sharded_adam_kernel): the group's
    parameters are re-homed into a flat region of the peer arena that mirrors its GradBucket (same order, same dense
    per-parameter layout), this rank keeps Adam's moments for its 1/world shard only, and ``step`` is ONE kernel that
    reduce-scatters the gradients, updates the shard and all-gathers the new parameters over NVLink."""

    def __init__(self, bucket: GradBucket, arena: "PeerArena") -> None:
        assert bucket.arena is arena
        self.bucket, self.arena = bucket, arena
        total = bucket.flat.numel()
        self.count = (total + 63) // 64 * 64  # the arena pads every region to 64 floats; the padding stays zero
        self.param_flat, self.param_offset = arena.take(total)
        off = 0
        with torch.no_grad():
            for p in bucket.params:
                v = torch.as_strided(self.param_flat, p.shape, p.stride(), storage_offset=self.param_flat.storage_offset() + off)
                v.copy_(p.data)
                p.data = v  # the model now reads / the kernel now writes the arena copy
                off += p.numel()
        n4 = self.count // 4
        self.lo4, self.hi4 = n4 * arena.rank // arena.world, n4 * (arena.rank + 1) // arena.world
        shard = max(4, (self.hi4 - self.lo4) * 4)
        dev = bucket.flat.device
        self.exp_avg = torch.zeros(shard, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(shard, dtype=torch.float32, device=dev)

    def step(self, lr: float, betas, eps: float, weight_decay: float, step: int, grad_scale: float,
             hyper_dev: Optional[torch.Tensor] = None, sparse: bool = False) -> None:
        """``sparse``: peers' gradient lines are read only where marked (``GradBucket.enable_touched_marks``); this rank's
        own shard must then hold the regularisers' gradient pre-multiplied by the world size (see TrainStep)."""
        from ctypes import c_void_p

        a = self.arena
        # a group of hundreds of MB needs more bytes in flight over NVLink than the 32 blocks that suit a 150 MB group
        # sharing the SMs with the proposal networks' backward
        blocks = a.blocks if self.count * 4 < (512 << 20) else max(a.blocks, 128)
        if sparse:
            a._lib.call("kp_peer_sharded_adam_sparse", a._ptrs, a.rank, a.world, int(self.bucket.arena_offset), int(self.param_offset),
                        int(self.bucket.touched_offset), int(self.count), c_void_p(self.exp_avg.data_ptr()),
                        c_void_p(self.exp_avg_sq.data_ptr()), float(lr), float(betas[0]), float(betas[1]), float(eps),
                        float(weight_decay), int(step), float(grad_scale), c_void_p(0 if hyper_dev is None else hyper_dev.data_ptr()),
                        blocks, c_void_p(torch.cuda.current_stream().cuda_stream))
            return
        a._lib.call("kp_peer_sharded_adam", a._ptrs, a.rank, a.world, int(self.bucket.arena_offset), int(self.param_offset),
                    int(self.count), c_void_p(self.exp_avg.data_ptr()), c_void_p(self.exp_avg_sq.data_ptr()), float(lr),
                    float(betas[0]), float(betas[1]), float(eps), float(weight_decay), int(step), float(grad_scale),
                    c_void_p(0 if hyper_dev is None else hyper_dev.data_ptr()), blocks,
                    c_void_p(torch.cuda.current_stream().cuda_stream))

    def shard_slice(self) -> Tuple[int, int]:
        """[begin, end) float range of the group this rank owns."""
        return self.lo4 * 4, self.hi4 * 4

    def gather_moments(self, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Full-length (padded) exp_avg / exp_avg_sq assembled from all ranks' shards (checkpointing; not on the step path)."""
        world = self.arena.world
        n4 = self.count // 4
        longest = max(n4 * (r + 1) // world - n4 * r // world for r in range(world)) * 4
        out = []
        for t in (self.exp_avg, self.exp_avg_sq):
            mine = torch.zeros(longest, dtype=torch.float32, device=t.device)
            mine[: (self.hi4 - self.lo4) * 4] = t[: (self.hi4 - self.lo4) * 4]
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine, group=group)
            full = torch.cat([parts[r][: (n4 * (r + 1) // world - n4 * r // world) * 4] for r in range(world)])
            out.append(full)
        return out[0], out[1]

    def scatter_moments(self, exp_avg_full: torch.Tensor, exp_avg_sq_full: torch.Tensor) -> None:
        b, e = self.shard_slice()
        self.exp_avg[: e - b].copy_(exp_avg_full[b:e])
        self.exp_avg_sq[: e - b].copy_(exp_avg_sq_full[b:e])


class PeerMemoryUnavailable(RuntimeError):
    """Raised (on every rank alike) when the peer-memory arenas cannot be set up on this node."""


class _RawCudaMemory:
    """``__cuda_array_interface__`` holder for memory the C library allocated (torch.as_tensor wraps it without a copy)."""

    def __init__(self, address: int, n_floats: int) -> None:
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (address, False), "version": 3,
                                         "strides": None}


class PeerArena:
    """One cudaMalloc'ed region per rank ([flag words | fp32 data]) that every other rank of the node maps through
    CUDA IPC, plus the in-place all-reduce kernel over it (csrc/peer_allreduce.cu).  Handles travel through the
    existing process group (``all_gather_object``); the data path itself is NVLink loads/stores issued by our kernel.
    Collective calls must be made by all ranks in the same order -- like any collective."""

    SIGNAL_BYTES = 16384  # KP_PEER_SIGNAL_BYTES

    def __init__(self, n_floats: int, group=None, blocks: int = 64) -> None:
        from ctypes import byref, c_void_p, create_string_buffer

        from . import _lib

        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerArena needs an initialised process group")
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise RuntimeError("PeerArena: one NVLink node, at most 8 ranks")
        self.blocks = int(blocks)
        self.n_floats = (int(n_floats) + 63) // 64 * 64
        self._lib = _lib
        # set-up failures (CUDA IPC not permitted in this container, no peer access, ...) are agreed on by ALL ranks
        # before anybody raises, so that the caller can fall back to NCCL consistently
        own, handle, err = c_void_p(), create_string_buffer(64), ""
        try:
            _lib.call("kp_peer_alloc", self.n_floats * 4, byref(own), handle)
        except RuntimeError as e:
            err = str(e)
        self._own = own.value or 0
        handles = [None] * self.world
        dist.all_gather_object(handles, (self.rank, err, handle.raw), group=group)
        self._ptrs = (c_void_p * self.world)()
        self._opened = []
        err = "; ".join(f"rank {r}: {e}" for r, e, _ in handles if e)
        if not err:
            for r, _e, raw in handles:
                if r == self.rank:
                    self._ptrs[r] = self._own
                    continue
                p = c_void_p()
                try:
                    _lib.call("kp_peer_open", create_string_buffer(raw, 64), byref(p))
                except RuntimeError as e:
                    err = str(e)
                    break
                self._ptrs[r] = p.value
                self._opened.append(p.value)
        errs = [None] * self.world
        dist.all_gather_object(errs, err, group=group)
        if any(errs):
            self.close()
            raise PeerMemoryUnavailable("; ".join(e for e in errs if e))
        self._holder = _RawCudaMemory(self._own + self.SIGNAL_BYTES, self.n_floats)
        self.data = torch.as_tensor(self._holder, device=torch.device("cuda", torch.cuda.current_device()))
        self._used = 0
        dist.barrier(group=group)  # every rank has mapped every arena before anybody's kernel touches one

    def take(self, n_floats: int) -> Tuple[torch.Tensor, int]:
        """Carve the next ``n_floats`` (rounded up to 64) out of the data region -> (tensor view, element offset)."""
        off = self._used
        if off + n_floats > self.n_floats:
            raise RuntimeError("PeerArena exhausted")
        self._used = off + (n_floats + 63) // 64 * 64
        return self.data[off: off + n_floats], off

    def all_reduce(self, begin: int, count: int) -> None:
        """Sum data[begin:begin+count] over all ranks, in place, on the current stream (one kernel launch)."""
        from ctypes import c_void_p

        self._lib.call("kp_peer_allreduce", self._ptrs, self.rank, self.world, int(begin), int(count), self.blocks,
                       c_void_p(torch.cuda.current_stream().cuda_stream))

    def error_word(self) -> int:
        """Non-zero if a cross-GPU barrier inside the kernel ever timed out (synchronises the device)."""
        from ctypes import byref, c_uint32, c_void_p

        torch.cuda.synchronize()
        w = c_uint32(0)
        self._lib.call("kp_peer_error", c_void_p(self._own), byref(w))
        return int(w.value)

    def close(self) -> None:
        from ctypes import c_void_p

        torch.cuda.synchronize()
        for p in self._opened:
            self._lib.call("kp_peer_close", c_void_p(p))
        self._opened = []
        if self._own:
            self.data = None
            self._lib.call("kp_peer_free", c_void_p(self._own))
            self._own = 0


def plane_write_ranges(plane_offsets, plane_sizes, lo: int, hi: int) -> List[Tuple[int, int]]:
    """Sparse gradient exchange: for planes occupying floats [off, off + size) of a flat bucket (``off`` None = the plane
    lives in another, densely exchanged bucket) the float4 range of each plane that lies inside this rank's shard
    [lo, hi) of the bucket -- what ``kp_plane_reg_fused_range`` may write.  Planes outside the shard get (0, 0), planes of
    other buckets their full range."""
    out = []
    for off, size in zip(plane_offsets, plane_sizes):
        if off is None:
            out.append((0, size // 4))
            continue
        a, b = max(lo, off) - off, min(hi, off + size) - off
        out.append((a // 4, b // 4) if b > a else (0, 0))
    return out


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
