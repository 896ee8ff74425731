"""One K-Planes training iteration, mirroring ``Trainer.train_iteration`` (NS/engine/trainer.py:382-412) and the
callback calls around it (:212, :221): callbacks BEFORE -> zero_grad -> forward (collider + get_outputs) ->
metrics -> loss dict -> backward -> [gradient all-reduce] -> Adam -> scheduler -> callbacks AFTER.

``use_cuda_graph=True`` captures that whole iteration (≈100 kernel launches) in a CUDA graph per sampler mode
("proposal networks updated" / "not updated") and replays it: the launch-bound Python/autograd dispatch then
costs one ``cudaGraphLaunch``.  Per-step scalars that the eager loop keeps on the host (Adam bias corrections,
cosine LR, proposal-weight anneal exponent) are looked up on the device from tables indexed by a device-resident
step counter, so a replay needs nothing from the host but the batch.
"""
from __future__ import annotations

import contextlib
import os
import sys
from typing import Dict, Optional

import numpy as np
import torch

from ..cameras.rays import RayBundle
from ..distributed import GradBucket, PeerArena, PeerMemoryUnavailable, ShardedAdamGroup, plane_write_ranges, world_info
from ..models.kplanes import KPlanesModel, TrainingCallbackLocation, scale_dict
from .optimizers import Optimizers, cosine_decay_factor


class TrainStep:
    REG_ADAM_AUTO_BYTES = 512 << 20  # fuse_reg_adam=None: planes at least this large use the fused regulariser + Adam pass

    def __init__(self, model: KPlanesModel, max_steps: int = 30000, lr: float = 1e-2, eps: float = 1e-12,
                 warm_up_end: int = 512, data_parallel: bool = False, use_cuda_graph: bool = False,
                 fuse_grad_accumulation: bool = True, overlap_branches: bool = True,
                 overlap_proposal_backward: Optional[bool] = None, allreduce_mode: str = "overlap",
                 allreduce_backend: str = "peer", fuse_regularizers: bool = True,
                 shard_optimizer: Optional[bool] = None, fuse_reg_adam: Optional[bool] = None,
                 sparse_grad_exchange: Optional[bool] = None, prioritize_main_stream: bool = True,
                 branch_small_kernels: bool = True) -> None:
        """``overlap_branches``: run the two branches of the step that do not depend on the main field's backward on
        their own CUDA streams (same arithmetic, same results): the plane regularisers (forward AND backward depend on
        the planes only) during the forward pass, and the back-propagation through the proposal networks (depends on
        the interlevel loss only) next to the decoder/scatter backward.  Needs the gradient sinks.
        ``overlap_proposal_backward``: None = automatic -- on for a single process; off under data parallelism, where
        the proposal backward is instead kept AFTER the field scatter so that it hides the field bucket's all-reduce.
        ``prioritize_main_stream``: capture the graphed step on a high-priority stream, so that the side branches (default
        priority) only get what the critical path leaves free.  ``branch_small_kernels``: the outputs no loss reads
        (accumulation, depths, median colour) and the distortion / interlevel terms run as parallel branches of the step
        (forward and backward) instead of one after the other in front of the loss head.
        ``allreduce_mode``: "overlap" (field bucket reduced on a communication stream as soon as the scatter is enqueued),
        "overlap-per-scale" (scatter launched per scale, finest first, and the finest scale reduced under the others),
        "after-backward" (both buckets reduced on the main stream after the whole backward).
        ``allreduce_backend``: "peer" -- gradient buckets live in NVLink peer memory and are summed by our in-place
        kernel (csrc/peer_allreduce.cu); "nccl" -- ``torch.distributed.all_reduce``.  If the peer arenas cannot be set
        up on this node (all ranks agree), a warning is printed and NCCL is used; ``self.allreduce_backend`` says which.
        ``shard_optimizer`` (None = on whenever the peer backend is active): ZeRO-1-style step fused with its
        collectives -- parameters also live in the peer arenas, every rank keeps Adam's moments for 1/world of each
        group only, and ONE kernel per group reduce-scatters the gradients, updates the owned shard and all-gathers the
        new parameters (``distributed.ShardedAdamGroup``).  Same NVLink bytes as the all-reduce; Adam's HBM traffic and
        optimizer memory drop by 1/world and the reduced gradient is never materialised.
        ``fuse_reg_adam`` (None = on when applicable -- gradient sinks, all six regularisers, CUDA, no gradient collective
        -- and the planes are HBM-resident, >= REG_ADAM_AUTO_BYTES): the plane regularisers are not evaluated as a branch of the step at all -- their gradient is computed
        from the pre-update planes INSIDE the optimizer's streaming pass (``kp_plane_reg_adam``),
        which also leaves the planes' gradient buffers zeroed for the next step; the six loss values come out of the same
        pass and are added to the reported loss after it.
        ``sparse_grad_exchange`` (sharded optimizer only; None = on when the field planes are HBM-resident, >=
        REG_ADAM_AUTO_BYTES): the field's scatter marks the 128-byte gradient lines it reduces into and the fused
        reduce-scatter reads a peer's line only if marked (one step's rays touch 15-30 % of the lines of the 16x / 32x
        scales); the regularisers' dense gradient is written by the shard's owner alone, pre-multiplied by the world size,
        and every rank clears its marked lines after the exchange, so the 2.3 GB bucket is neither memset nor pulled over
        NVLink in full.  Same sums in a different association order (atomics already make the order free)."""
        self.model = model
        self.max_steps, self.base_lr, self.warm_up_end = max_steps, lr, warm_up_end
        self.optimizers = Optimizers(model.get_param_groups(), lr=lr, eps=eps, warm_up_end=warm_up_end, max_steps=max_steps)
        self.callbacks = model.get_training_callbacks(None)
        self.step = 0
        self.rank, self.world = world_info()
        self.buckets: Dict[str, GradBucket] = {}
        self.arena, self.allreduce_backend = None, "nccl"
        self.reduce_grads = data_parallel and self.world > 1
        self.fuse_grad_accumulation = fuse_grad_accumulation
        if self.reduce_grads:
            # replicas start from rank 0's parameters, as DistributedDataParallel does for the reference
            # (NS/pipelines/base_pipeline.py:244-246)
            import torch.distributed as dist

            with torch.no_grad():
                for p in model.parameters():
                    if p.numel():
                        dist.broadcast(p.data, 0)
        if self.reduce_grads or fuse_grad_accumulation:
            # one flat gradient bucket per parameter group: the field bucket is complete as soon as the field's scatter
            # kernel has run, so its all-reduce overlaps the back-propagation through the proposal networks
            on_cuda = next(model.parameters()).is_cuda
            arena = None
            assert allreduce_backend in ("peer", "nccl")
            if self.reduce_grads and on_cuda and allreduce_backend == "peer":
                groups = model.get_param_groups()
                n = sum((sum(p.numel() for p in ps if p.requires_grad) + 63) // 64 * 64 for ps in groups.values())
                if shard_optimizer is None or shard_optimizer:
                    n *= 2  # the parameters get a region of their own next to the gradients
                    n += n // 64 + 1024  # + one "touched" byte per 128-byte gradient line (sparse exchange)
                try:
                    arena = PeerArena(n, blocks=int(os.environ.get("KP_PEER_BLOCKS", "32")))
                except PeerMemoryUnavailable as e:
                    if self.rank == 0:
                        print(f"[soccernerfs_b200] peer-memory all-reduce unavailable ({e}); using NCCL", file=sys.stderr)
            self.arena = arena
            self.allreduce_backend = "peer" if arena is not None else "nccl"
            self.buckets = {name: GradBucket(ps, arena=arena) for name, ps in model.get_param_groups().items()}
        self.sharded: Dict[str, ShardedAdamGroup] = {}
        self._last_sharded_lr: Dict[str, float] = {}
        if self.arena is not None and (shard_optimizer is None or shard_optimizer):
            self.sharded = {name: ShardedAdamGroup(b, self.arena) for name, b in self.buckets.items()}
        elif shard_optimizer:
            raise RuntimeError("shard_optimizer=True needs the peer-memory backend (data_parallel, world > 1, CUDA IPC)")
        on_cuda = next(model.parameters()).is_cuda
        self._sparse, self._bucket_dirty = False, True
        self._sparse_wanted = sparse_grad_exchange
        self.overlap = bool(overlap_branches and fuse_grad_accumulation and on_cuda)
        # fused regulariser sweep (values + gradient written into the sinks, no memset of the planes): ids of the planes
        # it writes in full, or None when not applicable
        self._reg_written = None
        if fuse_grad_accumulation and self.buckets and on_cuda and model.fused_regularizers_applicable() and fuse_regularizers:
            self._reg_written = frozenset(id(p) for p in model.regularized_planes() if p.requires_grad)
        # sparse gradient exchange of the field group (needs the fused regulariser sweep: it is what keeps the rest of the
        # bucket free of dense contributions)
        if "fields" in self.sharded and self._reg_written is not None and sparse_grad_exchange is not False:
            grids = getattr(model.field, "grids", None)
            planes = [p for g in grids for p in g] if grids is not None else []
            big = sum(p.numel() for p in planes) * 4 >= self.REG_ADAM_AUTO_BYTES
            if planes and (sparse_grad_exchange or big) and self.buckets["fields"].enable_touched_marks(planes):
                self._sparse = True
                self._setup_sparse_regularizers()
        if sparse_grad_exchange and not self._sparse:
            raise RuntimeError("sparse_grad_exchange=True needs the sharded optimizer, the fused regularisers and 32-feature planes "
                               "laid out on 128-byte lines")
        # (f1) regulariser stencil folded into the optimizer pass
        self._reg_adam = None
        if (fuse_reg_adam is None or fuse_reg_adam) and self._reg_written is not None and not self.reduce_grads:
            from .optimizers import PlaneRegAdamPlan

            plan = PlaneRegAdamPlan(model)
            # automatic: only where the planes are HBM-resident (the sweep costs 2 of ~9 passes over them).  While they
            # mostly fit the 126 MB L2 the regulariser sweep hides on its side stream and folding it into the optimizer
            # pass would put it ON the critical path.
            plane_bytes = sum(p.numel() for p in plan.planes) * 4
            if plan.supported and (fuse_reg_adam or plane_bytes >= self.REG_ADAM_AUTO_BYTES):
                self._reg_adam = plan
                self._reg_adam_first = True
        if fuse_reg_adam and self._reg_adam is None:
            raise RuntimeError("fuse_reg_adam=True needs CUDA, the gradient sinks, the fused regularisers (dynamic scene, all six "
                               "regulariser coefficients), plane feature dims in {4,8,16,32} and no gradient collective")
        self._field_ready = None
        self._scale_ready: Dict[int, "torch.cuda.Event"] = {}
        self._comm_stream = None
        self._first_span = None  # bucket range of the finest scale's planes: all-reduced first, under the other scatters
        assert allreduce_mode in ("overlap", "overlap-per-scale", "after-backward")
        if self.reduce_grads:
            # high priority: when the scatter finishes, the all-reduce kernel's blocks are placed BEFORE the proposal
            # backward's persistent blocks fill every SM (they would otherwise keep it out until they retire)
            self._comm_stream = torch.cuda.Stream(priority=-2) if on_cuda else None
        # The overlap hook marks "every gradient of the field bucket is in the stream" right after the scatter kernel was
        # enqueued.  That is only true with the gradient sinks (the kernels write the bucket themselves); without them
        # autograd's AccumulateGrad adds the returned gradients AFTER the hook, so the bucket is reduced after the backward.
        if self.reduce_grads and allreduce_mode != "after-backward" and fuse_grad_accumulation:
            hook = self._start_field_allreduce
            grids = getattr(model.field, "grids", None)
            if allreduce_mode == "overlap-per-scale" and self.overlap and grids is not None and len(grids) > 1:
                self._first_span = self.buckets["fields"].span_of(list(grids[len(grids) - 1]))
            if self._first_span is not None:
                hook = self._scale_scattered
            model.field._kp_post_backward = hook
        self._reg_stream = torch.cuda.Stream() if self.overlap else None
        if overlap_proposal_backward is None:
            overlap_proposal_backward = not self.reduce_grads
        self._prop_stream = torch.cuda.Stream() if (self.overlap and overlap_proposal_backward) else None
        model.proposal_sampler.side_stream = self._prop_stream
        # [0]: outputs no loss reads (accumulation / depths / median colour); [1..]: distortion and interlevel terms, forward
        # and (autograd replays a node on its forward's stream) backward, as parallel branches
        # (same priority as the captured step's stream: at the default priority they queue behind every block of the
        # regulariser sweep -- CUPTI, cfg3 at 2 ranks: the loss kernels waited 0.37 ms for it)
        self._branch_streams = ([torch.cuda.Stream(priority=-1 if prioritize_main_stream else 0) for _ in range(4)]
                                if (self.overlap and branch_small_kernels) else None)
        model._kp_branch_streams = self._branch_streams
        self.use_cuda_graph = use_cuda_graph
        # the captured step's own stream: above the side branches (priority 0), below the communication stream (-2)
        self._capture_stream = torch.cuda.Stream(priority=-1) if (on_cuda and use_cuda_graph and prioritize_main_stream) else None
        self._in_graph_body = False
        self.health_check_every = 1000  # steps between polls of the peer all-reduce's error word (a device sync)
        self._graphs: Dict[bool, torch.cuda.CUDAGraph] = {}
        self._graph_out: Dict[bool, Dict[str, torch.Tensor]] = {}
        self._seen: Dict[bool, int] = {}
        self._static: Optional[Dict[str, torch.Tensor]] = None

    def _setup_sparse_regularizers(self) -> None:
        """Per regularised plane: the float4 range of it that lies inside this rank's shard of the "fields" bucket (proposal
        planes: everything -- their bucket is exchanged densely) and the factor on its gradient (world for field planes)."""
        from ..model_components.losses import regularizer_plan

        model, dev = self.model, self.buckets["fields"].flat.device
        planes, _, _ = regularizer_plan(model.field.grids, [p.grids for p in model.proposal_networks])
        bucket, grp = self.buckets["fields"], self.sharded["fields"]
        lo, hi = grp.shard_slice()
        offsets = [bucket.param_offsets.get(id(p)) for p in planes]
        rng = plane_write_ranges(offsets, [p.numel() for p in planes], lo, hi)
        scale = [1.0 if off is None else float(self.world) for off in offsets]
        self._reg_range = torch.tensor(rng, dtype=torch.int64, device=dev)
        self._reg_scale = torch.tensor(scale, dtype=torch.float32, device=dev)
        # The loss VALUES are sharded like the gradient (kp_plane_reg_fused_shard): every rank sums the regulariser terms of
        # its own shard only -- 1/world of the 2.3 GB the full sweep reads (CUPTI, cfg3 at 2 ranks: the full sweep took
        # 2.0 ms and the loss head waited 0.9 ms for it) -- and the six scaled values are summed over the ranks in a few
        # floats of the arena after the exchange.  Planes of the densely exchanged group are swept in full by every rank:
        # their share is 1/world.
        self._reg_sum_scale = torch.tensor([1.0 / self.world if off is None else 1.0 for off in offsets], dtype=torch.float64,
                                           device=dev)
        self._reg_vals, self._reg_vals_off = self.arena.take(64)
        self._reg_vals.zero_()

    def _clean_field_bucket(self) -> None:
        """Before a sparse step that follows anything else (first step, a dense diagnostic iteration): the invariant "a line
        of the field bucket is zero unless marked" is re-established by one full memset."""
        b = self.buckets["fields"]
        b.flat.zero_()
        t = b.touched
        t.copy_(torch.where(t == 1, torch.zeros_like(t), t))
        self._bucket_dirty = False

    # ---- the iteration body (eager, or recorded into a graph) ------------------------------------------
    def _iteration(self, ray_bundle: RayBundle, batch: Dict[str, torch.Tensor], grad_scale_override=None):
        model = self.model
        if self.buckets:
            # planes whose gradient is written in full by the regulariser sweep, or (f1) left zeroed by the previous
            # step's fused regulariser + Adam pass, are not memset
            skip = self._reg_written
            if self._reg_adam is not None and self._reg_adam_first:
                skip, self._reg_adam_first = None, False  # first step: nothing has zeroed the planes' gradients yet
            sparse_now = self._sparse and self.reduce_grads and "fields" in self.sharded
            if sparse_now and self._bucket_dirty:
                self._clean_field_bucket()
            if self._sparse and not sparse_now:
                self._bucket_dirty = True  # a dense iteration (diagnostics) leaves reduced sums in the bucket
            for b in self.buckets.values():
                b.attach_zeroed(sink=self.fuse_grad_accumulation, skip=skip)
        else:
            self.optimizers.zero_grad_all()
        self._field_ready = None
        self._scale_ready = {}
        from .. import ops

        self._reg_mark = ops.PLANE_REG_BACKWARDS
        regs = None
        main = torch.cuda.current_stream() if self.overlap else None
        if self._reg_adam is not None:
            regs = {}  # evaluated inside the optimizer pass; merged into the loss dict after it
        elif self.overlap or self._reg_written:
            # regulariser branch: values and gradients (straight into the sinks) on its own stream, under the forward
            # pass; joined before the main backward, whose scatter kernels update the same gradients atomically
            if self.overlap:
                self._reg_stream.wait_stream(main)
            with torch.cuda.stream(self._reg_stream) if self.overlap else contextlib.nullcontext():
                if self._reg_written:
                    # one sweep per plane: loss values + gradient WRITTEN into the bucket (which was not memset there)
                    if self._sparse and self.reduce_grads and "fields" in self.sharded:
                        # field planes: only this rank's shard of the bucket, pre-multiplied by the world size; the values
                        # are this rank's share too and go to the arena slot that is summed over the ranks after the exchange
                        part = model.regularizers_into_grads(accumulate=False, write_range=self._reg_range, grad_scale=self._reg_scale,
                                                             sums_in_range=True, sum_scale=self._reg_sum_scale)
                        self._reg_names = list(part.keys())
                        if part:
                            self._reg_vals[: len(part)].copy_(torch.stack(list(part.values())))
                        regs = {}  # merged into the loss dict after the exchange (like the f1 path)
                    else:
                        regs = model.regularizers_into_grads(accumulate=False)
                else:
                    regs = model.regularizer_losses()
                    if regs:
                        regs = scale_dict(regs, model.config.loss_coefficients)
                        sum(regs.values()).backward()
                        regs = {k: v.detach() for k, v in regs.items()}
                if regs:
                    regs_vec = torch.stack(list(regs.values()))  # the loss head adds these into the total
        outputs = model(ray_bundle)
        if self._reg_adam is None and (self.overlap or self._reg_written):
            if self.overlap:
                main.wait_stream(self._reg_stream)
            if regs:
                outputs["_scaled_regularizers_vec"] = regs_vec
        metrics = model.get_metrics_dict(outputs, batch)
        loss_dict = model.get_loss_dict(outputs, batch, metrics, regularizers=regs)
        loss = getattr(loss_dict, "total", None)
        if loss is None:
            loss = sum(loss_dict.values())
        loss.backward()
        join_prop = None
        if self._prop_stream is not None and getattr(model.proposal_sampler, "side_stream_used", False):
            # the proposal networks' backward ran on the sampler's side stream: only THEIR optimizer group waits for it,
            # so the field group's optimizer pass (HBM-bound) can run next to its tail
            if self.reduce_grads or "proposal_networks" not in self.optimizers.optimizers:
                main.wait_stream(self._prop_stream)
            else:
                join_prop = {"proposal_networks": lambda: torch.cuda.current_stream().wait_stream(self._prop_stream)}
        grad_scale = 1.0
        if self.reduce_grads:
            main = torch.cuda.current_stream()
            fields, comm = self.buckets["fields"], self._comm_stream
            n_scales = len(getattr(model.field, "grids", ()))
            prop = self.buckets["proposal_networks"]
            # every collective goes through the communication stream, in one order on all ranks (the peer-memory
            # kernel's flag words, like an NCCL communicator, serve one collective at a time)
            if comm is None:  # CPU process group (gloo): nothing to overlap
                fields.all_reduce()
                prop.all_reduce()
                grad_scale = 1.0 / self.world
                self.optimizers.optimizer_step_all(grad_scale=grad_scale, use_device_hyper=self._in_graph_body)
                return self._finish(loss_dict, loss, metrics)
            if self.sharded:
                # reduce-scatter + Adam on the owned shard + all-gather of the new parameters: one kernel per group on the
                # communication stream.  The field group's starts as soon as the field's scatter is in the stream (its
                # in-kernel start barrier also guarantees that no rank still reads the field's parameters) and overlaps
                # the back-propagation through the proposal networks.
                if self._field_ready is not None:
                    comm.wait_event(self._field_ready)
                else:
                    comm.wait_stream(main)
                with torch.cuda.stream(comm):
                    self._sharded_step("fields")
                comm.wait_stream(main)
                with torch.cuda.stream(comm):
                    self._sharded_step("proposal_networks")
                    if self._sparse and getattr(self, "_reg_names", None):
                        self.arena.all_reduce(self._reg_vals_off, 64)  # the regularisers' values: sum of the ranks' shares
                main.wait_stream(comm)
                if self._sparse and getattr(self, "_reg_names", None):
                    vals = self._reg_vals[: len(self._reg_names)].clone()
                    loss_dict.update({n: vals[i] for i, n in enumerate(self._reg_names)})
                    loss = loss.detach() + vals.sum()
                return self._finish(loss_dict, loss, metrics)
            if self._first_span is not None and len(self._scale_ready) == n_scales:
                # finest scale (most of the bytes) while the other scales are still being scattered, the rest of the
                # bucket as soon as the last scatter is in the stream; both under the proposal networks' backward
                a, b = self._first_span
                comm.wait_event(self._scale_ready[n_scales - 1])
                with torch.cuda.stream(comm):
                    fields.all_reduce(span=(a, b))
                comm.wait_event(self._scale_ready[0])
                with torch.cuda.stream(comm):
                    fields.all_reduce(span=(0, a))
                    fields.all_reduce(span=(b, fields.flat.numel()))
            elif self._field_ready is not None:
                comm.wait_event(self._field_ready)  # fork: depends on the field scatter only
                with torch.cuda.stream(comm):
                    fields.all_reduce()
            else:
                comm.wait_stream(main)
                with torch.cuda.stream(comm):
                    fields.all_reduce()
            comm.wait_stream(main)  # the proposal networks' backward is complete
            with torch.cuda.stream(comm):
                prop.all_reduce()
            main.wait_stream(comm)  # join
            grad_scale = 1.0 / self.world
        self.optimizers.optimizer_step_all(grad_scale=grad_scale, use_device_hyper=self._in_graph_body, plane_reg=self._reg_adam,
                                           before=join_prop)
        if self._reg_adam is not None:
            regs = self._reg_adam.values()
            loss_dict.update(regs)
            loss = loss.detach() + torch.stack(list(regs.values())).sum()
        return self._finish(loss_dict, loss, metrics)

    def _sharded_step(self, name: str) -> None:
        opt = self.optimizers.optimizers[name]
        g = opt.param_groups[0]
        self._last_sharded_lr[name] = float(g["lr"])
        self.sharded[name].step(g["lr"], g["betas"], g["eps"], g["weight_decay"], self.step + 1, 1.0 / self.world,
                                hyper_dev=g.get("hyper_dev") if self._in_graph_body else None,
                                sparse=self._sparse and name == "fields")

    def _finish(self, loss_dict, loss, metrics):
        if self._branch_streams:
            torch.cuda.current_stream().wait_stream(self._branch_streams[0])  # the auxiliary outputs of get_outputs
        loss_dict = {k: v.detach() for k, v in loss_dict.items()}
        loss_dict["loss"] = loss.detach()
        loss_dict["psnr"] = metrics["psnr"]
        return loss_dict

    def _start_field_allreduce(self) -> None:
        """Called from the field's backward (autograd thread) right after its scatter kernel was enqueued: every
        gradient of the "fields" group (plane regularisers, decoder weights, scatter) is then in the stream, so an
        event recorded here marks the point from which the field bucket may be all-reduced.  The collective itself is
        issued by the main thread on a communication stream that waits only on this event, so -- eagerly and in the
        captured graph alike -- it overlaps the back-propagation through the proposal networks."""
        from .. import ops

        # only if the plane-regulariser backward (an independent autograd branch that also writes the field planes'
        # gradients) has already been enqueued; otherwise the bucket is reduced after the whole backward
        if self.reduce_grads and self._field_ready is None and ops.PLANE_REG_BACKWARDS > self._reg_mark:
            ev = torch.cuda.Event()
            ev.record()
            self._field_ready = ev

    def _scale_scattered(self, scale: int) -> None:
        """Per-scale variant of ``_start_field_allreduce``: the scatter of ``scale`` has just been enqueued (scales are
        scattered finest first).  Only used with the regulariser branch joined before the backward, so the planes'
        gradients of that scale are complete at this point of the stream."""
        ev = torch.cuda.Event()
        ev.record()
        self._scale_ready[scale] = ev

    _scale_scattered.per_scale = True

    def check_collective_health(self) -> None:
        """Raise if a cross-GPU barrier of the peer all-reduce ever timed out (the kernel traps in that case; this is the
        host-side report for the ranks that were not the one that trapped).  Synchronises the device."""
        if self.arena is not None and self.arena.error_word() != 0:
            raise RuntimeError("peer-memory all-reduce: a cross-GPU barrier timed out -- a rank fell behind by more than the "
                               "spin limit; gradients of this step are not trustworthy")

    def close(self) -> None:
        """Detach the gradient sinks from the parameters (another trainer / a plain torch optimizer may own the model
        next), report a failed collective, and release the peer arena."""
        for b in self.buckets.values():
            b.detach_sinks()
        if getattr(self.model.field, "_kp_post_backward", None) is not None:
            self.model.field._kp_post_backward = None
        self.model.proposal_sampler.side_stream = None
        self.model._kp_branch_streams = None
        self._prop_stream = None
        self._graphs.clear()
        self._graph_out.clear()
        if self.arena is not None:
            try:
                self.check_collective_health()
            finally:
                for p in self.model.parameters():
                    p.grad = None
                if self.sharded:  # the parameters live in the arena: move them back to ordinary tensors before it is freed
                    with torch.no_grad():
                        for grp in self.sharded.values():
                            for p in grp.bucket.params:
                                p.data = p.data.clone(memory_format=torch.preserve_format)
                    self.sharded = {}
                self.buckets = {}
                self.arena.close()
                self.arena = None

    def _run_callbacks(self, location: int) -> None:
        for cb in self.callbacks:
            cb.run_callback_at_location(self.step, location)

    def __call__(self, ray_bundle: RayBundle, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        self.model.train()
        self._run_callbacks(TrainingCallbackLocation.BEFORE_TRAIN_ITERATION)
        if self.use_cuda_graph:
            out = self._graphed(ray_bundle, batch)
        else:
            out = self._iteration(ray_bundle, batch)
        self.optimizers.scheduler_step_all(self.step)
        self._run_callbacks(TrainingCallbackLocation.AFTER_TRAIN_ITERATION)
        self.step += 1
        if self.arena is not None and self.step % self.health_check_every == 0:
            self.check_collective_health()
        return out

    # ---- checkpoint / resume (NS/engine/trainer.py:326-380) ----------------------------------------------------
    def state_dict(self) -> Dict:
        """``{"step", "pipeline", "optimizers", "schedulers", "scalers"}``: the top-level keys ``Trainer.save_checkpoint``
        writes (trainer.py:362-373), model tensors under the pipeline's ``_model.`` prefix with OUR parameter names.  This
        is this repo's resume format, NOT a reference checkpoint: ``step`` counts completed steps (the reference stores the
        last step index and resumes at step+1, trainer.py:343), the MLP keys are ours, and the reference's strict
        ``load_pipeline`` wants keys this path does not have (lpips, datamanager).  ``utils.checkpoint`` converts model
        tensors and optimizer state in both directions.  Optimizer step counts are taken from the trainer: a replayed
        graph advances them on the device."""
        optimizers = {}
        for name, opt in self.optimizers.optimizers.items():
            sd = opt.state_dict()
            for st in sd["state"].values():
                st["step"] = self.step
            sd["param_groups"] = [{k: v for k, v in g.items() if k != "hyper_dev"} for g in sd["param_groups"]]
            if name in self.sharded:  # assemble the full moments from the ranks' shards (a collective: call on every rank)
                grp = self.sharded[name]
                m_full, v_full = grp.gather_moments()
                index = {id(p): i for i, p in enumerate(opt.param_groups[0]["params"])}
                off = 0
                for p in grp.bucket.params:
                    sd["state"][index[id(p)]] = {
                        "step": self.step,
                        "exp_avg": torch.as_strided(m_full, p.shape, p.stride(), off).clone(memory_format=torch.preserve_format),
                        "exp_avg_sq": torch.as_strided(v_full, p.shape, p.stride(), off).clone(memory_format=torch.preserve_format)}
                    off += p.numel()
            optimizers[name] = sd
        return {
            "step": self.step,
            "pipeline": {f"_model.{k}": v for k, v in self.model.state_dict().items()},
            "optimizers": optimizers,
            "schedulers": {name: sch.state_dict() for name, sch in self.optimizers.schedulers.items()},
            "scalers": {},  # fp32 path: no GradScaler state (the reference stores its GradScaler's here)
        }

    def load_state_dict(self, state: Dict) -> None:
        """Resume from ``state_dict()`` (trainer.py:326-350).  Captured graphs are dropped: they hold the addresses of
        the optimizer state they were recorded with, and are re-captured on the next steps."""
        self.model.load_state_dict({k[len("_model."):]: v for k, v in state["pipeline"].items() if k.startswith("_model.")})
        for name, opt in self.optimizers.optimizers.items():
            keep = [g.get("hyper_dev") for g in opt.param_groups]
            opt.load_state_dict(state["optimizers"][name])
            for g, h in zip(opt.param_groups, keep):
                if h is not None:
                    g["hyper_dev"] = h
            if name in self.sharded:
                grp = self.sharded[name]
                m_full = torch.zeros(grp.count, dtype=torch.float32, device=grp.exp_avg.device)
                v_full = torch.zeros_like(m_full)
                off = 0
                for p in grp.bucket.params:
                    st = opt.state.get(p)
                    if st:
                        torch.as_strided(m_full, p.shape, p.stride(), off).copy_(st["exp_avg"])
                        torch.as_strided(v_full, p.shape, p.stride(), off).copy_(st["exp_avg_sq"])
                    off += p.numel()
                grp.scatter_moments(m_full, v_full)
                opt.state.clear()  # the shard owns the moments
        for name, sch in self.optimizers.schedulers.items():
            if name in state.get("schedulers", {}):
                sch.load_state_dict(state["schedulers"][name])
        self.step = int(state["step"])
        self.model.proposal_sampler._step = self.step  # the first `updated` decision after a resume uses the right step
        self._graphs.clear()
        self._graph_out.clear()
        self._seen.clear()
        if self._static is not None:
            self._step_t.fill_(self.step)

    # ---- CUDA-graph path ----------------------------------------------------------------------------------
    def _setup_graph_state(self, ray_bundle: RayBundle, batch) -> None:
        dev = ray_bundle.origins.device
        n = ray_bundle.origins.shape[0]
        self._static = {
            "origins": torch.empty(n, 3, device=dev), "directions": torch.empty(n, 3, device=dev),
            "pixel_area": torch.ones(n, 1, device=dev), "times": torch.empty(n, 1, device=dev),
            "image": torch.empty(n, 3, device=dev),
        }
        steps = np.arange(self.max_steps + 1)
        lr = np.array([self.base_lr * cosine_decay_factor(int(s), self.warm_up_end, self.max_steps) for s in steps])
        cfg = self.model.config
        frac = np.clip(steps / cfg.proposal_weights_anneal_max_num_iters, 0, 1)
        b = cfg.proposal_weights_anneal_slope
        anneal = (b * frac) / ((b - 1) * frac + 1) if cfg.use_proposal_weight_anneal else np.ones_like(frac)
        self._lr_table = torch.tensor(lr, dtype=torch.float64, device=dev)
        self._anneal_table = torch.tensor(anneal, dtype=torch.float32, device=dev)
        self._step_t = torch.full((), self.step, dtype=torch.int64, device=dev)
        self._anneal_t = torch.ones((), dtype=torch.float32, device=dev)
        self._grad_scale = 1.0 / self.world if self.reduce_grads else 1.0
        for opt in self.optimizers.optimizers.values():
            for group in opt.param_groups:
                group["hyper_dev"] = torch.zeros(3, dtype=torch.float32, device=dev)

    def _device_scalars(self) -> None:
        """In-graph prologue: anneal exponent and Adam scalars for the device-resident step counter (and the counter's
        increment) -- one kernel (kp_step_scalars) on CUDA."""
        if self._step_t.is_cuda:
            from ctypes import c_float, c_void_p

            from .. import _lib

            groups = [g for opt in self.optimizers.optimizers.values() for g in opt.param_groups]
            betas = (c_float * (2 * len(groups)))(*[float(b) for g in groups for b in g["betas"]])
            hyper = (c_void_p * len(groups))(*[g["hyper_dev"].data_ptr() for g in groups])
            _lib.call("kp_step_scalars", c_void_p(self._step_t.data_ptr()), c_void_p(self._lr_table.data_ptr()),
                      c_void_p(self._anneal_table.data_ptr()), int(self.max_steps), len(groups), betas, float(self._grad_scale),
                      c_void_p(self._anneal_t.data_ptr()), hyper, _lib.stream_ptr())
            return
        idx = self._step_t.clamp(max=self.max_steps).view(1)  # 1-d index: a 0-d tensor index would .item() (host sync)
        self._anneal_t.copy_(self._anneal_table.gather(0, idx).view(()))
        t = (self._step_t + 1).double()
        lr = self._lr_table.gather(0, idx).view(())
        for opt in self.optimizers.optimizers.values():
            for group in opt.param_groups:
                b1, b2 = group["betas"]
                bc1 = 1.0 - torch.pow(torch.full_like(t, b1), t)
                bc2 = 1.0 - torch.pow(torch.full_like(t, b2), t)
                hyper = torch.stack([lr / bc1, torch.rsqrt(bc2), torch.full_like(t, self._grad_scale)]).float()
                group["hyper_dev"].copy_(hyper)

    def _graph_body(self):
        s = self._static
        self._in_graph_body = True
        try:
            return self._graph_body_inner(s)
        finally:
            self._in_graph_body = False

    def _graph_body_inner(self, s):
        self._device_scalars()
        rb = RayBundle(origins=s["origins"], directions=s["directions"], pixel_area=s["pixel_area"], times=s["times"])
        out = self._iteration(rb, {"image": s["image"]})
        if not self._step_t.is_cuda:
            self._step_t += 1  # (on CUDA kp_step_scalars advanced the counter)
        return out

    def _graphable(self, ray_bundle: RayBundle, batch) -> bool:
        """The captured step has static buffers for origins / directions / times / image only.  Anything else a batch or
        bundle may carry and the model would USE (depth supervision and the directions_norm it needs, unknown per-ray
        metadata, camera indices for an appearance embedding, preset near / far bounds of a crop box, a static scene
        without times) runs through the eager iteration instead of being dropped silently.  What a datamanager's batch
        carries and the model never reads (kplanes.py:392-449 reads ``image`` and ``depth_image`` only: ``indices``,
        ``ist_weights``, ``mask`` ...) does not matter, nor does ``directions_norm`` without depth supervision."""
        if "image" not in batch or "depth_image" in batch:
            return False
        if ray_bundle.times is None or ray_bundle.nears is not None or ray_bundle.fars is not None:
            return False
        if ray_bundle.metadata and set(ray_bundle.metadata.keys()) - {"directions_norm"}:
            return False
        if getattr(self.model.field, "use_appearance_embedding", False):
            return False
        return True

    def _graphed(self, ray_bundle: RayBundle, batch) -> Dict[str, torch.Tensor]:
        sampler = self.model.proposal_sampler
        if not self._graphable(ray_bundle, batch):
            return self._iteration(ray_bundle, batch)
        if self._static is None:
            self._setup_graph_state(ray_bundle, batch)
        updated = bool(sampler._steps_since_update > sampler.update_sched(sampler._step) or sampler._step < 10)
        s = self._static
        s["origins"].copy_(ray_bundle.origins, non_blocking=True)
        s["directions"].copy_(ray_bundle.directions, non_blocking=True)
        s["times"].copy_(ray_bundle.times, non_blocking=True)
        s["image"].copy_(batch["image"], non_blocking=True)
        sampler._anneal = self._anneal_t  # device-resident exponent (the host callback's float is ignored here)
        seen = self._seen.get(updated, 0)
        self._seen[updated] = seen + 1
        if seen < 2:  # first two visits of a mode run eagerly (lazy initialisations, allocator warm-up)
            return self._graph_body()
        if updated not in self._graphs:
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=self._capture_stream, capture_error_mode="thread_local"):
                self._graph_out[updated] = self._graph_body()
            self._graphs[updated] = g
        self._graphs[updated].replay()
        if updated:
            sampler._steps_since_update = 0  # host-side state the captured Python would have set
        return self._graph_out[updated]
