#!/usr/bin/env python
"""bench.py -- train rays/sec of the K-Planes training step (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 5            # our arm (CUDA path through the C-ABI)
    python bench.py --impl reference --steps 3 --warmup 1     # the reference's CPU torch path (oracle port)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one full training iteration on one batch of synthetic rays of the broadcast-style scene shape:
AABB collider -> proposal sampling (256 + 128 samples, both proposal fields evaluated and trained every step)
-> K-Planes field -> compositing -> rgb / distortion / interlevel / plane regularisers -> backward ->
[all-reduce of the flat gradient buckets when N>1] -> Adam (lr 1e-2, eps 1e-12) -> cosine LR.

Headline (`value`, `e2e`, `roofline`): BASELINE.json configs[1] ("K-Planes default", cfg2: multiscale-res 1 2 4 8,
4096 rays/step, 48 samples).  The same JSON line also carries
  "cfg3"               the 32x preset (configs[2]: multiscale-res 1..32, T=100, sigma hidden 128, 64 samples; 2.3 GB of
                       planes, the HBM-bound regime) with its own value / e2e / per-scale roofline;
  "eval"               configs[4]: full 1920x1080 frames from the cfg3 field, tile-sharded over the ranks;
  "l2_peaks"           the achievable rate of the field kernels' own access pattern (random 128-byte lines) in L2 and in
                       HBM, measured live with kp_line_probe: the denominators of the per-scale fractions;
  "gpu_torch_baseline" the reference's own step as plain torch ops on the same GPU (N=1 only): the like-for-like "before";
  "pixel_sampler"      (N=1 only) the importance pixel sampler of the preset's shape (4096-ray batch, 10 % importance
                       pixels = 41 maps of 540x960 x 10 pixels): kp_importance_pixels on the device next to the reference's
                       per-image torch.multinomial loop on the host cores;
  "device_pipeline"    (N=1 only) the headline step fed by the device-resident datamanager (importance + uniform pixel
                       sampling on a 418-image cache in HBM, pixel gather, ray generation): rays/s of next_train + step,
                       serial and with the batch prefetched on a side stream;
  "cpu_baseline"       the oracle port on the host cores (N=1 only);
  "dp_check"           (N>1) parameters bit-identical across ranks after all steps, and the all-reduced gradient equal to
                       a single-rank gradient on the concatenated batch.
N>1: weak scaling, every rank owns its own 4096-ray slice of a global N*4096-ray batch.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

RAYS_PER_RANK = 4096
L2_BYTES = 126 * 1024 * 1024

WORKLOADS = {
    "cfg2": {
        "label": "kplanes-default(cfg2): multiscale-res 1 2 4 8, C=32, T=50, proposals [128^3,150]+[256^3,150] C=8, "
                 "256/128/48 samples, 4096 rays/rank",
        "oracle": "cfg2",
        "model": {},
    },
    "cfg3": {
        "label": "kplanes-32x(cfg3): multiscale-res 1 2 4 8 16 32, C=32, T=100, sigma hidden 128, no view dependence, "
                 "proposals [128^3,100]+[256^3,100] C=8, 256/128/64 samples, 4096 rays/rank",
        "oracle": "cfg3",
        "model": dict(spacetime_resolution=(64, 64, 64, 100), multiscale_res=(1, 2, 4, 8, 16, 32), sigma_net_hidden_dim=128,
                      disable_viewing_dependent=True, num_nerf_samples_per_ray=64, num_proposal_samples_per_ray=(256, 128),
                      proposal_net_args_list=[{"feature_dim": 8, "resolution": [128, 128, 128, 100]},
                                              {"feature_dim": 8, "resolution": [256, 256, 256, 100]}],
                      eval_num_rays_per_chunk=32768),
    },
}
WORKLOAD = WORKLOADS["cfg2"]["label"]
FIELD_KERNELS = ("kp_hexplane_fwd", "kp_hexplane_bwd", "kp_density_field_fwd", "kp_density_field_bwd")
OTHER_TIMED = ("kp_decoder_fwd_fused", "kp_color_net_bwd", "kp_sigma_net_bwd", "kp_sigma_net_fwd", "kp_color_net_fwd",
               "kp_adam_multi", "kp_plane_reg_multi_fwd", "kp_plane_reg_multi_bwd", "kp_plane_reg_fused", "kp_plane_reg_adam",
               "kp_peer_allreduce", "kp_peer_sharded_adam", "kp_peer_sharded_adam_sparse", "kp_hexplane_bwd_flags",
               "kp_plane_reg_fused_range")
ALIASES = {"kp_hexplane_bwd_flags": "kp_hexplane_bwd"}  # same kernel, with the touched-line marks switched on


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        # median of the samples taken UNDER LOAD (the idle samples before / after the timed region sit at the idle clock)
        busy = sorted(s for s in sm if mx is None or s >= 0.5 * mx) or sorted(sm)
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def _dist_setup():
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def _make_batches(n_steps: int, n_rays: int, seed: int):
    """Pinned host batches: (origins, directions, times, image) per step; one packed [n_rays,10] tensor each."""
    from soccernerfs_b200.data.synthetic import synthetic_rays

    gen = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_steps):
        o, d, t, _ = synthetic_rays(n_rays, gen)
        img = torch.rand(n_rays, 3, generator=gen)
        out.append(torch.cat([o, d, t, img], dim=-1).contiguous().pin_memory())
    return out


def _bundle(packed_dev):
    from soccernerfs_b200.cameras.rays import RayBundle

    n = packed_dev.shape[0]
    rb = RayBundle(origins=packed_dev[:, 0:3].contiguous(), directions=packed_dev[:, 3:6].contiguous(),
                   pixel_area=torch.ones(n, 1, device=packed_dev.device), times=packed_dev[:, 6:7].contiguous())
    return rb, {"image": packed_dev[:, 7:10].contiguous()}


def build_model(name: str, dev):
    """Same parameters on every rank (seed 42, like the reference's DDP broadcast of rank 0's model)."""
    from soccernerfs_b200.data.scene_box import SceneBox
    from soccernerfs_b200.data.synthetic import perturb_time_planes, synthetic_rays
    from soccernerfs_b200.models.kplanes import KPlanesModelConfig

    torch.manual_seed(42)
    _, _, _, aabb = synthetic_rays(4, torch.Generator().manual_seed(0))
    model = KPlanesModelConfig(**WORKLOADS[name]["model"]).setup(scene_box=SceneBox(aabb=aabb), num_train_data=19 * 25).to(dev)
    perturb_time_planes(model)
    return model


# ----------------------------------------------------------------------------------------------------------------------
# memory-hierarchy probe: achievable rate of the field kernels' access pattern in L2 and in HBM
# ----------------------------------------------------------------------------------------------------------------------
def line_probe_peaks(dev):
    """kp_line_probe over an L2-resident (64 MB) and an HBM-resident (4 GB) buffer, loads and red.v4 -> GB/s."""
    from ctypes import c_int64, byref, c_void_p

    from soccernerfs_b200 import _lib

    out = {"pattern": "8 lanes x 16 B = one 128-byte line at a pseudo-random texel, 8 independent lines in flight per lane",
           "l2_buffer_mb": 64, "hbm_buffer_mb": 4096}
    sink = torch.zeros(1, device=dev)
    blocks, iters = 148 * 8, 256
    for level, mb in (("l2", 64), ("hbm", 4096)):
        buf = torch.zeros(mb * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
        n_lines = buf.numel() // 32
        for mode, mname in ((0, "gather"), (1, "red")):
            best = None
            for rep in range(4):
                touched = c_int64(0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.call("kp_line_probe", c_void_p(buf.data_ptr()), n_lines, mode, blocks, iters, 1234 + rep, c_void_p(sink.data_ptr()),
                          byref(touched), _lib.stream_ptr())
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                if rep > 0:  # first launch warms the buffer into L2 (l2 level) / the TLBs
                    best = ms if best is None else min(best, ms)
            out[f"{level}_{mname}_gbs"] = touched.value * 128 / (best * 1e-3) / 1e9
        del buf
        torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------------------------------
# per-scale gather / scatter on the step's own samples
# ----------------------------------------------------------------------------------------------------------------------
def per_scale_probe(model, flush, peaks, probe, reps=3):
    from ctypes import c_void_p

    from soccernerfs_b200 import _lib, ops

    field = model.field
    pts = field._last_points
    ms_planes = [[ops.as_channel_last(p.detach()) for p in g] for g in field.grids]
    flat = [p for g in ms_planes for p in g]
    n_scales, n_planes = len(ms_planes), len(ms_planes[0])
    c, m = flat[0].shape[1], pts.M
    dev = flat[0].device
    gout = torch.randn(m, n_scales * c, device=dev)
    out1 = torch.empty(m, c, device=dev)
    pstruct = pts.struct()
    rows = []

    def timed(fn):
        best = None
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1)
            best = t if best is None else min(best, t)
        return best

    for k in range(n_scales):
        planes = ms_planes[k]
        plane_bytes = sum(p.numel() * 4 for p in planes)
        grads = [p.grad if (p.grad is not None and p.grad.stride() == q.stride()) else torch.zeros_like(q)
                 for p, q in zip(field.grids[k], planes)]
        targets = [None] * (n_scales * n_planes)
        targets[k * n_planes:(k + 1) * n_planes] = grads
        t_f = timed(lambda: _lib.call("kp_hexplane_fwd", ops._plane_ptrs(planes), ops._plane_hw(planes), 1, n_planes, c, pstruct, m, 1,
                                      0x3F, c_void_p(out1.data_ptr()), _lib.stream_ptr()))
        t_b = timed(lambda: _lib.call("kp_hexplane_bwd", ops._plane_ptrs(flat), ops._plane_ptrs(targets), ops._plane_hw(flat), n_scales,
                                      n_planes, c, pstruct, m, 1, 0x3F, c_void_p(gout.data_ptr()), _lib.stream_ptr()))
        algo_f = m * n_planes * 4 * c * 4
        # distinct 128-byte texel lines this step's samples touch at this scale: one more scatter into a zeroed
        # gradient copy, then count the lines that received anything.  Compulsory DRAM bytes of an HBM-resident scale:
        # gather = each distinct plane line once; scatter = plane line once (re-gather) + gradient line read-modify-write.
        zgrads = [torch.zeros_like(q) for q in planes]
        ztargets = [None] * (n_scales * n_planes)
        ztargets[k * n_planes:(k + 1) * n_planes] = zgrads
        _lib.call("kp_hexplane_bwd", ops._plane_ptrs(flat), ops._plane_ptrs(ztargets), ops._plane_hw(flat), n_scales, n_planes, c,
                  pstruct, m, 1, 0x3F, c_void_p(gout.data_ptr()), _lib.stream_ptr())
        lines = 0
        for zg in zgrads:  # physical layout is [H][W][C]: one texel = C contiguous floats
            lines += int((zg.permute(0, 2, 3, 1).reshape(-1, c) != 0).any(dim=1).sum())
        del zgrads, ztargets
        line_bytes = c * 4
        row = {"scale": int(field.multiscale_res_multipliers[k]), "plane_mb": plane_bytes / 2**20,
               "distinct_lines_touched": lines, "lines_in_scale": plane_bytes // line_bytes}
        for tag, t, algo, ws, l2key, hbmkey, compulsory in (
                ("gather", t_f, algo_f, plane_bytes, "l2_gather_gbs", "hbm_gather_gbs", lines * line_bytes),
                ("scatter", t_b, 2 * algo_f, 2 * plane_bytes, "l2_red_gbs", "hbm_red_gbs", 3 * lines * line_bytes)):
            level = "l2" if ws <= 0.6 * L2_BYTES else ("hbm" if ws >= 2 * L2_BYTES else "l2+hbm")
            gbs = algo / (t * 1e-3) / 1e9
            # what the bound resource of that level carries: L2 serves every algorithmic load; the reduction units carry
            # the red payload (= algo_f) while the re-gather of the same bytes rides along; HBM carries the distinct lines
            l2_bytes = algo_f
            r = {"ms": t, "algorithmic_gbs": gbs, "working_set_mb": ws / 2**20, "level": level,
                 "l2_level_gbs": l2_bytes / (t * 1e-3) / 1e9,
                 "frac_of_l2_probe": (l2_bytes / (t * 1e-3) / 1e9) / probe[l2key] if probe else None,
                 "compulsory_dram_mb": compulsory / 2**20, "compulsory_dram_gbs": compulsory / (t * 1e-3) / 1e9,
                 "frac_of_hbm_peak": (compulsory / (t * 1e-3) / 1e9) / peaks["hbm_gbs"]}
            r["frac"] = r["frac_of_hbm_peak"] if level == "hbm" else (r["frac_of_l2_probe"] if level == "l2" else
                                                                      max(r["frac_of_hbm_peak"], r["frac_of_l2_probe"] or 0.0))
            row[tag] = r
        rows.append(row)
    return rows


# ----------------------------------------------------------------------------------------------------------------------
# data-parallel correctness (N>1): the driver never runs tests/test_gpu_multi.py, so the bench asserts it
# ----------------------------------------------------------------------------------------------------------------------
@contextlib.contextmanager
def _global_rand(rank, world, seed, dev):
    """torch.rand([n, ...]) -> this rank's rows of a [world*n, ...] draw every rank generates identically (rank<0: all rows)."""
    real = torch.rand
    gen = torch.Generator(device=dev).manual_seed(seed)

    def fake(*size, **kw):
        if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)):
            size = tuple(size[0])
        n = size[0] if rank >= 0 else size[0] // world
        full = real((n * world,) + tuple(size[1:]), generator=gen, device=dev)
        return full if rank < 0 else full[rank * n:(rank + 1) * n].contiguous()

    torch.rand = fake
    try:
        yield
    finally:
        torch.rand = real


def dp_check(trainer, model, rank, world, dev):
    """(1) after all the steps so far the replicas hold bit-identical parameters; (2) the summed gradient of one more
    data-parallel step equals the gradient ONE rank computes on the concatenated batch; (3) with the sharded optimizer:
    one fused reduce-scatter + Adam + all-gather step equals torch's Adam formulas applied to that summed gradient on
    the owned shard.  The trainer's state is restored afterwards."""
    import torch.distributed as dist

    def checksum():
        acc = torch.zeros((), dtype=torch.int64, device=dev)
        for p in model.parameters():
            if p.numel():
                acc += p.detach().contiguous().view(-1).view(torch.int32).to(torch.int64).sum()
        return acc

    sums = [torch.zeros((), dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sums, checksum())
    identical = all(int(s) == int(sums[0]) for s in sums)
    from soccernerfs_b200.data.synthetic import synthetic_rays

    gen = torch.Generator().manual_seed(777)
    o, d, t, _ = synthetic_rays(RAYS_PER_RANK * world, gen)
    img = torch.rand(RAYS_PER_RANK * world, 3, generator=gen)
    packed = torch.cat([o, d, t, img], dim=-1).to(dev)
    keep = [p.detach().clone(memory_format=torch.preserve_format) for p in model.parameters()]

    def restore():
        with torch.no_grad():
            for p, q in zip(model.parameters(), keep):
                p.copy_(q)

    def run(ray_bundle, batch):
        # one iteration at the CURRENT trainer.step (no step / scheduler advance): the three runs below must see the same
        # proposal-weight anneal exponent, learning rate and Adam step number
        model.train()
        trainer._run_callbacks(1)  # TrainingCallbackLocation.BEFORE_TRAIN_ITERATION: sets the anneal exponent
        model.proposal_sampler._steps_since_update = 1
        trainer._iteration(ray_bundle, batch)

    was_graph, trainer.use_cuda_graph = trainer.use_cuda_graph, False
    sharded, trainer.sharded = trainer.sharded, {}  # (2) runs through the all-reduce path, optimizer switched off
    step_all = trainer.optimizers.optimizer_step_all
    trainer.optimizers.optimizer_step_all = lambda **kw: None
    sl = slice(rank * RAYS_PER_RANK, (rank + 1) * RAYS_PER_RANK)
    try:
        with _global_rand(rank, world, 99, dev):
            run(*_bundle(packed[sl].contiguous()))
        g_sum = {k: b.flat.clone() for k, b in trainer.buckets.items()}
        reduce_grads, trainer.reduce_grads = trainer.reduce_grads, False
        try:
            with _global_rand(-1, world, 99, dev):
                run(*_bundle(packed))
        finally:
            trainer.reduce_grads = reduce_grads
    finally:
        trainer.optimizers.optimizer_step_all = step_all
        trainer.sharded = sharded
    rel = {}
    for k, b in trainer.buckets.items():
        ref = b.flat
        rel[k] = float((g_sum[k] / world - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    worst = torch.tensor([max(rel.values())], dtype=torch.float64, device=dev)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    out = {"params_bit_identical_across_ranks": identical, "grad_vs_single_rank_on_concatenated_batch_rel": float(worst),
           "tolerance": 1e-4, "per_group": rel}
    ok = bool(identical and float(worst) < 1e-4)  # north_star's fp32 gradient bar (fp32 atomics: the summation order differs
    # between the two batch shapes; measured 1e-5 at 2 ranks, 7e-5 at 8 ranks x 4096 rays on cfg3)
    if sharded:  # (3)
        snap = {k: (g.param_flat.clone(), g.exp_avg.clone(), g.exp_avg_sq.clone()) for k, g in sharded.items()}
        with _global_rand(rank, world, 99, dev):
            run(*_bundle(packed[sl].contiguous()))
        worst3 = 0.0
        for k, g in sharded.items():
            grp = trainer.optimizers.optimizers[k].param_groups[0]
            b1, b2 = grp["betas"]
            step_no = trainer.step + 1  # Adam's 1-based step number of the iteration just run
            p0, m0, v0 = snap[k]
            lo, hi = g.shard_slice()
            n = hi - lo
            gr = torch.zeros(n, device=dev)
            avail = min(hi, g_sum[k].numel()) - lo
            if avail > 0:
                gr[:avail] = g_sum[k][lo:lo + avail] / world
            m1 = m0[:n] + (gr - m0[:n]) * (1.0 - b1)
            v1 = v0[:n] * b2 + (1.0 - b2) * gr * gr
            lr = trainer._last_sharded_lr[k]
            lr_over_bc1 = torch.tensor(lr / (1.0 - b1 ** step_no), dtype=torch.float32, device=dev)
            inv_sqrt_bc2 = torch.tensor(1.0 / (1.0 - b2 ** step_no) ** 0.5, dtype=torch.float32, device=dev)
            p_old = torch.zeros(n, device=dev)
            availp = min(hi, p0.numel()) - lo
            if availp > 0:
                p_old[:availp] = p0[lo:lo + availp]
            p1 = p_old - lr_over_bc1 * m1 / (torch.sqrt(v1) * inv_sqrt_bc2 + grp["eps"])
            got_p = torch.zeros(n, device=dev)
            if availp > 0:
                got_p[:availp] = g.param_flat[lo:lo + availp]
            e_m = float((g.exp_avg[:n] - m1).abs().max() / m1.abs().max().clamp_min(1e-30))
            e_v = float((g.exp_avg_sq[:n] - v1).abs().max() / v1.abs().max().clamp_min(1e-30))
            e_p = float((got_p - p1).abs().max() / max(lr, 1e-12))
            worst3 = max(worst3, e_m, e_v, e_p * 1e-2)  # parameters: within 1e-3 of one learning-rate step
        for k, g in sharded.items():  # put the optimizer state and the parameters back
            g.param_flat.copy_(snap[k][0])
            g.exp_avg.copy_(snap[k][1])
            g.exp_avg_sq.copy_(snap[k][2])
        w3 = torch.tensor([worst3], dtype=torch.float64, device=dev)
        dist.all_reduce(w3, op=dist.ReduceOp.MAX)
        out["sharded_adam_vs_torch_formulas_on_reduced_gradient"] = float(w3)
        ok = ok and float(w3) < 5e-5
    restore()
    trainer.use_cuda_graph = was_graph
    out["ok"] = ok
    return out


# ----------------------------------------------------------------------------------------------------------------------
# one training leg (cfg2 or cfg3)
# ----------------------------------------------------------------------------------------------------------------------
def train_leg(name, args, rank, world, local, dev, peaks, probe, keep_model=False):
    import torch.distributed as dist

    from soccernerfs_b200 import _lib
    from soccernerfs_b200.engine.trainer import TrainStep

    model = build_model(name, dev)
    torch.manual_seed(42 + rank)  # per-rank sampling randomness (NS/scripts/train.py:84)
    model.proposal_sampler.update_sched = lambda step: 0  # proposal networks evaluated with grad + trained EVERY step
    prop_overlap = {"auto": None, "on": True, "off": False}[args.prop_overlap]
    trainer = TrainStep(model, data_parallel=args.allreduce != "none", use_cuda_graph=not args.eager,
                        overlap_branches=not args.no_overlap, overlap_proposal_backward=prop_overlap,
                        allreduce_mode=args.allreduce if args.allreduce != "none" else "overlap",
                        allreduce_backend=args.allreduce_backend,
                        shard_optimizer={"auto": None, "on": True, "off": False}[args.shard_optimizer],
                        fuse_reg_adam={"auto": None, "on": True, "off": False}[args.reg_adam],
                        sparse_grad_exchange={"auto": None, "on": True, "off": False}[args.sparse_exchange],
                        branch_small_kernels=not args.no_branches, prioritize_main_stream=not args.no_stream_priority)
    n_steps = args.warmup + args.steps
    host = _make_batches(n_steps, RAYS_PER_RANK, seed=1000 + rank)
    resident = [h.to(dev) for h in host]
    flush = torch.empty(160 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_region(step_fn, steps=None, warmup=None):
        steps = args.steps if steps is None else steps
        warmup = args.warmup if warmup is None else warmup
        # the CUDA graph is captured on the third visit of a sampler mode: never let that fall into the timed steps,
        # whatever --warmup says (the contract asks for W >= 3 anyway)
        for i in range(max(0, 3 - warmup)):
            step_fn(i % n_steps)
        for i in range(warmup):
            step_fn(i % n_steps)
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for i in range(steps):
            flush.zero_()  # L2 flush between timed iterations (outside the per-step events)
            ev[i][0].record()
            step_fn((warmup + i) % n_steps)
            ev[i][1].record()
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        per_rank = [ms]
        if world > 1:
            allms = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
            dist.all_gather(allms, t)
            per_rank = [float(x) for x in allms]
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), sorted(p / steps for p in per_rank)

    # ---- arm 1: inputs resident in HBM --------------------------------------------------------------
    def step_resident(i):
        rb, batch = _bundle(resident[i])
        trainer(rb, batch)

    sampler = ClockSampler(local)
    sampler.start()  # every rank samples its own GPU
    total_ms, rank_ms = timed_region(step_resident)
    clocks = sampler.stop()
    if world > 1:
        allc = [None] * world
        dist.all_gather_object(allc, clocks)
        if rank == 0:
            clocks = dict(allc[0])
            clocks["per_rank_sm_mhz"] = [c.get("sm_mhz") for c in allc]
            clocks["reasons"] = sorted({r for c in allc for r in c.get("reasons", [])})
    # a graph replay launches the kernels captured once; count them from one eager iteration of the same step
    was_graph, trainer.use_cuda_graph = trainer.use_cuda_graph, False
    k0 = _lib.launch_count()
    step_resident(0)
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count() - k0
    trainer.use_cuda_graph = was_graph

    # ---- arm 2: end to end through the public API with host buffers -----------------------------------
    # every step: pinned host batch -> device (H2D), the step, and the step's loss -> pinned host memory (D2H).  Both
    # copies are stream-ordered and asynchronous (a training loop that logs the loss does not stall the GPU on it);
    # the region ends with a synchronize, after which every step's loss is on the host and is checked.
    loss_host = torch.zeros(n_steps, dtype=torch.float32).pin_memory()

    def step_e2e(i):
        packed = host[i].to(dev, non_blocking=True)  # pinned H2D inside the timed region
        rb, batch = _bundle(packed)
        out = trainer(rb, batch)
        loss_host[i].copy_(out["loss"], non_blocking=True)  # D2H read of the step's result

    e2e_ms, _ = timed_region(step_e2e)
    if not bool(torch.isfinite(loss_host).all()) or float(loss_host[args.warmup:].abs().min()) == 0.0:
        raise RuntimeError("end-to-end arm: a step's loss did not arrive on the host")
    last_loss = float(loss_host[n_steps - 1])

    leg = {}
    # ---- arm 3 (N=1, headline config): the data side on the device too ---------------------------------------
    # DynamicDataManager.next_train (importance + uniform pixel sampling on a resident 418-image cache, pixel gather, ray
    # generation) feeding the same graphed step: the whole training iteration of the preset with nothing crossing PCIe
    # but the loss.  An accessory of the line: a failure is reported inside it.
    if world == 1 and name == "cfg2" and "sampler" in set(args.legs.split(",")):
        try:
            leg["device_pipeline"] = _device_pipeline_arm(trainer, timed_region, loss_host, dev, args)
        except Exception as e:  # noqa: BLE001
            leg["device_pipeline"] = {"error": repr(e)}
            print(f"[bench] device pipeline arm failed: {e!r}", file=sys.stderr, flush=True)
    # ---- N>1: a longer run (the 20-step region is too short to separate ranks' skew from the collective) --------
    if world > 1 and not args.no_long_run:
        long_ms, long_rank = timed_region(step_resident, steps=100, warmup=3)
        leg["long_run"] = {"steps": 100, "ms_per_step": long_ms / 100, "rank_ms_per_step": {"min": long_rank[0], "median": long_rank[len(long_rank) // 2], "max": long_rank[-1]},
                           "value": RAYS_PER_RANK * world * 100 / (long_ms * 1e-3)}

    # ---- the reference's own proposal schedule after warm-up: proposal networks updated every 6th step ----------
    if not args.no_ref_schedule:
        model.proposal_sampler.update_sched = lambda step: 5
        model.proposal_sampler._step = max(model.proposal_sampler._step, 10)
        sched_steps = max(args.steps, 18)
        sched_ms, _ = timed_region(step_resident, steps=sched_steps, warmup=18)
        leg["reference_schedule"] = {"value": RAYS_PER_RANK * world * sched_steps / (sched_ms * 1e-3), "ms_per_step": sched_ms / sched_steps,
                                     "steps": sched_steps, "proposal_update": "every 6th step (update_sched = 5: NS/models/kplanes.py:254-259 after proposal_warmup)"}
        model.proposal_sampler.update_sched = lambda step: 0

    if world > 1 and trainer.reduce_grads:
        leg["dp_check"] = dp_check(trainer, model, rank, world, dev)

    # ---- per-kernel durations: CUDA events around the C-ABI calls on their launch stream, measured live in eager mode
    #      (events cannot be timed inside a graph replay), L2 flushed per step, each kernel alone on the GPU ------
    trainer.use_cuda_graph = False
    trainer.overlap, trainer._prop_stream, model.proposal_sampler.side_stream = False, None, None
    _lib.TIMED.update(FIELD_KERNELS + OTHER_TIMED)
    _lib.EVENTS.clear()
    k_steps = min(args.steps, 10)
    for i in range(k_steps):
        flush.zero_()
        step_resident(i % n_steps)
    torch.cuda.synchronize()
    _lib.TIMED.clear()
    kernel_ms = {}
    for kname, a, b in _lib.EVENTS:
        kernel_ms.setdefault(ALIASES.get(kname, kname), []).append(a.elapsed_time(b))
    _lib.EVENTS.clear()
    scales = per_scale_probe(model, flush, peaks, probe) if rank == 0 else None

    cfg = model.config
    n_field = RAYS_PER_RANK * cfg.num_nerf_samples_per_ray
    n_prop = RAYS_PER_RANK * sum(cfg.num_proposal_samples_per_ray)
    k_scales = len(cfg.multiscale_res)
    step_bytes = {"kp_hexplane_fwd": n_field * k_scales * 6 * 4 * cfg.feature_dim * 4,
                  "kp_hexplane_bwd": 2 * n_field * k_scales * 6 * 4 * cfg.feature_dim * 4,
                  "kp_density_field_fwd": n_prop * 6 * 4 * 8 * 4, "kp_density_field_bwd": 2 * n_prop * 6 * 4 * 8 * 4}
    per_kernel = {}
    for kname, ms_list in kernel_ms.items():
        per_kernel[kname] = {"ms_per_step": sum(ms_list) / k_steps, "launches_per_step": len(ms_list) / k_steps}
    for kname, nbytes in step_bytes.items():
        if kname in per_kernel and per_kernel[kname]["ms_per_step"] > 0:
            per_kernel[kname]["algorithmic_gbs"] = nbytes / (per_kernel[kname]["ms_per_step"] * 1e-3) / 1e9
    rays = RAYS_PER_RANK * world * args.steps
    leg.update({
        "workload": WORKLOADS[name]["label"], "value": rays / (total_ms * 1e-3), "unit": "rays/s", "ms_per_step": total_ms / args.steps,
        "rank_ms_per_step": {"min": rank_ms[0], "median": rank_ms[len(rank_ms) // 2], "max": rank_ms[-1]},
        "e2e": {"value": rays / (e2e_ms * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": host[0].numel() * 4 * world,
                "d2h_bytes_per_step": 4 * world, "ms_per_step": e2e_ms / args.steps, "last_loss": last_loss},
        "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
        "kernels": per_kernel, "per_scale": scales, "clocks": clocks,
        "plane_mb": sum(p.numel() for g in model.field.grids for p in g) * 4 / 2**20,
        "regularizers": ("folded into the optimizer pass (kp_plane_reg_adam)" if trainer._reg_adam is not None else
                         "one sweep into the gradient bucket (kp_plane_reg_fused) + kp_adam_multi"),
        "grad_allreduce": ("none (single rank)" if world == 1 else "disabled" if args.allreduce == "none" else
                           (f"{trainer.allreduce_backend}: fused reduce-scatter + sharded Adam + all-gather kernel"
                            + (" (field group: only the marked 128-byte gradient lines are pulled)" if trainer._sparse else "") if trainer.sharded
                            else f"{trainer.allreduce_backend} all-reduce ({args.allreduce}) + replicated Adam")),
    })
    if rank == 0 and scales:
        leg["roofline"] = _roofline(name, per_kernel, step_bytes, scales, peaks, probe)
    trainer.close()
    if keep_model:
        return leg, model
    del trainer, model, resident, flush
    torch.cuda.empty_cache()
    return leg, None


def _device_pipeline_arm(trainer, timed_region, loss_host, dev, args):
    """cfg2 fed by the device-resident datamanager: broadcast-style cameras (19 on a ring x 22 frames = 418 images of
    960x540, NS/data/dataparsers/broadcaststyle_dataparser.py:196-232), 2.6 GB of fp32 images + 0.4 GB of fp16 weight maps
    (1 % non-zero, IST-like) in HBM, is_pixel_ratio 0.1 after iters_to_start_is."""
    import math
    import types

    from soccernerfs_b200.cameras.cameras import Cameras
    from soccernerfs_b200.data.datamanagers.dynamic_datamanager import DynamicDataManager, DynamicDataManagerConfig
    from soccernerfs_b200.data.pixel_samplers import DynamicBasedPixelSampler

    n_cams, n_frames, h, w = 19, 22, 540, 960
    b = n_cams * n_frames
    c2w, times, ids = [], [], []
    for c in range(n_cams):
        ang = 2 * math.pi * c / n_cams
        pos = torch.tensor([math.cos(ang), math.sin(ang), 0.35])
        z = pos / pos.norm()
        x = torch.linalg.cross(torch.tensor([0.0, 0.0, 1.0]), z)
        x = x / x.norm()
        m = torch.stack([x, torch.linalg.cross(z, x), z, pos], dim=-1)
        for f in range(n_frames):
            c2w.append(m), times.append(f / (n_frames - 1)), ids.append(float(c))
    cams = Cameras(torch.stack(c2w), 800.0, 800.0, w / 2, h / 2, w, h, times=torch.tensor(times)[:, None], ids=torch.tensor(ids)[:, None])
    gen = torch.Generator(device=dev).manual_seed(12)
    images = torch.rand((b, h, w, 3), device=dev, generator=gen)
    maps = torch.where(torch.rand((b, h, w), device=dev, generator=gen) < 0.01,
                       torch.rand((b, h, w), device=dev, generator=gen) * 0.8 + 0.15, 0.0).half()
    cfg = DynamicDataManagerConfig(train_num_rays_per_batch=RAYS_PER_RANK, use_importance_sampling=False)
    state = types.SimpleNamespace(iters_to_start_ist=0, is_pixel_ratio=0.1)
    out = {"workload": f"{b} cached images of {w}x{h} (fp32, {images.numel() * 4 / 2**30:.1f} GB) + fp16 weight maps in HBM; "
                       f"{RAYS_PER_RANK} rays/step, 10 % importance pixels"}
    rays = RAYS_PER_RANK * args.steps
    for key, prefetch in (("serial", False), ("prefetch", True)):
        dm = DynamicDataManager(cfg, cams, images, device=dev, prefetch=prefetch)
        dm.image_cache.batch["ist_weights"] = maps  # synthetic maps (kp_ist_map's own cost is a once-per-cache cost)
        dm.train_pixel_sampler = DynamicBasedPixelSampler(RAYS_PER_RANK, dataset=state)

        def step_pipeline(i, dm=dm):
            rb, batch = dm.next_train(i)
            res = trainer(rb, batch)
            loss_host[i].copy_(res["loss"], non_blocking=True)

        ms, _ = timed_region(step_pipeline)
        out[key] = {"value": rays / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms / args.steps}
        torch.cuda.current_stream().synchronize()
        del dm
    if not bool(torch.isfinite(loss_host).all()):
        raise RuntimeError("device pipeline arm: non-finite loss")
    out["value"] = max(out["serial"]["value"], out["prefetch"]["value"])
    out["h2d_bytes_per_step"] = 41 * 3 * 4  # the (image, k, first row) table of the importance sampler
    del images, maps
    torch.cuda.empty_cache()
    return out


def _roofline(name, per_kernel, step_bytes, scales, peaks, probe):
    """Roofline of the dominant field kernel (the larger of gather / scatter), built from the level of the hierarchy each
    scale actually lives in ("reported per scale because the level of the hierarchy differs per scale").

    Bytes the bound resources carry, per scale k (B_k = units x 3072 B, SURVEY's per-scale B_fwd):
      * L2 level  -- gather: every load, B_k, at the measured random-line L2 load rate (`l2_gather_gbs`);
                     scatter: the reduction payload, B_k, at the measured random-line L2 `red.v4` rate (`l2_red_gbs`) --
                     the product-rule re-gather of the same B_k is served by L1/L2 next to it and is not the bound;
      * HBM level -- the DISTINCT 128-byte lines the step's samples touch (counted in this run from a scatter into a
                     zeroed copy): gather 1x, scatter 3x (plane line + gradient line read-modify-write), at
                     MEASURED_PEAKS.json hbm_gbs.  (bench steps run with a flushed L2, so this applies to every scale.)
    floor_k = max(L2 time, HBM time); `peak` = bytes / sum(floor_k); `achieved` = bytes / measured kernel time, so
    frac = sum(floor_k) / t <= 1 (gathers may exceed it through L1 hits between neighbouring samples).  `achieved` counts
    B_fwd x units per launch; SURVEY's 2 x B_fwd figure for the backward is kept in `survey_algorithmic_gbs`."""
    top = max(("kp_hexplane_fwd", "kp_hexplane_bwd"), key=lambda k: per_kernel.get(k, {"ms_per_step": 0})["ms_per_step"])
    tag, l2key = ("gather", "l2_gather_gbs") if top == "kp_hexplane_fwd" else ("scatter", "l2_red_gbs")
    top_ms = per_kernel[top]["ms_per_step"] / max(1.0, per_kernel[top]["launches_per_step"])
    payload = step_bytes["kp_hexplane_fwd"]  # B_fwd x units (all scales)
    per_scale_bytes = payload / len(scales)
    t_floor, t_hbm_bound, levels = 0.0, 0.0, []
    for row in scales:
        t_l2 = per_scale_bytes / (probe[l2key] * 1e9)
        t_hbm = row[tag]["compulsory_dram_mb"] * 2**20 / (peaks["hbm_gbs"] * 1e9)
        t_floor += max(t_l2, t_hbm)
        if t_hbm > t_l2:
            t_hbm_bound += t_hbm
        levels.append({"scale": row["scale"], "floor_ms": 1e3 * max(t_l2, t_hbm), "bound": "hbm" if t_hbm > t_l2 else "l2"})
    achieved = payload / (top_ms * 1e-3) / 1e9
    peak = payload / t_floor / 1e9
    traffic, tsrc = None, None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic_r2.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic, tsrc = tj.get(name, {}).get(top), tj.get("source")
    finest = scales[-1][tag]
    return {"bound": "hbm" if t_hbm_bound > 0.5 * t_floor else "l2", "kernel": top, "achieved": achieved, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": tsrc,
            "peak_source": "sum over scales of max(B_k / kp_line_probe %s (%.0f GB/s, measured in this run), distinct-line DRAM bytes / "
                           "MEASURED_PEAKS.json hbm_gbs (%.0f GB/s)); distinct lines counted in this run" % (l2key, probe[l2key], peaks["hbm_gbs"]),
            "algorithmic_bytes_per_launch": payload, "kernel_ms": top_ms, "floor_ms": 1e3 * t_floor, "per_scale_floor": levels,
            "survey_algorithmic_gbs": step_bytes[top] / (top_ms * 1e-3) / 1e9,
            "finest_scale": {"scale": scales[-1]["scale"], "kernel_ms": finest["ms"], "level": finest["level"],
                             "l2_level_gbs": finest["l2_level_gbs"], "frac_of_l2_probe": finest["frac_of_l2_probe"],
                             "compulsory_dram_gbs": finest["compulsory_dram_gbs"], "frac_of_hbm_peak": finest["frac_of_hbm_peak"]}}


# ----------------------------------------------------------------------------------------------------------------------
# eval leg: BASELINE configs[4]
# ----------------------------------------------------------------------------------------------------------------------
def eval_leg(model, frames, warmup, rank, world, dev, ray_tile=4, graph=True):
    """Full-frame inference with the 32x field: 1920x1080 rays per frame in chunks of 32768, rays generated on the device
    and finished tiles copied to pinned host frames (engine/frame_renderer.py).  One "step" = one frame; with N ranks the
    tiles are shared round-robin (no collective)."""
    import torch.distributed as dist

    from soccernerfs_b200 import _lib
    from soccernerfs_b200.cameras.cameras import Cameras
    from soccernerfs_b200.engine.frame_renderer import FrameRenderer

    model.eval()
    h, w = 1080, 1920
    n_frames = warmup + frames
    ang = torch.linspace(0.0, 1.0, n_frames) * 0.6  # a short camera path on the broadcast ring, looking at the origin
    pos = torch.stack([torch.cos(ang), torch.sin(ang), torch.full_like(ang, 0.35)], dim=-1)
    fwd = -pos / pos.norm(dim=-1, keepdim=True)
    right = torch.cross(fwd, torch.tensor([0.0, 0.0, 1.0]).expand_as(fwd), dim=-1)
    right = right / right.norm(dim=-1, keepdim=True)
    up = torch.cross(right, fwd, dim=-1)
    c2w = torch.cat([torch.stack([right, up, -fwd], dim=-1), pos[..., None]], dim=-1)  # camera looks along -z
    cams = Cameras(c2w.to(dev), 1600.0, 1600.0, w / 2, h / 2, w, h, times=torch.linspace(0, 1, n_frames).to(dev))
    renderer = FrameRenderer(model, cams, rank=rank, world=world, ray_tile=ray_tile, use_cuda_graph=graph)
    for i in range(warmup):
        renderer.render(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    k0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    checksum = 0.0
    for i in range(frames):
        frame = renderer.render(warmup + i)  # ends with the copy stream's synchronize: the frame is on the host
        checksum += float(frame["rgb"][::97, ::89].sum())
    e1.record()
    torch.cuda.synchronize()
    launches = _lib.launch_count() - k0
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    rays = h * w * frames
    return {"metric": "eval rays/sec (full-frame inference, device ray generation -> pinned host frames)",
            "value": rays / (ms * 1e-3), "unit": "rays/s", "frames": frames, "warmup": warmup, "ms_per_frame": ms / frames,
            "scaling": "strong", "workload": "kplanes-32x(cfg3 field) full-frame inference (BASELINE config 5): 1920x1080 rays/frame, "
                                             "chunk 32768, 256/128/64 samples, rays generated on the device",
            "gather_ray_tile": ray_tile, "launch": "one CUDA graph replay per tile" if graph else "eager",
            "parallelism": f"tile-sharded x{world} (round-robin chunks, no collective)",
            "d2h_bytes_per_frame": h * w * (3 + 1 + 1) * 4, "checksum": checksum, "gpu_launches": launches}


# ----------------------------------------------------------------------------------------------------------------------
# cfg4 leg: BASELINE configs[3] -- the samplers / compositing / loss kernels NeRFPlayer-nerfacto shares with K-Planes
# ----------------------------------------------------------------------------------------------------------------------
def cfg4_leg(rank, world, dev, steps, warmup, rays=RAYS_PER_RANK):
    """The part of the reference's ``nerfplayer-nerfacto`` training step that runs on THIS repo's kernels: piecewise
    lin-disp initial sampler with one jitter per ray, two PDF resampling rounds (256 -> 96 -> 48 samples,
    NS/models/nerfacto.py:88-90,125), get_weights, RGB / accumulation / expected-depth renderers, interlevel and
    distortion losses -- forward and backward -- on the unbounded stadium shape (near 0.05, far 1000,
    NS/models/nerfplayer_nerfacto.py:67-70).  The NeRFPlayer hash-grid fields themselves are outside SURVEY.md 8 and are
    stood in for by an analytic density / colour (a few elementwise torch ops, counted in the step); rays are sharded
    over the ranks with no collective.  Replayed from a CUDA graph like the training legs."""
    import torch.distributed as dist

    from soccernerfs_b200 import _lib
    from soccernerfs_b200.cameras.rays import RayBundle
    from soccernerfs_b200.model_components.losses import distortion_loss, interlevel_loss
    from soccernerfs_b200.model_components.ray_samplers import ProposalNetworkSampler
    from soccernerfs_b200.model_components.renderers import AccumulationRenderer, DepthRenderer, RGBRenderer
    from soccernerfs_b200.model_components.scene_colliders import NearFarCollider

    gen = torch.Generator().manual_seed(4000 + rank)
    # cameras on a ring of radius 60 m around a 105 x 68 m pitch, looking at players near the ground
    ang = torch.rand(rays, generator=gen) * 6.2831853
    origins = torch.stack([60 * torch.cos(ang), 45 * torch.sin(ang), 12 + 6 * torch.rand(rays, generator=gen)], -1)
    target = torch.stack([(torch.rand(rays, generator=gen) - 0.5) * 105, (torch.rand(rays, generator=gen) - 0.5) * 68,
                          torch.rand(rays, generator=gen) * 2.0], -1)
    directions = torch.nn.functional.normalize(target - origins, dim=-1)
    static = {"origins": origins.to(dev), "directions": directions.to(dev), "times": torch.rand(rays, 1, generator=gen).to(dev),
              "image": torch.rand(rays, 3, generator=gen).to(dev)}
    theta = [torch.full((1,), v, device=dev, requires_grad=True) for v in (0.6, 0.8, 1.0, 1.0)]  # stand-in field parameters
    freq = [torch.tensor(f, device=dev) for f in ([0.11, 0.07, 0.9], [0.13, 0.09, 1.1], [0.17, 0.05, 1.3])]

    def density_fn(level):
        def fn(positions):
            return theta[level] * torch.exp(-(positions * freq[level]).sin().square().sum(-1, keepdim=True)) * 0.05
        return fn

    sampler = ProposalNetworkSampler(num_proposal_samples_per_ray=(256, 96), num_nerf_samples_per_ray=48,
                                     num_proposal_network_iterations=2, single_jitter=True, update_sched=lambda step: 0).to(dev)
    sampler.train()
    collider = NearFarCollider(near_plane=0.05, far_plane=1000.0)
    collider.train()
    rgb_r, acc_r, depth_r = RGBRenderer("random"), AccumulationRenderer(), DepthRenderer("expected")
    out = {}

    def body():
        for t in theta:
            t.grad = None
        rb = RayBundle(origins=static["origins"], directions=static["directions"], pixel_area=torch.ones(rays, 1, device=dev),
                       times=static["times"])
        rb = collider.set_nears_and_fars(rb)
        rs, w_list, rs_list = sampler(rb, density_fns=[density_fn(0), density_fn(1)])
        pos = rs.frustums.get_positions()
        w = rs.get_weights(density_fn(2)(pos))
        w_list, rs_list = w_list + [w], rs_list + [rs]
        rgb = rgb_r(rgb=torch.sigmoid(theta[3] * (pos * 0.21).sin()), weights=w)
        out["accumulation"], out["depth"] = acc_r(weights=w), depth_r(weights=w, ray_samples=rs)
        loss = torch.nn.functional.mse_loss(rgb, static["image"]) + interlevel_loss(w_list, rs_list) \
            + 1e-3 * distortion_loss(w_list, rs_list)
        loss.backward()
        out["loss"] = loss.detach()

    # warm-up and capture on ONE non-default stream: autograd pins the leaves' AccumulateGrad nodes to the stream they were
    # first used on, and a capture must not touch the legacy default stream
    work = torch.cuda.Stream()
    work.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(work):
        for _ in range(3):
            body()
        torch.cuda.synchronize()
        k0 = _lib.launch_count()
        body()
        launches_per_step = _lib.launch_count() - k0
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=work, capture_error_mode="thread_local"):
        body()
    for _ in range(max(warmup, 3)):
        graph.replay()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    loss = float(out["loss"])
    if not (loss == loss and all(t.grad is not None and bool(torch.isfinite(t.grad).all()) for t in theta)):
        raise RuntimeError(f"cfg4 leg: non-finite loss / gradient (loss {loss})")
    return {"metric": "shared sampler + compositing + loss kernels, train rays/sec (fwd+bwd)", "value": world * rays / (ms * 1e-3),
            "unit": "rays/s", "ms_per_step": ms, "steps": steps, "scaling": "weak",
            "workload": "nerfplayer-nerfacto sampling/compositing (BASELINE config 4): piecewise lin-disp sampler + 2 PDF rounds, single "
                        f"jitter, 256/96/48 samples, near 0.05 / far 1000, {rays} rays/rank; analytic stand-in for the hash-grid fields",
            "parallelism": f"ray-sharded x{world}, no collective", "kp_launches_per_step": launches_per_step,
            "last_loss": loss, "launch": "replayed from a CUDA graph"}


# ----------------------------------------------------------------------------------------------------------------------
# baselines
# ----------------------------------------------------------------------------------------------------------------------
def gpu_torch_baseline(dev, steps=5, warmup=3, rays=RAYS_PER_RANK):
    """The reference's own step restated as plain torch ops (grid_sample / searchsorted / cumsum / autograd /
    torch.optim.Adam; the oracle, which follows the reference line by line) with every tensor on THIS GPU: what the
    reference's code path costs on a B200 without our kernels (and with an fp32 torch MLP for tiny-cuda-nn).  A reported
    baseline, never a product path."""
    from oracle import kplanes_oracle as ko

    gen = torch.Generator().manual_seed(42)
    origins, directions, times, aabb = ko.synthetic_rays(rays, gen)
    mp = ko.make_model_params("cfg2", gen, aabb)
    image = torch.rand(rays, 3, generator=gen)
    rands = [ko.make_rand(rays, mp, gen) for _ in range(steps + warmup)]
    for t in mp.tensors():
        t.data = t.data.to(dev)
    for obj in [mp.field] + list(mp.proposals):
        obj.aabb = obj.aabb.to(dev)
    origins, directions, times, image = (x.to(dev) for x in (origins, directions, times, image))
    rands = [{k: v.to(dev) for k, v in r.items()} for r in rands]
    opt = torch.optim.Adam(mp.tensors(), lr=1e-2, eps=1e-12)
    torch.set_default_device(dev)  # the oracle's factory calls (linspace / zeros / ones) then land on the GPU too
    try:
        def step(i):
            opt.zero_grad(set_to_none=True)
            ko.train_step(mp, origins, directions, times, image, rands[i])
            opt.step()

        for i in range(warmup):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(warmup + i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
    finally:
        torch.set_default_device("cpu")
    return {"value": rays / ms * 1e3, "unit": "rays/s", "ms_per_step": ms, "steps": steps, "warmup": warmup,
            "what": "reference step as plain torch CUDA ops on the same B200 (cfg2, fp32, eager; torch grid_sample / searchsorted / "
                    "cumsum / autograd / torch.optim.Adam; fp32 torch MLP in place of tiny-cuda-nn)"}


def pixel_sampler_leg(dev, reps=20, host_reps=2):
    """DynamicBasedPixelSampler.sample_method (NS/data/pixel_samplers.py:340-426) on the `ns-train k-planes` preset's shape
    (method_configs.py:491-511: 4096 rays, is_pixel_ratio 0.1; 540x960 maps after downscale 2): the device path
    (kp_importance_pixels, 6 launches, CUDA events) and the reference's host loop (41 torch.multinomial calls over 518 400
    categories) on the same fp16 maps, half of them sparse (IST-like, 1 % non-zero) and half dense (ISG-like)."""
    import random
    import types

    from soccernerfs_b200 import _lib
    from soccernerfs_b200.data.pixel_samplers import DynamicBasedPixelSampler

    gen = torch.Generator().manual_seed(4)
    b, h, w, n = 410, 540, 960, RAYS_PER_RANK  # >= 409 cached images: the walk visits 41 of them, 10 pixels each
    distinct = 42  # 42 distinct maps tiled over the cache keep the host-side setup short
    maps = torch.zeros(distinct, h * w, dtype=torch.float16)
    idx = torch.randint(0, h * w, (distinct, 5000), generator=gen)
    maps.scatter_(1, idx, (torch.rand(distinct, 5000, generator=gen) * 0.8 + 0.15).half())
    maps[distinct // 2:] = (torch.rand(distinct - distinct // 2, h * w, generator=gen) * 0.3 + 1e-3).half()
    maps = maps[torch.arange(b) % distinct].view(b, h, w)
    state = types.SimpleNamespace(iters_to_start_ist=0, is_pixel_ratio=0.1)
    host = DynamicBasedPixelSampler(n, dataset=state, device_sampler=False)
    random.seed(5)
    host.sample_method(n, b, h, w, batch={"ist_weights": maps, "iter_steps": 1})
    t0 = time.perf_counter()
    for _ in range(host_reps):
        ref = host.sample_method(n, b, h, w, batch={"ist_weights": maps, "iter_steps": 1})
    host_ms = (time.perf_counter() - t0) * 1e3 / host_reps
    dmaps = maps.to(dev)
    sampler = DynamicBasedPixelSampler(n, dataset=state)
    batch = {"ist_weights": dmaps, "iter_steps": 1}
    for _ in range(3):
        out = sampler.sample_method(n, b, h, w, batch=batch)
    torch.cuda.synchronize()
    k0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        out = sampler.sample_method(n, b, h, w, batch=batch)
    e1.record()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - w0) * 1e3 / reps
    dev_ms = e0.elapsed_time(e1) / reps
    num_ist = int(0.1 * n)
    o = out[:num_ist]
    ok = bool((dmaps[o[:, 0], o[:, 1], o[:, 2]] > 0).all()) and out.shape == ref.shape
    if not ok:
        raise RuntimeError("pixel sampler leg: an importance pixel with zero weight")
    # the whole caller side of a step on the device: DynamicDataManager.next_train = pixel sampler + pixel gather from the
    # resident image cache (2.5 GB fp32) + ray generation, nothing over PCIe
    from soccernerfs_b200.cameras.cameras import Cameras
    from soccernerfs_b200.data.datamanagers.dynamic_datamanager import DynamicDataManager, DynamicDataManagerConfig

    c2w = torch.eye(4)[:3][None].repeat(b, 1, 1)
    c2w[:, :3, 3] = torch.randn(b, 3, generator=gen)
    cams = Cameras(c2w, 1200.0, 1200.0, w / 2, h / 2, w, h, times=torch.rand(b, 1, generator=gen), ids=torch.arange(b).float()[:, None])
    dm = DynamicDataManager(DynamicDataManagerConfig(train_num_rays_per_batch=n, use_importance_sampling=False), cams,
                            torch.rand((b, h, w, 3), device=dev), device=dev)
    dm.image_cache.batch["ist_weights"] = dmaps
    dm.train_pixel_sampler = sampler
    for _ in range(3):
        rb, col = dm.next_train(0)
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    for _ in range(reps):
        rb, col = dm.next_train(0)
    torch.cuda.synchronize()
    next_train_ms = (time.perf_counter() - w0) * 1e3 / reps
    if rb.origins.shape != (n, 3) or col["image"].shape != (n, 3) or not bool((col["ist_weights"][:num_ist] > 0).all()):
        raise RuntimeError("pixel sampler leg: next_train returned an inconsistent batch")
    return {"next_train_ms": next_train_ms,
            "next_train": "DynamicDataManager.next_train on the device-resident cache: sampler + pixel gather + ray generation, wall clock","workload": f"{n}-ray batch, {num_ist} importance pixels = 41 of {b} maps ({h}x{w} fp16) x 10 pixels + uniform remainder",
            "device_ms": dev_ms, "device_wall_ms": wall_ms, "host_reference_ms": host_ms, "host_cores": torch.get_num_threads(),
            "speedup": host_ms / wall_ms, "launches_per_batch": (_lib.launch_count() - k0) / reps,
            "note": "same distribution, own random stream (Philox exponential race + radix select); the host path stays for bit-exact indices"}


def cpu_baseline(rays_per_step: int, steps: int, warmup: int):
    """The oracle port of the reference's CPU torch path (oracle/kplanes_oracle.py) timed on the host cores:
    same cfg2 model and scene shape, full fwd+bwd+Adam steps on a bounded number of steps."""
    from oracle import kplanes_oracle as ko

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    gen = torch.Generator().manual_seed(42)
    origins, directions, times, aabb = ko.synthetic_rays(rays_per_step, gen)
    mp = ko.make_model_params("cfg2", gen, aabb)
    opt = torch.optim.Adam(mp.tensors(), lr=1e-2, eps=1e-12)
    image = torch.rand(rays_per_step, 3, generator=gen)
    t_total = 0.0
    for i in range(warmup + steps):
        rand = ko.make_rand(rays_per_step, mp, gen)
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        ko.train_step(mp, origins, directions, times, image, rand)
        opt.step()
        if i >= warmup:
            t_total += time.perf_counter() - t0
    return {"value": rays_per_step * steps / t_total, "unit": "rays/s", "cores": cores, "kind": "port",
            "sample": f"{steps} full training steps of {rays_per_step} rays (cfg2), after {warmup} warm-up; torch CPU fp32, "
                      f"{cores} threads", "ms_per_step": 1e3 * t_total / steps}


def _config(world: int, extra=None):
    cfg = {"workload": WORKLOAD, "global_batch_rays": RAYS_PER_RANK * world, "parallelism": f"ray-sharded dp{world}",
           "proposal_update": "every step", "l2": "flushed between timed iterations (160 MB write)"}
    cfg.update(extra or {})
    return cfg


def run_ours(args):
    rank, world, local = _dist_setup()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a CUDA device: the product path has no CPU fallback")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    from soccernerfs_b200 import _lib

    _lib.load()
    peaks, peak_kind = _peaks()
    legs = set(args.legs.split(","))
    probe = line_probe_peaks(dev) if rank == 0 else None
    cfg2, _ = train_leg("cfg2", args, rank, world, local, dev, peaks, probe)
    line = None
    if rank == 0:
        line = {
            "metric": "train rays/sec (fwd+bwd+optimizer)", "value": cfg2["value"], "unit": "rays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cfg2["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": _config(world, {"launch": "eager" if args.eager else "whole step replayed from a CUDA graph",
                                      "grad_allreduce": cfg2["grad_allreduce"]}),
            "e2e": cfg2["e2e"], "gpu_launches": cfg2["gpu_launches"], "roofline": cfg2.get("roofline"), "kernels": cfg2["kernels"],
            "per_scale": cfg2["per_scale"], "rank_ms_per_step": cfg2["rank_ms_per_step"], "clocks": cfg2["clocks"],
            "l2_peaks": probe, "hbm_peak": {"gbs": peaks["hbm_gbs"], "source": peak_kind},
        }
        for k in ("long_run", "reference_schedule", "dp_check", "device_pipeline"):
            if k in cfg2:
                line[k] = cfg2[k]
    if "cfg3" in legs:
        cfg3, model3 = train_leg("cfg3", args, rank, world, local, dev, peaks, probe, keep_model="eval" in legs)
        if rank == 0:
            cfg3.pop("clocks", None)
            line["cfg3"] = cfg3
        if model3 is not None:
            ev = eval_leg(model3, frames=args.eval_frames, warmup=1, rank=rank, world=world, dev=dev, ray_tile=args.eval_ray_tile, graph=not args.eval_eager)
            if rank == 0:
                line["eval"] = ev
            del model3
            torch.cuda.empty_cache()
    if "cfg4" in legs:
        c4 = cfg4_leg(rank, world, dev, steps=max(args.steps, 20), warmup=args.warmup)
        if rank == 0:
            line["cfg4"] = c4
    if rank != 0:
        return
    if world == 1 and "torch" in legs:
        line["gpu_torch_baseline"] = gpu_torch_baseline(dev)
        line["gpu_torch_baseline"]["speedup_of_value"] = line["value"] / line["gpu_torch_baseline"]["value"]
    if world == 1 and "sampler" in legs:
        try:  # an accessory of the line: its failure is reported in the line, the headline still goes out
            line["pixel_sampler"] = pixel_sampler_leg(dev)
        except Exception as e:  # noqa: BLE001
            line["pixel_sampler"] = {"error": repr(e)}
            print(f"[bench] pixel_sampler leg failed: {e!r}", file=sys.stderr, flush=True)
        torch.cuda.empty_cache()
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(rays_per_step=RAYS_PER_RANK, steps=2, warmup=1)
    for leg in (line, line.get("cfg3") or {}):
        dp = leg.get("dp_check")
        if dp is not None and not dp["ok"]:  # the line still goes out, with ok: false in it
            print(f"[bench] DATA-PARALLEL CHECK FAILED: {dp}", file=sys.stderr, flush=True)
    print(json.dumps(line), flush=True)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (the oracle port: the reference is
    Python and cannot travel to the GPU box), all host threads, same config / metric / unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # The full 4096-ray step (about 2.3 s on 16 host cores: the plane regularisers and Adam stream all parameters every
    # step, so a smaller ray sample would under-state the CPU path's rays/s).  Only if the requested run would exceed
    # ~5 minutes is the number of timed steps bounded; the step itself is never shrunk.
    rays = RAYS_PER_RANK
    steps = max(1, min(args.steps, 100))
    warmup = min(args.warmup, 5)
    res = cpu_baseline(rays_per_step=rays, steps=steps, warmup=warmup)
    line = {
        "impl": "reference",
        "metric": "train rays/sec (fwd+bwd+optimizer)",
        "value": res["value"],
        "unit": "rays/s",
        "n_gpus": args.gpus,
        "steps": steps,
        "warmup": warmup,
        "ms_per_step": res["ms_per_step"],
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": _config(world),
        "cpu_baseline": {**res, "sample": f"{steps} steps of {rays} rays after {warmup} warm-up (of {args.steps} requested)"},
        "e2e": {"value": res["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--legs", default="cfg2,cfg3,eval,cfg4,torch,sampler",
                    help="comma list of the extra objects of the JSON line: cfg3 (32x training leg), eval (full-frame inference, "
                         "needs cfg3), cfg4 (the sampler / compositing kernels on the nerfplayer-nerfacto shape), torch (reference step as torch CUDA ops, N=1).  cfg2 (the headline) always runs.")
    ap.add_argument("--eval-frames", type=int, default=2)
    ap.add_argument("--eval-eager", action="store_true", help="full-frame inference: launch every tile's kernels from Python")
    ap.add_argument("--eval-ray-tile", type=int, default=4,
                    help="full-frame inference: neighbouring rays per warp of the gather (KpPoints.ray_tile; 0 = samples of one ray)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-long-run", action="store_true")
    ap.add_argument("--no-ref-schedule", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="keep the regulariser / proposal branches on the main stream")
    ap.add_argument("--prop-overlap", choices=["auto", "on", "off"], default="auto",
                    help="proposal-network backward on a side stream (auto: on for 1 GPU, off under data parallelism)")
    ap.add_argument("--allreduce", choices=["overlap", "overlap-per-scale", "after-backward", "none"], default="overlap",
                    help="gradient all-reduce scheduling under data parallelism ('none' is a diagnostic: ranks do not sync)")
    ap.add_argument("--allreduce-backend", choices=["peer", "nccl"], default="peer",
                    help="peer: our in-place NVLink peer-memory kernel; nccl: torch.distributed.all_reduce")
    ap.add_argument("--shard-optimizer", choices=["auto", "on", "off"], default="auto",
                    help="N>1, peer backend: reduce-scatter + Adam on the owned shard + all-gather in one kernel (auto = on)")
    ap.add_argument("--reg-adam", choices=["auto", "on", "off"], default="auto",
                    help="(f1) plane regularisers folded into the optimizer's streaming pass (auto: when the planes are HBM-resident)")
    ap.add_argument("--sparse-exchange", choices=["auto", "on", "off"], default="auto",
                    help="N>1, sharded optimizer: pull only the gradient lines the scatter marked (auto: HBM-resident planes)")
    ap.add_argument("--no-branches", action="store_true",
                    help="diagnostic: keep the small loss / auxiliary-output kernels on the main stream (round-2 behaviour)")
    ap.add_argument("--no-stream-priority", action="store_true",
                    help="diagnostic: capture the step on a default-priority stream (round-2 behaviour)")
    ap.add_argument("--eager", action="store_true", help="launch kernels from Python each step instead of a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            # CUDA graphs that captured NCCL work keep the communicator busy at interpreter shutdown: finish all GPU
            # work, then leave without the (hanging) communicator teardown.
            torch.cuda.synchronize()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)


if __name__ == "__main__":
    main()
