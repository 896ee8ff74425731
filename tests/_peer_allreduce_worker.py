"""Worker for tests/test_gpu_multi.py (launched with torch.distributed.run, one process per GPU)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main() -> None:
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from soccernerfs_b200.distributed import GradBucket, PeerArena

    dev = torch.device("cuda")
    n = 3_000_003  # not a multiple of 4: exercises the scalar tail
    arena = PeerArena(n + 1024, blocks=int(os.environ.get("KP_PEER_BLOCKS", "64")))
    buf, off = arena.take(n)
    gen = torch.Generator(device="cuda").manual_seed(100 + rank)
    for it in range(3):  # repeated calls: flags are monotonic
        x = torch.randn(n, device=dev, generator=gen)
        ref = x.clone()
        dist.all_reduce(ref)
        buf.copy_(x)
        arena.all_reduce(off, n)
        torch.cuda.synchronize()
        # NCCL's ring adds in a different order: equal up to fp32 rounding; all ranks bit-identical among themselves
        assert torch.allclose(buf, ref, rtol=1e-5, atol=1e-5), (it, float((buf - ref).abs().max()))
        mine = buf.clone()
        other = mine.clone()
        dist.broadcast(other, src=0)
        assert torch.equal(mine, other), "ranks disagree bitwise"
    # a sub-range leaves the rest untouched
    x = torch.randn(n, device=dev, generator=gen)
    buf.copy_(x)
    arena.all_reduce(off + 1024, 4096)
    ref = x.clone()
    part = x[1024:1024 + 4096].clone()
    dist.all_reduce(part)
    ref[1024:1024 + 4096] = part
    torch.cuda.synchronize()
    assert torch.allclose(buf, ref, rtol=1e-5, atol=1e-5)
    # CUDA-graph capture + replays
    static = torch.randn(n, device=dev, generator=gen)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        buf.copy_(static)
        arena.all_reduce(off, n)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        buf.copy_(static)
        arena.all_reduce(off, n)
    ref = static.clone()
    dist.all_reduce(ref)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert torch.allclose(buf, ref, rtol=1e-5, atol=1e-5)
    assert arena.error_word() == 0

    if os.environ.get("KP_PEER_TIMING"):
        big_n = 38_000_000  # ~ the cfg2 field bucket (152 MB)
        big = PeerArena(big_n, blocks=int(os.environ.get("KP_PEER_BLOCKS", "64")))
        t, o = big.take(big_n)
        y = torch.zeros(big_n, device=dev)
        cases = [(f"peer blocks={b}", b) for b in (16, 32, 64, 128, 148)] + [("nccl", 0)]
        for name, b in cases:
            if b:
                big.blocks = b
            fn = (lambda: big.all_reduce(o, big_n)) if b else (lambda: dist.all_reduce(y))
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            if rank == 0:
                print(f"ALLREDUCE_TIMING {name} world={world} {big_n * 4 / 1e6:.0f} MB: {ms:.3f} ms "
                      f"({big_n * 4 / ms / 1e6:.0f} GB/s algorithmic)")
        assert big.error_word() == 0

    if os.environ.get("KP_PEER_TIMING") == "only":
        if rank == 0:
            print("PEER_ALLREDUCE_OK")
        sys.stdout.flush()
        os._exit(0)

    # data-parallel TrainStep: peer-memory all-reduce vs NCCL -> same parameters
    from soccernerfs_b200.engine.trainer import TrainStep
    from tests.conftest import load_golden
    from tests.helpers import build_model, ray_bundle
    from tests.test_oracle_golden import load_tiny_model

    gold = load_golden("model_tiny")
    mp = load_tiny_model(gold)
    runs = {}
    # "peer" = in-place all-reduce + replicated Adam; "peer-sharded" = reduce-scatter + Adam on the shard + all-gather in
    # one kernel (the default under the peer backend)
    # "peer-sparse" = the sharded step with the sparse gradient exchange (scatter marks touched lines, only marked lines are
    # pulled, owner-only regulariser gradient, in-kernel clean-up instead of the bucket memset)
    for backend, graph in (("nccl", False), ("nccl-again", False), ("nccl", True), ("peer", False), ("peer", True),
                           ("peer-sharded", False), ("peer-sharded", True), ("peer-sparse", False), ("peer-sparse", True)):
        model = build_model("tiny", mp, gold["aabb"], "cuda")
        model.config.background_color_train = "black"
        model.proposal_sampler.initial_sampler.train_stratified = False
        model.proposal_sampler.pdf_sampler.train_stratified = False
        step = TrainStep(model, max_steps=100, warm_up_end=4, data_parallel=True, use_cuda_graph=graph,
                         allreduce_backend=backend.split("-")[0], allreduce_mode="overlap-per-scale" if graph else "overlap",
                         shard_optimizer=backend in ("peer-sharded", "peer-sparse"),
                         sparse_grad_exchange=backend == "peer-sparse")
        assert step.allreduce_backend == backend.split("-")[0]
        assert bool(step.sharded) == (backend in ("peer-sharded", "peer-sparse"))
        assert step._sparse == (backend == "peer-sparse")
        n_rays = gold["origins"].shape[0]
        lo, hi = rank * n_rays // world, (rank + 1) * n_rays // world  # every rank trains on its own rays
        losses = []
        for _ in range(6):
            rb = ray_bundle(gold["origins"][lo:hi], gold["directions"][lo:hi], gold["times"][lo:hi], "cuda")
            out = step(rb, {"image": gold["image"][lo:hi].to("cuda")})
            losses.append(float(out["loss"]))
        torch.cuda.synchronize()
        runs[(backend, graph)] = (losses, [p.detach().clone() for p in model.parameters()])
        if step.arena is not None:
            assert step.arena.error_word() == 0
        for p in runs[(backend, graph)][1]:  # replicas stay bit-identical
            q = p.clone()
            dist.broadcast(q, src=0)
            assert torch.equal(p, q), "replicas diverged"
        if backend == "peer-sharded" and not graph:  # checkpoint round trip of the sharded moments (a collective)
            sd = step.state_dict()
            st = sd["optimizers"]["fields"]["state"]
            assert len(st) > 0 and all(v["exp_avg"].shape == v["exp_avg_sq"].shape for v in st.values())
            before = {k: (g.exp_avg.clone(), g.exp_avg_sq.clone()) for k, g in step.sharded.items()}
            for g in step.sharded.values():
                g.exp_avg.zero_()
                g.exp_avg_sq.zero_()
            step.load_state_dict(sd)
            for k, g in step.sharded.items():
                assert torch.equal(g.exp_avg, before[k][0]) and torch.equal(g.exp_avg_sq, before[k][1]), k
        step.close()

    def param_err(a_params, b_params):
        return max(float((a - b).norm() / (a.norm() + 1e-12)) for a, b in zip(a_params, b_params) if a.numel())

    base_losses, base = runs[("nccl", False)]
    # run-to-run noise of the SAME configuration (atomics order, amplified by Adam's normalisation) sets the scale
    floor = param_err(base, runs[("nccl-again", False)][1])
    for key, (losses, params) in runs.items():
        err = param_err(base, params)
        if rank == 0:
            print(f"DP_STEP {key}: param err vs nccl-eager {err:.3e} (noise floor {floor:.3e}), losses {losses[-2:]}")
        for a, b in zip(base_losses, losses):
            assert abs(a - b) <= 2e-3 * abs(a), (key, base_losses, losses)
        assert err < 1e-3, (key, err, floor)
    dist.barrier()
    if rank == 0:
        print("PEER_ALLREDUCE_OK")
    sys.stdout.flush()
    os._exit(0)  # graphs captured collectives: skip the teardown


if __name__ == "__main__":
    main()
