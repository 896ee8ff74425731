"""world_size-2 gloo tests (CPU) of the ray-sharding / gradient-bucket logic used by the N>1 training step."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from soccernerfs_b200.distributed import GradBucket, round_robin_chunks, shard_slice


def test_shard_slice_and_chunks_cover_everything():
    for n, world in ((4096, 2), (4097, 4), (10, 8), (7, 3)):
        spans = [shard_slice(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1
    n_rays, chunk, world = 1920 * 1080, 32768, 8
    seen = sorted(c for r in range(world) for c in round_robin_chunks(n_rays, chunk, r, world))
    assert seen[0][0] == 0 and seen[-1][1] == n_rays and all(a[1] == b[0] for a, b in zip(seen, seen[1:]))
    counts = [len(round_robin_chunks(n_rays, chunk, r, world)) for r in range(world)]
    assert max(counts) - min(counts) <= 1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    # a "model": one channel-last plane + one matrix; a global batch sharded over ranks; loss = local mean
    plane = torch.nn.Parameter(torch.randn(1, 6, 5, 4).permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2))
    mat = torch.nn.Parameter(torch.randn(3, 4))
    x = torch.randn(16, 4)
    b, e = shard_slice(16, rank, world)
    bucket = GradBucket([plane, mat])
    bucket.attach_zeroed()
    loss = ((x[b:e] @ mat.t()) ** 2).mean() + (plane**2).mean()  # second term identical on all ranks (a regulariser)
    loss.backward()
    assert plane.grad.data_ptr() == bucket.views[0].data_ptr()  # autograd accumulated in place into the bucket
    bucket.all_reduce()
    g_plane, g_mat = plane.grad / world, mat.grad / world  # 1/world is folded into the Adam kernel in the product
    # single-process reference on the global batch
    p2, m2 = plane.detach().clone().requires_grad_(True), mat.detach().clone().requires_grad_(True)
    (((x @ m2.t()) ** 2).mean() + (p2**2).mean()).backward()
    ok = torch.allclose(g_plane, p2.grad, atol=1e-6) and torch.allclose(g_mat, m2.grad, atol=1e-6)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_grad_bucket_allreduce_equals_global_batch_gloo():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, 29531 + os.getpid() % 500, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def _span_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ps = [torch.nn.Parameter(torch.zeros(n)) for n in (8, 12, 4)]
    bucket = GradBucket(ps)
    bucket.attach_zeroed()
    bucket.flat.copy_(torch.arange(24, dtype=torch.float32) * (rank + 1))
    # reduce the middle parameter first (what the per-scale schedule does with the finest scale), then the rest
    a, b = bucket.span_of([ps[1]])
    bucket.all_reduce(span=(a, b))
    mid_only = bucket.flat.clone()
    bucket.all_reduce(span=(0, a))
    bucket.all_reduce(span=(b, 24))
    bucket.all_reduce(span=(5, 5))  # empty range: no collective issued
    factor = sum(r + 1 for r in range(world))
    base = torch.arange(24, dtype=torch.float32)
    ok = torch.equal(bucket.flat, base * factor)
    ok = ok and torch.equal(mid_only[a:b], base[a:b] * factor) and torch.equal(mid_only[:a], base[:a] * (rank + 1))
    ok = ok and ps[1].grad.data_ptr() == bucket.flat[a:].data_ptr()  # the views alias the flat buffer
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_grad_bucket_span_allreduce_gloo():
    """GradBucket.all_reduce(span=...): sub-ranges of the bucket can be reduced separately and in any order."""
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_span_worker, args=(world, 30100 + os.getpid() % 500, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_plane_write_ranges_partition_the_bucket():
    """Sparse exchange host logic: over all ranks the per-plane write ranges tile every plane of the sharded bucket exactly
    once (float4 granularity, shard boundaries inside planes), and planes of other buckets are always written in full."""
    from soccernerfs_b200.distributed import plane_write_ranges

    sizes = [64 * 64 * 32, 50 * 64 * 32, 128 * 128 * 32, 1024, 256 * 8 * 150]
    offsets, off = [], 0
    for i, n in enumerate(sizes):
        offsets.append(None if i == 4 else off)  # the last plane lives in another bucket
        off += 0 if i == 4 else n
    total = off + 8192  # MLP weights after the planes
    count = (total + 63) // 64 * 64
    for world in (2, 3, 8):
        n4 = count // 4
        covered = [torch.zeros(n // 4, dtype=torch.int32) for n in sizes]
        for rank in range(world):
            lo, hi = n4 * rank // world * 4, n4 * (rank + 1) // world * 4
            for k, (a, b) in enumerate(plane_write_ranges(offsets, sizes, lo, hi)):
                assert 0 <= a <= b <= sizes[k] // 4
                covered[k][a:b] += 1
        for k in range(4):
            assert bool((covered[k] == 1).all()), (world, k)
        assert bool((covered[4] == world).all())
