"""Kernel timeline of graph-replayed training steps (CUPTI through torch.profiler, nsys is absent): start, duration and
stream of every kernel of a few replays -> gpurun_out/<tag>_timeline.csv.  Shows what is on the critical path and the gaps
between dependent launches; not a bench number (profiler overhead)."""
import argparse, csv, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2")
ap.add_argument("--tag", default="tl")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--no-branches", action="store_true")
ap.add_argument("--no-stream-priority", action="store_true")
args = ap.parse_args()
from soccernerfs_b200.engine.trainer import TrainStep
from torch.profiler import profile, ProfilerActivity

rank, world, local = bench._dist_setup()  # (under torchrun: the data-parallel step; rank 0 writes its own timeline)
dev = torch.device("cuda", local)
model = bench.build_model(args.workload, dev)
torch.manual_seed(42 + rank)
model.proposal_sampler.update_sched = lambda step: 0
trainer = TrainStep(model, use_cuda_graph=True, data_parallel=world > 1, branch_small_kernels=not args.no_branches,
                    prioritize_main_stream=not args.no_stream_priority)
host = bench._make_batches(8, bench.RAYS_PER_RANK, seed=1000 + rank)
res = [h.to(dev) for h in host]
for i in range(8):
    trainer(*bench._bundle(res[i % 8]))
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(args.steps):
        trainer(*bench._bundle(res[i % 8]))
        torch.cuda.synchronize()
if world > 1:
    import torch.distributed as dist
    dist.barrier()
if rank != 0:
    torch.cuda.synchronize()
    os._exit(0)
rows = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        rows.append((e.time_range.start, e.time_range.end - e.time_range.start, getattr(e, "device_index", 0), e.name[:90]))
rows.sort()
t0 = rows[0][0] if rows else 0
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"{args.tag}_timeline.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["start_us", "dur_us", "dev", "name"])
    for s, d, dv, n in rows:
        w.writerow([f"{s - t0:.2f}", f"{d:.2f}", dv, n])
print("kernels:", len(rows))
try:
    prof.export_chrome_trace(os.path.join(ROOT, "gpurun_out", f"{args.tag}_trace.json"))
except Exception as ex:  # noqa: BLE001
    print("trace export failed:", ex)
if world > 1:
    sys.stdout.flush()
    os._exit(0)  # (captured graphs keep the peer arenas / communicator busy at interpreter shutdown, as in bench.py)
