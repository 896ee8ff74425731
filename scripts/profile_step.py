#!/usr/bin/env python
"""Run a few EAGER training steps of a bench workload between cudaProfilerStart/Stop, for ncu:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches_cfg3.csv python scripts/profile_step.py --workload cfg3 --steps 1
    ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:hexplane \
        -o gpurun_out/prof_cfg3_field python scripts/profile_step.py --workload cfg3 --steps 1

Side-stream overlap is off so that the launch list is one serial stream (shares of the step, not absolutes).
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2", choices=list(bench.WORKLOADS))
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--overlap", action="store_true")
    args = ap.parse_args()
    from soccernerfs_b200.engine.trainer import TrainStep

    dev = torch.device("cuda", 0)
    model = bench.build_model(args.workload, dev)
    model.proposal_sampler.update_sched = lambda step: 0
    trainer = TrainStep(model, use_cuda_graph=False, overlap_branches=args.overlap)
    host = bench._make_batches(args.warmup + args.steps, bench.RAYS_PER_RANK, seed=1000)
    flush = torch.empty(160 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    for i in range(args.warmup):
        trainer(*bench._bundle(host[i].to(dev)))
    torch.cuda.synchronize()
    packed = [host[args.warmup + i].to(dev) for i in range(args.steps)]
    flush.zero_()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for i in range(args.steps):
        trainer(*bench._bundle(packed[i]))
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
