"""One-off measurement ("also time the reference's torch-CUDA grid_sample path on 1 B200 ... as the
like-for-like 'before' number"): the oracle restatement of the reference's step (torch ops only: grid_sample,
searchsorted, cumsum, autograd, torch.optim.Adam) with every tensor on the GPU.  This is what the reference's own code
path costs on a B200 without tiny-cuda-nn's fused MLP (fp32 torch MLP instead) -- a reported baseline, never a product
path (it lives under tests/ because it executes the oracle).

    python tests/tools/torch_cuda_baseline.py [steps]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import kplanes_oracle as ko  # noqa: E402


def main() -> None:
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    rays = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    gen = torch.Generator().manual_seed(42)
    origins, directions, times, aabb = ko.synthetic_rays(rays, gen)
    mp = ko.make_model_params("cfg2", gen, aabb)
    image = torch.rand(rays, 3, generator=gen)
    rands = [ko.make_rand(rays, mp, gen) for _ in range(steps + 3)]
    dev = torch.device(os.environ.get("KP_BASELINE_DEVICE", "cuda"))  # "cpu" only to dry-run the script
    for t in mp.tensors():
        t.data = t.data.to(dev)
    for obj in [mp.field] + list(mp.proposals):
        obj.aabb = obj.aabb.to(dev)
    origins, directions, times, image = (x.to(dev) for x in (origins, directions, times, image))
    rands = [{k: v.to(dev) for k, v in r.items()} for r in rands]
    opt = torch.optim.Adam(mp.tensors(), lr=1e-2, eps=1e-12)
    torch.set_default_device(dev)  # the oracle's factory calls (linspace / zeros / ones) then land on the GPU too

    def step(i):
        opt.zero_grad(set_to_none=True)
        _, ld, _ = ko.train_step(mp, origins, directions, times, image, rands[i])
        opt.step()
        return ld

    for i in range(3):
        ld = step(i)
    import time

    if dev.type == "cuda":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        ld = step(3 + i)
    if dev.type == "cuda":
        torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / steps  # wall clock around synchronised ends: eager, launch-bound
    print(json.dumps({"what": "reference step restated in torch ops, all tensors on one B200 (fp32, eager)", "rays_per_step": rays,
                      "ms_per_step": ms, "rays_per_s": rays / ms * 1e3, "steps": steps,
                      "loss": float(sum(v for v in ld.values()))}))


if __name__ == "__main__":
    main()
