"""Diagnostic (GPU box): how large must the fragility windows be for the full-batch step comparison, and are the remaining
gradient differences explained by decoder ReLU flips?   python tests/tools/diag_fullbatch.py cfg2 4096"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import kplanes_oracle as ko  # noqa: E402
from oracle.fragility import fragile_rays  # noqa: E402
from tests.conftest import rel_err  # noqa: E402
from tests.helpers import build_model, train_step_cuda  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    smooth = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    dev = "cuda"
    gen = torch.Generator().manual_seed(11)
    origins, directions, times, aabb = ko.synthetic_rays(n, gen)
    mp = ko.make_model_params(cfg, gen, aabb, smooth_res=smooth)
    image = torch.rand(n, 3, generator=gen)
    rand = ko.make_rand(n, mp, gen)
    # (1) decoder masks on the oracle's own features
    with torch.no_grad():
        nears, fars = ko.aabb_collider(origins, directions, mp.field.aabb, 0.0)
        out = ko.model_forward(mp, origins, directions, times, nears, fars, rand, anneal=0.6, training=True)
    feats = out["features"]
    w1 = mp.field.sigma_w[0]
    pre = feats.double() @ w1.double().t()
    mag = feats.double().abs() @ w1.double().abs().t()
    from soccernerfs_b200 import ops

    f_dev = feats.to(dev).requires_grad_(True)
    ws = [w.detach().to(dev) for w in list(mp.field.sigma_w) + list(mp.field.color_w)]
    s = feats.shape[0] // n
    if ops.decoder_fused_supported(feats.shape[1], w1.shape[0], 64):
        dirs = directions.to(dev) if mp.field.view_dependent else None
        o, dens, rgb = ops.decoder_fused(f_dev, dirs, s, *ws)
        # h1 is saved for backward: recompute masks from a plain fp32 matmul on the GPU instead (same inputs)
    h1_gpu = torch.relu(f_dev.detach() @ ws[0].t())
    mism = ((h1_gpu.cpu() > 0) != (pre > 0))
    ratio = (pre.abs() / mag)[mism]
    print(f"[{cfg}] sigma layer-1 mask mismatches (torch GPU fp32 matmul vs fp64): {int(mism.sum())} of {mism.numel()}, "
          f"max |pre|/sum|terms| among them {float(ratio.max()) if ratio.numel() else 0:.2e}")
    # (2) gradient errors vs window
    for win in ((2e-6,) if smooth else (2e-6, 1e-5, 5e-5)):
        fragile, stats = fragile_rays(mp, origins, directions, times, rand, anneal=0.6, relu_window=win, edge_window=win, median_window=win)
        keep = ~fragile
        o_, d_, t_, im_ = origins[keep], directions[keep], times[keep], image[keep]
        r_ = {k: v[keep] for k, v in rand.items()}
        model = build_model(cfg, mp, aabb, dev)
        out_g, ld, grads = train_step_cuda(model, o_, d_, t_, im_, r_, 0.6, dev)
        ref_out, ref_ld, ref_grads = ko.train_step(mp, o_, d_, t_, im_, r_, anneal=0.6)
        errs = [rel_err(a.cpu(), b) for a, b in zip(grads, ref_grads)]
        worst = max(range(len(errs)), key=lambda i: errs[i])
        a, b = grads[worst].cpu().double(), ref_grads[worst].double()
        diff = (a - b).abs()
        mx = b.abs().max()
        print(f"[{cfg}] window {win:.0e}: flagged {stats}, kept {int(keep.sum())}; grad rel errs max {max(errs):.2e} (tensor {worst}, shape "
              f"{tuple(b.shape)}), entries > 1e-4*max: {int((diff > 1e-4 * mx).sum())} of {b.numel()}; outputs rgb {rel_err(out_g['rgb'].cpu(), ref_out['rgb'].detach()):.2e}")
        print("    per-tensor:", " ".join(f"{e:.1e}" for e in errs))
        del model
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
