import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import kplanes_oracle as ko
from tests.conftest import rel_err
from tests.helpers import build_model, train_step_cuda
gen = torch.Generator().manual_seed(11)
n=256
origins, directions, times, aabb = ko.synthetic_rays(n, gen)
mp = ko.make_model_params("cfg2", gen, aabb)
image = torch.rand(n, 3, generator=gen); rand = ko.make_rand(n, mp, gen)
model = build_model("cfg2", mp, aabb, "cuda")
out, ld, grads = train_step_cuda(model, origins, directions, times, image, rand, 0.6, "cuda")
ref_out, ref_ld, ref_grads = ko.train_step(mp, origins, directions, times, image, rand, anneal=0.6)
for lvl in range(2):
    a=out["inds_list"][lvl].cpu(); b=ref_out["inds_list"][lvl]
    print("inds level",lvl,"mismatch frac",float((a!=b).float().mean()), "n", int((a!=b).sum()))
for i in range(3):
    rs=out["ray_samples_list"][i]
    bins=torch.cat([rs.spacing_starts[...,0], rs.spacing_ends[...,-1:,0]],-1).cpu()
    print("bins",i,float((bins-ref_out["samples_list"][i].spacing_bins).abs().max()), "weights", rel_err(out["weights_list"][i].cpu(), ref_out["weights_list"][i].detach()))
names=[]
for p in range(2): names += [f"prop{p}.plane{j}" for j in range(6)]+[f"prop{p}.w1",f"prop{p}.w2"]
names += [f"field.s{k}.p{j}" for k in range(4) for j in range(6)] + ["sig.w1","sig.w2","col.w3","col.w4","col.w5"]
for nme,a,b in zip(names,grads,ref_grads):
    e=rel_err(a.cpu(),b)
    if e>2e-4: print(nme, e, float(b.abs().max()))
print("---- L2-relative error and outlier counts")
for nme,a,b in zip(names,grads,ref_grads):
    a=a.cpu().double(); b=b.double()
    l2=float((a-b).norm()/b.norm().clamp_min(1e-30)); mx=float(b.abs().max())
    nbad=int(((a-b).abs()>1e-4*mx).sum())
    if nbad: print(f"{nme:14s} l2rel {l2:.2e}  entries>1e-4*max: {nbad} of {b.numel()}")
# ReLU-boundary census in the oracle: how many hidden pre-activations sit within fp32 rounding of zero?
feats = ref_out["features"].detach()
pre1 = feats @ mp.field.sigma_w[0].t()
print("sigma pre-activations |pre| < 2e-6*rowmax:", int((pre1.abs() < 2e-6 * pre1.abs().max()).sum()), "of", pre1.numel())
