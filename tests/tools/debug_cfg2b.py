import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import kplanes_oracle as ko
from tests.conftest import rel_err
from tests.helpers import build_model, ray_bundle, rand_queue, rand_list
from soccernerfs_b200 import ops
from soccernerfs_b200.fields.base_field import FieldHeadNames as FH
gen = torch.Generator().manual_seed(11)
n=256
origins, directions, times, aabb = ko.synthetic_rays(n, gen)
mp = ko.make_model_params("cfg2", gen, aabb)
image = torch.rand(n, 3, generator=gen); rand = ko.make_rand(n, mp, gen)
model = build_model("cfg2", mp, aabb, "cuda"); model.train(); model.proposal_sampler.set_anneal(0.6)
cap = {}
orig_fwd = model.field.forward
def fwd(rs, **k):
    o = orig_fwd(rs, **k); cap["density"]=o[FH.DENSITY]; cap["rgb"]=o[FH.RGB]
    o[FH.DENSITY].retain_grad(); o[FH.RGB].retain_grad(); return o
model.field.forward = fwd
orig_sig = ops.sigma_net
def sig(feats, w1, w2):
    feats.retain_grad(); cap["feats"]=feats; return orig_sig(feats, w1, w2)
ops.sigma_net = sig
with rand_queue(rand_list(rand), "cuda"):
    out = model(ray_bundle(origins, directions, times, "cuda"))
ld = model.get_loss_dict(out, {"image": image.to("cuda")}, {})
sum(ld.values()).backward()
# oracle with retained grads
params = mp.tensors()
for p in params: p.requires_grad_(True); p.grad=None
nears, fars = ko.aabb_collider(origins, directions, mp.field.aabb, 0.0)
ro = ko.model_forward(mp, origins, directions, times, nears, fars, rand, anneal=0.6, training=True)
for k in ("density","rgb_samples","features"): ro[k].retain_grad()
rld = ko.model_loss_dict(mp, ro, image); sum(rld.values()).backward()
for a,b,name in ((cap["density"].grad, ro["density"].grad, "d_density"), (cap["rgb"].grad, ro["rgb_samples"].grad, "d_rgb"), (cap["feats"].grad, ro["features"].grad, "d_feats")):
    a=a.cpu().double().reshape(b.shape); b=b.double()
    diff=(a-b).abs(); mx=b.abs().max()
    bad=(diff>1e-4*mx)
    print(name, "rel", float(diff.max()/mx), "nbad", int(bad.sum()), "of", b.numel())
    if bad.any():
        idx=bad.nonzero()[:6]
        for ix in idx:
            t=tuple(ix.tolist()); print("   at",t,"gpu",float(a[t]),"ref",float(b[t]))
# forward values at those samples
d_g=cap["density"].detach().cpu().double().reshape(ro["density"].shape); d_r=ro["density"].detach().double()
print("density rel", float((d_g-d_r).abs().max()/d_r.abs().max()), "max density", float(d_r.max()))
w_g=out["weights_list"][2].detach().cpu().double(); w_r=ro["weights_list"][2].detach().double()
print("weights abs diff max", float((w_g-w_r).abs().max()))
a=cap["feats"].grad.cpu().double(); b=ro["features"].grad.double()
diff=(a-b).abs(); mx=b.abs().max()
rows=(diff>1e-4*mx).any(dim=1).nonzero().flatten().tolist()
print("bad rows:", rows)
for r in rows[:10]:
    cols=(diff[r]>1e-4*mx).nonzero().flatten().tolist()
    print(r, "tile",r//128,"row-in-tile",r%128,"ncols",len(cols),"cols head",cols[:8], "rowmax ref", float(b[r].abs().max()))
# hidden pre-activations of those rows in the oracle
feats=ro["features"].detach(); pre1=feats@mp.field.sigma_w[0].t()
o=torch.relu(pre1)@mp.field.sigma_w[1].t()
dirs=directions[:,None,:].expand(n,48,3).reshape(-1,3)
cin=torch.cat([ko.sh4((dirs+1)/2), o[:,:15]],-1)
pre2=cin@mp.field.color_w[0].t(); pre3=torch.relu(pre2)@mp.field.color_w[1].t()
for r in rows[:10]:
    print(r, "min|pre1|",float(pre1[r].abs().min()),"min|pre2|",float(pre2[r].abs().min()),"min|pre3|",float(pre3[r].abs().min()))
print("typical min|pre| over rows: ", float(pre1.abs().min(dim=1).values.median()), float(pre2.abs().min(dim=1).values.median()), float(pre3.abs().min(dim=1).values.median()))
