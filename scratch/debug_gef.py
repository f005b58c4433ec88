import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch, torch.nn.functional as F
from helpers import cast_data, load_golden, grad_digest
from oracle import xpainn_oracle as orc
import xequinet_b200 as xb
from xequinet_b200 import ops
from xequinet_b200.graph import graph_from_edge_index
DEV='cuda'
z, cfg, data = load_golden('mol_small')
model = xb.resolve_model("xpainn", **cfg.model_kwargs()); model.load_state_dict(orc.synthetic_state_dict(cfg, int(z["sd_seed"])), strict=False); model=model.to(DEV).train()
d = {k:(v.to(DEV)) for k,v in cast_data(data, torch.float32).items()}
out = model(d, compute_forces=True)
tE = torch.from_numpy(z["f64:target_energy"]).float().to(DEV); tF = torch.from_numpy(z["f64:target_forces"]).float().to(DEV)
loss = F.smooth_l1_loss(out["energy"], tE) + 100.0 * F.smooth_l1_loss(out["forces"], tF)
loss.backward()
for k,p in model.named_parameters():
    key=f"f64:gEF:sum:{k}"
    if key not in z.files: continue
    s,smp = grad_digest(p.grad)
    print(f"{k:45s} l2 got {s[1]:.5e} ref {z[key][1]:.5e} ratio {s[1]/z[key][1]:.4f}")

# isolated autograd-level double backward of the edge op
cfg = orc.CONFIG_DEFAULT
dd = orc.make_molecule_batch(6,(6,12),seed=2)
ei = dd['edge_index']; N = dd['pos'].shape[0]
g = torch.Generator().manual_seed(4); rnd=lambda *s: torch.randn(*s,generator=g,dtype=torch.float64)
t = dict(pos=dd["pos"].double(), s=rnd(N, cfg.H_msg), v=rnd(N, cfg.D), x=rnd(N, cfg.node_dim), V=rnd(N, cfg.D), W=0.3*rnd(cfg.H_msg,20), b=0.3*rnd(cfg.H_msg), freq=torch.pi*torch.arange(1,21,dtype=torch.float64)/5+0.1*rnd(20))
wx, wV, wp = rnd(N,cfg.node_dim), rnd(N,cfg.D), rnd(N,3)
def run(fn, tt, cm):
    req = {k: tt[k].clone().requires_grad_(True) for k in tt}
    xo, Vo = fn(req)
    E = (xo*wx_.to(xo)).sum() + (Vo*(wV_cm if cm else wV_).to(Vo)).sum()
    (gp,) = torch.autograd.grad(E, req['pos'], create_graph=True)
    L = ((gp*wp_.to(gp)).sum())**2 + E
    L.backward()
    return {k: req[k].grad for k in req}
wx_, wV_, wp_ = wx, wV, wp
wV_cm = orc.to_cm(wV, cfg)
ref = run(lambda r: orc.edge_message(r['x'], r['V'], r['s'], r['v'], r['pos'], r['W'], r['b'], r['freq'], ei, cfg), t, False)
graph = graph_from_edge_index(ei.to(DEV), N, dd['ptr'].numel()-1)
dims = ops.Dims(cfg.node_dim, *cfg.muls, 20, 5.0)
t32 = {k:(orc.to_cm(v,cfg) if k in ('v','V') else v).float().to(DEV) for k,v in t.items()}
wx_, wp_ = wx.float().to(DEV), wp.float().to(DEV); wV_cm = wV_cm.float().to(DEV)
got = run(lambda r: ops.edge_message(r['x'], r['V'], r['s'], r['v'], r['pos'], r['W'], r['b'], r['freq'], graph, dims), t32, True)
for k in ref:
    a = got[k].cpu().double(); 
    if k in ('v','V'): a = orc.from_cm(a, cfg)
    print(k, float((a-ref[k]).abs().max()/ref[k].abs().max()))
