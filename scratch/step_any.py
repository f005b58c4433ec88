"""Eager steps of one bench workload (for ncu launch lists): python scratch/step_any.py c5 [n_steps]"""
import sys; sys.path.insert(0, ".")
import torch
import xequinet_b200 as xb
from oracle import xpainn_oracle as orc
import bench
wl = sys.argv[1]; n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = bench.WORKLOADS[wl]; cfg = w["cfg"]; dev = "cuda"
model = xb.resolve_model("xpainn", **cfg.model_kwargs())
model.load_state_dict(orc.synthetic_state_dict(cfg, 1234), strict=False)
model = model.to(dev).train(w["train"])
params = list(model.parameters())
if not w["train"]:
    for p in params: p.requires_grad_(False)
opt = torch.optim.AdamW(params, lr=5e-4, fused=True) if w["train"] else None
tr = xb.NeighborTransform(cfg.cutoff)
d0 = {k: v.to(dev) for k, v in bench.make_batch(wl, w["n_mol"], 0).items()}
keys = ["pos", "atomic_numbers", "batch", "ptr"] + (["target_energy", "target_forces"] if w["train"] else []) + (["cell", "pbc"] if wl == "c5" else [])
for it in range(n_steps):
    torch.cuda.synchronize()
    if it == n_steps - 1: torch.cuda.profiler.start()
    d = tr({k: d0[k] for k in keys})
    out = model(d, compute_forces=w["forces"])
    if w["train"]:
        loss = bench.loss_fn(out, d, w["forces"]); opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
    torch.cuda.synchronize()
    if it == n_steps - 1: torch.cuda.profiler.stop()
print("E", float(out["energy"].sum()))
