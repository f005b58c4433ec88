"""Two eager c3 training steps (for ncu launch lists): python scratch/step_c3.py [n_steps]"""
import sys; sys.path.insert(0, ".")
import torch, torch.nn.functional as F
import xequinet_b200 as xb
from oracle import xpainn_oracle as orc
import bench
n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = "cuda"
cfg = orc.CONFIG_DEFAULT
model = xb.resolve_model("xpainn", **cfg.model_kwargs())
model.load_state_dict(orc.synthetic_state_dict(cfg, 1234), strict=False)
model = model.to(dev).train()
params = list(model.parameters())
opt = torch.optim.AdamW(params, lr=5e-4, fused=True)
tr = xb.NeighborTransform(cfg.cutoff)
d0 = {k: v.to(dev) for k, v in bench.make_batch("c3", 256, 0).items()}
for it in range(n_steps):
    torch.cuda.synchronize()
    if it == n_steps - 1:
        torch.cuda.profiler.start()
    d = tr({k: d0[k] for k in ["pos", "atomic_numbers", "batch", "ptr", "target_energy", "target_forces"]})
    out = model(d, compute_forces=True)
    loss = bench.loss_fn(out, d, True)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    torch.cuda.synchronize()
    if it == n_steps - 1:
        torch.cuda.profiler.stop()
print("loss", float(loss))
