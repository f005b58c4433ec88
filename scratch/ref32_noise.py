import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch, torch.nn.functional as F
from helpers import cast_data, load_golden, grad_digest, embed_table
from oracle import xpainn_oracle as orc
z, cfg, data = load_golden('mol_small')
def grads(dtype):
    sd = {k: v.requires_grad_(True) for k,v in orc.synthetic_state_dict(cfg, int(z["sd_seed"]), dtype).items()}
    out = orc.xpainn_energy_forces(sd, embed_table().to(dtype), cast_data(data, dtype), cfg, create_graph=True)
    tE = torch.from_numpy(z["f64:target_energy"]).to(dtype); tF = torch.from_numpy(z["f64:target_forces"]).to(dtype)
    loss = F.smooth_l1_loss(out["energy"], tE) + 100.0*F.smooth_l1_loss(out["forces"], tF)
    loss.backward()
    return {k: v.grad.double() for k,v in sd.items() if v.grad is not None}
g64, g32 = grads(torch.float64), grads(torch.float32)
for k in g64:
    n64 = g64[k].norm().item()
    print(f"{k:45s} rel l2 err {(g32[k]-g64[k]).norm().item()/max(n64,1e-30):.3e}  norm ratio {g32[k].norm().item()/max(n64,1e-30):.4f}")
