import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch, torch.nn.functional as F
from helpers import cast_data, load_golden, embed_table
from oracle import xpainn_oracle as orc
def grads(cfg, seed, data, tE, tF, dtype):
    sd = {k: v.requires_grad_(True) for k,v in orc.synthetic_state_dict(cfg, seed, dtype).items()}
    out = orc.xpainn_energy_forces(sd, embed_table().to(dtype), cast_data(data, dtype), cfg, create_graph=True)
    loss = F.smooth_l1_loss(out["energy"], tE.to(dtype)) + 100.0*F.smooth_l1_loss(out["forces"], tF.to(dtype))
    loss.backward()
    return {k: v.grad.double() for k,v in sd.items() if v.grad is not None}
for name in ['pbc_small','mol_small']:
    z, cfg, data = load_golden(name); data.pop('pbc',None)
    tE, tF = torch.from_numpy(z["f64:target_energy"]), torch.from_numpy(z["f64:target_forces"])
    for seed in [1234, 1, 2, 3, 4, 5, 6, 7]:
        g64, g32 = grads(cfg, seed, data, tE, tF, torch.float64), grads(cfg, seed, data, tE, tF, torch.float32)
        worst = max((g32[k]-g64[k]).norm().item()/max(g64[k].norm().item(),1e-30) for k in g64 if g64[k].norm()>0)
        print(name, seed, f"{worst:.3e}", flush=True)
