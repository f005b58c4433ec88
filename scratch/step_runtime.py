"""One E+F evaluation through the C inference runtime (for ncu launch lists): python scratch/step_runtime.py [c1|c5]"""
import sys; sys.path.insert(0, ".")
import torch
import xequinet_b200 as xb
from oracle import xpainn_oracle as orc
from xequinet_b200 import runtime
which = sys.argv[1] if len(sys.argv) > 1 else "c1"
cfg = orc.CONFIG_DEFAULT
d = orc.make_molecule_batch(64, 18, seed=0, with_edges=False) if which == "c1" else orc.make_water_box(15, seed=0)
model = xb.resolve_model("xpainn", **cfg.model_kwargs()); model.load_state_dict(orc.synthetic_state_dict(cfg, 1234), strict=False)
model = model.to("cuda").eval()
native = runtime.NativeModel(model)
data = xb.NeighborTransform(5.0)({k: v.to("cuda") for k, v in d.items() if torch.is_tensor(v)})
data.pop("pbc", None)
for it in range(3):
    torch.cuda.synchronize()
    if it == 2:
        torch.cuda.profiler.start()
    out = native(dict(data), compute_forces=True)
    torch.cuda.synchronize()
    if it == 2:
        torch.cuda.profiler.stop()
print("E", float(out["energy"].sum()))
