import sys; sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, numpy as np
from oracle import xpainn_oracle as orc
import xequinet_b200 as xb
import test_gpu_bench_shapes as T
cfgname = sys.argv[1] if len(sys.argv) > 1 else "c4"
cfg = orc.CONFIG_C4 if cfgname == "c4" else orc.CONFIG_DEFAULT
data = orc.make_molecule_batch(32, (30, 70), seed=0, z_table=orc._Z_SPICE, with_edges=False)
d64 = dict(data); d64["edge_index"] = orc.radius_graph(data["pos"], cfg.cutoff, data["batch"])
ref64, ref32 = T._oracle_ef(cfg, d64, torch.float64), T._oracle_ef(cfg, d64, torch.float32)
model = T._model(cfg)
d = xb.NeighborTransform(cfg.cutoff)(T._dev({k: data[k] for k in ("pos", "atomic_numbers", "batch", "ptr")}))
out = model(d, compute_forces=True)
ea = out["atomic_energies"].detach().cpu().double(); e64 = ref64["atomic_energies"].detach().double(); e32 = ref32["atomic_energies"].detach().double()
print(cfgname, "atomic: new-64 max %.2e mean %.2e bias %.2e | ref32-64 max %.2e mean %.2e bias %.2e | scale %.2f" % (
    (ea - e64).abs().max(), (ea - e64).abs().mean(), (ea - e64).mean(), (e32 - e64).abs().max(), (e32 - e64).abs().mean(), (e32 - e64).mean(), e64.abs().mean()))
E = out["energy"].detach().cpu().double(); E64 = ref64["energy"].detach().double(); E32 = ref32["energy"].detach().double()
print("molecular: new-64 max %.2e | ref32-64 max %.2e" % ((E - E64).abs().max(), (E32 - E64).abs().max()))
# segment-sum only: sum our atomic energies in fp64
Es = torch.zeros_like(E64).index_add_(0, data["batch"], ea)
print("our atomic energies summed in fp64 vs our energy: %.2e ; vs E64: %.2e" % ((Es - E).abs().max(), (Es - E64).abs().max()))
