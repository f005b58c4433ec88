import sys; sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, numpy as np
from oracle import xpainn_oracle as orc
import xequinet_b200 as xb
import test_gpu_bench_shapes as T
def run(name, cfg, data, pbc=False):
    d64 = dict(data)
    if pbc:
        n = torch.tensor([data["pos"].shape[0]])
        d64["edge_index"], d64["cell_offsets"] = orc.radius_graph_pbc(data["pos"], n, data["pbc"], data["cell"], cfg.cutoff)
    else:
        d64["edge_index"] = orc.radius_graph(data["pos"], cfg.cutoff, data["batch"])
    ref64, ref32 = T._oracle_ef(cfg, d64, torch.float64), T._oracle_ef(cfg, d64, torch.float32)
    model = T._model(cfg)
    keys_ = ("pos", "atomic_numbers", "batch", "ptr") + (("cell", "pbc") if pbc else ())
    d = xb.NeighborTransform(cfg.cutoff)(T._dev({k: data[k] for k in keys_})); d.pop("pbc", None)
    out = model(d, compute_forces=True)
    F64 = ref64["forces"].numpy()
    for lab, F in (("ours ", out["forces"].cpu().numpy()), ("ref32", ref32["forces"].numpy())):
        e = np.abs(F - F64).max(axis=1)
        print(f"{name} {lab}: N {len(e)} p50 {np.percentile(e,50):.1e} p90 {np.percentile(e,90):.1e} p99 {np.percentile(e,99):.1e} p99.9 {np.percentile(e,99.9):.1e} max {e.max():.1e}  >1e-4: {int((e>1e-4).sum())}  >1e-3: {int((e>1e-3).sum())}  |F| rms {np.sqrt((F64**2).mean()):.2f}")
run("c1", orc.CONFIG_DEFAULT, orc.make_molecule_batch(64, 18, seed=0, with_edges=False))
run("c3", orc.CONFIG_DEFAULT, orc.make_aspirin_batch(256, seed=0, with_edges=False))
run("c4", orc.CONFIG_C4, orc.make_molecule_batch(32, (30, 70), seed=0, z_table=orc._Z_SPICE, with_edges=False))
run("c5", orc.CONFIG_DEFAULT, orc.make_water_box(8, seed=1), pbc=True)
