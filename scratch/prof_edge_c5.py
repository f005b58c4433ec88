"""ncu target: forward + first-order edge kernels on the c5 water box (tile_mode 0: L2 gathers) or a c4-shaped batch."""
import sys; sys.path.insert(0, '.')
import torch
from oracle import xpainn_oracle as orc
import xequinet_b200 as xb
from xequinet_b200 import ops
which = sys.argv[1] if len(sys.argv) > 1 else "c5"
dev = 'cuda'
if which == "c5":
    cfg = orc.CONFIG_DEFAULT; d = orc.make_water_box(15, seed=0)
    g, _, _ = xb.build_graph(d["pos"].to(dev), 5.0, ptr=d["ptr"].to(dev), batch=d["batch"].to(dev), cell=d["cell"].to(dev), pbc=d["pbc"])
else:
    cfg = orc.CONFIG_C4; d = orc.make_molecule_batch(128, (30, 70), seed=0, z_table=orc._Z_SPICE, with_edges=False)
    g, _, _ = xb.build_graph(d["pos"].to(dev), 5.0, ptr=d["ptr"].to(dev), batch=d["batch"].to(dev))
N = g.n_nodes
dims = ops.Dims(cfg.node_dim, *cfg.muls, cfg.num_basis, cfg.cutoff)
r = lambda *s: torch.randn(*s, device=dev)
pos = d['pos'].to(dev); s, v, x, V = r(N, dims.H), r(N, dims.D), r(N, dims.node_dim), r(N, dims.D)
W, b = 0.3 * r(dims.H, 20), 0.3 * r(dims.H); freq = (torch.pi * torch.arange(1, 21, device=dev) / 5.0).float()
gx, gV = r(N, dims.node_dim), r(N, dims.D)
print(which, 'N', N, 'E', g.n_edges, 'tile_mode', g.tile_mode)
for _ in range(2):
    ops.edge_message_fwd_raw(g, dims, pos, s, v, x, V, W, b, freq)
    ops.edge_message_bwd_raw(g, dims, pos, s, v, W, b, freq, gx, gV, need_w=False)
torch.cuda.synchronize()
