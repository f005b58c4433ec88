"""Launches each hand-written edge kernel a few times on a c3-shaped (or large) batch: ncu target."""
import sys; sys.path.insert(0, '.')
import torch
from oracle import xpainn_oracle as orc
import xequinet_b200 as xb
from xequinet_b200 import ops
n_mol = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = orc.CONFIG_DEFAULT
d = orc.make_aspirin_batch(n_mol, seed=0, with_edges=False)
dev = 'cuda'
g, _, _ = xb.build_graph(d["pos"].to(dev), 5.0, ptr=d["ptr"].to(dev), batch=d["batch"].to(dev))
N = g.n_nodes
dims = ops.Dims(cfg.node_dim, *cfg.muls, cfg.num_basis, cfg.cutoff)
r = lambda *s: torch.randn(*s, device=dev)
pos = d['pos'].to(dev); s, v, x, V = r(N, dims.H), r(N, dims.D), r(N, dims.node_dim), r(N, dims.D)
W, b = 0.3 * r(dims.H, 20), 0.3 * r(dims.H); freq = (torch.pi * torch.arange(1, 21, device=dev) / 5.0).float()
gx, gV, a_s, a_v, a_p = r(N, dims.node_dim), r(N, dims.D), r(N, dims.H), r(N, dims.D), r(N, 3)
print('N', N, 'E', g.n_edges)
for _ in range(reps):
    ops.edge_message_fwd_raw(g, dims, pos, s, v, x, V, W, b, freq)
    ops.edge_message_bwd_raw(g, dims, pos, s, v, W, b, freq, gx, gV, need_w=True)
    ops.edge_message_bwdbwd_raw(g, dims, pos, s, v, W, b, freq, gx, gV, a_s, a_v, a_p)
torch.cuda.synchronize()
