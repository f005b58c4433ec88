"""Single-GPU emulation of rank 0 of a P-slab decomposition of the c5 box: edge kernel timings on the local graph."""
import sys; sys.path.insert(0, ".")
import torch
from oracle import xpainn_oracle as orc
import xequinet_b200 as xb
from xequinet_b200 import ops, domain
DEV = "cuda"; P = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = orc.CONFIG_DEFAULT
box = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in orc.make_water_box(15, seed=0).items()}
cell = box["cell"].reshape(3, 3)
pos = domain.wrap_into_cell(box["pos"], cell)
f0 = domain.fractional(pos, cell)[:, 0]
m = 5.0 / domain.perpendicular_width(cell, 0) * (1 + 1e-5)
own = torch.nonzero(f0 < 1.0 / P).flatten()
right = torch.nonzero((f0 >= 1.0 / P) & (f0 < 1.0 / P + m)).flatten()
left = torch.nonzero(f0 >= 1.0 - m).flatten()
pos_local = torch.cat([pos[own], pos[right], pos[left] - cell[0]]).contiguous()
g = domain.local_graph(pos_local, cell, 5.0, own.numel(), P)
gfull, _, _ = xb.build_graph(box["pos"], 5.0, ptr=box["ptr"], batch=box["batch"], cell=box["cell"], pbc=box["pbc"])
dims = ops.Dims(cfg.node_dim, *cfg.muls, cfg.num_basis, cfg.cutoff)
def tm(f, n=10):
    f(); torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
for name, gr, p in (("full", gfull, box["pos"]), (f"rank0 of {P}", g, pos_local)):
    N = gr.n_nodes; r = lambda *s: torch.randn(*s, device=DEV)
    s, v, x, V, gx, gV = r(N, dims.H), r(N, dims.D), r(N, dims.node_dim), r(N, dims.D), r(N, dims.node_dim), r(N, dims.D)
    W, b = 0.3 * r(dims.H, 20), 0.3 * r(dims.H); freq = (torch.pi * torch.arange(1, 21, device=DEV) / 5.0).float()
    t1 = tm(lambda: ops.edge_message_fwd_raw(gr, dims, p, s, v, x, V, W, b, freq))
    t2 = tm(lambda: ops.edge_message_bwd_raw(gr, dims, p, s, v, W, b, freq, gx, gV, need_w=False))
    rl = (gr.t_rowptr[1:] - gr.t_rowptr[:-1]).float()
    print(f"{name}: N {N} centers {gr.n_centers} E {gr.n_edges} n_tiles {gr.n_tiles} t_n_tiles {gr.t_n_tiles} fwd {t1:.3f} ms bwd {t2:.3f} ms; t-row len mean {float(rl.mean()):.1f} max {int(rl.max())} zero rows {int((rl == 0).sum())}")
