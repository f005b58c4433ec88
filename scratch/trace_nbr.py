"""Timeline of nbr_mma_kernel (debug build -DXEQ_TRACE): per chunk, work cycles of each warp of CTA 0 between
the barriers, and who arrives last.  usage: python scratch/trace_nbr.py [order]"""
import ctypes, sys; sys.path.insert(0, ".")
from pathlib import Path
import numpy as np, torch
from xequinet_b200 import _lib
_lib.LIB_PATH = Path("scratch/libxeq_trace.so").resolve()
from oracle import xpainn_oracle as orc
import xequinet_b200 as xb
from xequinet_b200 import ops
order = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = orc.CONFIG_DEFAULT
d = orc.make_aspirin_batch(256, seed=0, with_edges=False); dev = "cuda"
g, _, _ = xb.build_graph(d["pos"].to(dev), 5.0, ptr=d["ptr"].to(dev), batch=d["batch"].to(dev)); N = g.n_nodes
dims = ops.Dims(cfg.node_dim, *cfg.muls, cfg.num_basis, cfg.cutoff); r = lambda *s: torch.randn(*s, device=dev)
pos = d["pos"].to(dev); s, v = r(N, dims.H), r(N, dims.D)
W, b = 0.3 * r(dims.H, 20), 0.3 * r(dims.H); freq = (torch.pi * torch.arange(1, 21, device=dev) / 5.0).float()
gx, gV, a_s, a_v, a_p = r(N, dims.node_dim), r(N, dims.D), r(N, dims.H), r(N, dims.D), r(N, 3)
for _ in range(3):
    if order == 1: ops.edge_message_bwd_raw(g, dims, pos, s, v, W, b, freq, gx, gV, need_w=False)
    else: ops.edge_message_bwdbwd_raw(g, dims, pos, s, v, W, b, freq, gx, gV, a_s, a_v, a_p, need_g=False, need_w=False)
torch.cuda.synchronize()
fn = _lib.get().xeq_debug_nbr_trace; fn.restype = ctypes.c_int
buf = np.zeros((8, 2, 256), dtype=np.int64)
print("rc", fn(buf.ctypes.data_as(ctypes.c_void_p)))
t0 = buf[:, 0, :]; t1 = buf[:, 1, :]
n = int((t1[0] > 0).sum()); print("chunks traced", n)
names = ["L0"] * 4 + ["L1"] * 2 + ["L2", "producer"]
for c in range(2, min(n, 14)):
    work = t1[:, c] - t0[:, c]; last = int(np.argmax(t1[:, c]))
    period = t0[0, c + 1] - t0[0, c] if c + 1 < n else 0
    kinds = {}
    for w in range(8): kinds.setdefault(names[w], []).append(int(work[w]))
    print(c, "period", int(period), {k: max(vv) for k, vv in kinds.items()}, "last:", names[last], "spread", int(t1[:, c].max() - t1[:, c].min()))
w_all = (t1[:, 2:n] - t0[:, 2:n])
for k in ("L0", "L1", "L2", "producer"):
    idx = [i for i, nm in enumerate(names) if nm == k]
    print(k, "mean work", w_all[idx].mean().round(), "max", w_all[idx].max())
print("mean period", np.diff(t0[0, 2:n]).mean().round())
