"""Timeline of the warp-specialised forward kernel (debug build with -DXEQ_TRACE, scratch/libxeq_trace.so):
per chunk, when each warp of CTA 0 starts its work and when it arrives at the CTA barrier."""
import ctypes, os, sys; sys.path.insert(0, ".")
from pathlib import Path
import numpy as np, torch
from xequinet_b200 import _lib
_lib.LIB_PATH = Path("scratch/libxeq_trace.so").resolve()
from oracle import xpainn_oracle as orc
import xequinet_b200 as xb
from xequinet_b200 import ops
cfg = orc.CONFIG_DEFAULT
d = orc.make_aspirin_batch(256, seed=0, with_edges=False); dev = "cuda"
g, _, _ = xb.build_graph(d["pos"].to(dev), 5.0, ptr=d["ptr"].to(dev), batch=d["batch"].to(dev)); N = g.n_nodes
dims = ops.Dims(cfg.node_dim, *cfg.muls, cfg.num_basis, cfg.cutoff); r = lambda *s: torch.randn(*s, device=dev)
pos = d["pos"].to(dev); s, v, x, V = r(N, dims.H), r(N, dims.D), r(N, dims.node_dim), r(N, dims.D)
W, b = 0.3 * r(dims.H, 20), 0.3 * r(dims.H); freq = (torch.pi * torch.arange(1, 21, device=dev) / 5.0).float()
for _ in range(3): ops.edge_message_fwd_raw(g, dims, pos, s, v, x, V, W, b, freq)
torch.cuda.synchronize()
lib = _lib.get()
buf = np.zeros((21, 2, 256), dtype=np.int64)
fn = lib.xeq_debug_fwd_trace
fn.restype = ctypes.c_int
rc = fn(buf.ctypes.data_as(ctypes.c_void_p)); print("rc", rc)
t0 = buf[:, 0, :]; t1 = buf[:, 1, :]
n = int((t1[0] > 0).sum()); print("chunks traced", n)
base = t0[:, 0].min()
names = ["L0"] * 4 + ["L1"] * 2 + ["L2"] + ["cursor", "mma"] + ["radial"] * 12
print("chunk | per warp kind: work cycles (start->barrier arrival), and who arrives last")
for c in range(2, min(n, 30)):
    work = t1[:, c] - t0[:, c]
    last = int(np.argmax(t1[:, c]))
    period = t0[0, c + 1] - t0[0, c] if c + 1 < n else 0
    kinds = {}
    for w in range(21): kinds.setdefault(names[w], []).append(int(work[w]))
    print(c, "period", int(period), {k: max(vv) for k, vv in kinds.items()}, "last:", names[last], "spread", int(t1[:, c].max() - t1[:, c].min()))
w_all = (t1[:, 2:n] - t0[:, 2:n])
for k in ("L0", "L1", "L2", "cursor", "mma", "radial"):
    idx = [i for i, nm in enumerate(names) if nm == k]
    print(k, "mean work", w_all[idx].mean().round(), "max", w_all[idx].max())
print("mean period", np.diff(t0[0, 2:n]).mean().round())
