import sys; sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from oracle import xpainn_oracle as orc
import xequinet_b200 as xb
from xequinet_b200 import ops
exec(open("scratch/dev_wgrad.py").read().split("if mode == \"dump\":")[0].split("DEV = \"cuda\"")[1].replace("def inputs", "DEV='cuda'\ndef inputs", 1))
cfg = orc.CONFIG_DEFAULT; dims = ops.Dims(cfg.node_dim, *cfg.muls, cfg.num_basis, cfg.cutoff)
for name, d, g in graphs():
    t = inputs(g.n_nodes, dims); pos = d["pos"].to(DEV)
    print("case", name, flush=True)
    r1 = ops.edge_message_bwd_raw(g, dims, pos, t["s"], t["v"], t["W"], t["b"], t["freq"], t["gx"], t["gV"], need_w=True)
    torch.cuda.synchronize(); print(" first order done", flush=True)
    r2 = ops.edge_message_bwdbwd_raw(g, dims, pos, t["s"], t["v"], t["W"], t["b"], t["freq"], t["gx"], t["gV"], t["a_s"], t["a_v"], t["a_p"])
    torch.cuda.synchronize(); print(" second order done", flush=True)
