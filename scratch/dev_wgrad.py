"""Dev check of the weight-gradient kernels: new (edge_wgrad_ul.cu) vs round 1 (XEQ_WGRAD_R1=1), same process inputs.
usage: python scratch/dev_wgrad.py dump TAG | cmp TAG1 TAG2 | time"""
import sys, os; sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
mode = sys.argv[1]
if mode == "cmp":
    a, b = torch.load(f"/tmp/wg_{sys.argv[2]}.pt"), torch.load(f"/tmp/wg_{sys.argv[3]}.pt")
    bad = 0
    for k in a:
        x, y = a[k].double(), b[k].double()
        err = (x - y).abs().max().item() / max(y.abs().max().item(), 1e-30)
        ok = err < 3e-6
        bad += not ok
        print(f"{k:40s} rel {err:.2e} {'OK' if ok else 'FAIL'}")
    print("ALL OK" if not bad else f"{bad} FAIL")
    sys.exit(0)
from oracle import xpainn_oracle as orc
import xequinet_b200 as xb
from xequinet_b200 import ops
DEV = "cuda"
def inputs(N, dims, seed=0):
    g_ = torch.Generator(device=DEV).manual_seed(seed); r = lambda *s: torch.randn(*s, device=DEV, generator=g_)
    return dict(s=r(N, dims.H), v=r(N, dims.D), gx=r(N, dims.node_dim), gV=r(N, dims.D), a_s=r(N, dims.H), a_v=r(N, dims.D), a_p=r(N, 3),
                W=0.3 * r(dims.H, 20), b=0.3 * r(dims.H), freq=(torch.pi * torch.arange(1, 21, device=DEV) / 5.0).float() + 0.05 * r(20))
def graphs():
    d = orc.make_molecule_batch(30, (1, 23), seed=5)
    yield "tiles", d, xb.build_graph(d["pos"].to(DEV), 5.0, ptr=d["ptr"].to(DEV), batch=d["batch"].to(DEV))[0]
    d = orc.make_aspirin_batch(64, seed=0, with_edges=False)
    yield "aspirin64", d, xb.build_graph(d["pos"].to(DEV), 5.0, ptr=d["ptr"].to(DEV), batch=d["batch"].to(DEV))[0]
    d = orc.make_molecule_batch(12, (30, 70), seed=3)
    yield "mol30-70", d, xb.build_graph(d["pos"].to(DEV), 5.0, ptr=d["ptr"].to(DEV), batch=d["batch"].to(DEV))[0]
    d = orc.make_water_box(5, seed=0)
    yield "water375", d, xb.build_graph(d["pos"].to(DEV), 5.0, ptr=d["ptr"].to(DEV), batch=d["batch"].to(DEV), cell=d["cell"].to(DEV), pbc=d.get("pbc"))[0]
    d = orc.make_molecule_batch(3, (1, 2), seed=1)
    yield "tiny", d, xb.build_graph(d["pos"].to(DEV), 5.0, ptr=d["ptr"].to(DEV), batch=d["batch"].to(DEV))[0]
if mode == "dump":
    out = {}
    for cfg, cname in ((orc.CONFIG_DEFAULT, "c128"), (orc.CONFIG_C4, "c256")):
        dims = ops.Dims(cfg.node_dim, *cfg.muls, cfg.num_basis, cfg.cutoff)
        for name, d, g in graphs():
            t = inputs(g.n_nodes, dims); pos = d["pos"].to(DEV)
            r1 = ops.edge_message_bwd_raw(g, dims, pos, t["s"], t["v"], t["W"], t["b"], t["freq"], t["gx"], t["gV"], need_w=True)
            r2 = ops.edge_message_bwdbwd_raw(g, dims, pos, t["s"], t["v"], t["W"], t["b"], t["freq"], t["gx"], t["gV"], t["a_s"], t["a_v"], t["a_p"])
            r1b = ops.edge_message_bwd_raw(g, dims, pos, t["s"], t["v"], t["W"], t["b"], t["freq"], t["gx"], t["gV"], need_w=True)
            torch.cuda.synchronize()
            det = all(torch.equal(x, y) for x, y in zip(r1[3:], r1b[3:]))
            print(cname, name, "N", g.n_nodes, "E", g.n_edges, "tile_mode", g.tile_mode, "det", det, "finite", all(torch.isfinite(x).all().item() for x in r1[3:] + r2[5:]), flush=True)
            for i, k in enumerate(("gW", "gb", "gf")):
                out[f"{cname}/{name}/1/{k}"] = r1[3 + i].cpu(); out[f"{cname}/{name}/2/{k}"] = r2[5 + i].cpu()
    torch.save(out, f"/tmp/wg_{sys.argv[2]}.pt")
if mode == "time":
    cfg = orc.CONFIG_DEFAULT; dims = ops.Dims(cfg.node_dim, *cfg.muls, cfg.num_basis, cfg.cutoff)
    def tm(f, n=20):
        f(); torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b_ = torch.cuda.Event(enable_timing=True); a.record()
        for _ in range(n): f()
        b_.record(); torch.cuda.synchronize(); return a.elapsed_time(b_) / n
    for name, d in (("aspirin x 256", orc.make_aspirin_batch(256, seed=0, with_edges=False)), ("aspirin x 8192", orc.make_aspirin_batch(8192, seed=0, with_edges=False)), ("water 10125", orc.make_water_box(15, seed=0))):
        g = xb.build_graph(d["pos"].to(DEV), 5.0, ptr=d["ptr"].to(DEV), batch=d["batch"].to(DEV), cell=d["cell"].to(DEV) if "cell" in d else None, pbc=d.get("pbc"))[0]
        t = inputs(g.n_nodes, dims); pos = d["pos"].to(DEV)
        a = tm(lambda: ops.edge_message_bwd_raw(g, dims, pos, t["s"], t["v"], t["W"], t["b"], t["freq"], t["gx"], t["gV"], need_w=False))
        b = tm(lambda: ops.edge_message_bwd_raw(g, dims, pos, t["s"], t["v"], t["W"], t["b"], t["freq"], t["gx"], t["gV"], need_w=True))
        c = tm(lambda: ops.edge_message_bwdbwd_raw(g, dims, pos, t["s"], t["v"], t["W"], t["b"], t["freq"], t["gx"], t["gV"], t["a_s"], t["a_v"], t["a_p"], need_w=False))
        e = tm(lambda: ops.edge_message_bwdbwd_raw(g, dims, pos, t["s"], t["v"], t["W"], t["b"], t["freq"], t["gx"], t["gV"], t["a_s"], t["a_v"], t["a_p"], need_w=True))
        print(f"{name}: E {g.n_edges} wgrad1 {b - a:.4f} ms  wgrad2 {e - c:.4f} ms (bwd {a:.4f}, bwdbwd {c:.4f})", flush=True)
