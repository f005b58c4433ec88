"""Dev check of the first-order edge kernel (K2b): d/d(s, v, pos) vs fp64 autograd of the oracle on the parity-test cases, then timings."""
import sys, os; sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from oracle import xpainn_oracle as orc
import xequinet_b200 as xb
from xequinet_b200 import ops
import test_gpu_parity as T
DEV = "cuda"
what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("all", "check"):
    for cfg, cname in ((orc.CONFIG_DEFAULT, "c128"), (orc.CONFIG_C4, "c256")):
        for kind in ("mol", "iso", "pbc", "pbc2", "tiles"):
            if kind == "tiles":  # molecule tiles (staged window)
                d = orc.make_molecule_batch(30, (1, 23), seed=5); ei, co, cell = d["edge_index"], None, None
                N = d["pos"].shape[0]; g_ = torch.Generator().manual_seed(4); rnd = lambda *s: torch.randn(*s, generator=g_, dtype=torch.float64)
                t = dict(pos=d["pos"].double(), s=rnd(N, cfg.H_msg), v=rnd(N, cfg.D), x=rnd(N, cfg.node_dim), V=rnd(N, cfg.D), W=0.3 * rnd(cfg.H_msg, cfg.num_basis), b=0.3 * rnd(cfg.H_msg),
                         freq=torch.pi * torch.arange(1, cfg.num_basis + 1, dtype=torch.float64) / cfg.cutoff + 0.1 * rnd(cfg.num_basis), gx=rnd(N, cfg.node_dim), gV=rnd(N, cfg.D))
            else:
                d, ei, co, cell, t = T._edge_case(kind, cfg)
            N = d["pos"].shape[0]; G = d["ptr"].numel() - 1
            req = {k: t[k].clone().requires_grad_(True) for k in ("pos", "s", "v")}
            xo, Vo = orc.edge_message(t["x"], t["V"], req["s"], req["v"], req["pos"], t["W"], t["b"], t["freq"], ei, cfg,
                                      cell.double() if cell is not None else None, co, d["batch"])
            Phi = (t["gx"] * xo).sum() + (t["gV"] * Vo).sum()
            first = torch.autograd.grad(Phi, [req[k] for k in ("s", "v", "pos")])
            dims = ops.Dims(cfg.node_dim, *cfg.muls, cfg.num_basis, cfg.cutoff)
            graph = T.graph_from_edge_index(ei.to(DEV), N, G, cell_offsets=co.to(DEV) if co is not None else None,
                                            cell=cell.to(DEV) if cell is not None else None, batch=d["batch"].to(DEV),
                                            ptr=d["ptr"].to(DEV) if kind == "tiles" else None)
            f32 = lambda a: a.float().to(DEV).contiguous()
            cmf = lambda a: orc.to_cm(a, cfg).float().to(DEV).contiguous()
            pos, s, W, b, freq, gx = (f32(t[k]) for k in ("pos", "s", "W", "b", "freq", "gx"))
            v, gV = (cmf(t[k]) for k in ("v", "gV"))
            gs, gv, gpos, *_ = ops.edge_message_bwd_raw(graph, dims, pos, s, v, W, b, freq, gx, gV, need_w=False)
            torch.cuda.synchronize()
            errs = [T._rel(gs, first[0]), T._rel(orc.from_cm(gv.cpu(), cfg), first[1]), T._rel(gpos, first[2])]
            gs2, gv2, gpos2, *_ = ops.edge_message_bwd_raw(graph, dims, pos, s, v, W, b, freq, gx, gV, need_w=False)
            det = bool(torch.equal(gs, gs2) and torch.equal(gv, gv2) and torch.equal(gpos, gpos2))
            print(cname, kind, "N", N, "E", graph.n_edges, "tile_mode", graph.tile_mode, "rel err s %.2e v %.2e pos %.2e" % tuple(errs),
                  "OK" if max(errs) < 2e-5 else "FAIL", "det", det, flush=True)
if what in ("all", "time"):
    cfg = orc.CONFIG_DEFAULT
    cases = [("aspirin x 256", orc.make_aspirin_batch(256, seed=0, with_edges=False)), ("aspirin x 8192", orc.make_aspirin_batch(8192, seed=0, with_edges=False)),
             ("water box 10125", orc.make_water_box(15, seed=0))]
    for name, d in cases:
        g, _, _ = xb.build_graph(d["pos"].to(DEV), 5.0, ptr=d["ptr"].to(DEV), batch=d["batch"].to(DEV), cell=d["cell"].to(DEV) if "cell" in d else None,
                                 pbc=d.get("pbc")); N = g.n_nodes
        dims = ops.Dims(cfg.node_dim, *cfg.muls, cfg.num_basis, cfg.cutoff); r = lambda *s: torch.randn(*s, device=DEV)
        pos = d["pos"].to(DEV); s, v, gx, gV = r(N, dims.H), r(N, dims.D), r(N, dims.node_dim), r(N, dims.D)
        W, b = 0.3 * r(dims.H, 20), 0.3 * r(dims.H); freq = (torch.pi * torch.arange(1, 21, device=DEV) / 5.0).float()
        def tm(f, n=10):
            f(); torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b_ = torch.cuda.Event(enable_timing=True); a.record()
            for _ in range(n): f()
            b_.record(); torch.cuda.synchronize(); return a.elapsed_time(b_) / n
        ms = tm(lambda: ops.edge_message_bwd_raw(g, dims, pos, s, v, W, b, freq, gx, gV, need_w=False))
        H, D, C, E = dims.H, dims.D, dims.node_dim, g.n_edges
        byts = 4 * N * (2 * (H + D) + (C + D) + 3) + 12 * N + 4 * (N + 1) + 12 * E
        print(name, "N", N, "E", E, "tile_mode", g.tile_mode, "bwd ms %.4f" % ms, "GB/s %.1f" % (byts / ms / 1e6), "frac %.3f" % (byts / ms / 1e6 / 6548.2), flush=True)
if what == "prof":
    cfg = orc.CONFIG_DEFAULT
    nm = int(sys.argv[2])
    d = orc.make_aspirin_batch(nm, seed=0, with_edges=False) if nm > 0 else orc.make_water_box(15, seed=0)
    g, _, _ = xb.build_graph(d["pos"].to(DEV), 5.0, ptr=d["ptr"].to(DEV), batch=d["batch"].to(DEV), cell=d["cell"].to(DEV) if "cell" in d else None, pbc=d.get("pbc")); N = g.n_nodes
    dims = ops.Dims(cfg.node_dim, *cfg.muls, cfg.num_basis, cfg.cutoff); r = lambda *s: torch.randn(*s, device=DEV)
    pos = d["pos"].to(DEV); s, v, gx, gV = r(N, dims.H), r(N, dims.D), r(N, dims.node_dim), r(N, dims.D)
    W, b = 0.3 * r(dims.H, 20), 0.3 * r(dims.H); freq = (torch.pi * torch.arange(1, 21, device=DEV) / 5.0).float()
    for _ in range(3): ops.edge_message_bwd_raw(g, dims, pos, s, v, W, b, freq, gx, gV, need_w=False)
    torch.cuda.synchronize()
