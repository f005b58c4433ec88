"""Dev check of the forward edge kernel: values vs the fp64 oracle on the parity-test cases, then timings."""
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
        for kind in ("mol", "iso", "pbc", "pbc2"):
            d, ei, co, cell, t = T._edge_case(kind, cfg)
            N = d["pos"].shape[0]; G = d["ptr"].numel() - 1
            xo, Vo = orc.edge_message(t["x"], t["V"], t["s"], t["v"], t["pos"], t["W"], t["b"], t["freq"], ei, cfg,
                                      cell.double() if cell is not None else None, co, d["batch"])
            dims = ops.Dims(cfg.node_dim, *cfg.muls, cfg.num_basis, cfg.cutoff)
            graph = T.graph_from_edge_index(ei.to(DEV), N, G, cell_offsets=co.to(DEV) if co is not None else None,
                                            cell=cell.to(DEV) if cell is not None else None, batch=d["batch"].to(DEV))
            f32 = lambda a: a.float().to(DEV).contiguous()
            cmf = lambda a: orc.to_cm(a, cfg).float().to(DEV).contiguous()
            pos, s, W, b, freq = (f32(t[k]) for k in ("pos", "s", "W", "b", "freq"))
            v, V = (cmf(t[k]) for k in ("v", "V")); x = f32(t["x"])
            x_out, V_out = ops.edge_message_fwd_raw(graph, dims, pos, s, v, x, V, W, b, freq)
            torch.cuda.synchronize()
            ex, eV = T._rel(x_out, xo), T._rel(orc.from_cm(V_out.cpu(), cfg), Vo)
            x2, V2 = ops.edge_message_fwd_raw(graph, dims, pos, s, v, x, V, W, b, freq)
            print(cname, kind, "N", N, "E", graph.n_edges, "tile_mode", graph.tile_mode, "rel err x %.2e V %.2e" % (ex, eV),
                  "OK" if max(ex, eV) < 2e-5 else "FAIL", "det", bool(torch.equal(x_out, x2) and torch.equal(V_out, V2)), flush=True)
if what in ("all", "time"):
    cfg = orc.CONFIG_DEFAULT
    for nm in (256, 8192):
        d = orc.make_aspirin_batch(nm, seed=0, with_edges=False)
        g, _, _ = xb.build_graph(d["pos"].to(DEV), 5.0, ptr=d["ptr"].to(DEV), batch=d["batch"].to(DEV)); N = g.n_nodes
        dims = ops.Dims(cfg.node_dim, *cfg.muls, cfg.num_basis, cfg.cutoff); r = lambda *s: torch.randn(*s, device=DEV)
        pos = d["pos"].to(DEV); s, v, x, V = r(N, dims.H), r(N, dims.D), r(N, dims.node_dim), r(N, dims.D)
        W, b = 0.3 * r(dims.H, 20), 0.3 * r(dims.H); freq = (torch.pi * torch.arange(1, 21, device=DEV) / 5.0).float()
        def tm(f, n=20):
            f(); torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b_ = torch.cuda.Event(enable_timing=True); a.record()
            for _ in range(n): f()
            b_.record(); torch.cuda.synchronize(); return a.elapsed_time(b_) / n
        ms = tm(lambda: ops.edge_message_fwd_raw(g, dims, pos, s, v, x, V, W, b, freq))
        byts = 9104 * N + 4 * g.n_edges + 4
        print("aspirin x", nm, "N", N, "E", g.n_edges, "fwd ms %.4f" % ms, "GB/s %.1f" % (byts / ms / 1e6), "frac %.3f" % (byts / ms / 1e6 / 6548.2), flush=True)
if what == "dump":  # python dev_fwd.py dump <file>: forward on 64 aspirins + 40 mixed molecules with molecule tiles
    torch.manual_seed(0)
    outs = []
    cfg = orc.CONFIG_DEFAULT
    for d in (orc.make_aspirin_batch(64, seed=0, with_edges=False), orc.make_molecule_batch(40, (1, 30), seed=7)):
        g, _, _ = xb.build_graph(d["pos"].to(DEV), 5.0, ptr=d["ptr"].to(DEV), batch=d["batch"].to(DEV)); N = g.n_nodes
        dims = ops.Dims(cfg.node_dim, *cfg.muls, cfg.num_basis, cfg.cutoff); r = lambda *s: torch.randn(*s, device=DEV)
        pos = d["pos"].to(DEV); s, v, x, V = r(N, dims.H), r(N, dims.D), r(N, dims.node_dim), r(N, dims.D)
        W, b = 0.3 * r(dims.H, 20), 0.3 * r(dims.H); freq = (torch.pi * torch.arange(1, 21, device=DEV) / 5.0).float()
        xo, Vo = ops.edge_message_fwd_raw(g, dims, pos, s, v, x, V, W, b, freq)
        print("tile_mode", g.tile_mode, "N", N, "E", g.n_edges)
        outs += [xo.cpu(), Vo.cpu()]
    torch.save(outs, sys.argv[2])
if what == "cmp":
    a, b = torch.load(sys.argv[2]), torch.load(sys.argv[3])
    for i, (p, q) in enumerate(zip(a, b)):
        print(i, "max abs diff %.3e" % float((p - q).abs().max()), "scale %.3e" % float(q.abs().max()), "nan", bool(torch.isnan(p).any()))
if what == "prof":
    cfg = orc.CONFIG_DEFAULT
    nm = int(sys.argv[2])
    d = orc.make_aspirin_batch(nm, seed=0, with_edges=False)
    g, _, _ = xb.build_graph(d["pos"].to(DEV), 5.0, ptr=d["ptr"].to(DEV), batch=d["batch"].to(DEV)); N = g.n_nodes
    dims = ops.Dims(cfg.node_dim, *cfg.muls, cfg.num_basis, cfg.cutoff); r = lambda *s: torch.randn(*s, device=DEV)
    pos = d["pos"].to(DEV); s, v, x, V = r(N, dims.H), r(N, dims.D), r(N, dims.node_dim), r(N, dims.D)
    W, b = 0.3 * r(dims.H, 20), 0.3 * r(dims.H); freq = (torch.pi * torch.arange(1, 21, device=DEV) / 5.0).float()
    for _ in range(3): ops.edge_message_fwd_raw(g, dims, pos, s, v, x, V, W, b, freq)
    torch.cuda.synchronize()
