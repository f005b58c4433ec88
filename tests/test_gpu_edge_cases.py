"""Shapes that exercise the control paths of the round-2 edge kernels (mbarrier protocols must terminate on all of
them): batches that mix molecules inside and beyond the shared-memory window (23 rows), structures without any edge,
single atoms, tiles of exactly 23 / 24 nodes.  Molecule tiles (tile_mode 1: staged rows, TMA loader or in-kernel packer)
must give what edge-block tiles (tile_mode 0: L2 gathers) give -- per row the same sums in the same order: bit-identical."""
import pytest
import torch

from oracle import xpainn_oracle as orc

pytestmark = pytest.mark.gpu

import xequinet_b200 as xb  # noqa: E402
from xequinet_b200 import ops  # noqa: E402
from xequinet_b200.graph import build_graph, graph_from_edge_index  # noqa: E402

DEV = "cuda"


def _run(g, cfg, pos, seed=0):
    dims = ops.Dims(cfg.node_dim, *cfg.muls, cfg.num_basis, cfg.cutoff)
    gen = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=gen).to(DEV)
    N = g.n_nodes
    s, v, x, V, gx, gV = r(N, dims.H), r(N, dims.D), r(N, dims.node_dim), r(N, dims.D), r(N, dims.node_dim), r(N, dims.D)
    W, b = 0.3 * r(dims.H, 20), 0.3 * r(dims.H)
    freq = (torch.pi * torch.arange(1, 21) / 5.0).float().to(DEV)
    out = list(ops.edge_message_fwd_raw(g, dims, pos, s, v, x, V, W, b, freq))
    out += list(ops.edge_message_bwd_raw(g, dims, pos, s, v, W, b, freq, gx, gV, need_w=False))[:3]
    torch.cuda.synchronize()
    return out, (x, V)


@pytest.mark.parametrize("sizes", [(5, 40), (23, 23), (24, 24), (1, 64), (22, 25)])
@pytest.mark.parametrize("cfg", [orc.CONFIG_DEFAULT, orc.CONFIG_C4], ids=["c128", "c256"])
def test_molecule_tiles_equal_edge_block_tiles(sizes, cfg):
    d = orc.make_molecule_batch(20, sizes, seed=sizes[0] + sizes[1], with_edges=False)
    pos, batch, ptr = d["pos"].to(DEV), d["batch"].to(DEV), d["ptr"].to(DEV)
    g_mol, _, _ = build_graph(pos, cfg.cutoff, ptr=ptr, batch=batch)   # molecule tiles
    g_blk = graph_from_edge_index(g_mol.edge_index(), g_mol.n_nodes, g_mol.n_graphs, batch=batch)  # same graph, edge-block tiles
    assert g_mol.tile_mode == 1 and g_blk.tile_mode == 0 and torch.equal(g_mol.col, g_blk.col)
    a, _ = _run(g_mol, cfg, pos)
    b, _ = _run(g_blk, cfg, pos)
    for i, (p, q) in enumerate(zip(a, b)):
        assert torch.equal(p, q), f"output {i}"


def test_structures_without_edges():
    """Atoms further apart than the cutoff: no edge anywhere; the message is the residual, all derivatives vanish."""
    cfg = orc.CONFIG_DEFAULT
    for n, with_ptr in ((1, True), (7, True), (7, False), (300, False)):
        pos = (20.0 * torch.arange(n, dtype=torch.float32).reshape(-1, 1) * torch.tensor([[1.0, 0.3, 0.1]])).to(DEV)
        batch = torch.zeros(n, dtype=torch.long, device=DEV)
        if with_ptr:
            g, _, _ = build_graph(pos, cfg.cutoff, batch=batch, ptr=torch.tensor([0, n], device=DEV))
        else:
            g = graph_from_edge_index(torch.zeros((2, 0), dtype=torch.long, device=DEV), n, 1, batch=batch)
        assert g.n_edges == 0
        (x_out, V_out, gs, gv, gpos), (x, V) = _run(g, cfg, pos)
        assert torch.equal(x_out, x) and torch.equal(V_out, V)
        assert float(gs.abs().max()) == 0.0 and float(gv.abs().max()) == 0.0 and float(gpos.abs().max()) == 0.0


def test_model_on_a_single_atom_and_a_dimer():
    cfg = orc.CONFIG_DEFAULT
    model = xb.resolve_model("xpainn", **cfg.model_kwargs())
    model.load_state_dict(orc.synthetic_state_dict(cfg, 1234), strict=False)
    model = model.to(DEV).eval()
    for pos in ([[0.0, 0.0, 0.0]], [[0.0, 0.0, 0.0], [1.1, 0.0, 0.0]]):
        d = {"pos": torch.tensor(pos, device=DEV), "atomic_numbers": torch.tensor([8] * len(pos), dtype=torch.int32, device=DEV)}
        out = model(xb.NeighborTransform(cfg.cutoff)(d), compute_forces=True)
        assert out["energy"].shape == (1,) and out["forces"].shape == (len(pos), 3)
        assert bool(torch.isfinite(out["energy"]).all()) and bool(torch.isfinite(out["forces"]).all())
        assert float(out["forces"].sum(0).abs().max()) < 1e-5  # no net force


def _second_order(g, cfg, pos, seed=0, a_pos=True, a_sv=True):
    dims = ops.Dims(cfg.node_dim, *cfg.muls, cfg.num_basis, cfg.cutoff)
    gen = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=gen).to(DEV)
    N = g.n_nodes
    s, v, gx, gV, a_s, a_v, a_p = r(N, dims.H), r(N, dims.D), r(N, dims.node_dim), r(N, dims.D), r(N, dims.H), r(N, dims.D), r(N, 3)
    W, b = 0.3 * r(dims.H, 20), 0.3 * r(dims.H)
    freq = (torch.pi * torch.arange(1, 21) / 5.0).float().to(DEV)
    out = ops.edge_message_bwdbwd_raw(g, dims, pos, s, v, W, b, freq, gx, gV, a_s if a_sv else None, a_v if a_sv else None,
                                      a_p if a_pos else None)
    w1 = ops.edge_message_bwd_raw(g, dims, pos, s, v, W, b, freq, gx, gV, need_w=True)[3:]
    torch.cuda.synchronize()
    return list(out) + list(w1), dict(s=s, v=v, gx=gx, gV=gV, a_s=a_s, a_v=a_v, a_p=a_p, W=W, b=b, freq=freq, dims=dims)


@pytest.mark.parametrize("sizes", [(5, 40), (22, 25)])
@pytest.mark.parametrize("cfg", [orc.CONFIG_DEFAULT, orc.CONFIG_C4], ids=["c128", "c256"])
def test_second_order_tiles_equal_blocks_and_repeat_bitwise(sizes, cfg):
    """K2bb (JVP = two forward instances, reverse pass = two launches) and the weight-gradient kernels (one MMA-issuing warp
    serving three groups in a fixed order): molecule tiles == edge-block tiles for everything computed per row, and every
    output -- the weight gradients included, whose summation order depends on the tile walk -- repeats bit for bit."""
    d = orc.make_molecule_batch(16, sizes, seed=sizes[0], with_edges=False)
    pos, batch, ptr = d["pos"].to(DEV), d["batch"].to(DEV), d["ptr"].to(DEV)
    g_mol, _, _ = build_graph(pos, cfg.cutoff, ptr=ptr, batch=batch)
    g_blk = graph_from_edge_index(g_mol.edge_index(), g_mol.n_nodes, g_mol.n_graphs, batch=batch)
    a, _ = _second_order(g_mol, cfg, pos)
    a2, _ = _second_order(g_mol, cfg, pos)
    b, _ = _second_order(g_blk, cfg, pos)
    assert all(torch.equal(x, y) for x, y in zip(a, a2)), "second-order outputs are not reproducible"
    names = ["o_gx", "o_gV", "o_s", "o_v", "o_pos"]
    for k, name in enumerate(names):
        if name == "o_pos":  # per-edge records summed per node: same order in both tilings
            assert torch.allclose(a[k], b[k], rtol=0, atol=2e-5 * float(b[k].abs().max())), name
        else:
            assert torch.equal(a[k], b[k]), f"{name}: molecule tiles != edge-block tiles"
    for k in range(5, len(a)):  # weight gradients: different partial-sum order across tilings, fp32 noise only
        scale = float(b[k].abs().max()) + 1e-30
        assert float((a[k] - b[k]).abs().max()) <= 1e-4 * scale  # (frequency gradients: sums with cancellation)


def test_second_order_with_missing_tangents():
    """NULL a_pos (no geometry tangent: the tangent-filter launch is skipped) and NULL a_s / a_v (zero row tangents) give
    what explicit zero tensors give."""
    cfg = orc.CONFIG_DEFAULT
    d = orc.make_molecule_batch(6, (3, 30), seed=2, with_edges=False)
    pos, batch, ptr = d["pos"].to(DEV), d["batch"].to(DEV), d["ptr"].to(DEV)
    g, _, _ = build_graph(pos, cfg.cutoff, ptr=ptr, batch=batch)
    for a_pos, a_sv in ((False, True), (True, False)):
        got, t = _second_order(g, cfg, pos, seed=3, a_pos=a_pos, a_sv=a_sv)
        z = torch.zeros_like
        ref = ops.edge_message_bwdbwd_raw(g, t["dims"], pos, t["s"], t["v"], t["W"], t["b"], t["freq"], t["gx"], t["gV"],
                                          t["a_s"] if a_sv else z(t["a_s"]), t["a_v"] if a_sv else z(t["a_v"]),
                                          t["a_p"] if a_pos else z(t["a_p"]))
        for x, y in zip(got[:8], ref):
            assert torch.isfinite(x).all()
            assert torch.allclose(x, y, rtol=0, atol=1e-6 * (float(y.abs().max()) + 1e-30))
