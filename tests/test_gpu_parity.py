"""GPU parity tests (run on the B200 box): every CUDA entry point, called through the C ABI,
against the oracle on seeded inputs and against the golden vectors produced by the reference.

Tolerances: edge lists bit-exact after canonical sort; energies 1e-5 relative and forces
1e-4 eV/A absolute in fp32 (north_star), judged against the fp64 oracle with the fp32
noise-floor rule of SURVEY.md section 7 where the reference's own fp32 run exceeds 1e-4."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import cast_data, embed_table, force_gate, grad_digest, load_golden
from oracle import xpainn_oracle as orc

pytestmark = pytest.mark.gpu

import xequinet_b200 as xb  # noqa: E402
from xequinet_b200 import keys, ops  # noqa: E402
from xequinet_b200.graph import build_graph, graph_from_edge_index  # noqa: E402

DEV = "cuda"


def _dev(data):
    return {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in data.items()}


# ---------------------------------------------------------------------------------------
# K1 radius graph
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_mol,atoms,seed", [(64, 18, 0), (7, (1, 40), 3), (1, 200, 5), (3, 1, 1)])
def test_radius_graph_matches_oracle(n_mol, atoms, seed):
    d = orc.make_molecule_batch(n_mol, atoms, seed=seed)
    ei_ref, _ = orc.canonical_sort(d["edge_index"])
    ei = xb.radius_graph(d["pos"].to(DEV), 5.0, batch=d["batch"].to(DEV))
    assert ei.dtype == torch.int64
    assert torch.equal(ei.cpu(), ei_ref)  # K1 emits canonical order directly
    g, _, _ = build_graph(d["pos"].to(DEV), 5.0, batch=d["batch"].to(DEV))
    assert torch.equal(g.edge_index().cpu(), ei_ref)
    # transposed structure: slots grouped by neighbor, ordered by edge id
    order = torch.sort(ei_ref[1], stable=True)[1]
    assert torch.equal(g.t_eid[: g.n_edges].cpu().long(), order)
    assert torch.equal(g.t_row[: g.n_edges].cpu().long(), ei_ref[0][order])


def test_radius_graph_empty_and_isolated():
    pos = torch.tensor([[0.0, 0, 0], [100.0, 0, 0], [200.0, 0, 0]], device=DEV)
    ei = xb.radius_graph(pos, 5.0)
    assert ei.shape == (2, 0)
    ei = xb.radius_graph(torch.tensor([[0.0, 0, 0], [1.0, 0, 0], [50.0, 0, 0]], device=DEV), 5.0)
    assert ei.cpu().tolist() == [[0, 1], [1, 0]]


@pytest.mark.parametrize("name", ["pbc_small", "pbc_tiny", "pbc_slab", "pbc_two_graphs"])
def test_radius_graph_pbc_matches_reference_golden(name):
    z, cfg, data = load_golden(name)
    n = (data["ptr"][1:] - data["ptr"][:-1])
    ei, co = xb.radius_graph_pbc(data["pos"].to(DEV), n.to(DEV), data["pbc"].to(DEV), data["cell"].to(DEV), 5.0)
    assert torch.equal(ei.cpu(), data["edge_index"])
    assert torch.equal(co.cpu(), data["cell_offsets"])


def test_radius_graph_pbc_water_golden():
    z = np.load(__import__("helpers").GOLDEN / "water_edges.npz")
    pos, cell, pbc = (torch.from_numpy(z[k]).to(DEV) for k in ("pos", "cell", "pbc"))
    ei, co = xb.radius_graph_pbc(pos, torch.tensor([pos.shape[0]], device=DEV), pbc, cell, 5.0)
    assert np.array_equal(ei.cpu().numpy(), z["edge_index"].astype(np.int64))
    assert np.array_equal(co.cpu().numpy().astype(np.int8), z["cell_offsets"])


@pytest.mark.parametrize("shear", [0.0, 0.35])
def test_radius_graph_cell_list_matches_oracle(shear):
    """Large fully periodic box -> the cell-list path of K1 (csrc/radius_graph.cu); bit-exact against the
    restated reference search (data/radius_graph.py:35-192), cubic and triclinic cells, unwrapped input."""
    d = orc.make_water_box(8, seed=3)  # 1536 atoms
    cell = d["cell"].clone()
    if shear:
        cell[0, 1, 0] = shear * cell[0, 0, 0]
        cell[0, 2, 1] = -0.5 * shear * cell[0, 1, 1]
    n = torch.tensor([d["pos"].shape[0]])
    ei_ref, co_ref = orc.radius_graph_pbc(d["pos"], n, d["pbc"], cell, 5.0)
    ei_ref, co_ref = orc.canonical_sort(ei_ref, co_ref)
    ei, co = xb.radius_graph_pbc(d["pos"].to(DEV), n.to(DEV), d["pbc"].to(DEV), cell.to(DEV), 5.0)
    assert torch.equal(ei.cpu(), ei_ref)
    assert torch.equal(co.cpu(), co_ref)


def test_graph_from_unsorted_edge_index():
    d = orc.make_molecule_batch(5, (4, 12), seed=9)
    ei = d["edge_index"]
    perm = torch.randperm(ei.shape[1], generator=torch.Generator().manual_seed(0))
    g = graph_from_edge_index(ei[:, perm].to(DEV), d["pos"].shape[0])
    got = g.edge_index().cpu()
    ref, _ = orc.canonical_sort(ei)
    # rows are grouped by center; within a row the caller's order is kept
    assert torch.equal(torch.sort(got[0] * 10_000 + got[1])[0], ref[0] * 10_000 + ref[1])


# ---------------------------------------------------------------------------------------
# K2 / K2b / K2bb against fp64 autograd of the oracle's edge message
# ---------------------------------------------------------------------------------------
def _edge_case(kind, cfg):
    if kind == "mol":
        d = orc.make_molecule_batch(24, (6, 22), seed=2)
        ei, co, cell = d["edge_index"], None, None
    elif kind == "iso":  # many single atoms (rows and whole tiles without edges) next to small molecules
        d = orc.make_molecule_batch(40, (1, 4), seed=7)
        ei, co, cell = d["edge_index"], None, None
    elif kind == "pbc":
        d = orc.make_small_pbc(14, 6.5, seed=3, triclinic=True)
        ei, co = orc.radius_graph_pbc(d["pos"], torch.tensor([14]), d["pbc"], d["cell"], 5.0)
        cell = d["cell"]
    else:  # two periodic graphs
        z, _, d = load_golden("pbc_two_graphs")
        ei, co, cell = d["edge_index"], d["cell_offsets"], d["cell"]
    N = d["pos"].shape[0]
    g = torch.Generator().manual_seed(4)
    rnd = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    t = dict(pos=d["pos"].double(), s=rnd(N, cfg.H_msg), v=rnd(N, cfg.D), x=rnd(N, cfg.node_dim), V=rnd(N, cfg.D),
             W=0.3 * rnd(cfg.H_msg, cfg.num_basis), b=0.3 * rnd(cfg.H_msg),
             freq=torch.pi * torch.arange(1, cfg.num_basis + 1, dtype=torch.float64) / cfg.cutoff + 0.1 * rnd(cfg.num_basis),
             gx=rnd(N, cfg.node_dim), gV=rnd(N, cfg.D), a_s=rnd(N, cfg.H_msg), a_v=rnd(N, cfg.D), a_pos=rnd(N, 3))
    return d, ei, co, cell, t


def _rel(got, ref):
    ref = ref.double()
    return float((got.double().cpu() - ref).abs().max() / (ref.abs().max() + 1e-30))


@pytest.mark.parametrize("kind", ["mol", "pbc", "pbc2", "iso"])
@pytest.mark.parametrize("cfg", [orc.CONFIG_DEFAULT, orc.CONFIG_C4], ids=["c128", "c256"])
def test_edge_kernels_match_oracle_autograd(kind, cfg):
    d, ei, co, cell, t = _edge_case(kind, cfg)
    N = d["pos"].shape[0]
    G = d["ptr"].numel() - 1
    # fp64 oracle: value, first derivatives, second derivatives
    req = {k: t[k].clone().requires_grad_(True) for k in ("pos", "s", "v", "W", "b", "freq", "gx", "gV")}
    xo, Vo = orc.edge_message(t["x"], t["V"], req["s"], req["v"], req["pos"], req["W"], req["b"], req["freq"], ei, cfg,
                              cell.double() if cell is not None else None, co, d["batch"])
    Phi = (req["gx"] * xo).sum() + (req["gV"] * Vo).sum()
    first = torch.autograd.grad(Phi, [req[k] for k in ("s", "v", "pos", "W", "b", "freq")], create_graph=True)
    Psi = (t["a_s"] * first[0]).sum() + (t["a_v"] * first[1]).sum() + (t["a_pos"] * first[2]).sum()
    second = torch.autograd.grad(Psi, [req[k] for k in ("gx", "gV", "s", "v", "pos", "W", "b", "freq")])

    dims = ops.Dims(cfg.node_dim, *cfg.muls, cfg.num_basis, cfg.cutoff)
    graph = graph_from_edge_index(ei.to(DEV), N, G, cell_offsets=co.to(DEV) if co is not None else None,
                                  cell=cell.to(DEV) if cell is not None else None, batch=d["batch"].to(DEV))
    f32 = lambda a: a.float().to(DEV).contiguous()
    cmf = lambda a: orc.to_cm(a, cfg).float().to(DEV).contiguous()
    pos, s, W, b, freq, gx = (f32(t[k]) for k in ("pos", "s", "W", "b", "freq", "gx"))
    v, V, gV, a_v = (cmf(t[k]) for k in ("v", "V", "gV", "a_v"))
    x, a_s, a_pos = f32(t["x"]), f32(t["a_s"]), f32(t["a_pos"])
    tol = 2e-5  # fp32 kernels vs fp64 oracle, relative to the largest entry

    x_out, V_out = ops.edge_message_fwd_raw(graph, dims, pos, s, v, x, V, W, b, freq)
    assert _rel(x_out, xo.detach()) < tol
    assert _rel(orc.from_cm(V_out.cpu(), cfg), Vo.detach()) < tol

    gs, gv, gpos, gW, gb, gf = ops.edge_message_bwd_raw(graph, dims, pos, s, v, W, b, freq, gx, gV, need_w=True)
    for got, ref, name in zip((gs, orc.from_cm(gv.cpu(), cfg), gpos, gW, gb, gf), first, "s v pos W b f".split()):
        assert _rel(got, ref.detach()) < tol, f"first/{name}"
    # force-only variant (no weight gradients) takes a different kernel instantiation
    gs2, gv2, gpos2, *_ = ops.edge_message_bwd_raw(graph, dims, pos, s, v, W, b, freq, gx, gV, need_w=False)
    assert torch.equal(gs2, gs) and torch.equal(gv2, gv) and torch.equal(gpos2, gpos)

    o = ops.edge_message_bwdbwd_raw(graph, dims, pos, s, v, W, b, freq, gx, gV, a_s, a_v, a_pos)
    got = (o[0], orc.from_cm(o[1].cpu(), cfg), o[2], orc.from_cm(o[3].cpu(), cfg), o[4], o[5], o[6], o[7])
    for gt, ref, name in zip(got, second, "gx gV s v pos W b f".split()):
        assert _rel(gt, ref) < 5e-5, f"second/{name}"

    # deterministic: bitwise identical on a second run
    x_out2, V_out2 = ops.edge_message_fwd_raw(graph, dims, pos, s, v, x, V, W, b, freq)
    assert torch.equal(x_out, x_out2) and torch.equal(V_out, V_out2)


def test_layout_convert_roundtrip():
    cfg = orc.CONFIG_DEFAULT
    dims = ops.Dims(cfg.node_dim, *cfg.muls, cfg.num_basis, cfg.cutoff)
    V = torch.randn(33, cfg.D, device=DEV)
    Vc = ops.layout_convert(V, dims, to_cm=True)
    assert torch.equal(Vc.cpu(), orc.to_cm(V.cpu(), cfg))
    assert torch.equal(ops.layout_convert(Vc, dims, to_cm=False), V)


# ---------------------------------------------------------------------------------------
# whole model: energy / forces / parameter gradients
# ---------------------------------------------------------------------------------------
def _model(cfg, seed, train=False):
    model = xb.resolve_model("xpainn", **cfg.model_kwargs())
    model.load_state_dict(orc.synthetic_state_dict(cfg, seed), strict=False)
    model = model.to(DEV)
    return model.train() if train else model.eval()


def _force_gate(F_new, F_ref64, F_ref32):
    """helpers.force_gate: err_new <= max(1e-4, 1.5 * err_ref32) against the fp64 reference (percentile-wise on batches)."""
    force_gate(F_new, F_ref64, F_ref32)


@pytest.mark.parametrize("name", ["mol_small", "mol_c4_small", "pbc_small", "pbc_tiny", "pbc_slab", "pbc_two_graphs"])
def test_model_matches_reference_golden(name):
    z, cfg, data = load_golden(name)
    model = _model(cfg, int(z["sd_seed"]))
    d = _dev(cast_data(data, torch.float32))
    d.pop("pbc", None)
    out = model(d, compute_forces=True)
    assert set(out) == {"energy", "atomic_energies", "forces"}
    E, Ea, Fo = (out[k].detach().cpu().numpy() for k in ("energy", "atomic_energies", "forces"))
    np.testing.assert_allclose(E, z["f64:energy"], rtol=1e-5, atol=1e-6)
    # energies: 1e-5 relative (BASELINE.json north_star), taken against the scale of the atomic energies
    np.testing.assert_allclose(Ea, z["f64:atomic_energies"], rtol=1e-5, atol=1e-5 * np.abs(z["f64:atomic_energies"]).max())
    _force_gate(Fo, z["f64:forces"], z["f32:forces"])
    # direct fp32-vs-fp32 numbers as well (same tolerance as the reference's own noise allows)
    np.testing.assert_allclose(Fo, z["f32:forces"], rtol=0, atol=3e-4)


def test_model_matches_oracle_c1_shape():
    """config c1: 64 molecules x 18 atoms, E+F inference, with K1 building the graph."""
    cfg = orc.CONFIG_DEFAULT
    data = orc.make_molecule_batch(64, 18, seed=0, with_edges=False)
    sd = orc.synthetic_state_dict(cfg, 1234, torch.float64)
    d64 = cast_data(data, torch.float64)
    d64["edge_index"] = orc.radius_graph(data["pos"], 5.0, data["batch"])
    ref = orc.xpainn_energy_forces(sd, embed_table(), d64, cfg)
    sd32 = {k: v.float() for k, v in sd.items()}
    ref32 = orc.xpainn_energy_forces(sd32, embed_table().float(), cast_data(d64, torch.float32), cfg)
    model = _model(cfg, 1234)
    d = xb.NeighborTransform(5.0)(_dev(data))
    assert torch.equal(d["edge_index"].cpu(), orc.canonical_sort(d64["edge_index"])[0])
    out = model(d, compute_forces=True)
    np.testing.assert_allclose(out["energy"].detach().cpu().numpy(), ref["energy"].detach().numpy(), rtol=1e-5, atol=1e-6)
    _force_gate(out["forces"].cpu().numpy(), ref["forces"].numpy(), ref32["forces"].numpy())
    # physics: net force per molecule vanishes
    net = torch.zeros(64, 3, device=DEV).index_add(0, d["batch"], out["forces"])
    assert float(net.abs().max()) < 5e-5


def _oracle_param_grads(cfg, seed, data, tE, tF, use_forces, dtype):
    sd = {k: v.requires_grad_(True) for k, v in orc.synthetic_state_dict(cfg, seed, dtype).items()}
    out = orc.xpainn_energy_forces(sd, embed_table().to(dtype), cast_data(data, dtype), cfg, create_graph=True)
    loss = F.smooth_l1_loss(out["energy"], tE.to(dtype))
    if use_forces:
        loss = loss + 100.0 * F.smooth_l1_loss(out["forces"], tF.to(dtype))
    loss.backward()
    return float(loss.detach()), {k: v.grad.double() for k, v in sd.items() if v.grad is not None}


def _rel_l2(a, b):
    n = float(b.norm())
    return float((a.reshape(-1) - b.reshape(-1)).norm()) / n if n > 0 else float(a.abs().max())


def _case_data(name):
    if name == "mol_tiny":
        data = orc.make_molecule_batch(2, (6, 8), seed=13)
        g = torch.Generator().manual_seed(99)
        tE = torch.randn(2, generator=g, dtype=torch.float64)
        tF = torch.randn(data["pos"].shape[0], 3, generator=g, dtype=torch.float64)
        return orc.CONFIG_DEFAULT, data, tE, tF, 1234
    z, cfg, data = load_golden(name)
    data = dict(data)
    data.pop("pbc", None)
    return cfg, data, torch.from_numpy(z["f64:target_energy"]), torch.from_numpy(z["f64:target_forces"]), int(z["sd_seed"])


@pytest.mark.parametrize("name", ["mol_small", "pbc_small", "mol_tiny"])
@pytest.mark.parametrize("use_forces", [False, True], ids=["E-loss", "EF-loss"])
def test_param_grads_match_reference(name, use_forces):
    """Training semantics (utils/trainer.py:295-302): loss.backward() through forces -> K2bb.
    Every parameter gradient is compared with the fp64 oracle (itself pinned to the reference's
    golden digests in tests/test_oracle_golden.py).

    The double backward through Invariant's sqrt(q + 1e-10) (nn/o3layer.py:44) is ill-conditioned
    in fp32 whenever some |W| scalar falls below ~1e-4: the reference's own fp32 run is then
    1e-2..1e-1 away from its fp64 run (measured in scratch/scan_seeds.py).
    So the test first looks for a weight seed for which the reference itself is well conditioned
    in fp32 (err_ref32 < 5e-4) and demands 2e-3 there; if the input admits none, it falls back
    to the noise-floor gate err_new <= max(2e-3, 5 * err_ref32)."""
    cfg, data, tE, tF, golden_seed = _case_data(name)
    chosen = None
    for seed in [golden_seed, 1, 4, 5, 2, 6, 7]:
        loss64, g64 = _oracle_param_grads(cfg, seed, data, tE, tF, use_forces, torch.float64)
        _, g32 = _oracle_param_grads(cfg, seed, data, tE, tF, use_forces, torch.float32)
        err_ref = {k: _rel_l2(g32[k], g64[k]) for k in g64}
        if chosen is None:
            chosen = (seed, loss64, g64, err_ref)  # golden seed: the fallback
        if max(err_ref.values()) < 5e-4:
            chosen = (seed, loss64, g64, err_ref)
            break
    seed, loss64, g64, err_ref = chosen
    well_conditioned = max(err_ref.values()) < 5e-4
    print(f"{name}: seed {seed} (golden seed {golden_seed}), reference fp32-vs-fp64 worst {max(err_ref.values()):.2e}, tight={well_conditioned}")
    if seed != golden_seed:  # the golden seed is always tested too, with the noise-floor gate
        _check_param_grads(name, cfg, data, tE, tF, use_forces, golden_seed, tight=False)
    _check_param_grads(name, cfg, data, tE, tF, use_forces, seed, tight=well_conditioned, ref=(loss64, g64, err_ref))
    if not use_forces:
        assert well_conditioned  # first-order training gradients are always well conditioned


def _check_param_grads(name, cfg, data, tE, tF, use_forces, seed, tight, ref=None):
    if ref is None:
        loss64, g64 = _oracle_param_grads(cfg, seed, data, tE, tF, use_forces, torch.float64)
        _, g32 = _oracle_param_grads(cfg, seed, data, tE, tF, use_forces, torch.float32)
        err_ref = {k: _rel_l2(g32[k], g64[k]) for k in g64}
    else:
        loss64, g64, err_ref = ref
    well_conditioned = tight

    model = _model(cfg, seed, train=True)
    d = _dev(cast_data(data, torch.float32))
    out = model(d, compute_forces=use_forces)
    loss = F.smooth_l1_loss(out["energy"], tE.float().to(DEV))
    if use_forces:
        loss = loss + 100.0 * F.smooth_l1_loss(out["forces"], tF.float().to(DEV))
    np.testing.assert_allclose(loss.item(), loss64, rtol=2e-5)
    loss.backward()
    checked = 0
    for k, p in model.named_parameters():
        if k not in g64:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        err_new = _rel_l2(p.grad.detach().cpu().double(), g64[k])
        bound = 2e-3 if well_conditioned else max(2e-3, 5.0 * err_ref[k])
        assert err_new <= bound, (seed, k, err_new, err_ref[k])
        checked += 1
    assert checked > 50


# ---------------------------------------------------------------------------------------
# capacity mode (CUDA-graph replay / MD loops): same structure, edge count stays on the device
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("periodic", [False, True])
def test_static_graph_builder_matches_dynamic(periodic):
    from xequinet_b200.graph import StaticGraphBuilder

    if periodic:
        datas = [orc.make_small_pbc(14, 6.5, seed=s, triclinic=True) for s in (3, 4)]
    else:
        datas = [orc.make_aspirin_batch(6, seed=s, with_edges=False) for s in (0, 1)]
    d0 = _dev(datas[0])
    builder = StaticGraphBuilder(d0["pos"].shape[0], d0["ptr"], 5.0, edge_capacity=4096 if periodic else 3000,
                                 cell=d0.get("cell"), pbc=d0.get("pbc"))
    model = _model(orc.CONFIG_DEFAULT, 1234)
    for data in datas:  # the same builder (and buffers) serves every structure of that shape
        d = _dev(data)
        g_dyn, _, _ = build_graph(d["pos"], 5.0, ptr=d["ptr"], batch=d["batch"], cell=d.get("cell"), pbc=d.get("pbc"))
        g = builder.build(d["pos"])
        E = g_dyn.n_edges
        assert int(builder.overflow.item()) == 0 and int(g.rowptr[-1].item()) == E
        assert torch.equal(g.rowptr, g_dyn.rowptr) and torch.equal(g.col[:E], g_dyn.col[:E])
        assert torch.equal(g.t_rowptr, g_dyn.t_rowptr) and torch.equal(g.t_eid[:E], g_dyn.t_eid[:E])
        if periodic:
            assert torch.equal(g.offsets[:E], g_dyn.offsets[:E])
        outs = []
        for graph in (g_dyn, g):
            dd = {k: v for k, v in d.items() if k not in ("pbc",)}
            dd[keys.GRAPH] = graph
            dd["pos"] = d["pos"].clone()
            outs.append(model(dd, compute_forces=True))
        assert torch.equal(outs[0]["energy"], outs[1]["energy"]) and torch.equal(outs[0]["forces"], outs[1]["forces"])
    # overflow: raised on the host outside graph capture ...
    small = StaticGraphBuilder(d0["pos"].shape[0], d0["ptr"], 5.0, edge_capacity=64, cell=d0.get("cell"), pbc=d0.get("pbc"))
    with pytest.raises(RuntimeError, match="edge_capacity"):
        small.build(d0["pos"])
    # ... and memory-safe on the device when nobody looks (capture / check_overflow=False): the flag is raised, rowptr is
    # clamped to the capacity, and the edge kernels walk the truncated list without leaving the arrays
    g = small.build(d0["pos"], check_overflow=False)
    assert int(small.overflow.item()) == 1
    assert int(g.rowptr.max().item()) <= 64 and int(g.rowptr[-1].item()) == 64
    assert g.edge_index().shape[1] == 64
    dd = {k: v for k, v in d0.items() if k not in ("pbc",)}
    dd[keys.GRAPH] = g
    dd["pos"] = d0["pos"].clone()
    out = model(dd, compute_forces=True)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out["energy"]).all()) and bool(torch.isfinite(out["forces"]).all())
    with pytest.raises(ValueError):
        small.build(d0["pos"].double())


def test_stale_graph_cache_is_not_trusted():
    """nn/basic.py:67: `edge_index` in the dict is the source of truth.  A dict that went through the model once holds
    the derived CSR structure; replacing positions + edge_index (reference-style MD without this package's
    NeighborTransform) must give the energies of the NEW list, not of the cached one."""
    cfg = orc.CONFIG_DEFAULT
    model = _model(cfg, 1234)
    a = orc.make_molecule_batch(3, 9, seed=1)
    b = orc.make_molecule_batch(3, 9, seed=2)
    d = _dev({k: a[k] for k in ("pos", "atomic_numbers", "batch", "ptr", "edge_index")})
    e_a = model(d, compute_forces=True)["energy"].detach().clone()
    assert keys.GRAPH in d
    # a new structure with its own edge list, but the derived structure of the OLD list still in the dict (the other
    # entries the model wrote back -- atomic_energies accumulate, nn/output.py:120-123 -- are dropped as a caller would)
    d2 = _dev({k: b[k] for k in ("pos", "atomic_numbers", "batch", "ptr", "edge_index")})
    d2[keys.GRAPH] = d[keys.GRAPH]
    e_b = model(d2, compute_forces=True)["energy"].detach().clone()
    assert d2[keys.GRAPH] is not d[keys.GRAPH]
    fresh = _dev({k: b[k] for k in ("pos", "atomic_numbers", "batch", "ptr", "edge_index")})
    e_ref = model(fresh, compute_forces=True)["energy"].detach()
    assert torch.equal(e_b, e_ref) and not torch.equal(e_a, e_b)


# ---------------------------------------------------------------------------------------
# the two implementations of the filter contraction (tcgen05 / SIMT) agree
# ---------------------------------------------------------------------------------------
_SIMT_SCRIPT = r"""
import sys, torch
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
from test_gpu_parity import _edge_outputs
torch.save(_edge_outputs({n_mol}, {staged}), {out!r})
"""


def _edge_outputs(n_mol, staged):
    """All edge-kernel outputs (values, first and second derivatives) on a batch of molecules of 1..21 atoms
    (single atoms = rows and tiles without edges); `staged` selects molecule tiles (shared-memory row window)
    or edge-block tiles (direct gathers)."""
    cfg = orc.CONFIG_DEFAULT
    d = orc.make_molecule_batch(n_mol, (1, 21), seed=11, with_edges=False)
    kw = dict(ptr=d["ptr"].to(DEV)) if staged else {}
    g, _, _ = build_graph(d["pos"].to(DEV), cfg.cutoff, batch=d["batch"].to(DEV), **kw)
    N = g.n_nodes
    dims = ops.Dims(cfg.node_dim, *cfg.muls, cfg.num_basis, cfg.cutoff)
    gen = torch.Generator().manual_seed(5)
    r = lambda *s: torch.randn(*s, generator=gen).to(DEV)
    pos = d["pos"].to(DEV)
    s, v, x, V = r(N, dims.H), r(N, dims.D), r(N, dims.node_dim), r(N, dims.D)
    W, b = 0.3 * r(dims.H, 20), 0.3 * r(dims.H)
    freq = (torch.pi * torch.arange(1, 21) / 5.0).float().to(DEV)
    gx, gV, a_s, a_v, a_p = r(N, dims.node_dim), r(N, dims.D), r(N, dims.H), r(N, dims.D), r(N, 3)
    out = list(ops.edge_message_fwd_raw(g, dims, pos, s, v, x, V, W, b, freq))
    out += list(ops.edge_message_bwd_raw(g, dims, pos, s, v, W, b, freq, gx, gV, need_w=True))
    out += list(ops.edge_message_bwdbwd_raw(g, dims, pos, s, v, W, b, freq, gx, gV, a_s, a_v, a_p))
    torch.cuda.synchronize()
    return [o.cpu() for o in out]


@pytest.mark.parametrize("staged", [True, False], ids=["molecule-tiles", "edge-block-tiles"])
def test_tcgen05_and_simt_edge_kernels_agree(staged, tmp_path):
    """The tensor-core kernels of the product library (edge_fwd_ul.cu, edge_bwd_ul.cu, edge_bwd2_ul.cu, edge_wgrad_ul.cu) against an
    independent implementation: the round-1 SIMT kernels, compiled only into the test-only libxeq_b200_simt.so and run
    in a subprocess (XEQ_LIB + XEQ_EDGE_SIMT=1).  Same results to fp32 round-off on every output."""
    import os, subprocess, sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    ref_file = tmp_path / "simt.pt"
    script = _SIMT_SCRIPT.format(root=str(root), tests=str(root / "tests"), n_mol=40, staged=staged, out=str(ref_file))
    simt_lib = root / "xequinet_b200" / "libxeq_b200_simt.so"  # test-only build (xequinet_b200/build.py --simt)
    assert simt_lib.exists(), "build the test-only SIMT variant: python xequinet_b200/build.py --simt"
    env = dict(os.environ, XEQ_EDGE_SIMT="1", XEQ_LIB=str(simt_lib))
    subprocess.run([sys.executable, "-c", script], check=True, env=env, timeout=300)
    ref = torch.load(ref_file)
    got = _edge_outputs(40, staged)
    assert len(got) == len(ref) == 16
    for i, (a, bref) in enumerate(zip(got, ref)):
        assert _rel(a, bref) < 1e-5, f"output {i}"


# ---------------------------------------------------------------------------------------
# Verlet-skin neighbour-list reuse (MD loops)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("periodic", [False, True])
def test_skin_neighbor_list_gives_exact_energies_and_forces(periodic):
    """SkinNeighborTransform keeps a list built with cutoff + skin while atoms move less than skin / 2; edges
    beyond the cutoff contribute exactly zero, so E and F equal those of the exact list at every step."""
    cfg = orc.CONFIG_DEFAULT
    torch.manual_seed(0)
    if periodic:
        base = orc.make_water_box(4, seed=3)  # 192 atoms, L = 12.4 A
    else:
        base = orc.make_molecule_batch(6, (8, 20), seed=9, with_edges=False)
    model = xb.resolve_model("xpainn", **cfg.model_kwargs())
    model.load_state_dict(orc.synthetic_state_dict(cfg, 1234), strict=False)
    model = model.to(DEV).eval()
    exact = xb.NeighborTransform(cfg.cutoff)
    skin = xb.SkinNeighborTransform(cfg.cutoff, skin=1.0)
    pos0 = base["pos"].to(DEV)
    g = torch.Generator().manual_seed(1)
    step = 0.12 * torch.nn.functional.normalize(torch.randn(pos0.shape, generator=g), dim=-1).to(DEV)
    keys_in = [k for k in ("atomic_numbers", "batch", "ptr", "cell", "pbc") if k in base]
    for it in range(6):  # cumulative displacement 0 .. 0.6 A: crosses skin / 2 once
        pos = pos0 + it * step
        outs = []
        for tr in (exact, skin):
            d = {k: base[k].to(DEV) for k in keys_in}
            d["pos"] = pos.clone()
            outs.append(model(tr(d), compute_forces=True))
        e_ref, e_got = outs[0]["energy"].detach(), outs[1]["energy"].detach()
        assert float((e_got - e_ref).abs().max()) <= 2e-5 * max(1.0, float(e_ref.abs().max())), it
        assert float((outs[1]["forces"] - outs[0]["forces"]).abs().max()) < 1e-4, it
    assert skin.n_calls == 6 and 1 < skin.n_builds < 6


# ---------------------------------------------------------------------------------------
# virial through the strain trick (nn/basic.py:93-107, 162-199)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["mol_small", "pbc_small", "pbc_slab", "pbc_two_graphs"])
def test_virial_matches_reference_golden(name):
    """model(data, compute_forces=True, compute_virial=True): virial = -dE/dstrain against the reference's own
    fp64 virial (tests/golden/virial.npz); the forces of the same pass are those of the plain force pass."""
    from helpers import GOLDEN

    z, cfg, data = load_golden(name)
    ref = np.load(GOLDEN / "virial.npz")[f"{name}:virial"]
    model = _model(cfg, int(z["sd_seed"]))
    d = _dev(cast_data(data, torch.float32))
    d.pop("pbc", None)
    out = model(dict(d), compute_forces=True, compute_virial=True)
    assert set(out) == {"energy", "atomic_energies", "forces", "virial"}
    vir = out["virial"].detach().cpu().numpy()
    scale = max(1.0, float(np.abs(ref).max()))
    assert np.abs(vir - ref).max() <= 3e-4 * scale, (np.abs(vir - ref).max(), scale)
    np.testing.assert_allclose(vir, np.swapaxes(vir, 1, 2), atol=1e-6 * scale)  # symmetrised strain
    _force_gate(out["forces"].detach().cpu().numpy(), z["f64:forces"], z["f32:forces"])
    np.testing.assert_allclose(out["energy"].detach().cpu().numpy(), z["f64:energy"], rtol=1e-5, atol=1e-6)
    # virial alone (no forces requested)
    out2 = model(dict(d), compute_forces=False, compute_virial=True)
    assert "forces" not in out2
    np.testing.assert_allclose(out2["virial"].detach().cpu().numpy(), vir, rtol=0, atol=1e-6 * scale)


@pytest.mark.parametrize("name", ["mol_small", "pbc_small", "pbc_two_graphs"])
def test_virial_in_a_training_loss(name):
    """A virial term in the training loss (nn/basic.py:162-199 with training=True): loss.backward() differentiates the
    forces AND the strain derivative again -- for periodic structures that is the second derivative of the cell
    gradient (K2bb with the lattice tangent a_cell).  Every parameter gradient against the fp64 oracle."""
    z, cfg, data = load_golden(name)
    data = dict(data)
    data.pop("pbc", None)
    seed = int(z["sd_seed"])
    g = torch.Generator().manual_seed(7)
    G = data["ptr"].numel() - 1
    tV = torch.randn(G, 3, 3, generator=g, dtype=torch.float64)

    def loss_of(out, dtype):
        return (out["energy"].sum() + 0.5 * (out["forces"] ** 2).sum() + ((out["virial"] - tV.to(out["virial"])) ** 2).sum())

    grads = {}
    for dtype in (torch.float64, torch.float32):
        sd = {k: v.requires_grad_(True) for k, v in orc.synthetic_state_dict(cfg, seed, dtype).items()}
        out = orc.xpainn_energy_forces(sd, embed_table().to(dtype), cast_data(data, dtype), cfg, create_graph=True, compute_virial=True)
        loss = loss_of(out, dtype)
        loss.backward()
        grads[dtype] = (float(loss.detach()), {k: v.grad.double() for k, v in sd.items() if v.grad is not None})
    loss64, g64 = grads[torch.float64]
    _, g32 = grads[torch.float32]

    model = _model(cfg, seed, train=True)
    d = _dev(cast_data(data, torch.float32))
    out = model(dict(d), compute_forces=True, compute_virial=True)
    loss = out["energy"].sum() + 0.5 * (out["forces"] ** 2).sum() + ((out["virial"] - tV.float().to(DEV)) ** 2).sum()
    np.testing.assert_allclose(loss.item(), loss64, rtol=5e-5)
    loss.backward()
    checked = 0
    for k, p in model.named_parameters():
        if k not in g64:
            continue
        err_new, err_ref = _rel_l2(p.grad.detach().cpu().double(), g64[k]), _rel_l2(g32[k], g64[k])
        assert err_new <= max(2e-3, 5.0 * err_ref), (k, err_new, err_ref)
        checked += 1
    assert checked > 50
