"""CPU tests of the host-side mirror of the reference's module API: state_dict layout,
hyper-parameter handling and the node-level (cm-layout) blocks against the oracle."""
import pytest
import torch

import xequinet_b200 as xb
from oracle import xpainn_oracle as orc
from xequinet_b200.nn import cm
from xequinet_b200.nn.irreps import parse_irreps
from xequinet_b200.nn.layers import EquivariantLayerNorm, Invariant, O3Linear
from xequinet_b200.nn.xpainn import XPainnUpdate
from xequinet_b200 import keys


def test_parse_irreps():
    assert parse_irreps("128x0e + 64x1o + 32x2e") == (128, 64, 32)
    assert parse_irreps("256x0e+128x1o+64x2e") == (256, 128, 64)
    assert parse_irreps([(8, "0e"), (4, "1o")]) == (8, 4, 0)
    with pytest.raises(NotImplementedError):
        parse_irreps("8x0e + 4x3o")


@pytest.mark.parametrize("cfg", [orc.CONFIG_DEFAULT, orc.CONFIG_C4])
def test_state_dict_layout_matches_reference(cfg):
    """Parameter names/shapes of SURVEY.md 8a (S0), as observed on the reference model."""
    model = xb.resolve_model("xpainn", **cfg.model_kwargs())
    sd = model.state_dict()
    spec = orc.state_dict_spec(cfg)
    for name, shape, _ in spec:
        assert name in sd, name
        assert tuple(sd[name].shape) == shape, name
    learnable = {n for n, _ in model.named_parameters()}
    assert learnable == {n for n, _, _ in spec}
    assert sum(p.numel() for p in model.parameters()) == (865141 if cfg is orc.CONFIG_DEFAULT else 3339861)
    # e3nn-internal buffers keep their reference names
    for k in ("mods.message_0.o3norm.scalar_index", "mods.update_0.update_U.output_mask",
              "mods.update_0.invariant.tp.weight", "mods.message_0.rsh_conv.output_mask",
              "mods.embedding.embedding.0.embed_ten"):
        assert k in sd
    res = model.load_state_dict(orc.synthetic_state_dict(cfg), strict=False)
    assert not res.unexpected_keys


@pytest.mark.parametrize("name,cfg", [("default", orc.CONFIG_DEFAULT), ("c4", orc.CONFIG_C4)])
def test_state_dict_strict_round_trip(name, cfg):
    """tests/golden/state_dict_keys.json = the (name, shape, dtype) list of the REAL reference model's state_dict, in
    its own order (oracle/make_golden_state_dict.py; 137 entries incl. the e3nn-internal buffers).  A state_dict of
    exactly that shape loads with strict=True, and what this model saves has exactly those entries."""
    import json
    from pathlib import Path

    spec = json.loads((Path(__file__).resolve().parent / "golden" / "state_dict_keys.json").read_text())[name]
    assert len(spec) == 137
    g = torch.Generator().manual_seed(0)
    ref_shaped = {}
    for k, shape, dtype in spec:
        dt = getattr(torch, dtype)
        ref_shaped[k] = (torch.randn(*shape, generator=g, dtype=dt) if dt.is_floating_point
                         else torch.arange(shape[0], dtype=dt) if shape else torch.zeros((), dtype=dt))
    model = xb.resolve_model("xpainn", **cfg.model_kwargs())
    res = model.load_state_dict(ref_shaped, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    saved = model.state_dict()
    assert [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in saved.items()] == spec
    for k, v in saved.items():
        assert torch.equal(v, ref_shaped[k]), k
    twin = xb.resolve_model("xpainn", **cfg.model_kwargs())
    twin.load_state_dict(saved, strict=True)


def test_defaults_and_unsupported_options():
    m = xb.resolve_model("xpainn")
    assert m.cutoff_radius == 5.0 and list(m.mods)[:3] == ["embedding", "message_0", "update_0"]
    assert list(m.mods)[-1] == "output_energy" and m.extra_properties == ["energy", "atomic_energies"]
    with pytest.raises(NotImplementedError):
        xb.resolve_model("painn")
    with pytest.raises(NotImplementedError):
        xb.resolve_model("xpainn", rbf_kernel="gaussian")


@pytest.mark.gpu
def test_cm_blocks_match_oracle():
    """EquivariantLayerNorm / Invariant / gate expansion on the cm layout (fused kernels) vs the oracle."""
    cfg = orc.XPaiNNConfig(node_dim=32, muls=(32, 32, 32))
    g = torch.Generator().manual_seed(0)
    N = 70
    V = torch.randn(N, cfg.D, generator=g, dtype=torch.float64)  # e3nn layout
    Vc = orc.to_cm(V, cfg).float().cuda()
    gam, bet = torch.randn(cfg.M, generator=g, dtype=torch.float64), torch.randn(32, generator=g, dtype=torch.float64)
    ln = EquivariantLayerNorm(cfg.muls)
    ln.affine_weight.data, ln.affine_bias.data = gam.float(), bet.float()
    ln = ln.cuda()
    back = lambda t: orc.from_cm(t.detach().cpu().double(), cfg)
    torch.testing.assert_close(back(ln(Vc)), orc.equivariant_layer_norm(V, gam, bet, cfg), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(Invariant(cfg.muls)(Vc).cpu().double(), orc.invariant(V, cfg), rtol=1e-5, atol=1e-5)
    gate = torch.randn(N, cfg.M, generator=g, dtype=torch.float64)
    torch.testing.assert_close(orc.from_cm(cm.expand_gate(gate, cfg.muls) * Vc.cpu().double(), cfg), orc.expand_gate(gate, cfg) * V)


def test_dense_blocks_have_no_cpu_path():
    """The dense contractions (K3) run on the tcgen05 kernels only: host tensors must raise."""
    lin = O3Linear((16, 8, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lin(torch.randn(3, 16 + 24 + 20))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        EquivariantLayerNorm((32, 32, 32))(torch.randn(3, 32 * 9))
    from xequinet_b200.nn.layers import Linear
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Linear(8, 8)(torch.randn(3, 8))


@pytest.mark.gpu
def test_o3_linear_matches_oracle():
    cfg = orc.XPaiNNConfig(node_dim=32, muls=(32, 32, 32))
    g = torch.Generator().manual_seed(0)
    V = torch.randn(300, cfg.D, generator=g, dtype=torch.float64)  # e3nn layout
    lin = O3Linear(cfg.muls)
    lin.bias.data = torch.randn(32, generator=g)
    ref = orc.o3_linear(V, lin.weight.double(), lin.bias.double(), cfg)
    out = lin.cuda()(orc.to_cm(V, cfg).float().cuda())
    torch.testing.assert_close(orc.from_cm(out.cpu().double(), cfg), ref, rtol=0, atol=2e-5)


@pytest.mark.gpu
def test_update_block_matches_oracle_layer():
    """XPainnUpdate (cm layout, tcgen05 GEMMs) against the oracle's update math in fp64."""
    cfg = orc.XPaiNNConfig(node_dim=32, muls=(32, 32, 32), action_blocks=1)
    sd = orc.synthetic_state_dict(cfg, 7, torch.float64)
    upd = XPainnUpdate(node_dim=32, node_irreps=cfg.irreps_str)
    upd.load_state_dict({k[len("mods.update_0."):]: v.float() for k, v in sd.items() if k.startswith("mods.update_0.")}, strict=False)
    upd = upd.cuda()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(5, 32, generator=g, dtype=torch.float64)
    V = torch.randn(5, cfg.D, generator=g, dtype=torch.float64)
    out = upd({keys.NODE_INVARIANT: x.float().cuda(), keys.NODE_EQUIVARIANT: orc.to_cm(V, cfg).float().cuda()})
    out = {k: v.detach().cpu().double() for k, v in out.items()}
    import torch.nn.functional as F
    p = "mods.update_0."
    xn = F.layer_norm(x, (32,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    vn = orc.equivariant_layer_norm(V, sd[p + "o3norm.affine_weight"], sd[p + "o3norm.affine_bias"], cfg)
    U = orc.o3_linear(vn, sd[p + "update_U.weight"], sd[p + "update_U.bias"], cfg)
    W = orc.o3_linear(vn, sd[p + "update_V.weight"], sd[p + "update_V.bias"], cfg)
    a = F.linear(F.silu(F.linear(torch.cat([xn, orc.invariant(W, cfg)], 1), sd[p + "update_mlp.0.weight"], sd[p + "update_mlp.0.bias"])),
                 sd[p + "update_mlp.2.weight"], sd[p + "update_mlp.2.bias"])
    M, C = cfg.M, 32
    x_ref = x + a[:, M:M + C] * F.linear(orc.irrep_dot(U, W, cfg), sd[p + "dot_lin.weight"]) + a[:, M + C:]
    V_ref = V + U * orc.expand_gate(a[:, :M], cfg)
    torch.testing.assert_close(out[keys.NODE_INVARIANT], x_ref, rtol=1e-5, atol=2e-5)
    torch.testing.assert_close(orc.from_cm(out[keys.NODE_EQUIVARIANT], cfg), V_ref, rtol=1e-5, atol=2e-5)


def test_skin_neighbor_list_rebuild_policy():
    """Verlet-skin reuse (graph.SkinNeighborTransform): the rebuild decision is pure tensor logic, exercised here
    with a stubbed builder (K1 itself needs a GPU)."""
    import torch
    from xequinet_b200 import keys
    from xequinet_b200.graph import SkinNeighborTransform

    class Stub(SkinNeighborTransform):
        def _build(self, data):
            out = dict(data)
            out[keys.EDGE_INDEX] = torch.zeros(2, 0, dtype=torch.long)
            out[keys.GRAPH] = ("graph", self.n_builds)
            return out

    tr = Stub(5.0, skin=1.0)
    torch.manual_seed(5)
    pos = torch.randn(10, 3)
    cell = (torch.eye(3) * 20.0).reshape(1, 3, 3)
    d = tr({keys.POSITIONS: pos.clone(), keys.CELL: cell.clone(), keys.PBC: torch.ones(1, 3, dtype=torch.bool)})
    assert tr.n_builds == 1 and d[keys.GRAPH] == ("graph", 0)
    # small displacements (< skin / 2): the list is kept
    moved = pos + 0.2 * torch.nn.functional.normalize(torch.randn(10, 3), dim=-1)
    d = tr({keys.POSITIONS: moved, keys.CELL: cell.clone(), keys.PBC: torch.ones(1, 3, dtype=torch.bool)})
    assert tr.n_builds == 1 and d[keys.GRAPH] == ("graph", 0) and tr.n_calls == 2
    # one atom beyond skin / 2: rebuild, and the reference positions move with it
    far = moved.clone()
    far[3] = pos[3] + torch.tensor([0.6, 0.0, 0.0])  # 0.6 A from where the list was built
    assert tr.needs_rebuild({keys.POSITIONS: far, keys.CELL: cell})
    tr({keys.POSITIONS: far, keys.CELL: cell.clone()})
    assert tr.n_builds == 2
    assert not tr.needs_rebuild({keys.POSITIONS: far + 0.1, keys.CELL: cell})  # |(0.1, 0.1, 0.1)| = 0.17 < 0.5
    # a changed cell or a different number of atoms forces a rebuild
    assert tr.needs_rebuild({keys.POSITIONS: far, keys.CELL: cell * 1.01})
    assert tr.needs_rebuild({keys.POSITIONS: far[:9], keys.CELL: cell})
    with pytest.raises(ValueError):
        SkinNeighborTransform(5.0, skin=-1.0)
