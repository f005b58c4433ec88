"""CPU test of the HOST logic of the read-out heads / conditioning modules (slicing of the cm layout, weight-block
bookkeeping, padding of narrow Linear layers, per-graph sums): the kernel entry points are replaced by torch
stand-ins (test-only monkeypatching -- the product ops have no CPU path) and the modules are fed the oracle's
features; results against the oracle's restatement, itself pinned to the reference (tests/test_oracle_heads.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import GOLDEN, cast_data, embed_table, load_golden
from oracle import xpainn_oracle as orc

MODES = ["energy", "scalar", "charges", "dipole", "polar", "spatial"]


@pytest.fixture
def torch_kernels(monkeypatch):
    from xequinet_b200 import gemm, nodeops, ops

    def mm(A, B, ta=False, tb=False, bias=None, alpha=1.0):
        C = alpha * ((A.t() if ta else A) @ (B.t() if tb else B))
        return C if bias is None else C + bias

    monkeypatch.setattr(gemm, "mm", mm)
    monkeypatch.setattr(gemm, "linear", lambda x, w, b=None: F.linear(x, w, b))
    monkeypatch.setattr(nodeops, "silu", F.silu)
    monkeypatch.setattr(gemm, "rowdot_raw", lambda x, w, b=None: x @ w.reshape(-1) + (b.reshape(-1)[0] if b is not None else 0.0))
    monkeypatch.setattr(gemm, "outer_raw", lambda g, w: g.reshape(-1, 1) * w.reshape(1, -1))
    monkeypatch.setattr(gemm, "wsum_raw", lambda g, x: g.reshape(-1) @ x)
    monkeypatch.setattr(gemm, "colsum_raw", lambda g: g.sum(0))
    monkeypatch.setattr(ops, "segment_sum", lambda src, ptr, batch: torch.zeros(ptr.numel() - 1, dtype=src.dtype).index_add(0, batch, src))


def test_heads_and_conditioning_host_logic(torch_kernels):
    import xequinet_b200 as xb
    from xequinet_b200 import keys

    z, cfg, data = load_golden("heads_mol")
    data = cast_data(data, torch.float64)
    spec = orc.heads_state_dict_spec(cfg, True, True, MODES)
    sd = orc.synthetic_state_dict(cfg, 31, torch.float64, spec=spec)
    mass = torch.from_numpy(np.load(GOLDEN.parent.parent / "xequinet_b200" / "data" / "atom_mass.npy"))
    ref = orc.xpainn_heads(sd, embed_table(), data, cfg, MODES, atom_mass=mass)

    model = xb.resolve_model("xpainn", charge_embed=True, spin_embed=True, output_modes=MODES).double()
    own = model.state_dict()
    model.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)
    # features entering the heads: the oracle's, with the conditioning modules run by the product classes
    plain = {k: v for k, v in data.items() if k not in ("charge", "spin")}
    sd_plain = {k: v for k, v in sd.items() if "_embedding." not in k}
    x0 = F.linear(embed_table()[data["atomic_numbers"].long()], sd["mods.embedding.embedding.1.weight"],
                  sd["mods.embedding.embedding.1.bias"])
    d = {keys.NODE_INVARIANT: x0, keys.BATCH: data["batch"], keys.BATCH_PTR: data["ptr"],
         "_xeq_ptr32": data["ptr"].to(torch.int32), keys.TOTAL_CHARGE: data["charge"], keys.TOTAL_SPIN: data["spin"]}
    d = model.mods["spin_embedding"](model.mods["charge_embedding"](d))
    G = data["ptr"].numel() - 1
    x1 = orc.spin_embedding(sd, "mods.spin_embedding.", orc.charge_embedding(sd, "mods.charge_embedding.", x0, data["charge"],
                            data["batch"], G), data["spin"], data["batch"], G)
    np.testing.assert_allclose(d[keys.NODE_INVARIANT].detach().numpy(), x1.numpy(), rtol=1e-12, atol=1e-13)

    x, V, batch, G = orc.xpainn_features(sd, embed_table(), data, cfg)
    d.update({keys.NODE_INVARIANT: x, keys.NODE_EQUIVARIANT: orc.to_cm(V, cfg), keys.POSITIONS: data["pos"],
              keys.ATOMIC_NUMBERS: data["atomic_numbers"]})
    for mode in MODES:
        d = model.mods[f"output_{mode}"](d)
    for k, v in ref.items():
        # the mass table is a float32 buffer (as the reference's torch.Tensor(ATOM_MASS)), the oracle's is fp64
        np.testing.assert_allclose(d[k].detach().numpy(), v.numpy(), rtol=1e-7 if k == "spatial_extent" else 1e-11,
                                   atol=1e-12, err_msg=k)


def test_narrow_linear_is_padded_to_the_kernel_granularity(torch_kernels, monkeypatch):
    from xequinet_b200 import gemm
    from xequinet_b200.nn.layers import Linear

    seen = []
    real = gemm.linear
    monkeypatch.setattr(gemm, "linear", lambda x, w, b=None: (seen.append((tuple(x.shape), tuple(w.shape))), real(x, w, b))[1])
    torch.manual_seed(0)
    for fin, fout, bias in ((64, 1, True), (2, 128, False), (1, 128, False), (64, 2, True), (128, 64, True)):
        lin = Linear(fin, fout, bias=bias).double()
        x = torch.randn(7, fin, dtype=torch.float64, requires_grad=True)
        y = lin(x)
        np.testing.assert_allclose(y.detach().numpy(), F.linear(x, lin.weight, lin.bias).detach().numpy(), rtol=1e-13, atol=1e-14)
        if fout == 1:  # one output feature: the row-dot kernels, not a GEMM tile
            assert not seen
        else:
            assert seen[-1][0][1] % 4 == 0 and seen[-1][1][0] % 4 == 0 and seen[-1][1][1] % 4 == 0
        y.sum().backward()
        assert lin.weight.grad.shape == (fout, fin) and x.grad.shape == (7, fin)


def test_linear_to_scalar_formulas_are_closed_under_differentiation(monkeypatch):
    """rowdot / outer / wsum (xeq_rowdot, xeq_outer, xeq_colsum_weighted): the derivative FORMULAS of the one-output
    Linear, with the three kernels replaced by torch stand-ins, against torch autograd of F.linear to second order
    (force training differentiates the read-out twice)."""
    from xequinet_b200 import gemm

    monkeypatch.setattr(gemm, "rowdot_raw", lambda x, w, b=None: x @ w.reshape(-1) + (b.reshape(-1)[0] if b is not None else 0.0))
    monkeypatch.setattr(gemm, "outer_raw", lambda g, w: g.reshape(-1, 1) * w.reshape(1, -1))
    monkeypatch.setattr(gemm, "wsum_raw", lambda g, x: g.reshape(-1) @ x)
    monkeypatch.setattr(gemm, "colsum_raw", lambda g: g.sum(0))
    gen = torch.Generator().manual_seed(3)
    x0 = torch.randn(9, 64, generator=gen, dtype=torch.float64)
    w0 = torch.randn(1, 64, generator=gen, dtype=torch.float64)
    b0 = torch.randn(1, generator=gen, dtype=torch.float64)
    r1 = torch.randn(9, 1, generator=gen, dtype=torch.float64)
    r2 = torch.randn(9, 64, generator=gen, dtype=torch.float64)

    def run(fn):
        x, w, b = (t.clone().requires_grad_(True) for t in (x0, w0, b0))
        y = fn(torch.tanh(x), w, b)
        (gx,) = torch.autograd.grad((y * r1).sum(), x, create_graph=True)   # "forces"
        loss = (y ** 2).sum() + (gx * r2).sum() + (gx ** 2).sum()            # loss on values and on the first derivative
        return [y.detach(), gx.detach(), *torch.autograd.grad(loss, (x, w, b))]

    for a, b in zip(run(gemm.linear_to_scalar), run(F.linear)):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-12, atol=1e-12)
