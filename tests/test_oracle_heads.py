"""CPU tests: the oracle's restatement of the optional conditioning modules (nn/electronic.py) and read-out heads
(nn/output.py) against golden vectors produced by the reference's own code (oracle/make_golden_heads.py), and the
state_dict contract of the host-side mirrors."""
import json

import numpy as np
import torch

import xequinet_b200 as xb
from helpers import GOLDEN, cast_data, embed_table, grad_digest, load_golden
from oracle import xpainn_oracle as orc

MODES = ["energy", "scalar", "charges", "dipole", "polar"]
OUT_KEYS = ["energy", "atomic_energies", "scalar_output", "atomic_charges", "dipole", "polarizability"]


def _sd(z, cfg, dtype, modes=MODES):
    return orc.synthetic_state_dict(cfg, int(z["sd_seed"]), dtype, spec=orc.heads_state_dict_spec(cfg, True, True, modes))


def test_oracle_heads_match_reference_fp64():
    z, cfg, data = load_golden("heads_mol")
    out = orc.xpainn_heads(_sd(z, cfg, torch.float64), embed_table(), cast_data(data, torch.float64), cfg, MODES)
    for k in OUT_KEYS:
        np.testing.assert_allclose(out[k].numpy(), z["f64:" + k], rtol=1e-11, atol=1e-11, err_msg=k)


def test_oracle_conditioned_forces_match_reference():
    z, cfg, data = load_golden("heads_mol")
    out = orc.xpainn_energy_forces(_sd(z, cfg, torch.float64), embed_table(), cast_data(data, torch.float64), cfg)
    np.testing.assert_allclose(out["energy"].detach().numpy(), z["f64:energy"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(out["forces"].numpy(), z["f64:forces"], rtol=0, atol=1e-10)
    # the conditioning changes the result: without the charge / spin inputs the energies differ
    plain = {k: v for k, v in cast_data(data, torch.float64).items() if k not in ("charge", "spin")}
    e0 = orc.xpainn_energy_forces(_sd(z, cfg, torch.float64), embed_table(), plain, cfg)["energy"]
    assert float((e0.detach() - torch.from_numpy(z["f64:energy"])).abs().max()) > 1e-4


def test_oracle_head_param_grads_match_reference():
    z, cfg, data = load_golden("heads_mol")
    sd = {k: v.requires_grad_(True) for k, v in _sd(z, cfg, torch.float64).items()}
    out = orc.xpainn_heads(sd, embed_table(), cast_data(data, torch.float64), cfg, MODES)
    loss = sum((out[k] * torch.from_numpy(z["cot:" + k])).sum() for k in sorted(OUT_KEYS))
    np.testing.assert_allclose(loss.item(), float(z["loss_heads"]), rtol=1e-11)
    loss.backward()
    n = 0
    for k, p in sd.items():
        key = f"gH:sum:{k}"
        if key not in z.files:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        s, smp = grad_digest(p.grad)
        np.testing.assert_allclose(s, z[key], rtol=1e-8, atol=1e-10, err_msg=k)
        np.testing.assert_allclose(smp, z[f"gH:smp:{k}"], rtol=1e-8, atol=1e-10, err_msg=k)
        n += 1
    assert n > 100


def test_oracle_spatial_extent_properties():
    """SpatialOut cannot be pinned (the reference's own forward raises on a shape mismatch, nn/output.py:364): the
    restatement is checked through what the quantity must satisfy -- invariance under translation and rotation."""
    z, cfg, data = load_golden("heads_mol")
    sd = orc.synthetic_state_dict(cfg, 7, torch.float64, spec=orc.heads_state_dict_spec(cfg, False, False, ["spatial"]))
    mass = torch.from_numpy(np.load(GOLDEN.parent.parent / "xequinet_b200" / "data" / "atom_mass.npy"))
    d = {k: v for k, v in cast_data(data, torch.float64).items() if k not in ("charge", "spin")}
    a = orc.xpainn_heads(sd, embed_table(), d, cfg, ["spatial"], atom_mass=mass)["spatial_extent"]
    q, _ = torch.linalg.qr(torch.randn(3, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(1)))
    d2 = dict(d, pos=d["pos"] @ q.T + torch.tensor([1.0, -2.0, 0.5], dtype=torch.float64))
    b = orc.xpainn_heads(sd, embed_table(), d2, cfg, ["spatial"], atom_mass=mass)["spatial_extent"]
    assert a.shape == (int(data["ptr"].numel() - 1), 1) and float(a.abs().min()) > 0
    np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-9)


def test_heads_state_dict_strict_round_trip():
    """tests/golden/state_dict_keys_heads.json = the REAL reference model's state_dict entries with charge_embed,
    spin_embed and the five pinned heads; a state_dict of that shape loads with strict=True and saves identically."""
    spec = json.loads((GOLDEN / "state_dict_keys_heads.json").read_text())["heads"]
    g = torch.Generator().manual_seed(0)
    ref_shaped = {}
    for k, shape, dtype in spec:
        dt = getattr(torch, dtype)
        ref_shaped[k] = (torch.randn(*shape, generator=g, dtype=dt) if dt.is_floating_point
                         else torch.arange(shape[0], dtype=dt) if shape else torch.zeros((), dtype=dt))
    model = xb.resolve_model("xpainn", charge_embed=True, spin_embed=True, output_modes=MODES)
    res = model.load_state_dict(ref_shaped, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    saved = model.state_dict()
    assert [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in saved.items()] == spec
    assert model.extra_properties == ["energy", "atomic_energies", "scalar_output", "atomic_charges", "dipole",
                                      "polarizability"]
    assert list(model.mods)[:3] == ["embedding", "charge_embedding", "spin_embedding"]
