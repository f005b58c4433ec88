"""GPU parity of the optional conditioning modules (nn/electronic.py) and read-out heads (nn/output.py) -- SURVEY.md
8f rank 4 -- against the golden vectors of the reference's own code (tests/golden/heads_mol.npz) and the fp64 oracle."""
import json

import numpy as np
import pytest
import torch

from helpers import GOLDEN, cast_data, embed_table, force_gate, load_golden
from oracle import xpainn_oracle as orc

pytestmark = pytest.mark.gpu

import xequinet_b200 as xb  # noqa: E402

DEV = "cuda"
MODES = ["energy", "scalar", "charges", "dipole", "polar"]
OUT_KEYS = ["energy", "atomic_energies", "scalar_output", "atomic_charges", "dipole", "polarizability"]


def _dev(data):
    return {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in data.items()}


def _model(z, cfg, modes, train=False, charge=True, spin=True):
    spec = orc.heads_state_dict_spec(cfg, True, True, MODES)
    sd = orc.synthetic_state_dict(cfg, int(z["sd_seed"]), torch.float32, spec=spec)
    if "spatial" in modes:
        sd.update(orc.synthetic_state_dict(cfg, 7, torch.float32, spec=orc.heads_state_dict_spec(cfg, False, False, ["spatial"])))
    model = xb.resolve_model("xpainn", charge_embed=charge, spin_embed=spin, output_modes=modes, **cfg.model_kwargs())
    own = model.state_dict()
    res = model.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)
    assert all(not own[k].numel() or "output_mask" in k or k.endswith("embed_ten") or "scalar_index" in k or "masses" in k
               for k in res.missing_keys), res.missing_keys
    model = model.to(DEV)
    return model.train() if train else model.eval()


def _gate(name, got, ref64, ref32, rel=2e-5):
    """|ours - fp64| <= max(rel * scale, 3 x the reference's own fp32 error), scale = max |fp64 value|."""
    scale = np.abs(ref64).max()
    err, err_ref = np.abs(got - ref64).max(), np.abs(ref32 - ref64).max()
    assert err <= max(rel * scale, 3.0 * err_ref), (name, err, err_ref, scale)


def test_heads_match_reference_golden():
    z, cfg, data = load_golden("heads_mol")
    model = _model(z, cfg, MODES)
    d = _dev(cast_data(data, torch.float32))
    out = model(d, compute_forces=False)
    assert set(out) == set(OUT_KEYS)
    for k in OUT_KEYS:
        assert out[k].shape == z["f64:" + k].shape, k
        _gate(k, out[k].detach().cpu().numpy().astype(np.float64), z["f64:" + k], z["f32:" + k].astype(np.float64))


def test_conditioned_forces_match_reference_golden():
    """Charge / spin conditioning changes the node scalars every message layer filters: E/F of the conditioned model."""
    z, cfg, data = load_golden("heads_mol")
    model = _model(z, cfg, ["energy", "dipole"])
    out = model(_dev(cast_data(data, torch.float32)), compute_forces=True)
    np.testing.assert_allclose(out["energy"].detach().cpu().numpy(), z["f64:energy"], rtol=1e-5, atol=1e-6)
    force_gate(out["forces"].cpu().numpy(), z["f64:forces"], z["f32:forces"])
    _gate("dipole", out["dipole"].detach().cpu().numpy().astype(np.float64), z["f64:dipole"], z["f32:dipole"].astype(np.float64))
    # a model built with the conditioning modules ignores them when the inputs are absent (nn/electronic.py:32-33)
    plain = {k: v for k, v in _dev(cast_data(data, torch.float32)).items() if k not in ("charge", "spin")}
    e0 = model(plain, compute_forces=False)["energy"]
    assert float((e0.detach().cpu() - torch.from_numpy(z["f64:energy"])).abs().max()) > 1e-4


def test_head_param_grads_match_reference_golden():
    """loss = sum_k <out_k, r_k> over every head with the golden cotangents; every parameter gradient against the
    reference's fp64 digests (sum, norm, strided sample)."""
    z, cfg, data = load_golden("heads_mol")
    model = _model(z, cfg, MODES, train=True)
    out = model(_dev(cast_data(data, torch.float32)), compute_forces=False)
    loss = sum((out[k] * torch.from_numpy(z["cot:" + k]).float().to(DEV)).sum() for k in sorted(OUT_KEYS))
    np.testing.assert_allclose(loss.item(), float(z["loss_heads"]), rtol=2e-5)
    loss.backward()
    n = 0
    # absolute floor for gradients that vanish analytically (the last bias of the charge-conserving head is removed
    # by the conservation step, nn/output.py:165-177: fp64 leaves 1e-15, fp32 1e-7): 1e-6 of the largest gradient norm
    floor = 1e-6 * max(float(z[f][1]) for f in z.files if f.startswith("gH:sum:"))
    for k, p in model.named_parameters():
        key = f"gH:sum:{k}"
        if key not in z.files:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        g = p.grad.detach().double().reshape(-1).cpu()
        ref_norm = float(z[key][1])
        smp = g[:: max(1, g.numel() // 64)][:64].numpy()
        assert abs(float(g.norm()) - ref_norm) <= 2e-3 * ref_norm + floor, (k, float(g.norm()), ref_norm)
        assert np.abs(smp - z[f"gH:smp:{k}"]).max() <= 2e-3 * max(np.abs(z[f"gH:smp:{k}"]).max(), ref_norm / np.sqrt(g.numel())) + floor, k
        n += 1
    assert n > 100


def test_spatial_extent_matches_oracle():
    """SpatialOut: unpinned in the reference (its forward raises, nn/output.py:364) -- against the fp64 restatement."""
    z, cfg, data = load_golden("heads_mol")
    data = {k: v for k, v in data.items() if k not in ("charge", "spin")}
    model = _model(z, cfg, ["spatial"], charge=False, spin=False)
    out = model(_dev(cast_data(data, torch.float32)), compute_forces=False)
    sd = orc.synthetic_state_dict(cfg, int(z["sd_seed"]), torch.float64, spec=orc.heads_state_dict_spec(cfg, True, True, MODES))
    sd.update(orc.synthetic_state_dict(cfg, 7, torch.float64, spec=orc.heads_state_dict_spec(cfg, False, False, ["spatial"])))
    sd = {k: v for k, v in sd.items() if "_embedding." not in k}
    mass = torch.from_numpy(np.load(GOLDEN.parent.parent / "xequinet_b200" / "data" / "atom_mass.npy"))
    ref = orc.xpainn_heads(sd, embed_table(), cast_data(data, torch.float64), cfg, ["spatial"], atom_mass=mass)
    np.testing.assert_allclose(out["spatial_extent"].detach().cpu().numpy(), ref["spatial_extent"].numpy(), rtol=5e-5)


def test_unsupported_head_raises():
    with pytest.raises(NotImplementedError):
        xb.resolve_model("xpainn", output_modes=["cartesian"])
