"""GPU tests of the deployment wrappers (xequinet/interface/jit_model.py:12-237): the engine's
unit system in, the engine's unit system out, on the same kernels as the plain model."""
import numpy as np
import pytest
import torch

from helpers import cast_data, load_golden
from oracle import xpainn_oracle as orc

pytestmark = pytest.mark.gpu

import xequinet_b200 as xb  # noqa: E402
from xequinet_b200 import jit_model, keys, units  # noqa: E402

DEV = "cuda"
EV_KCAL = 23.060547830619026
BOHR = 0.5291772111941798


@pytest.fixture(autouse=True)
def _units():
    saved = dict(units.DEFAULT_UNITS_MAP)
    units.set_default_units({keys.TOTAL_ENERGY: "eV", keys.TOTAL_CHARGE: "e", keys.DIPOLE: "e*Angstrom"})
    yield
    units.DEFAULT_UNITS_MAP.clear()
    units.DEFAULT_UNITS_MAP.update(saved)


def _dev(data):
    return {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in data.items()}


def _load(model, cfg, seed):
    model.load_state_dict(orc.synthetic_state_dict(cfg, seed), strict=False)
    return model.to(DEV).eval()


@pytest.mark.parametrize("name", ["mol_small", "pbc_small"])
def test_lammps_wrapper_units(name):
    z, cfg, data = load_golden(name)
    data = {k: v for k, v in cast_data(data, torch.float32).items() if k != "pbc"}
    if name == "mol_small":  # the wrapper handles ONE structure (no batch): first molecule of the golden batch
        n0 = int(data["ptr"][1])
        keep = data["edge_index"][0] < n0
        data = {"pos": data["pos"][:n0], "atomic_numbers": data["atomic_numbers"][:n0], "edge_index": data["edge_index"][:, keep]}
    base = _load(xb.resolve_model("xpainn", **cfg.model_kwargs()), cfg, int(z["sd_seed"]))
    virial = name == "pbc_small"
    ref = base(_dev(dict(data)), compute_forces=True, compute_virial=virial)
    for style, e_fac, len_fac in (("metal", 1.0, 1.0), ("real", EV_KCAL, 1.0), ("electron", 1.0 / 27.211386245988, BOHR)):
        m = _load(jit_model.resolve_jit_model("lmp", style, **cfg.model_kwargs()), cfg, int(z["sd_seed"]))
        d = _dev(dict(data))
        d["pos"] = d["pos"] / len_fac  # the engine's coordinates
        if "cell" in d:
            d["cell"] = d["cell"] / len_fac
        out = m(d, compute_forces=True, compute_virial=virial)
        assert set(out) == set(ref)
        torch.testing.assert_close(out["energy"], ref["energy"] * e_fac, rtol=2e-6 if len_fac == 1.0 else 1e-5, atol=1e-7)
        # Bohr coordinates are rounded to fp32 before the wrapper scales them back: positions differ by an ulp
        f_tol = (2e-6 if len_fac == 1.0 else 5e-5) * e_fac * len_fac
        torch.testing.assert_close(out["forces"], ref["forces"] * (e_fac * len_fac), rtol=1e-5, atol=f_tol)
        if virial:
            torch.testing.assert_close(out["virial"], ref["virial"] * e_fac, rtol=1e-5, atol=(5e-6 if len_fac == 1.0 else 1e-4) * e_fac)
        assert m.cutoff_radius == pytest.approx(5.0 / len_fac)
    if name == "mol_small":  # and the golden of the reference itself (first molecule: independent of the others)
        np.testing.assert_allclose(ref["energy"].detach().cpu().numpy(), z["f64:energy"][:1], rtol=1e-5, atol=1e-6)


def test_gromacs_wrapper_energy_and_engine_side_forces():
    z, cfg, data = load_golden("pbc_small")
    data = cast_data(data, torch.float32)
    base = _load(xb.resolve_model("xpainn", **cfg.model_kwargs()), cfg, int(z["sd_seed"]))
    ref = base(_dev({k: v for k, v in data.items() if k != "pbc"}), compute_forces=True)
    m = _load(jit_model.resolve_jit_model("gmx", **cfg.model_kwargs()), cfg, int(z["sd_seed"]))
    pos_nm = (data["pos"] / 10.0).to(DEV).requires_grad_(True)
    e = m(pos_nm, data["atomic_numbers"].to(DEV), box=(data["cell"][0] / 10.0).to(DEV), pbc=data["pbc"].reshape(-1)[:3].to(DEV))
    torch.testing.assert_close(e, ref["energy"] * m.energy_unit_factor, rtol=2e-6, atol=1e-5)
    np.testing.assert_allclose(e.detach().cpu().numpy() / m.energy_unit_factor, z["f64:energy"], rtol=1e-5, atol=1e-6)
    f = -torch.autograd.grad(e.sum(), pos_nm)[0]  # what GROMACS does with the returned energy
    torch.testing.assert_close(f, ref["forces"] * m.forces_unit_factor, rtol=1e-5, atol=2e-3)
    # no box: an isolated cluster, same as the plain model with K1's open-boundary list
    d = orc.make_molecule_batch(1, 14, seed=3, with_edges=False)
    e2 = m((d["pos"] / 10.0).to(DEV), d["atomic_numbers"].to(DEV))
    r2 = base(xb.NeighborTransform(5.0)(_dev({"pos": d["pos"], "atomic_numbers": d["atomic_numbers"]})), compute_forces=False)
    torch.testing.assert_close(e2, r2["energy"] * m.energy_unit_factor, rtol=2e-6, atol=1e-5)


def test_dipole_wrapper_matches_reference_golden():
    z, cfg, data = load_golden("heads_mol")
    data = cast_data(data, torch.float32)
    n0 = int(data["ptr"][1])
    keep = data["edge_index"][0] < n0
    spec = orc.heads_state_dict_spec(cfg, True, True, ["energy", "scalar", "charges", "dipole", "polar"])
    sd = orc.synthetic_state_dict(cfg, int(z["sd_seed"]), torch.float32, spec=spec)
    charge = int(data["charge"][0])
    m = jit_model.resolve_jit_model("dipole", "metal", net_charge=charge, charge_embed=True, output_modes=["dipole"],
                                    **cfg.model_kwargs())
    own = m.state_dict()
    m.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)
    m = m.to(DEV).eval()
    out = m({"pos": data["pos"][:n0].to(DEV), "atomic_numbers": data["atomic_numbers"][:n0].to(DEV),
             "edge_index": data["edge_index"][:, keep].to(DEV)})
    assert set(out) == {"dipole"}
    # the golden model also has the spin embedding; molecule 0 has spin 0 -> value = linear_v(0) = 0: identical
    assert float(data["spin"][0]) == 0.0
    ref64, ref32 = z["f64:dipole"][:1], z["f32:dipole"][:1].astype(np.float64)
    err = np.abs(out["dipole"].detach().cpu().numpy() - ref64).max()
    assert err <= max(2e-5 * np.abs(ref64).max(), 3 * np.abs(ref32 - ref64).max()), err
