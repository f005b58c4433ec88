"""CPU tests: the unit table / conversions (xequinet/utils/qc.py:13-148) against the reference's own table
(tests/golden/units.json, oracle/make_golden_heads.py) and the host logic of the deployment wrappers
(xequinet/interface/jit_model.py)."""
import json

import pytest

from helpers import GOLDEN
from xequinet_b200 import jit_model, keys, units


@pytest.fixture(autouse=True)
def _restore_default_units():
    saved = dict(units.DEFAULT_UNITS_MAP)
    yield
    units.DEFAULT_UNITS_MAP.clear()
    units.DEFAULT_UNITS_MAP.update(saved)


def test_unit_table_matches_reference():
    ref = json.loads((GOLDEN / "units.json").read_text())
    assert set(ref) == set(units.UNITS)
    for k, v in ref.items():
        assert units.UNITS[k] == pytest.approx(v, rel=1e-14), k


def test_unit_expressions():
    ref = json.loads((GOLDEN / "units.json").read_text())
    assert units.unit_conversion("eV", "kcal/mol") == pytest.approx(ref["eV"] / (ref["kcal"] / ref["mol"]), rel=1e-14)
    assert units.unit_conversion("eV/Angstrom", "kJ/(mol*nm)") == pytest.approx(
        ref["eV"] / ref["Angstrom"] / (ref["kJ"] / (ref["mol"] * ref["nm"])), rel=1e-14)
    assert units.unit_conversion("eV/Angstrom^3", "GPa") == pytest.approx(160.21766, rel=1e-6)
    assert units.unit_conversion("kcal/mol/Angstrom", "kcal/mol/Angstrom") == 1.0
    assert units.unit_conversion(None, "eV") == 1.0
    assert units.eval_unit("2*Bohr**2") == 2.0
    for bad in ("__import__('os')", "eV+", "parsec", "eV/", "(eV", "1e3*eV"):
        assert not units.check_unit(bad)
        with pytest.raises(ValueError):
            units.eval_unit(bad)


def test_set_default_units_derives_gradient_units():
    units.set_default_units({keys.TOTAL_ENERGY: "kcal/mol", keys.TOTAL_CHARGE: "e"})
    d = units.get_default_units()
    assert d[keys.FORCES] == "kcal/mol/Angstrom" and d[keys.VIRIAL] == "kcal/mol/Angstrom^3"
    assert d[keys.ATOMIC_CHARGES] == "e" and d["base_energy"] == "kcal/mol"
    for bad in ({keys.FORCES: "eV/Angstrom"}, {"base_energy": "eV"}, {keys.ATOMIC_CHARGES: "e"}, {keys.TOTAL_ENERGY: "furlong"}):
        with pytest.raises(ValueError):
            units.set_default_units(bad)


def test_deployment_wrappers_unit_factors():
    units.set_default_units({keys.TOTAL_ENERGY: "eV"})
    m = jit_model.resolve_jit_model("lmp", "real", net_charge=-1, charge_embed=True)
    assert m.pos_unit_factor == 1.0 and m.cutoff_radius == 5.0 and m.net_charge == -1
    assert m.energy_unit_factor == pytest.approx(23.0605478, rel=1e-8) == m.forces_unit_factor
    assert list(m.mods)[:2] == ["embedding", "charge_embedding"]
    m = jit_model.resolve_jit_model("lmp", "electron", cutoff=4.0)
    assert m.pos_unit_factor == pytest.approx(0.529177211, rel=1e-8)
    assert m.cutoff_radius == pytest.approx(4.0 / 0.529177211, rel=1e-8)  # the ENGINE's neighbour-list cutoff, in Bohr
    assert m.mods["embedding"].rbf.cutoff == 4.0                           # the model's own cutoff stays in Angstrom
    g = jit_model.resolve_jit_model("gmx")
    assert g.pos_unit_factor == pytest.approx(10.0) and g.energy_unit_factor == pytest.approx(96.4853321, rel=1e-8)
    assert g.forces_unit_factor == pytest.approx(964.853321, rel=1e-8)
    d = jit_model.resolve_jit_model("dipole", "metal", output_modes=["dipole"])
    assert d.dipole_unit_factor == 1.0 and d.extra_properties == ["dipole"]
    with pytest.raises(NotImplementedError):
        jit_model.resolve_jit_model("openmm")
    meta = jit_model.deployment_metadata(m)
    assert meta["n_species"] == "87" and meta["periodic_table"].split()[1] == "H" and meta["periodic_table"].split()[-1] == "Rn"
    assert float(meta["cutoff_radius"]) == m.cutoff_radius


def test_export_and_load_deployment_round_trip(tmp_path):
    import torch

    import xequinet_b200 as xb

    base = xb.resolve_model("xpainn", action_blocks=1)
    ckpt = tmp_path / "model.pt"
    torch.save({"model": base.state_dict(), "config": {"model_name": "xpainn", "model_kwargs": {"action_blocks": 1},
                                                       "default_units": {keys.TOTAL_ENERGY: "eV"}}}, ckpt)
    meta = jit_model.export_deployment(str(ckpt), str(tmp_path / "model-lmp-real.pt"), mode="lmp", unit_style="real")
    assert meta["cutoff_radius"] == "5.0"
    m = jit_model.load_deployment(str(tmp_path / "model-lmp-real.pt"), device="cpu")
    assert isinstance(m, jit_model.XPaiNNLMP) and m.energy_unit_factor == pytest.approx(23.0605478, rel=1e-8)
    for (k, a), (_, b) in zip(base.state_dict().items(), m.state_dict().items()):
        assert torch.equal(a, b), k
