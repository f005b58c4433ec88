"""MD front end (xequinet_b200/interface.py <- interface/ase_calculator.py:20-118): the calculator protocol on an
atoms-like object, Verlet-skin reuse and CUDA-graph replay give the results of a plain model call."""
import numpy as np
import pytest
import torch

from oracle import xpainn_oracle as orc

pytestmark = pytest.mark.gpu

import xequinet_b200 as xb  # noqa: E402
from xequinet_b200.interface import XequiCalculator, full_3x3_to_voigt_6_stress  # noqa: E402

DEV = "cuda"


class FakeAtoms:
    """The part of ase.Atoms the calculator touches."""

    def __init__(self, positions, numbers, cell=None, pbc=(False, False, False)):
        self.positions, self.numbers = np.array(positions, dtype=np.float64), np.array(numbers)
        self.cell, self.pbc = (None if cell is None else np.array(cell, dtype=np.float64)), np.array(pbc, dtype=bool)

    def get_positions(self): return self.positions
    def get_atomic_numbers(self): return self.numbers
    def get_cell(self): return self.cell if self.cell is not None else np.zeros((3, 3))
    def get_pbc(self): return self.pbc
    def get_volume(self): return float(abs(np.linalg.det(self.cell)))


def _model():
    cfg = orc.CONFIG_DEFAULT
    m = xb.resolve_model("xpainn", **cfg.model_kwargs())
    m.load_state_dict(orc.synthetic_state_dict(cfg, 1234), strict=False)
    return m.to(DEV).eval()


def _direct(model, atoms, virial=False):
    d = {"pos": torch.tensor(atoms.positions, dtype=torch.float32, device=DEV),
         "atomic_numbers": torch.tensor(atoms.numbers, dtype=torch.int32, device=DEV)}
    if atoms.pbc.any():
        d["cell"] = torch.tensor(atoms.cell, dtype=torch.float32, device=DEV).reshape(1, 3, 3)
        d["pbc"] = torch.tensor(atoms.pbc).reshape(1, 3).to(DEV)
    d = xb.NeighborTransform(5.0)(d)
    d.pop("pbc", None)
    return model(d, compute_forces=True, compute_virial=virial)


@pytest.mark.parametrize("graph_replay", [False, True])
def test_calculator_md_loop_matches_plain_model_calls(graph_replay):
    model = _model()
    mol = orc.make_molecule_batch(1, 20, seed=4, with_edges=False)
    atoms = FakeAtoms(mol["pos"].numpy(), mol["atomic_numbers"].numpy())
    calc = XequiCalculator(model=model, skin=1.0, graph_replay=graph_replay)
    rng = np.random.default_rng(0)
    for it in range(5):  # a short "trajectory": small random displacements, inside the skin
        atoms.positions = atoms.positions + 0.05 * rng.standard_normal(atoms.positions.shape)
        calc.calculate(atoms, ["energy", "forces"])
        ref = _direct(model, atoms)
        assert abs(calc.results["energy"] - float(ref["energy"])) <= 2e-5 * max(1.0, abs(float(ref["energy"])))
        assert np.abs(calc.results["forces"] - ref["forces"].cpu().numpy()).max() < 1e-4
        assert calc.results["energies"].shape == (20,) and calc.results["forces"].shape == (20, 3)
    if not graph_replay:
        assert calc.transform.n_calls == 5 and calc.transform.n_builds < 5  # the list was reused


def test_calculator_stress_of_a_periodic_cell():
    model = _model()
    box = orc.make_small_pbc(14, 6.5, seed=3, triclinic=True)
    atoms = FakeAtoms(box["pos"].numpy(), box["atomic_numbers"].numpy(), cell=box["cell"].numpy().reshape(3, 3), pbc=(True, True, True))
    calc = XequiCalculator(model=model, skin=0.5)
    stress = calc.get_stress(atoms)
    ref = _direct(model, atoms, virial=True)
    vir = ref["virial"].detach().cpu().numpy().reshape(3, 3)
    np.testing.assert_allclose(stress, full_3x3_to_voigt_6_stress(vir) / atoms.get_volume(), rtol=0, atol=2e-5 * max(1.0, np.abs(vir).max()))
    np.testing.assert_allclose(calc.results["forces"], ref["forces"].cpu().numpy(), rtol=0, atol=1e-4)
    # fixed cell, graph replay: same energies and forces
    calc2 = XequiCalculator(model=model, graph_replay=True)
    f2 = calc2.get_forces(atoms)
    np.testing.assert_allclose(f2, ref["forces"].cpu().numpy(), rtol=0, atol=1e-4)
    assert abs(calc2.results["energy"] - float(ref["energy"])) <= 2e-5 * max(1.0, abs(float(ref["energy"])))
