"""Size-independent properties of the oracle (the e3nn restatements have no golden vectors of
their own, so they are also pinned by what the physics demands), float64 on CPU:
  * E(3) symmetry: energies invariant, forces and virial equivariant under rotations, reflections, translations;
  * permutation of the atoms permutes the forces;
  * forces are the finite-difference gradient of the energy; the virial is the finite-difference strain derivative;
  * lattice-vector shifts of single atoms in a periodic cell change nothing (with the neighbour list rebuilt);
  * the net force on every isolated molecule vanishes."""
import numpy as np
import pytest
import torch

from helpers import embed_table
from oracle import xpainn_oracle as orc

CFG = orc.XPaiNNConfig(node_dim=32, muls=(32, 16, 8))


def _setup(periodic, seed=0):
    sd = orc.synthetic_state_dict(CFG, 77, torch.float64)
    if periodic:
        d = orc.make_small_pbc(8, 6.0, seed=seed, dtype=torch.float64, triclinic=True)
        d["ptr"] = torch.tensor([0, d["pos"].shape[0]])
        d["batch"] = torch.zeros(d["pos"].shape[0], dtype=torch.long)
    else:
        d = orc.make_molecule_batch(3, (5, 9), seed=seed, dtype=torch.float64, with_edges=False)
    return sd, d


def _run(sd, d, virial=False):
    d = dict(d)
    if "cell" in d:
        n = torch.tensor([d["pos"].shape[0]])
        d["edge_index"], d["cell_offsets"] = orc.radius_graph_pbc(d["pos"], n, d["pbc"], d["cell"], CFG.cutoff)
    else:
        d["edge_index"] = orc.radius_graph(d["pos"], CFG.cutoff, d["batch"])
    return orc.xpainn_energy_forces(sd, embed_table().double(), d, CFG, compute_virial=virial)


def _random_orthogonal(seed, reflect):
    g = torch.Generator().manual_seed(seed)
    q, r = torch.linalg.qr(torch.randn(3, 3, generator=g, dtype=torch.float64))
    q = q * torch.sign(torch.diagonal(r))
    if (torch.det(q) < 0) != reflect:
        q[:, 0] = -q[:, 0]
    return q


@pytest.mark.parametrize("periodic", [False, True])
@pytest.mark.parametrize("reflect", [False, True])
def test_e3_symmetry(periodic, reflect):
    sd, d = _setup(periodic)
    ref = _run(sd, d, virial=True)
    R = _random_orthogonal(5, reflect)
    shift = torch.tensor([0.3, -1.1, 2.0], dtype=torch.float64)
    d2 = dict(d)
    d2["pos"] = d["pos"] @ R.T + shift
    if periodic:
        d2["cell"] = d["cell"] @ R.T
    out = _run(sd, d2, virial=True)
    torch.testing.assert_close(out["energy"], ref["energy"], rtol=1e-11, atol=1e-11)
    torch.testing.assert_close(out["forces"], ref["forces"] @ R.T, rtol=1e-9, atol=1e-10)
    # the strain derivative picks up the translation for non-periodic graphs only through sum_n F_n = 0
    torch.testing.assert_close(out["virial"], R @ ref["virial"] @ R.T, rtol=1e-8, atol=1e-9)


def test_permutation_equivariance_and_zero_net_force():
    sd, d = _setup(False, seed=3)
    ref = _run(sd, d)
    for gi in range(d["ptr"].numel() - 1):
        sl = slice(int(d["ptr"][gi]), int(d["ptr"][gi + 1]))
        assert float(ref["forces"][sl].sum(0).abs().max()) < 1e-11
    # permute the atoms inside the first molecule
    n0 = int(d["ptr"][1])
    perm = torch.cat([torch.randperm(n0, generator=torch.Generator().manual_seed(1)), torch.arange(n0, d["pos"].shape[0])])
    d2 = dict(d)
    d2["pos"], d2["atomic_numbers"] = d["pos"][perm], d["atomic_numbers"][perm]
    out = _run(sd, d2)
    torch.testing.assert_close(out["energy"], ref["energy"], rtol=1e-11, atol=1e-11)
    torch.testing.assert_close(out["forces"], ref["forces"][perm], rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("periodic", [False, True])
def test_forces_and_virial_are_finite_difference_derivatives(periodic):
    sd, d = _setup(periodic, seed=2)
    ref = _run(sd, d, virial=True)
    h = 2e-6  # Invariant's sqrt(q + 1e-10) makes the third derivatives large: keep the truncation error of the central difference small
    g = torch.Generator().manual_seed(9)
    for _ in range(3):  # random displacement directions
        dx = torch.randn(d["pos"].shape, generator=g, dtype=torch.float64)
        ep = _run(sd, {**d, "pos": d["pos"] + h * dx})["energy"].sum()
        em = _run(sd, {**d, "pos": d["pos"] - h * dx})["energy"].sum()
        fd = (ep - em) / (2 * h)
        an = -(ref["forces"] * dx).sum()
        assert abs(float(fd - an)) < 1e-6 * max(1.0, abs(float(an)))
    # strain: pos -> pos (1 + eps), cell -> cell (1 + eps) with a symmetric eps
    eps = torch.randn(3, 3, generator=g, dtype=torch.float64)
    eps = 0.5 * (eps + eps.T)

    def strained(sign):
        d2 = dict(d)
        d2["pos"] = d["pos"] + sign * h * d["pos"] @ eps
        if periodic:
            d2["cell"] = d["cell"] + sign * h * d["cell"] @ eps
        return _run(sd, d2)["energy"].sum()

    fd = (strained(+1) - strained(-1)) / (2 * h)
    an = -(ref["virial"].sum(0) * eps).sum()
    assert abs(float(fd - an)) < 1e-6 * max(1.0, abs(float(an)))


def test_lattice_shift_of_single_atoms_changes_nothing():
    sd, d = _setup(True, seed=4)
    ref = _run(sd, d)
    d2 = dict(d)
    pos = d["pos"].clone()
    cell = d["cell"].reshape(3, 3)
    pos[1] += cell[0]
    pos[4] -= cell[1] + 2 * cell[2]
    d2["pos"] = pos
    out = _run(sd, d2)
    torch.testing.assert_close(out["energy"], ref["energy"], rtol=1e-11, atol=1e-11)
    torch.testing.assert_close(out["forces"], ref["forces"], rtol=1e-9, atol=1e-10)
