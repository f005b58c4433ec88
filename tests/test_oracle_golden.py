"""CPU tests: the oracle restatement (oracle/xpainn_oracle.py) against the golden vectors
produced by the reference's own code (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import xpainn_oracle as orc
from helpers import cast_data, embed_table, grad_digest, load_golden

CASES = ["mol_small", "mol_c4_small", "pbc_small", "pbc_tiny", "pbc_slab", "pbc_two_graphs"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_fp64(name):
    z, cfg, data = load_golden(name)
    sd = orc.synthetic_state_dict(cfg, int(z["sd_seed"]), torch.float64)
    out = orc.xpainn_energy_forces(sd, embed_table(), cast_data(data, torch.float64), cfg)
    np.testing.assert_allclose(out["energy"].detach().numpy(), z["f64:energy"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(out["atomic_energies"].detach().numpy(), z["f64:atomic_energies"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(out["forces"].numpy(), z["f64:forces"], rtol=0, atol=1e-10)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_fp32(name):
    # tolerance stated by north_star: 1e-5 relative on energies, 1e-4 eV/A absolute on forces
    z, cfg, data = load_golden(name)
    sd = orc.synthetic_state_dict(cfg, int(z["sd_seed"]), torch.float32)
    out = orc.xpainn_energy_forces(sd, embed_table().float(), cast_data(data, torch.float32), cfg)
    np.testing.assert_allclose(out["energy"].detach().numpy(), z["f32:energy"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out["forces"].numpy(), z["f32:forces"], rtol=0, atol=1e-4)
    # and both fp32 runs sit at the fp32 noise floor around the fp64 reference
    np.testing.assert_allclose(out["forces"].numpy(), z["f64:forces"], rtol=0, atol=2e-4)


@pytest.mark.parametrize("name", ["mol_small", "pbc_small"])
@pytest.mark.parametrize("use_forces", [False, True])
def test_oracle_param_grads_match_reference(name, use_forces):
    """Training step semantics (utils/trainer.py:295-302): loss.backward() through forces
    (double backward).  Compared through compact per-tensor digests."""
    z, cfg, data = load_golden(name)
    sd = orc.synthetic_state_dict(cfg, int(z["sd_seed"]), torch.float64)
    sd = {k: v.requires_grad_(True) for k, v in sd.items()}
    out = orc.xpainn_energy_forces(sd, embed_table(), cast_data(data, torch.float64), cfg, create_graph=True)
    tE, tF = torch.from_numpy(z["f64:target_energy"]), torch.from_numpy(z["f64:target_forces"])
    loss = F.smooth_l1_loss(out["energy"], tE)
    if use_forces:
        loss = loss + 100.0 * F.smooth_l1_loss(out["forces"], tF)
    tag = "gEF" if use_forces else "gE"
    np.testing.assert_allclose(loss.item(), float(z[f"f64:loss_{tag}"]), rtol=1e-11)
    loss.backward()
    n = 0
    for k, p in sd.items():
        key = f"f64:{tag}:sum:{k}"
        if key not in z.files:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        s, smp = grad_digest(p.grad)
        np.testing.assert_allclose(s, z[key], rtol=1e-8, atol=1e-11, err_msg=k)
        np.testing.assert_allclose(smp, z[f"f64:{tag}:smp:{k}"], rtol=1e-8, atol=1e-11, err_msg=k)
        n += 1
    assert n > 50


def test_oracle_pbc_edges_match_reference_water():
    z = np.load(__import__("helpers").GOLDEN / "water_edges.npz")
    pos, cell, pbc = (torch.from_numpy(z[k]) for k in ("pos", "cell", "pbc"))
    ei, co = orc.radius_graph_pbc(pos, torch.tensor([pos.shape[0]]), pbc, cell, 5.0)
    assert np.array_equal(ei.numpy(), z["edge_index"].astype(np.int64))
    assert np.array_equal(co.numpy().astype(np.int8), z["cell_offsets"])


def test_radius_graph_symmetric_and_loop_free():
    d = orc.make_molecule_batch(5, (6, 20), seed=11)
    ei = d["edge_index"]
    assert (ei[0] != ei[1]).all()
    fwd = set(map(tuple, ei.t().tolist()))
    assert all((b, a) in fwd for a, b in fwd)
    assert (d["batch"][ei[0]] == d["batch"][ei[1]]).all()


@pytest.mark.parametrize("name", ["mol_small", "pbc_small", "pbc_slab", "pbc_two_graphs"])
def test_oracle_virial_matches_reference(name):
    """The oracle's strain trick (nn/basic.py:93-107, 162-199) against the reference's own virial
    (tests/golden/virial.npz, oracle/make_golden_virial.py), float64."""
    import numpy as np
    from helpers import GOLDEN

    z, cfg, data = load_golden(name)
    ref = np.load(GOLDEN / "virial.npz")[f"{name}:virial"]
    sd = orc.synthetic_state_dict(cfg, int(z["sd_seed"]), torch.float64)
    d = cast_data(data, torch.float64)
    out = orc.xpainn_energy_forces(sd, embed_table().double(), d, cfg, compute_virial=True)
    np.testing.assert_allclose(out["virial"].detach().numpy(), ref, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(out["forces"].detach().numpy(), z["f64:forces"], rtol=1e-9, atol=1e-9)
    # identity used by the B200 path: virial = -(sum_n pos_n (x) dE/dpos_n + cell^T dE/dcell), symmetrised
    # (for non-periodic graphs the cell term vanishes: the virial follows from the forces alone)
    if "cell" not in d:
        G = d["ptr"].numel() - 1
        M = torch.zeros(G, 3, 3, dtype=torch.float64).index_add(0, d["batch"], torch.einsum("na,nb->nab", d["pos"], -out["forces"].detach()))
        np.testing.assert_allclose((-0.5 * (M + M.transpose(1, 2))).numpy(), ref, rtol=1e-9, atol=1e-9)
