"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the REAL reference
(/root/reference/xequinet/nn/*.py, data/radius_graph.py, unmodified, through
oracle/ref_stubs.py) on seeded synthetic inputs, and exports the reference's embedding
table (utils/pre_computed/*.pt, a data artefact) to xequinet_b200/data/.

Run in the build container only (needs /root/reference):
    python oracle/make_golden.py
The fixtures travel to the GPU box; this script and the reference do not need to.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ref_stubs  # noqa: E402
from oracle import xpainn_oracle as orc  # noqa: E402

GOLD = ROOT / "tests" / "golden"


def export_embedding_tables():
    out = ROOT / "xequinet_b200" / "data"
    out.mkdir(parents=True, exist_ok=True)
    for aux in ("aux28", "aux56"):
        torch.set_default_dtype(torch.float64)
        ten = ref_stubs.get_embedding_tensor("gfn2-xtb", aux)
        torch.set_default_dtype(torch.float32)
        np.save(out / f"gfn2-xtb_{aux}.npy", ten.numpy())
        print("embedding", aux, tuple(ten.shape))


def reference_model(cfg: orc.XPaiNNConfig, sd, dtype):
    torch.set_default_dtype(dtype)
    try:
        resolve_model = ref_stubs.reference_resolve_model()
        model = resolve_model("xpainn", **cfg.model_kwargs())
        missing, unexpected = model.load_state_dict({k: v.to(dtype) for k, v in sd.items()}, strict=False)
        assert not unexpected, unexpected
        assert all(("output_mask" in k) or k.endswith(".weight") and "tp" in k or "embed_ten" in k or "scalar_index" in k
                   or "rsh_conv" in k or "scalar_mul" in k for k in missing), missing
    finally:
        torch.set_default_dtype(torch.float32)
    return model


def loss_fn(out, tE, tF, wE=1.0, wF=100.0):
    loss = wE * F.smooth_l1_loss(out["energy"], tE)
    if tF is not None:
        loss = loss + wF * F.smooth_l1_loss(out["forces"], tF)
    return loss


def grad_digest(named_grads):
    """Compact per-tensor digest: [sum, l2, strided sample] (full grads would be MBs)."""
    dig = {}
    for k, g in named_grads.items():
        g = g.detach().double().reshape(-1)
        dig["sum:" + k] = np.array([g.sum().item(), g.norm().item()])
        dig["smp:" + k] = g[:: max(1, g.numel() // 64)][:64].numpy()
    return dig


def run_reference(cfg, sd, data, dtype, train: bool):
    model = reference_model(cfg, sd, dtype)
    d = {k: (v.to(dtype) if v.is_floating_point() else v.clone()) for k, v in data.items()}
    d.pop("pbc", None)
    res = {}
    model.eval()
    torch.set_default_dtype(dtype)  # node_equivariant is created in the default dtype (xpainn.py:77-80)
    try:
        out = model(dict(d), compute_forces=True, compute_virial=False)
        res["energy"] = out["energy"].detach().numpy()
        res["atomic_energies"] = out["atomic_energies"].detach().numpy()
        res["forces"] = out["forces"].detach().numpy()
        if train:
            g = torch.Generator().manual_seed(99)
            tE = torch.randn(out["energy"].shape, generator=g, dtype=torch.float64).to(dtype)
            tF = torch.randn(out["forces"].shape, generator=g, dtype=torch.float64).to(dtype)
            res["target_energy"], res["target_forces"] = tE.numpy(), tF.numpy()
            model.train()
            for tag, use_f in (("gE", False), ("gEF", True)):
                model.zero_grad()
                dd = dict(d)
                dd["pos"] = d["pos"].clone()
                o = model(dd, compute_forces=use_f, compute_virial=False)
                loss = loss_fn(o, tE, tF if use_f else None)
                loss.backward()
                res[f"loss_{tag}"] = np.array(loss.item())
                grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
                for k, v in grad_digest(grads).items():
                    res[f"{tag}:{k}"] = v
    finally:
        torch.set_default_dtype(torch.float32)
    return res


def save_case(name, cfg, data, seed_sd, train=True, dtypes=(torch.float64, torch.float32)):
    sd = orc.synthetic_state_dict(cfg, seed_sd, torch.float64)
    blob = {"sd_seed": np.array(seed_sd), "cfg_node_dim": np.array(cfg.node_dim), "cfg_muls": np.array(cfg.muls)}
    for k, v in data.items():
        blob["in:" + k] = v.numpy()
    for dt in dtypes:
        tag = "f64" if dt == torch.float64 else "f32"
        res = run_reference(cfg, sd, data, dt, train and dt == torch.float64)
        for k, v in res.items():
            blob[f"{tag}:{k}"] = v
        print(name, tag, "E[:3]=", res["energy"][:3], "|F|max=", np.abs(res["forces"]).max())
    np.savez_compressed(GOLD / f"{name}.npz", **blob)


def main():
    GOLD.mkdir(parents=True, exist_ok=True)
    export_embedding_tables()
    rg_pbc, rg_single = ref_stubs.reference_radius_graph_pbc()

    # --- molecules, default widths -------------------------------------------------
    data = orc.make_molecule_batch(6, (8, 14), seed=3)
    save_case("mol_small", orc.CONFIG_DEFAULT, data, seed_sd=1234)

    # --- molecules, c4 widths (256 channels) ----------------------------------------
    data = orc.make_molecule_batch(2, (10, 12), seed=5, z_table=orc._Z_SPICE)
    save_case("mol_c4_small", orc.CONFIG_C4, data, seed_sd=4321, train=False)

    # --- periodic cells: edges from the reference's own radius_graph_pbc ------------
    for name, d in (
        ("pbc_small", orc.make_small_pbc(12, 6.0, seed=1, triclinic=True)),
        ("pbc_tiny", orc.make_small_pbc(3, 4.0, seed=2, triclinic=False)),
        ("pbc_slab", orc.make_small_pbc(10, 7.0, seed=4, triclinic=True, pbc=(True, True, False))),
    ):
        n = torch.tensor([d["pos"].shape[0]])
        ei, co = rg_pbc(d["pos"], n, d["pbc"], d["cell"], 5.0)
        ei, co = orc.canonical_sort(ei, co)
        ei2, co2 = orc.radius_graph_pbc(d["pos"], n, d["pbc"], d["cell"], 5.0)
        assert torch.equal(ei, ei2) and torch.equal(co, co2), name
        d["edge_index"], d["cell_offsets"] = ei, co
        print(name, "edges", ei.shape[1], "self-image", int((ei[0] == ei[1]).sum()))
        save_case(name, orc.CONFIG_DEFAULT, d, seed_sd=1234, train=(name == "pbc_small"))

    # --- two-graph periodic batch (multi-graph branch, nn/basic.py:124-128) ----------
    a = orc.make_small_pbc(9, 6.5, seed=7, triclinic=True)
    b = orc.make_small_pbc(7, 5.5, seed=8, triclinic=False)
    d = {
        "pos": torch.cat([a["pos"], b["pos"]]),
        "atomic_numbers": torch.cat([a["atomic_numbers"], b["atomic_numbers"]]),
        "batch": torch.cat([torch.zeros(9, dtype=torch.long), torch.ones(7, dtype=torch.long)]),
        "ptr": torch.tensor([0, 9, 16]),
        "cell": torch.cat([a["cell"], b["cell"]]),
        "pbc": torch.tensor([[True, True, True]] * 2),
    }
    ei, co = rg_pbc(d["pos"], torch.tensor([9, 7]), d["pbc"], d["cell"], 5.0)
    ei, co = orc.canonical_sort(ei, co)
    ei2, co2 = orc.radius_graph_pbc(d["pos"], torch.tensor([9, 7]), d["pbc"], d["cell"], 5.0)
    assert torch.equal(ei, ei2) and torch.equal(co, co2)
    d["edge_index"], d["cell_offsets"] = ei, co
    save_case("pbc_two_graphs", orc.CONFIG_DEFAULT, d, seed_sd=1234, train=False)

    # --- edge lists only: reduced water box, from the reference's radius_graph_pbc ---
    w = orc.make_water_box(n_side=5, seed=0)
    n = torch.tensor([w["pos"].shape[0]])
    ei, co = rg_pbc(w["pos"], n, w["pbc"], w["cell"], 5.0)
    ei, co = orc.canonical_sort(ei, co)
    ei2, co2 = orc.radius_graph_pbc(w["pos"], n, w["pbc"], w["cell"], 5.0)
    assert torch.equal(ei, ei2) and torch.equal(co, co2)
    np.savez_compressed(
        GOLD / "water_edges.npz",
        pos=w["pos"].numpy(), cell=w["cell"].numpy(), pbc=w["pbc"].numpy(),
        edge_index=ei.to(torch.int32).numpy(), cell_offsets=co.to(torch.int8).numpy(),
    )
    print("water_edges", ei.shape)


if __name__ == "__main__":
    main()
