"""TEST INFRASTRUCTURE ONLY.  Golden vectors for the optional conditioning modules and read-out heads
(SURVEY.md 8f rank 4): runs the REAL reference (nn/model.py, nn/electronic.py, nn/output.py, unmodified, through
oracle/ref_stubs.py) with charge_embed / spin_embed and output_modes = energy, scalar, charges, dipole, polar
on seeded inputs (SpatialOut: see MODES), exports the reference's unit table (utils/qc.py:13-71) to
tests/golden/units.json, and exports the reference's atomic-mass table (utils/qc.py:181-190, a data artefact) to
xequinet_b200/data/atom_mass.npy.

Run in the build container only (needs /root/reference):
    python oracle/make_golden_heads.py
"""
from __future__ import annotations

import json
import re
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ref_stubs  # noqa: E402
from oracle import xpainn_oracle as orc  # noqa: E402
from oracle.make_golden import grad_digest  # noqa: E402

GOLD = ROOT / "tests" / "golden"
# `spatial` (SpatialOut) cannot be pinned: the reference multiplies masses [N] by pos [N, 3] (nn/output.py:364), which
# raises for every N != 3; the restatement follows the evident intent (masses as a column) and is unpinned for it.
MODES = ["energy", "scalar", "charges", "dipole", "polar"]


def reference_atom_mass() -> torch.Tensor:
    """The ATOM_MASS literal of utils/qc.py, read as data (the module itself imports pyscf)."""
    src = (ref_stubs.REFERENCE_ROOT / "xequinet" / "utils" / "qc.py").read_text()
    body = re.search(r"ATOM_MASS = torch\.Tensor\(\[(.*?)\]\)", src, re.S).group(1)
    return torch.tensor([float(t) for t in re.findall(r"[0-9]+\.?[0-9]*", body)], dtype=torch.float64)


def reference_units_table() -> dict:
    """The table `gen_units_dict` of utils/qc.py:13-71 builds, obtained by executing that one function."""
    import ast
    from math import pi

    src = (ref_stubs.REFERENCE_ROOT / "xequinet" / "utils" / "qc.py").read_text()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "gen_units_dict")
    ns = {"pi": pi}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "qc.py", "exec"), ns)
    return ns["gen_units_dict"]()


def _scatter_add_sum(src, index, dim=0, out=None, dim_size=None):
    """torch_scatter.scatter_sum as the wheel implements it (broadcast index + `scatter_add_`): unlike the `index_add`
    of ref_stubs it keeps no reference to `src`, which AtomicChargesOut then updates in place (nn/output.py:177)."""
    assert dim == 0
    if dim_size is None:
        dim_size = int(index.max().item()) + 1 if index.numel() else 0
    idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
    return src.new_zeros((dim_size,) + tuple(src.shape[1:])).scatter_add_(0, idx, src)


def build(cfg, sd, modes, dtype):
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        resolve_model = ref_stubs.reference_resolve_model()
        sys.modules["xequinet.utils.qc"].ATOM_MASS = reference_atom_mass().tolist()
        import xequinet.nn.output as ref_output
        ref_output.scatter_sum = _scatter_add_sum
        model = resolve_model("xpainn", charge_embed=True, spin_embed=True, output_modes=modes, **cfg.model_kwargs())
        want = {k: v.to(dtype) for k, v in sd.items() if not k.startswith("mods.output_") or k.split(".")[1][7:] in modes}
        missing, unexpected = model.load_state_dict(want, strict=False)
        assert not unexpected, unexpected
        assert all(not dict(model.named_parameters()).get(k, torch.zeros(0)).numel() or "embed_ten" in k for k in missing), missing
    finally:
        torch.set_default_dtype(prev)
    return model.eval()


def main():
    mass = reference_atom_mass()
    np.save(ROOT / "xequinet_b200" / "data" / "atom_mass.npy", mass.numpy())
    print("atom_mass", tuple(mass.shape))

    json.dump(reference_units_table(), open(GOLD / "units.json", "w"), indent=0, sort_keys=True)

    cfg = orc.CONFIG_DEFAULT
    spec = orc.heads_state_dict_spec(cfg, True, True, MODES)
    sd = orc.synthetic_state_dict(cfg, 2718, torch.float64, spec=spec)
    data = orc.make_molecule_batch(5, (6, 13), seed=11)
    G = data["ptr"].numel() - 1
    data["charge"] = torch.tensor([0, 1, -2, 0, 1][:G], dtype=torch.long)
    data["spin"] = torch.tensor([0, 1, 0, 2, 1][:G], dtype=torch.float64).reshape(-1, 1)
    blob = {"sd_seed": np.array(2718), "cfg_node_dim": np.array(cfg.node_dim), "cfg_muls": np.array(cfg.muls)}
    for k, v in data.items():
        blob["in:" + k] = v.numpy()

    model = build(cfg, sd, MODES, torch.float32)
    keys_spec = [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in model.state_dict().items()]
    json.dump({"heads": keys_spec}, open(GOLD / "state_dict_keys_heads.json", "w"))
    for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        torch.set_default_dtype(dt)
        try:
            d = {k: (v.to(dt) if v.is_floating_point() else v.clone()) for k, v in data.items()}
            d.pop("pbc", None)
            # every head, no gradient properties
            out = build(cfg, sd, MODES, dt)(dict(d, pos=d["pos"].clone()), compute_forces=False, compute_virial=False)
            for k, v in out.items():
                blob[f"{tag}:{k}"] = v.detach().numpy()
            # forces of the charge / spin conditioned model
            m2 = build(cfg, sd, ["energy", "dipole"], dt)
            o2 = m2(dict(d, pos=d["pos"].clone()), compute_forces=True, compute_virial=False)
            blob[f"{tag}:forces"] = o2["forces"].detach().numpy()
            assert np.allclose(o2["energy"].detach().numpy(), blob[f"{tag}:energy"])
            if dt == torch.float64:
                # parameter gradients of a loss over every head (fixed random cotangents)
                m3 = build(cfg, sd, MODES, dt).train()
                o3 = m3(dict(d, pos=d["pos"].clone()), compute_forces=False, compute_virial=False)
                g = torch.Generator().manual_seed(5)
                loss = 0.0
                for k in sorted(o3):
                    r = torch.randn(o3[k].shape, generator=g, dtype=torch.float64)
                    blob["cot:" + k] = r.numpy()
                    loss = loss + (o3[k] * r).sum()
                loss.backward()
                blob["loss_heads"] = np.array(loss.item())
                grads = {k: p.grad for k, p in m3.named_parameters() if p.grad is not None}
                for k, v in grad_digest(grads).items():
                    blob["gH:" + k] = v
        finally:
            torch.set_default_dtype(torch.float32)
        print(tag, {k: np.asarray(v).reshape(-1)[:2] for k, v in blob.items() if k.startswith(tag + ":")})
    np.savez_compressed(GOLD / "heads_mol.npz", **blob)


if __name__ == "__main__":
    main()
