"""TEST INFRASTRUCTURE ONLY.  Adds tests/golden/virial.npz: energies, forces and VIRIALS of the real reference
(BaseModel.forward(compute_forces=True, compute_virial=True): the strain trick of nn/basic.py:93-107, 162-199,
run unmodified through oracle/ref_stubs.py) on the inputs of the existing fixtures, in float64.

Run in the build container only (needs /root/reference):
    python oracle/make_golden_virial.py
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import xpainn_oracle as orc  # noqa: E402
from oracle.make_golden import GOLD, reference_model  # noqa: E402

CASES = ("mol_small", "pbc_small", "pbc_slab", "pbc_two_graphs")


def main():
    blob = {}
    for name in CASES:
        z = np.load(GOLD / f"{name}.npz")
        cfg = orc.XPaiNNConfig(node_dim=int(z["cfg_node_dim"]), muls=tuple(int(v) for v in z["cfg_muls"]))
        data = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in:")}
        sd = orc.synthetic_state_dict(cfg, int(z["sd_seed"]), torch.float64)
        model = reference_model(cfg, sd, torch.float64).eval()
        d = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in data.items()}
        d.pop("pbc", None)
        torch.set_default_dtype(torch.float64)
        try:
            out = model(dict(d), compute_forces=True, compute_virial=True)
        finally:
            torch.set_default_dtype(torch.float32)
        assert np.allclose(out["energy"].detach().numpy(), z["f64:energy"], rtol=1e-12, atol=1e-12), name
        assert np.allclose(out["forces"].detach().numpy(), z["f64:forces"], rtol=1e-10, atol=1e-10), name
        blob[f"{name}:virial"] = out["virial"].detach().numpy()
        print(name, "virial", out["virial"].detach().numpy().reshape(-1, 9)[0][:4])
    np.savez_compressed(GOLD / "virial.npz", **blob)


if __name__ == "__main__":
    main()
