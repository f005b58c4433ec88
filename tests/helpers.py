"""Shared helpers for the tests: golden-fixture loading, oracle wrappers."""
from pathlib import Path

import numpy as np
import torch

from oracle import xpainn_oracle as orc

GOLDEN = Path(__file__).resolve().parent / "golden"


def embed_table(aux="aux56"):
    p = Path(__file__).resolve().parent.parent / "xequinet_b200" / "data" / f"gfn2-xtb_{aux}.npy"
    return torch.from_numpy(np.load(p))


def load_golden(name):
    z = np.load(GOLDEN / f"{name}.npz")
    cfg = orc.XPaiNNConfig(node_dim=int(z["cfg_node_dim"]), muls=tuple(int(v) for v in z["cfg_muls"]))
    data = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in:")}
    return z, cfg, data


def cast_data(data, dtype):
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in data.items()}


def grad_digest(g):
    g = g.detach().double().reshape(-1).cpu()
    return np.array([g.sum().item(), g.norm().item()]), g[:: max(1, g.numel() // 64)][:64].numpy()
