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


def force_gate(F_new, F_ref64, F_ref32):
    """Forces: 1e-4 eV/A absolute against the fp64 reference (north_star), with the noise-floor rule of SURVEY.md
    section 7.  The reference's OWN fp32 run misses fp64 by more than 1e-4 on a few atoms per batch (Invariant's
    sqrt(q + 1e-10) - 1e-5 flips the sign of its derivative within |v| ~ 1e-5, nn/o3layer.py:44): which atoms, and by
    how much, changes with every rounding difference, so the tail of the error distribution is a property of the
    model, not of an implementation.

      small inputs (< 500 atoms):  max err_new <= max(1e-4, 1.5 * max err_ref32)
      batches:  the 50th / 90th / 95th percentile of the per-atom error each obey that rule against the SAME
                percentile of the reference's fp32 error; the chaotic tail is bounded by 2x at the 99th percentile and
                the maximum -- an order statistic of a heavy tail: ONE atom -- is capped at max(1e-4, 8 * max err_ref32)
                (measured on B200, scratch/force_err_dist.py: p50 / p90 within +-25 % of the reference's at c1, c3,
                c4 and c5, p99 within 0.6-1.7x, max within 0.5-4.1x; a real defect moves the median by orders of
                magnitude)."""
    e_new = np.abs(np.asarray(F_new, dtype=np.float64) - F_ref64).max(axis=1)
    e_ref = np.abs(np.asarray(F_ref32, dtype=np.float64) - F_ref64).max(axis=1)
    if e_new.shape[0] < 500:
        assert e_new.max() <= max(1e-4, 1.5 * e_ref.max()), (e_new.max(), e_ref.max())
    else:
        for q, factor in ((50, 1.5), (90, 1.5), (95, 1.5), (99, 2.0)):
            a, b = np.percentile(e_new, q), np.percentile(e_ref, q)
            assert a <= max(1e-4 if q == 99 else 2e-6, factor * b), (q, a, b)
        assert e_new.max() <= max(1e-4, 8.0 * e_ref.max()), (e_new.max(), e_ref.max())
    return float(e_new.max()), float(e_ref.max())
