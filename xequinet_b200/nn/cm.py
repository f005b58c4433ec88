"""Node-level per-irrep helpers on the component-major ("cm") layout
[mul0 | 3 x mul1 | 5 x mul2] (include/xeq_b200.h).  Plain differentiable torch views; the
E-sized work lives in the CUDA kernels."""
from __future__ import annotations

import torch


def split(V: torch.Tensor, muls):
    m0, m1, m2 = muls
    N = V.shape[0]
    return V[:, :m0], V[:, m0 : m0 + 3 * m1].reshape(N, 3, m1), V[:, m0 + 3 * m1 :].reshape(N, 5, m2)


def join(p0, p1, p2):
    N = p0.shape[0]
    return torch.cat([p0, p1.reshape(N, -1), p2.reshape(N, -1)], dim=1)


def expand_gate(g: torch.Tensor, muls) -> torch.Tensor:
    """Per-irrep gate [*, M] -> per-component [*, D] (ElementwiseTensorProduct with 'Mx0e')."""
    m0, m1, m2 = muls
    return torch.cat([g[..., :m0], g[..., m0 : m0 + m1].repeat(1, 3), g[..., m0 + m1 :].repeat(1, 5)], dim=-1)


def irrep_dot(a: torch.Tensor, b: torch.Tensor, muls) -> torch.Tensor:
    """Per-irrep sum_m a*b -> [N, M] (EquivariantDot / Invariant(squared=True), nn/o3layer.py:23-29,104-109)."""
    a0, a1, a2 = split(a, muls)
    b0, b1, b2 = split(b, muls)
    return torch.cat([a0 * b0, (a1 * b1).sum(1), (a2 * b2).sum(1)], dim=1)
