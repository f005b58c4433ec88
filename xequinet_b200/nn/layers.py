"""Host-side mirrors of the reference's NN blocks for the XPaiNN path, with identical
parameter / buffer names so that state_dicts are interchangeable:
  SphericalBesselj0, CosineCutoff        <- xequinet/nn/rbf.py:43-57, 134-152
  Invariant, EquivariantDot, EquivariantLayerNorm <- xequinet/nn/o3layer.py:11-44, 78-171
  O3Linear                               <- e3nn o3.Linear as used at nn/xpainn.py:186-187
  Int2c1eEmbedding                       <- xequinet/nn/basic.py:34-57
Equivariant tensors are held in the cm layout; the parameters are layout independent."""
from __future__ import annotations

import math
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import gemm, nodeops
from . import cm
from .irreps import irreps_dim, num_irreps

_DATA = Path(__file__).resolve().parent.parent / "data"


class SiLU(nn.SiLU):
    """SiLU as one fused kernel per derivative order (csrc/node_update.cu)."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return nodeops.silu(x)


class LayerNorm(nn.LayerNorm):
    """nn.LayerNorm(C) (same parameters) on the fused irreps-norm kernel: it is the "Cx0e" case of
    EquivariantLayerNorm (csrc/node_norm.cu)."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return nodeops.layer_norm(x, self.weight, self.bias, self.eps)

    def with_passthrough(self, x: torch.Tensor):
        """(norm(x), x) -- nodeops._IrrepsNorm: feed the residual from the second output."""
        return nodeops.layer_norm_pass(x, self.weight, self.bias, self.eps)


def resolve_activation(activation: str) -> nn.Module:
    """xequinet/nn/basic.py:241-262; only SiLU is on the XPaiNN path (basic.py:255-256)."""
    table = {"silu": SiLU, "relu": nn.ReLU, "leakyrelu": nn.LeakyReLU, "softplus": nn.Softplus,
             "sigmoid": nn.Sigmoid, "tanh": nn.Tanh, "identity": nn.Identity}
    if activation.lower() not in table:
        raise NotImplementedError(f"Unsupported activation function {activation}")
    return table[activation.lower()]()


class Linear(nn.Linear):
    """nn.Linear (same parameters / state_dict names) evaluated by the tcgen05 GEMM kernel (K3,
    csrc/node_gemm.cu) together with all of its derivatives.  Widths that are not multiples of 4
    (only the 64 -> 1 energy read-out, nn/output.py:107-111) stay a torch op."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.in_features % 4 or self.out_features % 4 or x.dim() != 2:
            return F.linear(x, self.weight, self.bias)
        return gemm.linear(x, self.weight, self.bias)


class _E3nnBuffers(nn.Module):
    """Carries the non-trainable state_dict entries a real e3nn TensorProduct contributes
    (`weight` empty, `output_mask` ones) so checkpoints load with strict=True."""

    def __init__(self, out_dim: int):
        super().__init__()
        self.register_buffer("weight", torch.Tensor())
        self.register_buffer("output_mask", torch.ones(out_dim))


class _TPHolder(nn.Module):
    def __init__(self, out_dim: int):
        super().__init__()
        self.tp = _E3nnBuffers(out_dim)


class SphericalBesselj0(nn.Module):
    """Holder of the learnable Bessel frequencies (nn/rbf.py:143-144); the basis itself is
    evaluated inside the fused edge kernel."""

    def __init__(self, num_basis: int, cutoff: float, eps: float = 1e-5):
        super().__init__()
        self.num_basis, self.cutoff, self.eps = num_basis, cutoff, eps
        freq = math.pi * torch.arange(1, num_basis + 1) / cutoff
        self.freq = nn.Parameter(freq.view(1, -1))
        self.coeff = math.sqrt(2 / cutoff)


class CosineCutoff(nn.Module):
    def __init__(self, cutoff: float):
        super().__init__()
        self.cutoff = cutoff


def get_embedding_tensor(embed_basis: str = "gfn2-xtb", aux_basis: str = "aux56") -> torch.Tensor:
    """utils/qc.py:222-237; the table itself is the reference's data artefact
    (utils/pre_computed/*.pt) exported by oracle/make_golden.py."""
    path = _DATA / f"{embed_basis}_{aux_basis}.npy"
    if not path.exists():
        raise FileNotFoundError(f"no embedding table for {embed_basis}/{aux_basis} ({path})")
    return torch.from_numpy(np.load(path)).to(torch.get_default_dtype())


class Int2c1eEmbedding(nn.Module):
    def __init__(self, embed_basis: str = "gfn2-xtb", aux_basis: str = "aux28"):
        super().__init__()
        embed_ten = get_embedding_tensor(embed_basis, aux_basis)
        self.register_buffer("embed_ten", embed_ten)
        self.embed_dim = embed_ten.shape[1]

    def forward(self, at_no: torch.Tensor) -> torch.Tensor:
        return self.embed_ten[at_no.long()]


class Invariant(_TPHolder):
    """sqrt(sum_m x^2 + eps^2) - eps per irrep, or the squared norm (nn/o3layer.py:40-44)."""

    def __init__(self, muls, squared: bool = False, eps: float = 1e-5):
        super().__init__(num_irreps(muls))
        self.muls, self.squared, self.eps = tuple(muls), squared, eps

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.squared:
            return nodeops.invariant_dot(x, x, self.muls)[1]
        return nodeops.invariant_dot(x, x, self.muls)[0]


class EquivariantDot(_TPHolder):
    def __init__(self, muls):
        super().__init__(num_irreps(muls))
        self.muls = tuple(muls)

    def forward(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        return nodeops.invariant_dot(a, b, self.muls)[1]


class EquivariantLayerNorm(nn.Module):
    """nn/o3layer.py:112-171 on the cm layout."""

    def __init__(self, muls, affine: bool = True, eps: float = 1e-5):
        super().__init__()
        self.muls = tuple(muls)
        self.dim = irreps_dim(muls)
        self.num_scalar = muls[0]
        self.num_features = num_irreps(muls)
        self.register_buffer("scalar_index", torch.arange(self.num_scalar, dtype=torch.long))
        self.invariant = Invariant(muls, squared=True)
        self.scalar_mul = _E3nnBuffers(self.dim)
        weight, bias = torch.ones(self.num_features), torch.zeros(self.num_scalar)
        if affine:
            self.affine_weight = nn.Parameter(weight)
            self.affine_bias = nn.Parameter(bias)
        else:
            self.register_buffer("affine_weight", weight)
            self.register_buffer("affine_bias", bias)
        self.eps = eps

    def forward(self, V: torch.Tensor) -> torch.Tensor:
        assert V.shape[-1] == self.dim, "Input tensor must have the same last dimension as the irreps"
        return nodeops.irreps_norm(V, self.affine_weight, self.affine_bias, self.muls, self.eps)

    def with_passthrough(self, V: torch.Tensor):
        """(norm(V), V) -- nodeops._IrrepsNorm: feed the residual from the second output."""
        return nodeops.irreps_norm_pass(V, self.affine_weight, self.affine_bias, self.muls, self.eps)


class O3Linear(nn.Module):
    """e3nn o3.Linear(irreps, irreps, biases=True): per l, out[w,m] = sum_u W_l[u,w] in[u,m]/sqrt(mul_l)
    (+ bias on 0e); flat weight in (0e,1o,2e) order, each block row-major [u,w]."""

    def __init__(self, muls, biases: bool = True):
        super().__init__()
        self.muls = tuple(muls)
        self.weight = nn.Parameter(torch.randn(sum(m * m for m in muls)))
        self.bias = nn.Parameter(torch.zeros(muls[0] if biases else 0))
        self.register_buffer("output_mask", torch.ones(irreps_dim(muls)))

    def blocks(self):
        m0, m1, m2 = self.muls
        w = self.weight
        return (w[: m0 * m0].view(m0, m0), w[m0 * m0 : m0 * m0 + m1 * m1].view(m1, m1),
                w[m0 * m0 + m1 * m1 :].view(m2, m2))

    def forward(self, V: torch.Tensor) -> torch.Tensor:
        # one grouped tcgen05 launch: a problem per (l, m) block of the cm layout
        return gemm.irreps_linear(V, self.weight, self.bias if self.bias.numel() else None, self.muls)
