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
    csrc/node_gemm.cu) together with all of its derivatives.  The kernel addresses operands in 16-byte
    units.  A single output feature (the 64 -> 1 energy read-out, nn/output.py:107-111) is a row-dot
    kernel with kernel-backed derivatives (gemm.linear_to_scalar); other widths that are not multiples of 4
    (the 2 -> C / 1 -> C projections of the charge / spin embedding, nn/electronic.py:25-26, the 64 -> 2
    polarizability read-out) are zero-padded to the next multiple of 4 and the result sliced: no linear
    layer of the path is a library call."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        lead = None
        if x.dim() != 2:
            lead, x = x.shape[:-1], x.reshape(-1, x.shape[-1])
        w, b = self.weight, self.bias
        if self.out_features == 1:
            y = gemm.linear_to_scalar(x, w, b)
            return y if lead is None else y.reshape(*lead, 1)
        pk, pn = -self.in_features % 4, -self.out_features % 4
        if pk:
            x, w = F.pad(x, (0, pk)), F.pad(w, (0, pk))
        if pn:
            w = F.pad(w, (0, 0, 0, pn))
            b = F.pad(b, (0, pn)) if b is not None else None
        y = gemm.linear(x, w, b)
        if pn:
            y = y[:, : self.out_features]
        return y if lead is None else y.reshape(*lead, self.out_features)


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


class O3LinearMap(nn.Module):
    """e3nn o3.Linear between DIFFERENT multiplicities, as the read-out heads use it (nn/output.py:218-222,
    282-286): every l present on both sides gets a path out[w,m] = sum_u W_l[u,w] in[u,m] / sqrt(mul_in_l); flat
    weight with the blocks in ascending l, each row-major [u,w]; bias on the 0e outputs.  Both sides are held in
    the cm layout, so a path is (2l+1) plain GEMMs on column slices (K3); 1-wide outputs are a weighted row sum."""

    def __init__(self, muls_in, muls_out, biases: bool = False):
        super().__init__()
        self.muls_in, self.muls_out = tuple(muls_in), tuple(muls_out)
        self.paths = [l for l in range(3) if self.muls_in[l] and self.muls_out[l]]
        self.weight = nn.Parameter(torch.randn(sum(self.muls_in[l] * self.muls_out[l] for l in self.paths)))
        nb = self.muls_out[0] if biases else 0
        if nb:
            self.bias = nn.Parameter(torch.zeros(nb))
        else:  # e3nn registers an empty buffer when there is no bias: the state_dict key exists either way
            self.register_buffer("bias", torch.Tensor())
        self.register_buffer("output_mask", torch.ones(irreps_dim(self.muls_out)))

    def forward(self, V: torch.Tensor) -> torch.Tensor:
        N = V.shape[0]
        outs, woff, ioff = [], 0, 0
        for l in range(3):
            mi, mo, d = self.muls_in[l], self.muls_out[l], 2 * l + 1
            if mo and l not in self.paths:
                outs.append(V.new_zeros(N, d * mo))
            elif mo:
                W = self.weight[woff : woff + mi * mo].view(mi, mo)
                woff += mi * mo
                alpha = 1.0 / math.sqrt(mi)
                if mo % 4 == 0 and mi % 4 == 0:
                    bias = self.bias if (l == 0 and self.bias.numel()) else None
                    blk = [gemm.mm(V[:, ioff + m * mi : ioff + (m + 1) * mi], W, alpha=alpha, bias=bias)
                           for m in range(d)]
                    outs.append(torch.cat(blk, dim=1) if d > 1 else blk[0])
                else:
                    y = (V[:, ioff : ioff + d * mi].reshape(N, d, mi, 1) * W.view(1, 1, mi, mo)).sum(2) * alpha
                    if l == 0 and self.bias.numel():
                        y = y + self.bias
                    outs.append(y.reshape(N, d * mo))
            ioff += d * mi
        return torch.cat(outs, dim=1) if len(outs) > 1 else outs[0]


class Gate(nn.Module):
    """nn/o3layer.py:47-75 (refine=False): x * act'(Invariant(x)) per irrep with act' = activation / x
    (silu -> sigmoid, nn/basic.py:244-246), on the cm layout."""

    def __init__(self, muls, activation: str = "silu"):
        super().__init__()
        self.muls = tuple(muls)
        self.invariant = _TPHolder(num_irreps(muls))
        div_x = {"silu": "sigmoid", "relu": "identity", "leakyrelu": "identity"}
        name = activation.lower()
        self.activation = resolve_activation(div_x.get(name, name))
        self.scalar_mul = _E3nnBuffers(irreps_dim(muls))
        self.eps = 1e-5

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        inv = torch.sqrt(cm.irrep_dot(x, x, self.muls) + self.eps * self.eps) - self.eps  # nn/o3layer.py:40-44
        return x * cm.expand_gate(self.activation(inv), self.muls)


class ResidualLayer(nn.Module):
    """nn/basic.py:11-31: (x + mlp(x)) / sqrt(2), mlp = n_layers x (bias-free Linear, activation) with ONE shared
    activation module (state_dict keys mlp.0.weight, mlp.2.weight, ...)."""

    def __init__(self, node_dim: int = 128, n_layers: int = 2, activation: str = "silu"):
        super().__init__()
        act_fn = resolve_activation(activation)
        self.mlp = nn.Sequential()
        for _ in range(n_layers):
            self.mlp.append(Linear(node_dim, node_dim, bias=False))
            self.mlp.append(act_fn)
        self.inv_sqrt_2 = 1 / math.sqrt(2)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.inv_sqrt_2 * (x + self.mlp(x))
