"""XPaiNN blocks behind the reference's module API (xequinet/nn/xpainn.py):
XEmbedding :14-83, XPainnMessage :86-161, XPainnUpdate :164-231 -- same constructor
arguments, parameter names and data-dict protocol; the E-sized work of the message block is
one call into the fused CUDA edge kernel."""
from __future__ import annotations

from typing import Dict, Iterable

import torch
import torch.nn as nn

from .. import keys, nodeops, ops
from . import cm
from .irreps import irreps_dim, num_irreps, parse_irreps
from .layers import (CosineCutoff, LayerNorm, Linear, EquivariantDot, EquivariantLayerNorm, Int2c1eEmbedding, Invariant, O3Linear,
                     SphericalBesselj0, _E3nnBuffers, resolve_activation)


class XEmbedding(nn.Module):
    def __init__(self, node_dim: int = 128, node_irreps: Iterable = "128x0e + 64x1o + 32x2e",
                 embed_basis: str = "gfn2-xtb", aux_basis: str = "aux56", num_basis: int = 20,
                 rbf_kernel: str = "bessel", cutoff: float = 5.0, cutoff_fn: str = "cosine") -> None:
        super().__init__()
        self.node_dim = node_dim
        self.muls = parse_irreps(node_irreps)
        self.node_num_irreps = num_irreps(self.muls)
        if embed_basis == "one-hot":
            self.embedding = nn.Embedding(100, self.node_dim, padding_idx=0)
        else:
            int2c1e = Int2c1eEmbedding(embed_basis, aux_basis)
            self.embedding = nn.Sequential(int2c1e, Linear(int2c1e.embed_dim, self.node_dim))
            nn.init.zeros_(self.embedding[1].bias)
        if rbf_kernel != "bessel" or cutoff_fn != "cosine":
            raise NotImplementedError("the B200 edge kernel implements the bessel basis with the cosine cutoff "
                                      "(the XPaiNN defaults, nn/model.py:62-64)")
        self.rbf = SphericalBesselj0(num_basis, cutoff)
        self.cutoff_fn = CosineCutoff(cutoff)

    def forward(self, data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        at_no = data[keys.ATOMIC_NUMBERS]
        x = self.embedding(at_no.long() if isinstance(self.embedding, nn.Embedding) else at_no)
        data[keys.NODE_INVARIANT] = x
        # rbf / cutoff / spherical harmonics are never materialised: the edge kernel evaluates
        # them from positions; it only needs the (learnable, shared) frequencies.
        data[keys.RBF_FREQ] = self.rbf.freq
        data[keys.RBF_CUTOFF] = self.rbf.cutoff
        data[keys.NODE_EQUIVARIANT] = torch.zeros((x.shape[0], irreps_dim(self.muls)), device=x.device, dtype=x.dtype)
        return data


def _norm_pass(norm: nn.Module, t: torch.Tensor):
    """(norm(t), t) with t routed through the norm's pass-through output when the norm is one of ours."""
    if hasattr(norm, "with_passthrough"):
        return norm.with_passthrough(t)
    return norm(t), t


class XPainnMessage(nn.Module):
    def __init__(self, node_dim: int = 128, node_irreps: Iterable = "128x0e + 64x1o + 32x2e", num_basis: int = 20,
                 activation: str = "silu", layer_norm: bool = True) -> None:
        super().__init__()
        self.node_dim = node_dim
        self.muls = parse_irreps(node_irreps)
        self.node_num_irreps = num_irreps(self.muls)
        self.hidden_dim = self.node_dim + self.node_num_irreps * 2
        self.num_basis = num_basis
        self.scalar_mlp = nn.Sequential(
            Linear(self.node_dim, self.node_dim),
            resolve_activation(activation),
            Linear(self.node_dim, self.hidden_dim),
        )
        self.rbf_lin = Linear(self.num_basis, self.hidden_dim, bias=True)
        self.rsh_conv = _E3nnBuffers(irreps_dim(self.muls))
        self.norm = LayerNorm(self.node_dim) if layer_norm else nn.Identity()
        self.o3norm = EquivariantLayerNorm(self.muls) if layer_norm else nn.Identity()
        self._dims = None

    def forward(self, data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        x, V = data[keys.NODE_INVARIANT], data[keys.NODE_EQUIVARIANT]
        cutoff = float(data[keys.RBF_CUTOFF])
        if self._dims is None or self._dims.cutoff != cutoff:
            self._dims = ops.Dims(self.node_dim, *self.muls, self.num_basis, cutoff)
        # x and V feed both their norm and the residual inside the edge kernel: take the residual from the norm's
        # pass-through output, so that the two gradient contributions are summed inside the norm's backward kernel
        xn, x = _norm_pass(self.norm, x)
        v, V = _norm_pass(self.o3norm, V)
        s = self.scalar_mlp(xn)
        plan = data.get(keys.HALO)
        if plan is not None:
            # spatially sharded run (xequinet_b200/domain.py): the rows of boundary atoms travel to the ranks
            # that ghost them -- one exchange of [s | v] per layer; node-side work stays on owned atoms
            from ..domain import halo_gather
            H = s.shape[1]
            sv = halo_gather(torch.cat([s, v], dim=1), plan)
            s, v = sv[:, :H], sv[:, H:]
            pad = (0, 0, 0, plan.n_ghost)
            x_loc, V_loc = torch.nn.functional.pad(x, pad), torch.nn.functional.pad(V, pad)
            x_new, V_new = ops.edge_message(x_loc, V_loc, s, v, data[keys.POSITIONS], self.rbf_lin.weight,
                                            self.rbf_lin.bias, data[keys.RBF_FREQ], data[keys.GRAPH], self._dims)
            data[keys.NODE_INVARIANT] = x_new[: plan.n_owned]
            data[keys.NODE_EQUIVARIANT] = V_new[: plan.n_owned]
            return data
        # with compute_virial the kernels differentiate the strained positions / cell (nn/basic.py:99-107)
        x_new, V_new = ops.edge_message(x, V, s, v, data.get(keys.POS_EFF, data[keys.POSITIONS]), self.rbf_lin.weight,
                                        self.rbf_lin.bias, data[keys.RBF_FREQ], data[keys.GRAPH], self._dims,
                                        cell=data.get(keys.CELL_EFF))
        data[keys.NODE_INVARIANT] = x_new
        data[keys.NODE_EQUIVARIANT] = V_new
        return data


class XPainnUpdate(nn.Module):
    def __init__(self, node_dim: int = 128, node_irreps: Iterable = "128x0e + 64x1o + 32x2e", activation: str = "silu",
                 layer_norm: bool = True) -> None:
        super().__init__()
        self.node_dim = node_dim
        self.muls = parse_irreps(node_irreps)
        self.node_num_irreps = num_irreps(self.muls)
        self.hidden_dim = self.node_dim * 2 + self.node_num_irreps
        self.update_U = O3Linear(self.muls, biases=True)
        self.update_V = O3Linear(self.muls, biases=True)
        self.invariant = Invariant(self.muls)
        self.equidot = EquivariantDot(self.muls)
        self.dot_lin = Linear(self.node_num_irreps, self.node_dim, bias=False)
        self.rsh_conv = _E3nnBuffers(irreps_dim(self.muls))
        self.update_mlp = nn.Sequential(
            Linear(self.node_dim + self.node_num_irreps, self.node_dim),
            resolve_activation(activation),
            Linear(self.node_dim, self.hidden_dim),
        )
        self.norm = LayerNorm(self.node_dim) if layer_norm else nn.Identity()
        self.o3norm = EquivariantLayerNorm(self.muls) if layer_norm else nn.Identity()

    def forward(self, data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        x, V = data[keys.NODE_INVARIANT], data[keys.NODE_EQUIVARIANT]
        M, C = self.node_num_irreps, self.node_dim
        xn, x = _norm_pass(self.norm, x)
        vn, V = _norm_pass(self.o3norm, V)
        U = self.update_U(vn)
        W = self.update_V(vn)
        n, t0, U = nodeops.invariant_dot_pass(U, W, self.muls)  # Invariant(W), EquivariantDot(U, W): one kernel; U passes through
        a = self.update_mlp(nodeops.cat2(xn, n))  # [a_vv | a_sv | a_ss]
        t = self.dot_lin(t0)
        # x + a_sv * t + a_ss ,  V + expand(a_vv) * U : one kernel
        data[keys.NODE_INVARIANT], data[keys.NODE_EQUIVARIANT] = nodeops.gate_residual(a, U, t, x, V, self.muls)
        return data
