"""Read-out heads behind the reference's API (xequinet/nn/output.py): EnergyOut :79-128 (the hot path),
ScalarOut :27-76, AtomicChargesOut :131-180, DipoleOut :183-243, PolarOut :246-327, SpatialOut :330-373.
Same constructor arguments, parameter names and data-dict keys.  Every Linear / o3.Linear runs on K3, per-graph
sums on the segment-sum kernel; equivariant features arrive in the cm layout (a 1-multiplicity irrep has the same
component order in both layouts, so the outputs need no conversion)."""
from __future__ import annotations

import math
from pathlib import Path
from typing import Dict, Iterable, List, Optional

import numpy as np

import torch
import torch.nn as nn

from .. import keys, ops
from .irreps import parse_irreps
from .layers import Gate, Linear, O3LinearMap, _E3nnBuffers, resolve_activation


class OutputModule(nn.Module):
    extra_properties: List[str]


class EnergyOut(OutputModule):
    def __init__(self, node_dim: int = 128, hidden_dim: int = 64, activation: str = "silu", node_shift: float = 0.0,
                 node_scale: float = 1.0, **kwargs) -> None:
        super().__init__()
        self.node_dim, self.hidden_dim = node_dim, hidden_dim
        final_linear = Linear(self.hidden_dim, 1)
        final_linear.weight.data *= node_scale  # nn/output.py:104-106: baked in at construction
        nn.init.constant_(final_linear.bias, node_shift)
        self.out_mlp = nn.Sequential(Linear(self.node_dim, self.hidden_dim), resolve_activation(activation),
                                     final_linear)
        self.extra_properties = [keys.TOTAL_ENERGY, keys.ATOMIC_ENERGIES]

    def forward(self, data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        atom_eng_out = self.out_mlp(data[keys.NODE_INVARIANT]).reshape(-1)
        if keys.ATOMIC_ENERGIES in data:
            atomic_energies = data[keys.ATOMIC_ENERGIES] + atom_eng_out
        else:
            atomic_energies = atom_eng_out
        # scatter_sum(atomic_energies, batch) (nn/output.py:124) as a contiguous segment sum
        data[keys.TOTAL_ENERGY] = ops.segment_sum(atomic_energies, data["_xeq_ptr32"], data[keys.BATCH])
        data[keys.ATOMIC_ENERGIES] = atomic_energies
        return data


def _mlp(node_dim: int, hidden_dim: int, out_dim: int, activation: str, zero_bias: bool = True) -> nn.Sequential:
    mlp = nn.Sequential(Linear(node_dim, hidden_dim), resolve_activation(activation), Linear(hidden_dim, out_dim))
    if zero_bias:
        nn.init.zeros_(mlp[0].bias)
        nn.init.zeros_(mlp[2].bias)
    return mlp


def segment_sum_cols(src: torch.Tensor, data: Dict[str, torch.Tensor]) -> torch.Tensor:
    """scatter_sum(src [N, c], batch) -> [G, c]: one contiguous segment sum per column."""
    ptr, batch = data["_xeq_ptr32"], data[keys.BATCH]
    if src.dim() == 1:
        return ops.segment_sum(src, ptr, batch)
    return torch.stack([ops.segment_sum(src[:, c].contiguous(), ptr, batch) for c in range(src.shape[1])], dim=1)


def _atoms_per_graph(data: Dict[str, torch.Tensor], dtype) -> torch.Tensor:
    ptr = data[keys.BATCH_PTR]
    return (ptr[1:] - ptr[:-1]).to(dtype)


class ScalarOut(OutputModule):
    def __init__(self, node_dim: int = 128, hidden_dim: int = 64, activation: str = "silu", node_shift: float = 0.0,
                 node_scale: float = 1.0, reduce_op: Optional[str] = "sum", output_field: str = keys.SCALAR_OUTPUT,
                 **kwargs) -> None:
        super().__init__()
        if reduce_op not in (None, "sum", "add", "mean"):
            raise NotImplementedError(f"reduce_op {reduce_op!r}: the segment kernel sums (sum / mean / None)")
        self.node_dim, self.hidden_dim = node_dim, hidden_dim
        self.out_mlp = _mlp(node_dim, hidden_dim, 1, activation, zero_bias=False)
        self.out_mlp[2].weight.data *= node_scale
        nn.init.constant_(self.out_mlp[2].bias, node_shift)
        self.reduce_op, self.output_field = reduce_op, output_field
        self.extra_properties = [output_field]

    def forward(self, data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        res = self.out_mlp(data[keys.NODE_INVARIANT]).reshape(-1)
        if self.reduce_op is not None:
            res = segment_sum_cols(res, data)
            if self.reduce_op == "mean":
                res = res / _atoms_per_graph(data, res.dtype).clamp(min=1)
        data[self.output_field] = res
        return data


class AtomicChargesOut(OutputModule):
    def __init__(self, node_dim: int = 128, hidden_dim: int = 64, activation: str = "silu", conservation: bool = True,
                 **kwargs) -> None:
        super().__init__()
        self.node_dim, self.hidden_dim = node_dim, hidden_dim
        self.out_mlp = _mlp(node_dim, hidden_dim, 1, activation)
        self.conservation = conservation
        self.extra_properties = [keys.ATOMIC_CHARGES]

    def forward(self, data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        batch = data[keys.BATCH]
        q = self.out_mlp(data[keys.NODE_INVARIANT]).reshape(-1)
        if self.conservation:  # spread the missing charge evenly over the atoms of each graph (nn/output.py:165-177)
            raw_total = segment_sum_cols(q, data)
            total = data[keys.TOTAL_CHARGE].to(q.dtype).reshape(-1) if keys.TOTAL_CHARGE in data else torch.zeros_like(raw_total)
            delta = (total - raw_total) / _atoms_per_graph(data, q.dtype)
            q = q + delta.index_select(0, batch)
        data[keys.ATOMIC_CHARGES] = q
        return data


class DipoleOut(OutputModule):
    def __init__(self, node_dim: int = 128, node_irreps: Iterable = "128x0e + 64x1o + 32x2e", hidden_dim: int = 64,
                 hidden_irreps: Iterable = "32x1o", activation: str = "silu", magnitude: bool = False, **kwargs) -> None:
        super().__init__()
        self.node_dim, self.hidden_dim = node_dim, hidden_dim
        self.muls, self.hidden_muls = parse_irreps(node_irreps), parse_irreps(hidden_irreps)
        self.scalar_out_mlp = _mlp(node_dim, hidden_dim, 1, activation)
        self.equi_out_mlp = nn.Sequential(O3LinearMap(self.muls, self.hidden_muls),
                                          Gate(self.hidden_muls, activation=activation),
                                          O3LinearMap(self.hidden_muls, (0, 1, 0)))
        self.magnitude = magnitude
        self.extra_properties = [keys.DIPOLE if not magnitude else keys.DIPOLE_MAGNITUDE]

    def forward(self, data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        equi_out = self.equi_out_mlp(data[keys.NODE_EQUIVARIANT])[:, [2, 0, 1]]  # [y, z, x] -> [x, y, z]
        scalar_out = self.scalar_out_mlp(data[keys.NODE_INVARIANT])
        dipole = segment_sum_cols(equi_out * scalar_out, data)
        data[keys.DIPOLE] = dipole
        if self.magnitude:
            data[keys.DIPOLE_MAGNITUDE] = torch.linalg.norm(dipole, dim=-1)
        return data


class PolarOut(OutputModule):
    def __init__(self, node_dim: int = 128, node_irreps: Iterable = "128x0e + 64x1o + 32x2e", hidden_dim: int = 64,
                 hidden_irreps: Iterable = "64x0e + 16x2e", activation: str = "silu", isotropic: bool = False,
                 **kwargs) -> None:
        super().__init__()
        self.node_dim, self.hidden_dim = node_dim, hidden_dim
        self.muls, self.hidden_muls = parse_irreps(node_irreps), parse_irreps(hidden_irreps)
        self.scalar_out_mlp = _mlp(node_dim, hidden_dim, 2, activation)
        self.equi_out_mlp = nn.Sequential(O3LinearMap(self.muls, self.hidden_muls, biases=True),
                                          Gate(self.hidden_muls, activation=activation),
                                          O3LinearMap(self.hidden_muls, (1, 0, 1), biases=True))
        self.rsh_conv = _E3nnBuffers(6)
        self.isotropic = isotropic
        self.extra_properties = [keys.POLARIZABILITY if not isotropic else keys.ISO_POLARIZABILITY]

    def forward(self, data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        equi_out = self.equi_out_mlp(data[keys.NODE_EQUIVARIANT])  # [N, 1 + 5]
        scalar_out = self.scalar_out_mlp(data[keys.NODE_INVARIANT])  # [N, 2]
        # ElementwiseTensorProduct("1x0e + 1x2e", "2x0e") (nn/output.py:287): each irrep times its own scalar
        src = torch.cat([equi_out[:, :1] * scalar_out[:, :1], equi_out[:, 1:] * scalar_out[:, 1:2]], dim=1)
        polar = segment_sum_cols(src, data)
        zero_order, d = polar[:, 0], polar[:, 1:6]
        d_norm = torch.linalg.norm(d, dim=-1)
        dxy, dyz, dz2, dzx, dx2_y2 = d.unbind(dim=1)
        r3 = 1 / math.sqrt(3)
        xx = r3 * (d_norm - dz2) + dx2_y2 + zero_order  # nn/output.py:305-322
        yy = r3 * (d_norm - dz2) - dx2_y2 + zero_order
        zz = r3 * (d_norm + 2 * dz2) + zero_order
        polarizability = torch.stack([torch.stack([xx, dxy, dzx], dim=-1), torch.stack([dxy, yy, dyz], dim=-1),
                                      torch.stack([dzx, dyz, zz], dim=-1)], dim=-2)
        data[keys.POLARIZABILITY] = polarizability
        if self.isotropic:
            data[keys.ISO_POLARIZABILITY] = torch.diagonal(polarizability, dim1=-2, dim2=-1).mean(dim=-1)
        return data


def atomic_masses() -> torch.Tensor:
    """utils/qc.py:181-190, exported as a data artefact by oracle/make_golden_heads.py."""
    path = Path(__file__).resolve().parent.parent / "data" / "atom_mass.npy"
    return torch.from_numpy(np.load(path)).to(torch.get_default_dtype())


class SpatialOut(OutputModule):
    masses: torch.Tensor

    def __init__(self, node_dim: int = 128, hidden_dim: int = 64, activation: str = "silu", **kwargs) -> None:
        super().__init__()
        self.node_dim, self.hidden_dim = node_dim, hidden_dim
        self.scalar_out_mlp = _mlp(node_dim, hidden_dim, 1, activation)
        self.register_buffer("masses", atomic_masses())
        self.extra_properties = [keys.SPATIAL_EXTENT]

    def forward(self, data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        batch, pos = data[keys.BATCH], data[keys.POSITIONS]
        masses = self.masses[data[keys.ATOMIC_NUMBERS].long()].unsqueeze(-1)
        centroids = segment_sum_cols(masses * pos, data) / segment_sum_cols(masses, data)
        rel = pos - centroids.index_select(0, batch)  # the reference shifts `pos` in place (nn/output.py:364)
        scalar_out = self.scalar_out_mlp(data[keys.NODE_INVARIANT])
        spatial = torch.square(rel).sum(dim=1, keepdim=True)
        data[keys.SPATIAL_EXTENT] = segment_sum_cols(scalar_out * spatial, data)
        return data


def resolve_output(mode: str, **kwargs) -> OutputModule:
    """nn/output.py:468-480.  `cartesian` (CartTensorOut: self-mix tensor products + Wigner-3j reduced tensor
    products, nn/xe3net.py) is outside the XPaiNN path this package accelerates."""
    output_factory = {
        "scalar": ScalarOut,
        "energy": EnergyOut,
        "charges": AtomicChargesOut,
        "atomic_charges": AtomicChargesOut,
        "dipole": DipoleOut,
        "polar": PolarOut,
        "spatial": SpatialOut,
    }
    if mode not in output_factory:
        raise NotImplementedError(f"output mode {mode!r} is not available on the B200 path "
                                  f"(available: {sorted(output_factory)})")
    return output_factory[mode](**kwargs)
