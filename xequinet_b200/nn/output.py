"""Energy head behind the reference's API (xequinet/nn/output.py:79-128)."""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn as nn

from .. import keys, ops
from .layers import Linear, resolve_activation


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


def resolve_output(mode: str, **kwargs) -> OutputModule:
    """nn/output.py:468-480; only the energy head is on the accelerated path."""
    if mode != "energy":
        raise NotImplementedError(f"output mode {mode!r} is outside the B200 hot path (energy/forces only)")
    return EnergyOut(**kwargs)
