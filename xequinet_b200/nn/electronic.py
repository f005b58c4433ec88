"""Total-charge / total-spin conditioning of the node scalars behind the reference's module API
(xequinet/nn/electronic.py:13-90): an attention-like pooling of a per-graph embedding onto the atoms.
Same constructor arguments and parameter names; every Linear runs on K3, the per-graph normalisation
(`scatter_sum(attn, batch)`) on the segment-sum kernel."""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import keys, ops
from .layers import Linear, ResidualLayer


class _GraphConditioning(nn.Module):
    """Shared arithmetic of ChargeEmbedding / SpinEmbedding (nn/electronic.py:31-51, 71-90)."""

    n_feat: int

    def __init__(self, node_dim: int = 128, activation: str = "silu") -> None:
        super().__init__()
        self.node_dim = node_dim
        self.scale_factor = 1 / math.sqrt(node_dim)
        self.linear_q = Linear(node_dim, node_dim)
        self.linear_k = Linear(self.n_feat, node_dim, bias=False)
        self.linear_v = Linear(self.n_feat, node_dim, bias=False)
        self.residual = ResidualLayer(node_dim=node_dim, n_layers=2, activation=activation)

    def _condition(self, data: Dict[str, torch.Tensor], feat: torch.Tensor) -> Dict[str, torch.Tensor]:
        batch = data[keys.BATCH]
        x = data[keys.NODE_INVARIANT]
        norm = torch.maximum(feat, torch.ones_like(feat))
        query = self.linear_q(x)
        key = self.linear_k(feat / norm).index_select(0, batch)
        value = self.linear_v(feat).index_select(0, batch)
        dot = torch.sum(query * key, dim=-1)
        attn = F.softplus(dot * self.scale_factor)
        attn_sum = ops.segment_sum(attn, data["_xeq_ptr32"], batch).index_select(0, batch)
        embed = self.residual(value * (attn / attn_sum).unsqueeze(-1))
        data[keys.NODE_INVARIANT] = x + embed
        return data


class ChargeEmbedding(_GraphConditioning):
    n_feat = 2  # positive and negative charge are different features (nn/electronic.py:24-26)

    def forward(self, data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        if keys.TOTAL_CHARGE not in data:
            return data
        charge = data[keys.TOTAL_CHARGE].to(data[keys.NODE_INVARIANT].dtype).reshape(-1)
        return self._condition(data, F.relu(torch.stack([charge, -charge], dim=-1)))


class SpinEmbedding(_GraphConditioning):
    n_feat = 1  # spin is positive only (nn/electronic.py:64-66)

    def forward(self, data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        if keys.TOTAL_SPIN not in data:
            return data
        spin = data[keys.TOTAL_SPIN].to(data[keys.NODE_INVARIANT].dtype).reshape(-1, 1)
        return self._condition(data, spin)
