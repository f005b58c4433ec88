"""Edge preparation and gradient properties (xequinet/nn/basic.py:60-238) for the B200 path.

The reference materialises edge vectors / lengths here; on this path they are recomputed
inside the fused edge kernel, so this step only resolves the CSR neighbour structure and the
batch bookkeeping, and marks `pos` for differentiation."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .. import keys, ops
from ..graph import NeighborGraph, graph_from_edge_index


def compute_edge_data(data: Dict[str, torch.Tensor], compute_forces: bool = True, compute_virial: bool = False):
    pos = data[keys.POSITIONS]
    if compute_virial:
        raise NotImplementedError("compute_virial (strain derivative, nn/basic.py:93-107) is not on the B200 path yet")
    if not pos.is_cuda:
        raise RuntimeError("xequinet_b200 models run on CUDA tensors only: there is no CPU fallback")
    if pos.dtype != torch.float32:
        raise RuntimeError("xequinet_b200 kernels compute in fp32")
    N = pos.shape[0]
    if keys.BATCH not in data:  # nn/basic.py:70-77
        data[keys.BATCH] = torch.zeros(N, dtype=torch.long, device=pos.device)
        data[keys.BATCH_PTR] = torch.tensor([0, N], dtype=torch.long, device=pos.device)
    n_graphs = data[keys.BATCH_PTR].numel() - 1
    data["_xeq_ptr32"] = data[keys.BATCH_PTR].to(torch.int32).contiguous()
    has_cell = keys.CELL in data
    graph: Optional[NeighborGraph] = data.get(keys.GRAPH)
    if graph is None or graph.n_nodes != N:
        graph = graph_from_edge_index(
            data[keys.EDGE_INDEX], N, n_graphs,
            cell_offsets=data.get(keys.CELL_OFFSETS) if has_cell else None,
            cell=data[keys.CELL] if has_cell else None,
            batch=data[keys.BATCH],
            ptr=data[keys.BATCH_PTR],
        )
        data[keys.GRAPH] = graph
    if compute_forces:
        pos.requires_grad_()  # nn/basic.py:90-91
    return data


def compute_forces_only(energy: torch.Tensor, pos: torch.Tensor, training: bool = True) -> torch.Tensor:
    """nn/basic.py:143-159: the backward pass runs K2b (and records K2bb when training)."""
    grad_outputs: List[Optional[torch.Tensor]] = [torch.ones_like(energy)]
    with ops.param_grads(False):  # only d/dpos is requested here
        pos_grad = torch.autograd.grad(outputs=[energy], inputs=[pos], grad_outputs=grad_outputs,
                                       retain_graph=training, create_graph=training, allow_unused=True)[0]
    if pos_grad is None:
        pos_grad = torch.zeros_like(pos)
    return -1.0 * pos_grad


def compute_properties(data, compute_forces: bool = True, compute_virial: bool = False, training: bool = True,
                       extra_properties: Optional[List[str]] = None):
    results = {}
    if compute_forces:
        results[keys.FORCES] = compute_forces_only(data[keys.TOTAL_ENERGY], data[keys.POSITIONS], training)
    if extra_properties is not None:
        results.update({k: data[k] for k in extra_properties})
    return results
