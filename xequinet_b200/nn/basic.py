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
    if graph is not None and graph.n_nodes == N and graph._src is not None:
        # a structure derived from an edge list is only a cache of THAT list: `edge_index` / `cell_offsets` / `cell`
        # in the dict stay the source of truth, as in the reference (nn/basic.py:67,119-128)
        if not graph.derived_from(data.get(keys.EDGE_INDEX), data.get(keys.CELL_OFFSETS) if has_cell else None,
                                  data[keys.CELL] if has_cell else None):
            graph = None
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
    data.pop(keys.POS_EFF, None)
    data.pop(keys.CELL_EFF, None)
    strain = torch.zeros((n_graphs, 3, 3), dtype=pos.dtype, device=pos.device)
    if compute_virial:
        # nn/basic.py:99-107: positions and cell displaced by a symmetrised per-graph strain (evaluated at zero).
        # The edge kernels differentiate with respect to the displaced positions and -- for periodic cells -- the
        # displaced cell (K2b returns dE/dcell from its per-edge d/dr records); autograd carries both to `strain`.
        if keys.HALO in data:
            raise NotImplementedError("compute_virial is not available for spatially sharded runs")
        strain.requires_grad_()
        symm = 0.5 * (strain + strain.transpose(1, 2))
        batch = data[keys.BATCH].long()
        data[keys.POS_EFF] = pos + torch.bmm(pos.unsqueeze(1), symm.index_select(0, batch)).squeeze(1)
        if has_cell:
            cell = data[keys.CELL].reshape(-1, 3, 3).to(pos.dtype)
            data[keys.CELL_EFF] = cell + torch.bmm(cell, symm)
            graph.seg_ptr = data["_xeq_ptr32"]
    data[keys.STRAIN] = strain
    return data


def compute_forces_only(energy: torch.Tensor, pos: torch.Tensor, training: bool = True) -> torch.Tensor:
    """nn/basic.py:143-159: the backward pass runs K2b (and records K2bb when training)."""
    grad_outputs: List[Optional[torch.Tensor]] = [torch.ones_like(energy)]
    # only d/dpos is requested: the backward formulas ask the engine (xequinet_b200/_state.py) and skip weight gradients
    pos_grad = torch.autograd.grad(outputs=[energy], inputs=[pos], grad_outputs=grad_outputs,
                                   retain_graph=training, create_graph=training, allow_unused=True)[0]
    if pos_grad is None:
        pos_grad = torch.zeros_like(pos)
    return -1.0 * pos_grad


def compute_virial_and_forces(energy: torch.Tensor, pos: torch.Tensor, strain: torch.Tensor, want_forces: bool,
                              training: bool, periodic: bool):
    """nn/basic.py:162-199: virial = -dE/dstrain (and forces = -dE/dpos from the same backward pass)."""
    inputs = ([pos] if want_forces else []) + [strain]
    grads = torch.autograd.grad(outputs=[energy], inputs=inputs, grad_outputs=[torch.ones_like(energy)],
                                retain_graph=training, create_graph=training, allow_unused=True)
    grads = [g if g is not None else torch.zeros_like(t) for g, t in zip(grads, inputs)]
    forces = -1.0 * grads[0] if want_forces else None
    return forces, -1.0 * grads[-1]


def compute_properties(data, compute_forces: bool = True, compute_virial: bool = False, training: bool = True,
                       extra_properties: Optional[List[str]] = None):
    results = {}
    if compute_virial:
        forces, virial = compute_virial_and_forces(data[keys.TOTAL_ENERGY], data[keys.POSITIONS], data[keys.STRAIN],
                                                   compute_forces, training, keys.CELL_EFF in data)
        if compute_forces:
            results[keys.FORCES] = forces
        results[keys.VIRIAL] = virial
    elif compute_forces:
        results[keys.FORCES] = compute_forces_only(data[keys.TOTAL_ENERGY], data[keys.POSITIONS], training)
    if extra_properties is not None:
        results.update({k: data[k] for k in extra_properties})
    return results
