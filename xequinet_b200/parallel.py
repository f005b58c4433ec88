"""Data-parallel plumbing for molecule batches (first row): graphs never share edges, so ranks
take disjoint molecules and the only collective of a training step is ONE all-reduce of the gradients.

  shard_molecules(n_atoms_per_molecule, world, cutoff_neighbors=None)
        balanced contiguous-free assignment of molecules to ranks by estimated EDGE count (the edge kernels are
        the cost; reference: DistributedSampler + batch_size // world_size, run/train.py:100-121, which balances
        molecule counts only -- c4 has 30..70 atoms per molecule)
  allreduce_gradients(params, group=None, average=True)
        flat single bucket (865 141 fp32 = 3.5 MB at the defaults: latency-bound, so one launch), NCCL over
        NVLink on the GPU box, gloo in the CPU tests; replaces DDP's bucketed hooks (run/train.py:185-190)
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def estimate_edges(n_atoms: int, cutoff_neighbors: Optional[int] = None) -> int:
    """Directed edges of a molecule of n atoms: all pairs for small molecules (everything is inside a 5 A
    cutoff up to ~20 atoms), else n * cutoff_neighbors."""
    full = n_atoms * (n_atoms - 1)
    if cutoff_neighbors is None:
        return full
    return min(full, n_atoms * cutoff_neighbors)


def shard_molecules(n_atoms_per_molecule: Sequence[int], world: int, cutoff_neighbors: Optional[int] = None) -> List[List[int]]:
    """Molecule indices per rank, balanced by estimated edge count (longest-processing-time greedy: molecules in
    decreasing cost go to the currently lightest rank; ties by rank index, so the result is deterministic).
    Every rank's list is sorted, every molecule appears exactly once."""
    if world < 1:
        raise ValueError("world must be >= 1")
    cost = [estimate_edges(int(n), cutoff_neighbors) + int(n) for n in n_atoms_per_molecule]
    order = sorted(range(len(cost)), key=lambda i: (-cost[i], i))
    load = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += cost[i]
    return [sorted(x) for x in out]


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group=None, average: bool = True) -> None:
    """Sum (or average) the .grad of every parameter over the ranks with one flat all-reduce, in place.
    Parameters without a gradient contribute zeros (all ranks must pass the same parameter list)."""
    params = [p for p in params]
    if not params:
        return
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    dist.all_reduce(flat, group=group)
    if average:
        flat /= world
    off = 0
    for p in params:
        n = p.numel()
        piece = flat[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = piece.clone()
        else:
            p.grad.copy_(piece)
        off += n
