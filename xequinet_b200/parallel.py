"""Data-parallel plumbing for molecule batches (first row): graphs never share edges, so ranks
take disjoint molecules and the only collective of a training step is ONE all-reduce of the gradients.

  shard_molecules(n_atoms_per_molecule, world, cutoff_neighbors=None)
        balanced contiguous-free assignment of molecules to ranks by estimated EDGE count (the edge kernels are
        the cost; reference: DistributedSampler + batch_size // world_size, run/train.py:100-121, which balances
        molecule counts only -- c4 has 30..70 atoms per molecule)
  allreduce_gradients(params, group=None, average=True)
        flat single bucket (865 141 fp32 = 3.5 MB at the defaults: latency-bound, so one launch), NCCL over
        NVLink on the GPU box, gloo in the CPU tests; replaces DDP's bucketed hooks (run/train.py:185-190)
  FlatGradients(params, n_buckets=3, group=None, average=True)
        the training-loop form of the same collective: ONE flat fp32 buffer owns every gradient (each p.grad is a
        view into it, so there is no gather / scatter copy around the collective), split into contiguous buckets in
        reverse parameter order; a bucket is all-reduced on a side stream as soon as autograd has finalised its
        last gradient, i.e. while the backward pass of the earlier layers is still running.  Capturable: inside
        a CUDA-graph capture the side stream is forked from / joined to the capturing stream.
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


class FlatGradients:
    """Flat gradient storage + bucketed, overlapped all-reduce (see the module docstring).

        flat = FlatGradients(model.parameters())
        ...
        flat.zero()              # start of the step: p.grad = None, autograd ASSIGNS fresh gradients (no accumulate kernels)
        loss.backward()          # a bucket is gathered (one multi-tensor copy) and all-reduced as soon as it is complete
        flat.finish()            # join the side stream; every p.grad is now a view of the averaged flat buffer
        optimizer.step()

    Per bucket the step costs one `_foreach_copy_` launch and one all-reduce (three buckets: six launches per step,
    against ~75 gather + ~75 scatter copies of a cat / split formulation, or ~75 accumulate kernels when autograd is
    made to add into preallocated views).  With world size 1 (or no process group) nothing is hooked: finish() is a
    no-op.  Gradients finalise in reverse order of use (read-out first, embedding last), so bucket 0 holds the LAST
    parameters of the list.  Capturable: under CUDA-graph capture the side stream is forked from / joined to the
    capturing stream and the gradient tensors autograd produced during capture keep their addresses on replay."""

    def __init__(self, params: Iterable[torch.nn.Parameter], n_buckets: int = 3, group=None, average: bool = True):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatGradients: no trainable parameters")
        self.group, self.average = group, average
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        p0 = self.params[0]
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=p0.dtype, device=p0.device)
        off = 0
        self._range = {}
        self._view = {}
        for p in self.params:
            n = p.numel()
            self._range[id(p)] = (off, off + n)
            self._view[id(p)] = self.flat[off:off + n].view_as(p)
            off += n
        # contiguous buckets of roughly equal size over the REVERSED parameter list
        n_buckets = max(1, min(int(n_buckets), len(self.params)))
        target = total / n_buckets
        self.buckets: List[List[torch.nn.Parameter]] = [[]]
        acc = 0
        for p in reversed(self.params):
            if acc >= target * len(self.buckets) and len(self.buckets) < n_buckets:
                self.buckets.append([])
            self.buckets[-1].append(p)
            acc += p.numel()
        self._bucket_of = {id(p): b for b, ps in enumerate(self.buckets) for p in ps}
        self._slice = [(min(self._range[id(p)][0] for p in ps), max(self._range[id(p)][1] for p in ps)) for ps in self.buckets]
        self._pending = [0] * len(self.buckets)
        self._side = torch.cuda.Stream(device=p0.device) if p0.is_cuda else None
        self._launched = [False] * len(self.buckets)
        if self.world > 1:
            for p in self.params:
                p.register_post_accumulate_grad_hook(self._hook)
        self.zero()

    # -- step protocol ------------------------------------------------------------------------------------
    def zero(self) -> None:
        for p in self.params:
            p.grad = None
        self._pending = [len(ps) for ps in self.buckets]
        self._launched = [False] * len(self.buckets)

    def _reduce_bucket(self, b: int) -> None:
        ps = [p for p in self.buckets[b] if p.grad is not None]
        missing = [p for p in self.buckets[b] if p.grad is None]
        if ps:
            torch._foreach_copy_([self._view[id(p)] for p in ps], [p.grad for p in ps])  # one multi-tensor launch
            for p in ps:
                p.grad = self._view[id(p)]
        for p in missing:  # no gradient on this rank: contributes zeros, receives the other ranks' share
            self._view[id(p)].zero_()
            p.grad = self._view[id(p)]
        lo, hi = self._slice[b]
        piece = self.flat[lo:hi]
        if self._side is not None:
            self._side.wait_stream(torch.cuda.current_stream(piece.device))
            with torch.cuda.stream(self._side):
                if self.average and dist.get_backend(self.group) == "nccl":
                    dist.all_reduce(piece, op=dist.ReduceOp.AVG, group=self.group)  # averaged inside the collective
                else:
                    dist.all_reduce(piece, group=self.group)
                    if self.average:
                        piece.div_(self.world)
        else:
            dist.all_reduce(piece, group=self.group)
            if self.average:
                piece.div_(self.world)
        self._launched[b] = True

    def _hook(self, p: torch.nn.Parameter) -> None:
        b = self._bucket_of[id(p)]
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._reduce_bucket(b)

    def finish(self) -> None:
        """All buckets reduced and visible to the current stream.  Parameters that received no gradient in this
        backward pass never fire their hook: their buckets are reduced here."""
        if self.world == 1:
            return
        for b in range(len(self.buckets)):
            if not self._launched[b]:
                self._reduce_bucket(b)
        if self._side is not None:
            torch.cuda.current_stream(self.flat.device).wait_stream(self._side)
