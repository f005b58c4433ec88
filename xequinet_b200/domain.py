"""Spatial domain sharding with halo exchange for one large periodic cell (BASELINE.json configs[4],
SURVEY.md 8e: "not in the reference -- new design").  One process per GPU; NCCL (or gloo on CPU for the
host-logic tests) only moves halo rows, there is no other data-path collective.

Decomposition: P slabs along lattice axis 0 (fractional coordinate f0 in [r/P, (r+1)/P) -> rank r);
the slab width must be >= the cutoff, so every neighbour of an owned atom is owned by the rank itself or
by rank r-1 / r+1 (periodic).  Each rank works on  [owned atoms | ghost atoms]:
  * ghost POSITIONS arrive with the lattice shift of the periodic wrap already applied by the sender,
    so axis 0 is an open axis of the local problem (K1 runs with pbc = (False, True, True));
  * K1 rows are kept for owned centers only (ghost rows are dropped: no redundant edge work);
  * per message layer one exchange ships the filter input s|v rows of boundary atoms to the ranks that
    ghost them (nn/xpainn.py:142,151 gathers them by neighbour); its autograd backward ships the ghost-row
    gradients back and ADDS them onto the owners in a fixed order (deterministic);
  * the position exchange is differentiable too, so  -dE_total/dpos_owned  (nn/basic.py:150-156) picks up
    the contributions of edges that live on the neighbouring ranks.
Energies are per-rank partial sums over owned atoms; forces stay sharded."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from . import keys


@dataclass
class HaloPlan:
    rank: int
    world: int
    n_owned: int
    n_ghost: int
    send_idx: torch.Tensor          # [n_send] owned-row indices, grouped by destination rank
    send_shift: torch.Tensor        # [n_send, 3] lattice shift the sender adds to POSITIONS
    send_splits: List[int]          # rows per destination rank
    recv_splits: List[int]          # rows per source rank
    segments: List[slice]           # contiguous pieces of send_idx with unique indices (deterministic add)
    group: Optional[object] = None

    @property
    def n_local(self) -> int:
        return self.n_owned + self.n_ghost


def _all_to_all(out: torch.Tensor, inp: torch.Tensor, out_splits, in_splits, group):
    dist.all_to_all_single(out, inp, output_split_sizes=list(out_splits), input_split_sizes=list(in_splits), group=group)


class _HaloGather(torch.autograd.Function):
    """[n_owned, F] -> [n_owned + n_ghost, F]: ghost rows = rows of the ranks that own them."""

    @staticmethod
    def forward(ctx, t_owned, plan: HaloPlan, shifted: bool):
        ctx.plan = plan
        t_owned = t_owned.contiguous()
        send = t_owned.index_select(0, plan.send_idx)
        if shifted:
            send = send + plan.send_shift.to(send.dtype)
        recv = t_owned.new_empty((plan.n_ghost,) + tuple(t_owned.shape[1:]))
        _all_to_all(recv, send.contiguous(), plan.recv_splits, plan.send_splits, plan.group)
        return torch.cat([t_owned, recv], dim=0)

    @staticmethod
    def backward(ctx, g_local):
        return _HaloScatterAdd.apply(g_local, ctx.plan), None, None


class _HaloScatterAdd(torch.autograd.Function):
    """Adjoint of _HaloGather: ghost-row values travel back and are added onto the owners' rows."""

    @staticmethod
    def forward(ctx, g_local, plan: HaloPlan):
        ctx.plan = plan
        g_local = g_local.contiguous()
        g_ghost = g_local[plan.n_owned:]
        back = g_local.new_empty((plan.send_idx.numel(),) + tuple(g_local.shape[1:]))
        _all_to_all(back, g_ghost.contiguous(), plan.send_splits, plan.recv_splits, plan.group)
        out = g_local[: plan.n_owned].clone()
        for seg in plan.segments:  # indices are unique inside a segment -> order of the adds is fixed
            if seg.stop > seg.start:
                out.index_add_(0, plan.send_idx[seg], back[seg])
        return out

    @staticmethod
    def backward(ctx, g_owned):
        return _HaloGather.apply(g_owned, ctx.plan, False), None


def halo_gather(t_owned: torch.Tensor, plan: HaloPlan, shifted: bool = False) -> torch.Tensor:
    return _HaloGather.apply(t_owned, plan, shifted)


def fractional(pos: torch.Tensor, cell: torch.Tensor) -> torch.Tensor:
    """pos = f @ cell (rows of `cell` are the lattice vectors, data/radius_graph.py:6-32)."""
    return pos @ torch.linalg.inv(cell.to(pos.dtype))


def wrap_into_cell(pos: torch.Tensor, cell: torch.Tensor) -> torch.Tensor:
    f = fractional(pos.detach(), cell)
    return pos - torch.floor(f) @ cell.to(pos.dtype)  # constant lattice shifts: gradients pass through


def perpendicular_width(cell: torch.Tensor, axis: int = 0) -> float:
    c = cell.detach().double().cpu()
    vol = abs(float(torch.det(c)))
    n = torch.linalg.cross(c[(axis + 1) % 3], c[(axis + 2) % 3])
    return vol / float(torch.linalg.norm(n))


def max_slabs(cell: torch.Tensor, cutoff: float) -> int:
    return max(1, int(perpendicular_width(cell, 0) / cutoff))


def owner_of(pos_wrapped: torch.Tensor, cell: torch.Tensor, world: int) -> torch.Tensor:
    f0 = fractional(pos_wrapped, cell)[:, 0]
    return torch.clamp((f0 * world).floor().long(), 0, world - 1)


def plan_slabs(pos_owned: torch.Tensor, cell: torch.Tensor, cutoff: float, rank: int, world: int, group=None) -> HaloPlan:
    """Send / receive lists of this step (positions wrapped into the cell; all ranks call this together).
    One small all-to-all of row counts (host-synchronised, like any neighbour-list rebuild).  `cutoff` may include a
    Verlet skin: the plan then stays valid while no atom has moved more than skin / 2 (ShardedStep)."""
    if world > max_slabs(cell, cutoff):
        raise ValueError(f"{world} slabs are thinner than the cutoff: at most {max_slabs(cell, cutoff)} ranks for this cell")
    dev = pos_owned.device
    n_owned = pos_owned.shape[0]
    cell = cell.reshape(3, 3)
    f0 = fractional(pos_owned.detach(), cell)[:, 0]
    margin = cutoff / perpendicular_width(cell, 0) * (1.0 + 1e-5)
    lo, hi = rank / world, (rank + 1) / world
    left = torch.nonzero(f0 < lo + margin).flatten()     # -> rank-1 ; it sees them beyond its upper face
    right = torch.nonzero(f0 >= hi - margin).flatten()   # -> rank+1 ; it sees them below its lower face
    a0 = cell[0].to(pos_owned.dtype)
    left_shift = a0 if rank == 0 else torch.zeros_like(a0)
    right_shift = -a0 if rank == world - 1 else torch.zeros_like(a0)
    dst_left, dst_right = (rank - 1) % world, (rank + 1) % world
    per_dst: Dict[int, list] = {}
    per_dst.setdefault(dst_left, []).append((left, left_shift))
    per_dst.setdefault(dst_right, []).append((right, right_shift))
    idx_parts, shift_parts, splits, segments, off = [], [], [0] * world, [], 0
    for dst in sorted(per_dst):
        for idx, sh in per_dst[dst]:
            idx_parts.append(idx)
            shift_parts.append(sh.unsqueeze(0).expand(idx.numel(), 3))
            splits[dst] += int(idx.numel())
            segments.append(slice(off, off + int(idx.numel())))
            off += int(idx.numel())
    send_idx = torch.cat(idx_parts) if idx_parts else torch.zeros(0, dtype=torch.long, device=dev)
    send_shift = torch.cat(shift_parts) if shift_parts else torch.zeros((0, 3), dtype=pos_owned.dtype, device=dev)
    counts = torch.tensor(splits, dtype=torch.long, device=dev)
    recv_counts = torch.empty_like(counts)
    dist.all_to_all_single(recv_counts, counts, group=group)
    recv_splits = [int(v) for v in recv_counts.tolist()]
    return HaloPlan(rank, world, n_owned, sum(recv_splits), send_idx, send_shift.contiguous(), splits, recv_splits, segments, group)


# ------------------------------------------------------------------------------------------
# model driver
# ------------------------------------------------------------------------------------------
def shard_atoms(data: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    """Owned atoms of `rank` out of a replicated single-graph periodic structure (synthetic benches and
    tests; an MD driver would keep atoms resident on their ranks and migrate them)."""
    cell = data[keys.CELL].reshape(3, 3)
    pos = wrap_into_cell(data[keys.POSITIONS], cell)
    mine = torch.nonzero(owner_of(pos, cell, world) == rank).flatten()
    out = {keys.POSITIONS: pos[mine].contiguous(), keys.ATOMIC_NUMBERS: data[keys.ATOMIC_NUMBERS][mine].contiguous(),
           keys.CELL: data[keys.CELL], "global_index": mine}
    return out


def local_graph(pos_local: torch.Tensor, cell: torch.Tensor, cutoff: float, n_owned: int, world: int = 0):
    """K1 on [owned | ghosts], rows of ghost centers dropped.  Axis 0 is open for the local problem; it is
    presented to K1 as a periodic axis of a STRETCHED lattice vector a0' = t * a0 whose images are further
    than the cutoff apart (t = slab + both margins + one more margin), so that the all-periodic cell-list
    path applies: every edge then has a zero offset along axis 0 and the lattice vectors that the edge
    kernel multiplies by non-zero offsets (a1, a2) are the true ones."""
    from .graph import NeighborGraph, build_graph

    dev = pos_local.device
    N = pos_local.shape[0]
    ptr = torch.tensor([0, N], dtype=torch.int32, device=dev)
    cell = cell.reshape(3, 3)
    if world > 0:
        margin = cutoff / perpendicular_width(cell, 0)
        t = 1.0 / world + 3.05 * margin
        cell_k1 = torch.stack([cell[0] * t, cell[1], cell[2]]).reshape(1, 3, 3)
        pbc = [True, True, True]
    else:
        cell_k1, pbc = cell.reshape(1, 3, 3), [False, True, True]
    g, _, _ = build_graph(pos_local, cutoff, ptr=ptr, cell=cell_k1, pbc=pbc)
    rowptr = g.rowptr.clone()
    e_owned = int(rowptr[n_owned].item())
    rowptr[n_owned:] = e_owned
    return NeighborGraph(N, 1, rowptr, g.col[:e_owned], g.offsets[:e_owned] if g.offsets is not None else None, g.cell, None,
                         n_centers=n_owned)


def energy_forces_sharded(model, owned: Dict[str, torch.Tensor], rank: int, world: int, group=None, compute_forces: bool = True):
    """E+F of one periodic structure sharded over `world` ranks.  `owned` holds this rank's atoms
    (positions wrapped into the cell).  Returns {"energy": partial sum over owned atoms [1],
    "atomic_energies" [n_owned], "forces" [n_owned, 3] = -dE_total/dpos_owned}."""
    cutoff = float(model.cutoff_radius)
    cell = owned[keys.CELL].reshape(3, 3)
    pos_owned = owned[keys.POSITIONS]
    if compute_forces:
        pos_owned = pos_owned.detach().requires_grad_()
    n_owned = pos_owned.shape[0]
    plan = plan_slabs(pos_owned, cell, cutoff, rank, world, group)
    pos_local = halo_gather(pos_owned, plan, shifted=True)
    graph = local_graph(pos_local.detach(), cell, cutoff, n_owned, world)
    dev = pos_owned.device
    data = {
        keys.POSITIONS: pos_local, keys.ATOMIC_NUMBERS: owned[keys.ATOMIC_NUMBERS], keys.CELL: owned[keys.CELL].reshape(1, 3, 3),
        keys.BATCH: torch.zeros(n_owned, dtype=torch.long, device=dev),
        keys.BATCH_PTR: torch.tensor([0, n_owned], dtype=torch.long, device=dev),
        keys.GRAPH: graph, keys.HALO: plan,
    }
    out = model(data, compute_forces=False)
    res = {keys.TOTAL_ENERGY: out[keys.TOTAL_ENERGY], keys.ATOMIC_ENERGIES: out[keys.ATOMIC_ENERGIES]}
    if compute_forces:
        (g,) = torch.autograd.grad([out[keys.TOTAL_ENERGY].sum()], [pos_owned], retain_graph=model.training,
                                   create_graph=model.training)
        res[keys.FORCES] = -g
    return res


class ShardedStep:
    """The sharded E+F step as ONE CUDA graph per rank (strong scaling of an MD-style loop on a fixed cell).

    The eager path above re-plans the halo and sizes the neighbour list on the host every step; at ~1-5 k atoms per
    rank that host work and ~600 eager launches cost more than the kernels.  Here the halo plan (who sends which rows
    to whom) is built once with `cutoff + skin` margins and reused while atoms have moved less than skin / 2 -- ghost
    atoms beyond the cutoff only add edges whose filter value is exactly zero -- K1 runs in capacity mode on
    [owned | ghost] positions, ghost rows are dropped on the device, and the whole step (position halo -> K1 -> model
    with one s|v halo exchange per layer -> forces with the reverse exchanges -> energy all-reduce) is captured once.
    Replays need no host synchronisation; NCCL all-to-alls are graph nodes.

        step = ShardedStep(model, owned, rank, world, skin=0.5)      # all ranks together
        out = step(pos_owned)        # {"energy": total energy of the box [1], "forces": [n_owned, 3]} (static tensors)
    """

    def __init__(self, model, owned: Dict[str, torch.Tensor], rank: int, world: int, skin: float = 0.5, group=None,
                 capacity_margin: float = 1.2, warmup: int = 2):
        from .graph import NeighborGraph, StaticGraphBuilder, build_graph

        self.model, self.rank, self.world, self.group = model, rank, world, group
        cutoff = float(model.cutoff_radius)
        cell = owned[keys.CELL].reshape(3, 3)
        self.pos = owned[keys.POSITIONS].detach().clone().requires_grad_()
        dev = self.pos.device
        self.n_owned = n_owned = self.pos.shape[0]
        self.plan = plan = plan_slabs(self.pos, cell, cutoff + skin, rank, world, group)
        self.pos_ref = self.pos.detach().clone()
        self.skin = float(skin)
        n_local = plan.n_local
        # K1 on [owned | ghosts]: axis 0 presented as a stretched periodic axis (see local_graph)
        margin = (cutoff + skin) / perpendicular_width(cell, 0)
        t = 1.0 / world + 3.05 * margin
        cell_k1 = torch.stack([cell[0] * t, cell[1], cell[2]]).reshape(1, 3, 3)
        ptr = torch.tensor([0, n_local], dtype=torch.int32, device=dev)
        with torch.no_grad():
            pos_local = halo_gather(self.pos.detach(), plan, shifted=True)
        g0, _, _ = build_graph(pos_local, cutoff, ptr=ptr, cell=cell_k1, pbc=[True, True, True])
        # rows of ghost centers are emptied on the device after every build: only the first n_owned nodes are centers
        self.builder = StaticGraphBuilder(n_local, ptr, cutoff, int(g0.n_edges * capacity_margin) + 1024, cell=cell_k1,
                                          pbc=[True, True, True], n_centers=n_owned)
        self.z = owned[keys.ATOMIC_NUMBERS]
        self.cell1 = owned[keys.CELL].reshape(1, 3, 3)
        self.batch = torch.zeros(n_owned, dtype=torch.long, device=dev)
        self.bptr = torch.tensor([0, n_owned], dtype=torch.long, device=dev)
        self.out: Dict[str, torch.Tensor] = {}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self._body()
        torch.cuda.synchronize()

    def _body(self) -> Dict[str, torch.Tensor]:
        self.pos.grad = None
        plan, n_owned = self.plan, self.n_owned
        pos_local = halo_gather(self.pos, plan, shifted=True)
        g = self.builder.build(pos_local.detach().contiguous(), check_overflow=False)
        with torch.no_grad():  # drop the rows of ghost centers: no redundant edge work, owners only
            g.rowptr[n_owned:] = g.rowptr[n_owned]
        g.transpose()
        data = {keys.POSITIONS: pos_local, keys.ATOMIC_NUMBERS: self.z, keys.CELL: self.cell1, keys.BATCH: self.batch,
                keys.BATCH_PTR: self.bptr, keys.GRAPH: g, keys.HALO: plan}
        out = self.model(data, compute_forces=False)
        e = out[keys.TOTAL_ENERGY]
        (grad,) = torch.autograd.grad([e.sum()], [self.pos])
        e_tot = e.detach().sum().reshape(1).clone()
        if self.world > 1:
            dist.all_reduce(e_tot, group=self.group)
        return {keys.TOTAL_ENERGY: e_tot, keys.FORCES: -grad}

    def __call__(self, pos_owned: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        if pos_owned is not None:
            with torch.no_grad():
                self.pos.copy_(pos_owned, non_blocking=True)
        self.graph.replay()
        return self.out

    def needs_replan(self, pos_owned: torch.Tensor) -> bool:
        """Host synchronisation: True when some atom has moved more than skin / 2 since the plan was built (the caller
        then builds a new ShardedStep; all ranks must take the same decision -- all-reduce the flag)."""
        d2 = ((pos_owned.detach() - self.pos_ref) ** 2).sum(-1).max()
        flag = (d2 > (0.5 * self.skin) ** 2).to(torch.int32).reshape(1)
        if self.world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
        return bool(flag.item())

    def check(self) -> None:
        if int(self.builder.overflow.item()) != 0:
            raise RuntimeError("ShardedStep: edge capacity exceeded")
