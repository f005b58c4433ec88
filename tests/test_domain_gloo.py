"""Spatial sharding host logic (xequinet_b200/domain.py) on CPU with the gloo backend, world_size 2 and 3:
 * the union of the per-rank edge sets (owned centers over owned + ghost atoms, axis 0 open, ghost positions
   carrying their lattice shift) equals the reference's global periodic radius graph (oracle restatement
   of data/radius_graph.py:35-192) edge for edge, compared through the edge displacement vectors;
 * halo gather / scatter-add are adjoint (forward, backward and double backward through the collectives)."""
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _edges_as_keys(center_gid, nbr_gid, disp):
    """one int64 key per edge: (center, neighbor, displacement rounded to 1e-3 A)"""
    q = torch.round(disp.double() * 1000).long() + 100_000
    return (((center_gid.long() * 20_000 + nbr_gid.long()) * 200_001 + q[:, 0]) * 200_001 + q[:, 1]) * 200_001 + q[:, 2]


def _worker(rank: int, world: int, port: int, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import xpainn_oracle as orc
        from xequinet_b200 import domain, keys

        cutoff = 5.0
        box = orc.make_water_box(6, seed=5, dtype=torch.float64)  # 648 atoms, L = 18.6 A (3 slabs of 6.2 A)
        cell = box["cell"].reshape(3, 3)
        if world == 3:  # triclinic variant
            cell = cell.clone()
            cell[1, 0] = 0.2 * cell[0, 0]
        box["cell"] = cell.reshape(1, 3, 3)
        owned = domain.shard_atoms(box, rank, world)
        pos_owned = owned[keys.POSITIONS]
        plan = domain.plan_slabs(pos_owned, cell, cutoff, rank, world)
        pos_local = domain.halo_gather(pos_owned, plan, shifted=True)
        gid_local = domain.halo_gather(owned["global_index"].double().unsqueeze(1), plan).squeeze(1).long()
        # local graph with the oracle: axis 0 open, owned centers only
        n_loc = torch.tensor([pos_local.shape[0]])
        ei, co = orc.radius_graph_pbc(pos_local, n_loc, torch.tensor([[False, True, True]]), cell.reshape(1, 3, 3), cutoff)
        keep = ei[0] < plan.n_owned
        ei, co = ei[:, keep], co[keep]
        disp = pos_local[ei[0]] - pos_local[ei[1]] - co.double() @ cell
        keys_local = _edges_as_keys(gid_local[ei[0]], gid_local[ei[1]], disp)
        gathered = [None] * world
        dist.all_gather_object(gathered, keys_local)
        # adjointness: <gather(x), y> == <x, scatter_add(y)>, also through autograd (backward + double backward)
        g = torch.Generator().manual_seed(7 + rank)
        x = torch.randn(plan.n_owned, 5, generator=g, dtype=torch.float64, requires_grad=True)
        y = torch.randn(plan.n_local, 5, generator=g, dtype=torch.float64, requires_grad=True)
        lhs = (domain.halo_gather(x, plan) * y.detach()).sum()
        rhs = (x.detach() * domain._HaloScatterAdd.apply(y, plan)).sum()
        tot = torch.stack([lhs.detach(), rhs.detach()])
        dist.all_reduce(tot)
        out = domain.halo_gather(x, plan)
        (gx,) = torch.autograd.grad((out ** 2).sum(), x, create_graph=True)   # backward = scatter-add collective
        (ggx,) = torch.autograd.grad((gx ** 3).sum(), x)                      # double backward = gather collective
        # reference for the derivative checks: multiplicity m_i = 1 + number of ranks ghosting atom i
        mult = torch.ones(plan.n_owned, dtype=torch.float64)
        mult.index_add_(0, plan.send_idx, torch.ones(plan.send_idx.numel(), dtype=torch.float64))
        gx_ref = 2 * x.detach() * mult.unsqueeze(1)
        ggx_ref = 3 * gx_ref ** 2 * 2 * mult.unsqueeze(1)
        ok_grad = bool(torch.allclose(gx.detach(), gx_ref) and torch.allclose(ggx, ggx_ref))
        if rank == 0:
            ei_g, co_g = orc.radius_graph_pbc(box["pos"], torch.tensor([box["pos"].shape[0]]), box["pbc"], box["cell"], cutoff)
            disp_g = box["pos"][ei_g[0]] - box["pos"][ei_g[1]] - co_g.double() @ cell
            ref = torch.sort(_edges_as_keys(ei_g[0], ei_g[1], disp_g))[0]
            got = torch.sort(torch.cat(gathered))[0]
            q.put({"n_ref": int(ref.numel()), "n_got": int(got.numel()), "equal": bool(ref.numel() == got.numel() and torch.equal(ref, got)),
                   "adjoint": float((tot[0] - tot[1]).abs() / tot[0].abs()), "grad": ok_grad, "ghosts": plan.n_ghost})
        else:
            assert ok_grad
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_decomposition_matches_global_graph(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["ghosts"] > 0
    assert res["n_ref"] == res["n_got"] and res["equal"], res
    assert res["adjoint"] < 1e-12 and res["grad"], res
