"""Molecule-batch data parallelism (xequinet_b200/parallel.py) on CPU: the sharding is a balanced partition, and
the flat gradient all-reduce over gloo (world_size 2) reproduces the single-process gradient of the CPU oracle
on the union of the shards."""
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def test_shard_molecules_is_a_balanced_partition():
    from xequinet_b200 import parallel

    g = torch.Generator().manual_seed(3)
    sizes = torch.randint(30, 71, (128,), generator=g).tolist()  # c4-shaped: 30..70 atoms
    for world in (1, 2, 4, 8):
        shards = parallel.shard_molecules(sizes, world, cutoff_neighbors=40)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(len(sizes)))
        loads = [sum(parallel.estimate_edges(sizes[i], 40) + sizes[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(parallel.estimate_edges(n, 40) + n for n in sizes)
        assert shards == parallel.shard_molecules(sizes, world, cutoff_neighbors=40)  # deterministic
    assert parallel.shard_molecules([], 4) == [[], [], [], []]
    with pytest.raises(ValueError):
        parallel.shard_molecules(sizes, 0)


def _loss_and_grads(sd, table, batch, cfg, orc):
    import torch.nn.functional as F

    d = dict(batch)
    d["edge_index"] = orc.radius_graph(d["pos"], cfg.cutoff, d["batch"])
    out = orc.xpainn_energy_forces(sd, table, d, cfg, create_graph=True)
    # sums (not means): the global loss is the sum of the per-rank losses
    loss = F.smooth_l1_loss(out["energy"], d["target_energy"], reduction="sum") + \
        F.smooth_l1_loss(out["forces"], d["target_forces"], reduction="sum")
    names = sorted(sd)
    grads = torch.autograd.grad(loss, [sd[k] for k in names], allow_unused=True)
    return names, [g if g is not None else torch.zeros_like(sd[k]) for g, k in zip(grads, names)]


def _make(n_mol, seed):
    import numpy as np
    from oracle import xpainn_oracle as orc

    cfg = orc.CONFIG_DEFAULT
    table = torch.from_numpy(np.load(ROOT / "xequinet_b200" / "data" / "gfn2-xtb_aux56.npy")).double()
    sd = {k: v.double().requires_grad_(True) for k, v in orc.synthetic_state_dict(cfg, 1234).items() if v.is_floating_point()}
    sd_all = dict(orc.synthetic_state_dict(cfg, 1234))
    sd_all.update(sd)
    b = orc.make_molecule_batch(n_mol, (3, 9), seed=seed, with_edges=False)
    b = {k: (v.double() if v.is_floating_point() else v) for k, v in b.items()}
    g = torch.Generator().manual_seed(seed + 50)
    b["target_energy"] = torch.randn(n_mol, generator=g, dtype=torch.float64)
    b["target_forces"] = torch.randn(b["pos"].shape[0], 3, generator=g, dtype=torch.float64)
    return cfg, table, sd_all, b, orc


def _select(batch, mols):
    """sub-batch of the given molecule indices"""
    ptr = batch["ptr"]
    idx = torch.cat([torch.arange(int(ptr[m]), int(ptr[m + 1])) for m in mols]) if mols else torch.zeros(0, dtype=torch.long)
    n = torch.tensor([int(ptr[m + 1] - ptr[m]) for m in mols], dtype=torch.long)
    out = {"pos": batch["pos"][idx], "atomic_numbers": batch["atomic_numbers"][idx],
           "batch": torch.repeat_interleave(torch.arange(len(mols)), n),
           "ptr": torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(n, 0)]),
           "target_energy": batch["target_energy"][torch.tensor(mols, dtype=torch.long)],
           "target_forces": batch["target_forces"][idx]}
    return out


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from xequinet_b200 import parallel

        torch.set_num_threads(2)
        cfg, table, sd, batch, orc = _make(6, seed=21)
        sizes = torch.diff(batch["ptr"]).tolist()
        mine = parallel.shard_molecules(sizes, world)[rank]
        names, grads = _loss_and_grads(sd, table, _select(batch, mine), cfg, orc)
        params = [torch.nn.Parameter(sd[k].detach().clone()) for k in names]
        for p, g in zip(params, grads):
            p.grad = g.clone()
        parallel.allreduce_gradients(params, average=False)
        q.put((rank, {k: p.grad.detach().numpy().copy() for k, p in zip(names, params)}))  # by value (no fd sharing)
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_matches_single_process():
    world, port = 2, 29641
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg, table, sd, batch, orc = _make(6, seed=21)
    names, ref = _loss_and_grads(sd, table, batch, cfg, orc)
    for r in range(world):
        for k, g in zip(names, ref):
            got = torch.from_numpy(results[r][k])
            assert torch.allclose(got, g, rtol=1e-9, atol=1e-11), (r, k)


def _flat_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from xequinet_b200 import parallel

        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.SiLU(), torch.nn.Linear(7, 3), torch.nn.Linear(3, 1)).double()
        unused = torch.nn.Parameter(torch.ones(4, dtype=torch.float64))  # never receives a gradient
        params = list(net.parameters()) + [unused]
        flat = parallel.FlatGradients(params, n_buckets=3)
        assert 2 <= len(flat.buckets) <= 3 and sum(len(b) for b in flat.buckets) == len(params)
        out = {}
        for step in range(2):  # two steps: zero() must reset the views and the bucket bookkeeping
            x = torch.randn(6, 5, dtype=torch.float64, generator=torch.Generator().manual_seed(10 * step + rank))
            flat.zero()
            assert all(p.grad is None for p in params)
            net(x).pow(2).sum().backward()
            flat.finish()
            assert all(p.grad.data_ptr() == flat.flat.data_ptr() + 8 * flat._range[id(p)][0] for p in params)  # views of the flat buffer
            out[step] = [p.grad.detach().clone().numpy() for p in params]
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_flat_gradients_bucketed_allreduce_gloo():
    world, port = 2, 29643
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_flat_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.SiLU(), torch.nn.Linear(7, 3), torch.nn.Linear(3, 1)).double()
    for step in range(2):
        ref = None
        for rank in range(world):
            x = torch.randn(6, 5, dtype=torch.float64, generator=torch.Generator().manual_seed(10 * step + rank))
            net.zero_grad(set_to_none=True)
            net(x).pow(2).sum().backward()
            g = [p.grad.clone() for p in net.parameters()]
            ref = g if ref is None else [a + b for a, b in zip(ref, g)]
        ref = [a / world for a in ref] + [torch.zeros(4, dtype=torch.float64)]
        for rank in range(world):
            for got, want in zip(results[rank][step], ref):
                assert torch.allclose(torch.from_numpy(got), want, rtol=1e-12, atol=1e-14)
