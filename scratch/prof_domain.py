"""torchrun --nproc-per-node P scratch/prof_domain.py : phase timings of the sharded c5 step (rank 0 prints)."""
import os, sys, time; sys.path.insert(0, ".")
import torch, torch.distributed as dist
import xequinet_b200 as xb
from xequinet_b200 import domain, keys, ops
from oracle import xpainn_oracle as orc
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"])); dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
cfg = orc.CONFIG_DEFAULT
model = xb.resolve_model("xpainn", **cfg.model_kwargs()); model.load_state_dict(orc.synthetic_state_dict(cfg, 1234), strict=False)
model = model.to(dev).eval()
for p in model.parameters(): p.requires_grad_(False)
box = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in orc.make_water_box(15, seed=0).items()}
owned = domain.shard_atoms(box, rank, world)
cutoff = 5.0; cell = owned["cell"].reshape(3, 3)
def phases():
    T = {}
    def mark(name, t0):
        torch.cuda.synchronize(); T[name] = T.get(name, 0) + (time.perf_counter() - t0) * 1e3
    t = time.perf_counter(); pos_owned = owned["pos"].detach().requires_grad_()
    plan = domain.plan_slabs(pos_owned, cell, cutoff, rank, world); mark("plan", t)
    t = time.perf_counter(); pos_local = domain.halo_gather(pos_owned, plan, shifted=True); mark("halo_pos", t)
    t = time.perf_counter(); graph = domain.local_graph(pos_local.detach(), cell, cutoff, plan.n_owned, world); mark("local_graph", t)
    n_owned = plan.n_owned
    data = {keys.POSITIONS: pos_local, keys.ATOMIC_NUMBERS: owned["atomic_numbers"], keys.CELL: owned["cell"].reshape(1, 3, 3),
            keys.BATCH: torch.zeros(n_owned, dtype=torch.long, device=dev), keys.BATCH_PTR: torch.tensor([0, n_owned], dtype=torch.long, device=dev),
            keys.GRAPH: graph, keys.HALO: plan}
    from xequinet_b200.nn.basic import compute_edge_data
    t = time.perf_counter(); data = compute_edge_data(data, False, False); mark("fwd.edge_data", t)
    for name, mod in model.mods.items():
        t = time.perf_counter(); data = mod(data); mark("fwd." + name, t)
    out = {"energy": data["energy"]}
    t = time.perf_counter()
    if True:
        (g,) = torch.autograd.grad([out["energy"].sum()], [pos_owned])
    mark("backward", t)
    return T, plan, graph
for _ in range(3): phases()
acc = {}
for _ in range(5):
    T, plan, graph = phases()
    for k, v in T.items(): acc[k] = acc.get(k, 0) + v / 5
if rank == 0:
    print(f"P={world} owned {plan.n_owned} ghosts {plan.n_ghost} edges {graph.n_edges}: " + "  ".join(f"{k} {v:.2f} ms" for k, v in acc.items()), flush=True)
torch.cuda.synchronize(); sys.stdout.flush(); os._exit(0)
