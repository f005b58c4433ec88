#!/usr/bin/env python
"""bench.py -- molecules/s for XPaiNN energy+forces on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4|c5] [--impl reference]

One "step" is one pass of the hot path over one batch of synthetic molecules:
  c3 (default; the configuration the metric is quoted on at 1/2/4/8 GPUs): aspirin-shaped
     (21 atoms) energy+force TRAINING step -- neighbour list (K1), forward, forces with
     create_graph (K2/K2b), smooth-L1 E+F loss, backward through the forces (K2bb), gradient
     all-reduce (N > 1), AdamW -- on 256 molecules per GPU (weak scaling).
  c1: E+F inference, 64 x 18 atoms.   c2: energy-only training, 256 x 18 atoms.
  c4: E+F inference, 256 channels, 128 drug-like molecules (30..70 atoms) per GPU.
  c5: periodic water box (~10k atoms), E+F inference incl. neighbour rebuild; N > 1 = spatial slabs with
      halo exchange (xequinet_b200/domain.py), strong scaling of one box; value counts water molecules.

Output: ONE JSON line (rank 0).  `value` = whole-job molecules/s with the inputs resident in
HBM; `e2e` = the same through the public API with pinned-host inputs copied in and the loss /
energies read back every step; `roofline` = the dominant hand-written kernel of the step
(CUDA events on the launching stream, algorithmic bytes per DESIGN.md) against the measured
HBM peak; `roofline_throughput` = the fused edge forward kernel on a working set >> L2 (8192 molecules);
`cpu_baseline` = the CPU oracle (port of the reference path) on the host cores.
`--impl reference` times that CPU path alone.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from oracle import xpainn_oracle as orc  # noqa: E402  (cpu_baseline / reference arm only)

METRIC = "molecules/sec energy+forces"
UNIT = "molecules/s"

WORKLOADS = {
    "c1": dict(desc="XPaiNN E+F inference, QM9-shaped 64 x 18 atoms", cfg=orc.CONFIG_DEFAULT, n_mol=64, train=False, forces=True),
    "c2": dict(desc="XPaiNN energy-only training, QM9-shaped 256 x 18 atoms, fp32", cfg=orc.CONFIG_DEFAULT, n_mol=256, train=True, forces=False),
    "c3": dict(desc="XPaiNN E+F training (double-backward force loss), MD17/aspirin-shaped 256 x 21 atoms per GPU", cfg=orc.CONFIG_DEFAULT, n_mol=256, train=True, forces=True),
    "c4": dict(desc="XPaiNN-256ch E+F inference, SPICE-shaped 128 x (30..70) atoms per GPU", cfg=orc.CONFIG_C4, n_mol=128, train=False, forces=True),
    "c5": dict(desc="periodic water box 10125 atoms, PBC radius graph + E+F inference", cfg=orc.CONFIG_DEFAULT, n_mol=1, train=False, forces=True),
}


def make_batch(workload: str, n_mol: int, seed: int):
    if workload in ("c1", "c2"):
        d = orc.make_molecule_batch(n_mol, 18, seed=seed, with_edges=False)
    elif workload == "c3":
        d = orc.make_aspirin_batch(n_mol, seed=seed, with_edges=False)
    elif workload == "c4":
        d = orc.make_molecule_batch(n_mol, (30, 70), seed=seed, z_table=orc._Z_SPICE, with_edges=False)
    else:
        d = orc.make_water_box(15, seed=seed)
    g = torch.Generator().manual_seed(1000 + seed)
    G = d["ptr"].numel() - 1
    d["target_energy"] = torch.randn(G, generator=g)
    d["target_forces"] = torch.randn(d["pos"].shape[0], 3, generator=g)
    return d


def loss_fn(out, batch, forces: bool):
    loss = F.smooth_l1_loss(out["energy"], batch["target_energy"])
    if forces:
        loss = loss + 100.0 * F.smooth_l1_loss(out["forces"], batch["target_forces"])
    return loss


# ----------------------------------------------------------------------------------------------
# CPU oracle arm (cpu_baseline and --impl reference)
# ----------------------------------------------------------------------------------------------
def cpu_step_fn(workload: str, n_mol: int):
    """Returns (step callable, molecules per step, description) for the CPU oracle."""
    w = WORKLOADS[workload]
    cfg = w["cfg"]
    torch.set_num_threads(os.cpu_count() or 1)
    table = torch.from_numpy(np.load(ROOT / "xequinet_b200" / "data" / "gfn2-xtb_aux56.npy")).float()
    sd = {k: v.requires_grad_(w["train"]) for k, v in orc.synthetic_state_dict(cfg, 1234).items()}
    if workload == "c5":
        batch = orc.make_water_box(6, seed=0)  # 648 atoms: bounded sample of the box
        n = torch.tensor([batch["pos"].shape[0]])
    else:
        batch = make_batch(workload, n_mol, 0)
    opt = torch.optim.AdamW(list(sd.values()), lr=5e-4) if w["train"] else None

    def step():
        d = dict(batch)
        if workload == "c5":
            d["edge_index"], d["cell_offsets"] = orc.radius_graph_pbc(d["pos"], n, d["pbc"], d["cell"], cfg.cutoff)
        else:
            d["edge_index"] = orc.radius_graph(d["pos"], cfg.cutoff, d["batch"])
        if w["forces"]:
            out = orc.xpainn_energy_forces(sd, table, d, cfg, create_graph=w["train"])
        else:
            e, ea = orc.xpainn_energy(sd, table, d, cfg)
            out = {"energy": e}
        if w["train"]:
            opt.zero_grad(set_to_none=True)
            loss_fn(out, d, w["forces"]).backward()
            opt.step()
        return out

    mols = batch["ptr"].numel() - 1 if workload != "c5" else batch["pos"].shape[0] // 3  # c5 counts water molecules
    return step, float(mols), f"{mols} molecule(s), {batch['pos'].shape[0]} atoms per step"


def time_torch_gpu(workload: str, n_mol: int, dev, steps: int = 5, warmup: int = 2):
    """SURVEY.md 8d "reference-on-GPU" line: the SAME oracle (the reference's algorithm as stock PyTorch ops: index_select /
    index_add, E-sized intermediates, autograd for forces and the double backward) on the B200, full workload size, edge
    list precomputed and not timed.  A reported baseline next to cpu_baseline -- what running the reference's code path
    on this GPU through PyTorch costs; rank 0, single GPU only."""
    import xequinet_b200 as xb

    w = WORKLOADS[workload]
    cfg = w["cfg"]
    table = torch.from_numpy(np.load(ROOT / "xequinet_b200" / "data" / "gfn2-xtb_aux56.npy")).float().to(dev)
    sd = {k: v.to(dev).requires_grad_(w["train"]) for k, v in orc.synthetic_state_dict(cfg, 1234).items()}
    batch = make_batch(workload, n_mol, 0)
    d0 = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    # the edge list comes from K1 (bit-exact against the oracle's search, tests/test_gpu_parity.py) and is not timed
    if workload == "c5":
        n = torch.tensor([batch["pos"].shape[0]], device=dev)
        d0["edge_index"], d0["cell_offsets"] = xb.radius_graph_pbc(d0["pos"], n, d0["pbc"], d0["cell"], cfg.cutoff)
    else:
        d0["edge_index"] = xb.radius_graph(d0["pos"], cfg.cutoff, batch=d0["batch"])
    batch["edge_index"] = d0["edge_index"]
    opt = torch.optim.AdamW(list(sd.values()), lr=5e-4, fused=True) if w["train"] else None

    def step():
        d = dict(d0)
        if w["forces"]:
            out = orc.xpainn_energy_forces(sd, table, d, cfg, create_graph=w["train"])
        else:
            out = {"energy": orc.xpainn_energy(sd, table, d, cfg)[0]}
        if w["train"]:
            opt.zero_grad(set_to_none=True)
            loss_fn(out, d, w["forces"]).backward()
            opt.step()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    mols = batch["ptr"].numel() - 1 if workload != "c5" else batch["pos"].shape[0] // 3
    peak_gb = torch.cuda.max_memory_allocated() / 2**30
    return mols / (ms * 1e-3), ms, f"{mols} molecule(s), {batch['pos'].shape[0]} atoms, {batch['edge_index'].shape[1]} edges per step"


def time_cpu(workload: str, n_mol: int, steps: int, warmup: int):
    step, mols, sample = cpu_step_fn(workload, n_mol)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return mols / dt, dt * 1e3, sample


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.stop_flag = threading.Event()
        self.samples = []

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for s in self.samples:
            try:
                sm.append(float(s[1]))
                mx = max(mx, float(s[2]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if s[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        # the GPU idles between the first samples; "under load" = upper half of the samples
        sm.sort()
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# roofline bookkeeping (DESIGN.md "algorithmic bytes")
# ----------------------------------------------------------------------------------------------
def algorithmic_bytes(kind: str, cfg: orc.XPaiNNConfig, N: int, E: int, periodic: bool) -> float:
    C, D, H = cfg.node_dim, cfg.D, cfg.H_msg
    csr = 4 * (N + 1) + 4 * E + (4 * E if periodic else 0)
    if kind == "edge_fwd":  # read s, v once; read+write residual x, V; pos; CSR  
        return 4 * N * (H + D) + 8 * N * (C + D) + 12 * N + csr
    if kind in ("edge_bwd", "edge_bwd_wgrad"):  # read s, v, gx, gV; write gs, gv, gpos; pos; transposed CSR (+ eid)
        return 4 * N * (2 * (H + D) + (C + D) + 3) + 12 * N + csr + 8 * E
    if kind == "edge_bwdbwd":  # JVP pass + reverse pass: read s, v, a_s, a_v (twice), gx, gV; write o_gx, o_gV, o_s, o_v, o_pos
        return 4 * N * (4 * (H + D) + 2 * (C + D) + (H + D) + 3) + 36 * N + 2 * csr + 8 * E
    raise KeyError(kind)


# DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels behind each timed op, from the
# committed `ncu --set full` capture of the final round-2 kernels at the c3 shape (N = 5376, E = 106068):
# profiles/r02_edge_ncu_full.md.  Only valid for that shape; other workloads report traffic = null.  (Writes mostly
# stay in the 126 MB L2 at this size.)
NCU_TRAFFIC_C3 = {
    "edge_fwd": 36.41e6 + 0.03e6,                                                    # center_fwd_ul_kernel<.., 0> (rows packed in-kernel)
    "edge_bwd": 36.84e6 + 0.32e6,                                                    # nbr_bwd_ul_kernel
    "edge_bwd_wgrad": (36.84e6 + 0.32e6) + (36.77e6 + 0.03e6),                       # + wgrad_ul_kernel<1>
    # JVP = center_fwd_ul<.., 1> + <.., 2>; reverse pass = nbr_bwd2_ul<2> + <3>; wgrad_ul_kernel<2>
    "edge_bwdbwd": (40.64e6 + 0.02e6) + (36.48e6 + 0.0) + (59.65e6 + 2.32e6) + (38.19e6 + 0.04e6) + (59.54e6 + 1.17e6),
}
# the same for the throughput probe (8192 molecules): center_fwd_ul_kernel, profiles/r02_edge_ncu_full.md (final capture)
NCU_TRAFFIC_PROBE_FWD = 1.162e9 + 0.396e9


def edge_throughput_probe(cfg, dev, peak, n_mol=8192, reps=10):
    """The fused edge FORWARD kernel on a batch whose working set (s, v, x, V rows: 1.4 GB at 8192 aspirin-shaped
    molecules) is >> the 126 MB L2 -- the "roofline (throughput) mode" of SURVEY.md 8d: at the named batch sizes
    the working set is L2-resident and HBM bandwidth is not the binding resource.  CUDA events around `reps`
    launches of xeq_edge_message_fwd through the C ABI; algorithmic bytes as in algorithmic_bytes().  Returns a
    roofline-shaped dict, or None when anything goes wrong (this probe must never break the bench line)."""
    try:
        import xequinet_b200 as xb
        from xequinet_b200 import ops

        d = orc.make_aspirin_batch(n_mol, seed=0, with_edges=False)
        g, _, _ = xb.build_graph(d["pos"].to(dev), cfg.cutoff, ptr=d["ptr"].to(dev), batch=d["batch"].to(dev))
        N, E = g.n_nodes, g.n_edges
        dims = ops.Dims(cfg.node_dim, *cfg.muls, cfg.num_basis, cfg.cutoff)
        r = lambda *shape: torch.randn(*shape, device=dev)
        pos = d["pos"].to(dev)
        s, v, x, V = r(N, dims.H), r(N, dims.D), r(N, dims.node_dim), r(N, dims.D)
        W, b = 0.3 * r(dims.H, cfg.num_basis), 0.3 * r(dims.H)
        freq = (torch.pi * torch.arange(1, cfg.num_basis + 1, device=dev) / cfg.cutoff).float()
        for _ in range(3):
            ops.edge_message_fwd_raw(g, dims, pos, s, v, x, V, W, b, freq)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(reps):
            ops.edge_message_fwd_raw(g, dims, pos, s, v, x, V, W, b, freq)
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / reps
        achieved = algorithmic_bytes("edge_fwd", cfg, N, E, False) / (ms * 1e-3) / 1e9
        return {"kernel": "edge_fwd", "launch": "xeq_edge_message_fwd: center_fwd_ul_kernel (window rows packed in-kernel)", "bound": "hbm",
                "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 5),
                "traffic": NCU_TRAFFIC_PROBE_FWD if n_mol == 8192 else None,
                "n_nodes": N, "n_edges": E, "mean_launch_ms": round(ms, 4),
                "working_set": f"{n_mol} aspirin-shaped molecules: node rows {4 * N * (dims.H + 2 * dims.D + 2 * dims.node_dim) / 1e9:.2f} GB >> L2"}
    except Exception as exc:  # pragma: no cover
        return {"error": str(exc)[:200]}


def measured_hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch.distributed as dist

    import xequinet_b200 as xb
    from xequinet_b200 import _lib, ops, parallel

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w = WORKLOADS[args.workload]
    cfg, train, forces = w["cfg"], w["train"], w["forces"]
    n_mol = args.molecules or w["n_mol"]

    model = xb.resolve_model("xpainn", **cfg.model_kwargs())
    model.load_state_dict(orc.synthetic_state_dict(cfg, 1234), strict=False)  # random-init weights
    model = model.to(dev)
    model.train(train)
    params = [p for p in model.parameters()]
    if not train:
        for p in params:
            p.requires_grad_(False)
    opt = torch.optim.AdamW(params, lr=5e-4, fused=True) if train else None
    # N > 1: one flat gradient buffer (every p.grad is a view), bucketed all-reduce overlapped with the backward pass
    flat = parallel.FlatGradients(params, n_buckets=3) if (train and world > 1) else None
    transform = xb.NeighborTransform(cfg.cutoff)

    # distinct synthetic batches per rank, pinned on the host
    n_batches = 4 if args.workload != "c5" else 1
    host = []
    for b in range(n_batches):
        d = make_batch(args.workload, n_mol, seed=100 * rank + b)
        host.append({k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in d.items()})
    resident = [{k: v.to(dev) for k, v in d.items()} for d in host]
    h2d_keys = ["pos", "atomic_numbers", "batch", "ptr"] + (["target_energy", "target_forces"] if train else []) + \
               (["cell", "pbc"] if args.workload == "c5" else [])
    loss_host = torch.zeros(1).pin_memory()
    energy_host = torch.zeros(host[0]["ptr"].numel() - 1).pin_memory()

    sharded = args.workload == "c5" and world > 1
    if sharded:
        # spatial domain sharding with halo exchange (xequinet_b200/domain.py): strong scaling of ONE box
        from xequinet_b200 import domain
        own_host = []
        for d in host:
            o = domain.shard_atoms({k: v for k, v in d.items() if torch.is_tensor(v)}, rank, world)
            own_host.append({k: v.pin_memory() for k, v in o.items()})
        own_res = [{k: v.to(dev) for k, v in o.items()} for o in own_host]
        for a, b in zip(host, own_host):
            a["_owned"] = b
        for a, b in zip(resident, own_res):
            a["_owned"] = b
        energy_host = torch.zeros(1).pin_memory()
        forces_own_host = torch.zeros(own_host[0]["pos"].shape).pin_memory()

    def step_sharded(batch, e2e: bool):
        o = batch["_owned"]
        if e2e:
            o = {k: o[k].to(dev, non_blocking=True) for k in ("pos", "atomic_numbers", "cell")}
        out = domain.energy_forces_sharded(model, o, rank, world)
        e_tot = out["energy"].detach().sum().reshape(1)
        dist.all_reduce(e_tot)  # total energy of the box (forces stay sharded)
        if e2e:
            energy_host.copy_(e_tot, non_blocking=True)
            forces_own_host.copy_(out["forces"], non_blocking=True)  # this rank's share of the forces
        return e_tot

    def step(batch, e2e: bool):
        if sharded:
            return step_sharded(batch, e2e)
        if e2e:
            d = {k: batch[k].to(dev, non_blocking=True) for k in h2d_keys}
        else:
            d = {k: batch[k] for k in h2d_keys}
        d = transform(d)  # K1 (+ transposed CSR)
        out = model(d, compute_forces=forces)
        if train:
            loss = loss_fn(out, d, forces)
            if flat is not None:
                flat.zero()
            else:
                opt.zero_grad(set_to_none=True)
            loss.backward()  # N > 1: buckets are all-reduced (NCCL over NVLink) as their last gradient is written
            if flat is not None:
                flat.finish()
            opt.step()
            result = loss.detach()
            if e2e:
                loss_host.copy_(result.reshape(1), non_blocking=True)
        else:
            result = out["energy"]
            if e2e:
                energy_host.copy_(result.detach(), non_blocking=True)
                if forces:
                    forces_host[: out["forces"].shape[0]].copy_(out["forces"], non_blocking=True)  # the metric is energy + forces
        return result

    flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    # ---- CUDA-graph replay of the whole step (xequinet_b200.replay.CapturedStep: K1 in capacity mode + model + loss
    # + backward + gradient all-reduce + AdamW captured once; static input buffers, no host work inside the step;
    # tests/test_gpu_bench_shapes.py checks replay == eager bit for bit).  A captured graph needs one static batch
    # structure: workloads whose batches differ in size (c4: 30..70 atoms per molecule) are replayed eagerly.
    native_runtime = False
    same_shape = all(d["pos"].shape == resident[0]["pos"].shape and torch.equal(d["ptr"], resident[0]["ptr"]) for d in resident)
    use_graph = (not args.eager) and same_shape
    graph_step = None
    forces_host = torch.zeros(max(d["pos"].shape[0] for d in host), 3).pin_memory() if (forces and not train) else None
    copied_keys = list(h2d_keys)
    if use_graph and sharded:
        # one CUDA graph per rank (domain.ShardedStep): skin-padded halo plan built once, K1 in capacity mode, the
        # per-layer NCCL all-to-alls are graph nodes; a step copies the rank's positions in and E / F out
        captured = domain.ShardedStep(model, own_res[0], rank, world, skin=0.5)
        copied_keys = ["pos"]

        def graph_step(batch, e2e: bool):
            out = captured(batch["_owned"]["pos"])
            if e2e:
                energy_host.copy_(out["energy"], non_blocking=True)
                forces_own_host.copy_(out["forces"], non_blocking=True)
            return out["energy"]

    elif use_graph:
        from xequinet_b200.graph import build_graph
        from xequinet_b200.replay import CapturedStep
        e_max = 0
        for d in resident:
            g_dyn, _, _ = build_graph(d["pos"], cfg.cutoff, ptr=d["ptr"], batch=d["batch"], cell=d.get("cell"), pbc=d.get("pbc"))
            e_max = max(e_max, g_dyn.n_edges)
        if train:
            opt = torch.optim.AdamW(params, lr=5e-4, fused=True, capturable=True)
        step_model = model
        if not train:
            # inference: the captured step runs the C inference runtime (xequinet_b200.runtime.NativeModel ->
            # xeq_model_energy_forces_mt: forward + force pass scheduled inside the library, independent branches on a
            # second stream); checked bit for bit against the autograd module path before it is timed
            from xequinet_b200 import runtime
            step_model = runtime.NativeModel(model)
            chk_in = transform({k: resident[0][k] for k in h2d_keys})
            chk_mod = model(dict(chk_in), compute_forces=forces)
            chk_nat = step_model(dict(chk_in), compute_forces=forces)
            assert all(torch.equal(chk_nat[k], chk_mod[k].detach()) for k in chk_nat), "inference runtime differs from the module path"
            native_runtime = True
        captured = CapturedStep(step_model, {k: resident[0][k] for k in h2d_keys}, compute_forces=forces,
                                loss_fn=(lambda out, d: loss_fn(out, d, forces)) if train else None, optimizer=opt,
                                flat_grads=flat, edge_capacity=int(e_max * 1.15) + 1024, input_keys=h2d_keys)
        copied_keys = [k for k in h2d_keys if k not in ("ptr", "batch", "pbc")]  # the batch structure is static
        if not train:  # the timed path gives the eager result, bit for bit
            ref = {k: v.clone() for k, v in captured.eager(resident[0]).items()}
            got = captured(resident[0])
            assert all(torch.equal(got[k], ref[k]) for k in ref), "CUDA-graph replay differs from the eager step"

        def graph_step(batch, e2e: bool):
            out = captured(batch)  # H2D of the step's inputs (pinned host -> static buffers) + one graph launch
            r = out["loss"] if train else out["energy"]
            if e2e:
                if train:
                    loss_host.copy_(r.reshape(-1), non_blocking=True)
                else:
                    energy_host.copy_(r.reshape(-1), non_blocking=True)
                    if forces:
                        forces_host.copy_(out["forces"], non_blocking=True)
            return r

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    profile_calls = {}

    def timed(e2e: bool, steps: int, warmup: int, record_kernels: bool, fn=None):
        fn = fn or step
        src = host if e2e else resident
        for i in range(warmup):
            fn(src[i % n_batches], e2e)
        barrier()
        ops.KernelTimer.records = []
        ops.KernelTimer.enabled = record_kernels
        if record_kernels:
            _lib.start_profile()  # per-entry-point call profile of the timed steps only
        launches0 = _lib.get().xeq_launch_count()
        total_ms = 0.0
        for i in range(steps):
            flush_buf.zero_()  # L2 flush between timed iterations (not timed)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            fn(src[i % n_batches], e2e)
            b.record()
            torch.cuda.synchronize()
            total_ms += a.elapsed_time(b)
        barrier()
        ops.KernelTimer.enabled = False
        if record_kernels:
            profile_calls.clear()
            profile_calls.update(_lib.stop_profile())
        launches = _lib.get().xeq_launch_count() - launches0
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max over ranks
        return float(t.item()), launches

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    if use_graph:
        total_ms, _ = timed(False, args.steps, args.warmup, False, graph_step)
        e2e_ms, _ = timed(True, args.steps, max(3, args.warmup // 2), False, graph_step)
        captured.check()  # edge capacity of the captured graph never exceeded
        # kernel-level timings and the launch count come from an eager replica of the same step
        if train:
            opt = torch.optim.AdamW(params, lr=5e-4, fused=True)
        kern_steps = min(args.steps, 5)
        eager_ms, launches = timed(False, kern_steps, 3, True)
        calls = dict(profile_calls)
        kern = ops.KernelTimer.summary()
        launches = launches * args.steps // kern_steps
    else:
        # eagerly launched INFERENCE steps go through the C inference runtime (xequinet_b200.runtime.NativeModel ->
        # xeq_model_energy_forces: forward + force pass scheduled inside the library, bit-identical to the module path):
        # no autograd graph and no per-op Python between the ~110 launches of a step
        headline = step
        if not train and not sharded and not args.eager:
            from xequinet_b200 import runtime
            native = runtime.NativeModel(model)
            chk_mod = model(transform({k: resident[0][k] for k in h2d_keys}), compute_forces=forces)
            chk_nat = native(transform({k: resident[0][k] for k in h2d_keys}), compute_forces=forces)
            assert all(torch.equal(chk_nat[k], chk_mod[k].detach()) for k in chk_nat), "inference runtime differs from the module path"

            def headline(batch, e2e: bool):
                d = {k: (batch[k].to(dev, non_blocking=True) if e2e else batch[k]) for k in h2d_keys}
                out = native(transform(d), compute_forces=forces)
                if e2e:
                    energy_host.copy_(out["energy"], non_blocking=True)
                    if forces:
                        forces_host[: out["forces"].shape[0]].copy_(out["forces"], non_blocking=True)
                return out["energy"]

            native_runtime = True
        total_ms, launches = timed(False, args.steps, args.warmup, False, headline)
        e2e_ms, _ = timed(True, args.steps, max(3, args.warmup // 2), False, headline)
        # kernel-level timings from a separate pass: the per-call events and the call profile cost host time, which
        # an eagerly launched step is bound by
        kern_steps = min(args.steps, 5)
        eager_ms, _ = timed(False, kern_steps, 3, True)
        calls = dict(profile_calls)
        kern = ops.KernelTimer.summary()
    if sampler:
        sampler.stop_flag.set()
        sampler.join(timeout=2)

    if args.workload == "c5":
        n_units = host[0]["pos"].shape[0] // 3  # water molecules in the box
        mols_per_step = n_units if sharded else n_units * world
    else:
        mols_per_step = n_mol * world
    ms_per_step = total_ms / args.steps
    value = mols_per_step / (ms_per_step * 1e-3)
    e2e_value = mols_per_step / (e2e_ms / args.steps * 1e-3)
    # what a step really copies (ragged workloads: the mean over the batches)
    h2d = sum(d[k].numel() * d[k].element_size() for d in host for k in copied_keys) // len(host)
    d2h = 4 if train else energy_host.numel() * 4 + (sum(d["pos"].numel() for d in host) // len(host) * 4 if forces else 0)
    if sharded:
        h2d = sum(host[0]["_owned"][k].numel() * host[0]["_owned"][k].element_size() for k in (copied_keys if use_graph else ("pos", "atomic_numbers", "cell")))
        d2h = 4 + forces_own_host.numel() * 4

    def hard_exit():
        # NCCL teardown with captured graphs alive can hang; all timed work is done, so leave
        # without running destructors (rank 0 still has the CPU baseline and the print to do).
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)

    if rank != 0:
        torch.cuda.synchronize()
        hard_exit()

    # roofline of the dominant hand-written kernel (by share of the step)
    peak, peak_src = measured_hbm_peak()
    periodic = args.workload == "c5"
    kernels = {}
    dom, dom_share = None, -1.0
    for kind, (cnt, mean_ms, N, E) in kern.items():
        by = algorithmic_bytes(kind, cfg, N, E, periodic)
        share = (cnt / kern_steps) * mean_ms / (total_ms / args.steps)
        kernels[kind] = {"launches_per_step": cnt / kern_steps, "mean_ms": round(mean_ms, 5), "share_of_step": round(share, 4),
                         "algorithmic_GB_s": round(by / (mean_ms * 1e-3) / 1e9, 2)}
        if share > dom_share:
            dom, dom_share = kind, share
    roofline = None
    if dom is not None:
        cnt, mean_ms, N, E = kern[dom]
        achieved = algorithmic_bytes(dom, cfg, N, E, periodic) / (mean_ms * 1e-3) / 1e9
        traffic = NCU_TRAFFIC_C3.get(dom) if (args.workload == "c3" and N == 5376) else None
        launch_of = {"edge_fwd": "xeq_edge_message_fwd: center_fwd_ul_kernel (+ pack_fwd_kernel when tiles exceed the window)",
                     "edge_bwd": "xeq_edge_message_bwd: nbr_bwd_ul_kernel + pos_grad",
                     "edge_bwd_wgrad": "xeq_edge_message_bwd with weight gradients: nbr_bwd_ul_kernel + wgrad_ul_kernel<1> + reductions",
                     "edge_bwdbwd": "xeq_edge_message_bwdbwd: JVP = center_fwd_ul_kernel<.., 1> + <.., 2>; reverse = nbr_bwd2_ul_kernel<.., 2> + <.., 3>; wgrad_ul_kernel<2> + reductions"}
        roofline = {"kernel": dom, "launch": launch_of.get(dom), "bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                    "frac": round(achieved / peak, 5), "traffic": traffic, "peak_source": peak_src,
                    "n_nodes": N, "n_edges": E, "mean_launch_ms": round(mean_ms, 5)}

    # step shares by kernel group (C-ABI entry points bracketed by CUDA events in the eager replica of the step; what
    # is left is torch's own elementwise / optimizer / NCCL kernels and launch gaps) and the tensor-pipe roofline of K3
    group_of = lambda n: ("edge" if n.startswith("xeq_edge") else "gemm" if n.startswith("xeq_gemm") else
                          "graph" if (n.startswith("xeq_radius") or n.startswith("xeq_csr")) else "node")
    groups = {}
    for name, (cnt, ms, fl) in calls.items():
        c, t, f = groups.get(group_of(name), (0, 0.0, 0.0))
        groups[group_of(name)] = (c + cnt, t + ms, f + fl)
    eager_step_ms = eager_ms / kern_steps
    step_shares = {k: {"calls_per_step": round(c / kern_steps, 1), "ms_per_step": round(t / kern_steps, 4),
                       "share_of_eager_step": round(t / kern_steps / eager_step_ms, 4)} for k, (c, t, f) in sorted(groups.items())}
    step_shares["torch_and_gaps"] = {"ms_per_step": round(eager_step_ms - sum(t for _, t, _ in groups.values()) / kern_steps, 4)}
    step_shares["eager_step_ms"] = round(eager_step_ms, 4)
    roofline_tensor = None
    if "gemm" in groups and groups["gemm"][1] > 0:
        c, t, f = groups["gemm"]
        tf32_peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["bf16_tflops"] / 2 if (ROOT / "MEASURED_PEAKS.json").exists() else 1590.0 / 2
        achieved = 3.0 * f / (t * 1e-3) / 1e12  # three TF32 products per fp32 product
        roofline_tensor = {"kernel": "gemm_tf32x3_kernel (K3: nn.Linear / o3.Linear and their derivatives)", "bound": "tensor",
                           "achieved": round(achieved, 2), "peak": round(tf32_peak, 1), "unit": "TFLOP/s", "frac": round(achieved / tf32_peak, 5),
                           "peak_source": "MEASURED_PEAKS.json bf16 burst / 2 (kind::tf32 runs at half the bf16 rate)",
                           "fp32_equivalent_TFLOP_s": round(achieved / 3.0, 2), "launches_per_step": round(c / kern_steps, 1),
                           "mean_launch_us": round(t / c * 1e3, 2)}

    # the fused edge forward kernel with a working set >> L2 (SURVEY.md 8d "throughput mode"), default widths only
    roofline_throughput = edge_throughput_probe(cfg, dev, peak) if (cfg is orc.CONFIG_DEFAULT and world == 1) else None

    # CPU oracle on the host cores, bounded sample
    cpu_mols = {"c1": 64, "c2": 64, "c3": 32, "c4": 8, "c5": 1}[args.workload]
    if args.no_cpu_baseline:
        cpu_val, cpu_ms, cpu_sample = 0.0, 0.0, "skipped (--no-cpu-baseline)"
    else:
        cpu_val, cpu_ms, cpu_sample = time_cpu(args.workload, cpu_mols, steps=2, warmup=1)

    torch_gpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            tg_val, tg_ms, tg_sample = time_torch_gpu(args.workload, n_mol, dev)
            torch_gpu = {"value": round(tg_val, 2), "unit": UNIT, "ms_per_step": round(tg_ms, 3), "kind": "port on stock PyTorch CUDA ops",
                         "sample": f"oracle/xpainn_oracle.py on this GPU (eager, edge list not timed), {tg_sample}"}
        except Exception as e:  # a baseline must never take the bench line down (e.g. out of memory at c5)
            torch_gpu = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
            torch.cuda.empty_cache()

    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (seeded generators of SURVEY.md 8d), random-init weights",
        "config": {"workload": f"{args.workload}: {w['desc']}", "molecules_per_gpu": n_mol,
                   "atoms_per_gpu": int(host[0]["pos"].shape[0]), "l2": "flushed between timed steps (256 MB write, untimed)",
                   "parallelism": (f"dp{world}" if args.workload != "c5" else
                                   (f"spatial slabs x{world} + per-layer halo exchange (NCCL all-to-all), one CUDA graph per rank" if sharded else "single GPU")),
                   "execution": ("whole step replayed as one CUDA graph per rank (xequinet_b200.domain.ShardedStep)" if sharded else
                                 ("whole step replayed as one CUDA graph (xequinet_b200.replay.CapturedStep, K1 in capacity mode)" +
                                  (", E+F through the C inference runtime xeq_model_energy_forces_mt (bit-identical to the module path)"
                                   if native_runtime else ""))) if use_graph else
                                ("eager launches" if args.eager else
                                 ("eager launches (batch shapes vary: no static graph)" +
                                  (", E+F through the C inference runtime xeq_model_energy_forces (bit-identical to the module path)"
                                   if native_runtime else "")))},
        "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "roofline_throughput": roofline_throughput,
        "roofline_tensor": roofline_tensor,
        "kernels": kernels,
        "step_shares": step_shares,
        "cpu_baseline": {"value": round(cpu_val, 3), "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                         "sample": f"CPU oracle (oracle/xpainn_oracle.py), {cpu_sample}, {round(cpu_ms, 1)} ms/step, 2 timed steps"},
        "torch_gpu_baseline": torch_gpu,
        "clocks": sampler.summary() if sampler else None,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        hard_exit()


def run_reference(args):
    """The reference's CPU path (oracle port: the reference itself is Python + uninstallable
    third-party wheels, it cannot travel to the GPU box) on all host cores, bounded sample."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cpu_mols = {"c1": 64, "c2": 64, "c3": 32, "c4": 8, "c5": 1}[args.workload]
    steps, warmup = args.steps, args.warmup  # each step is a bounded sample (32 molecules at c3: ~0.3 s on 16 cores)
    val, ms, sample = time_cpu(args.workload, cpu_mols, steps=steps, warmup=warmup)
    w = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": METRIC, "value": round(val, 3), "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", 1)),
        "steps": steps, "warmup": warmup, "ms_per_step": round(ms, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w['desc']}", "sample": sample},
        "cpu_baseline": {"value": round(val, 3), "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": round(val, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--molecules", type=int, default=0, help="molecules per GPU (default: the workload's)")
    ap.add_argument("--eager", action="store_true", help="no CUDA-graph capture of the step")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle leg (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
