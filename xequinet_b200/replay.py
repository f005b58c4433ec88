"""Whole-step CUDA-graph replay: neighbour list (K1, capacity mode) + model + forces (+ loss, backward, gradient
all-reduce, optimizer) captured once and replayed without host work.

This is how the small-molecule workloads are meant to be run (at 1-5 k atoms everything is
launch / latency bound): one `CapturedStep` per batch *shape* (atom count and molecule boundaries are static, positions,
species and targets change every step), used by bench.py for its timed region and by MD-style loops
(interface/ase_calculator.py:75-118 rebuilds the list and syncs to the host at every step).

    step = CapturedStep(model, example_batch, compute_forces=True)            # inference
    out = step(batch)            # {'energy': ..., 'forces': ...}: STATIC tensors, overwritten by the next call

    step = CapturedStep(model, example_batch, loss_fn=loss_fn, optimizer=opt) # training (optimizer: capturable=True)
    out = step(batch)            # {'loss': ..., 'energy': ..., 'forces': ...}

`step.eager(batch)` runs exactly the same work without the graph (tests compare the two bit for bit);
`step.check()` raises if a replayed structure had more edges than the captured capacity."""
from __future__ import annotations

from typing import Callable, Dict, Iterable, Optional

import torch

from . import keys
from .graph import StaticGraphBuilder, build_graph

_STRUCTURE_KEYS = (keys.BATCH, keys.BATCH_PTR, keys.PBC)


class CapturedStep:
    def __init__(self, model: torch.nn.Module, example: Dict[str, torch.Tensor], *, compute_forces: bool = True,
                 loss_fn: Optional[Callable] = None, optimizer: Optional[torch.optim.Optimizer] = None,
                 flat_grads=None, edge_capacity: Optional[int] = None, capacity_margin: float = 1.15,
                 input_keys: Optional[Iterable[str]] = None, warmup: int = 3, capture: bool = True):
        pos = example[keys.POSITIONS]
        if not pos.is_cuda:
            raise RuntimeError("CapturedStep needs CUDA tensors: xequinet_b200 has no CPU fallback")
        if (loss_fn is None) != (optimizer is None):
            raise ValueError("training needs both loss_fn and optimizer")
        self.model, self.forces, self.loss_fn, self.opt, self.flat = model, compute_forces, loss_fn, optimizer, flat_grads
        self.train = loss_fn is not None
        cutoff = float(model.cutoff_radius)
        if input_keys is None:
            input_keys = [k for k, v in example.items() if torch.is_tensor(v) and not k.startswith("_xeq")
                          and k not in (keys.EDGE_INDEX, keys.CELL_OFFSETS)]
        self.input_keys = list(input_keys)
        self.static = {k: example[k].detach().clone() for k in self.input_keys}  # detached: an earlier eager call may have marked pos for differentiation
        if keys.BATCH_PTR not in self.static:
            n = pos.shape[0]
            self.static[keys.BATCH] = torch.zeros(n, dtype=torch.long, device=pos.device)
            self.static[keys.BATCH_PTR] = torch.tensor([0, n], dtype=torch.long, device=pos.device)
        if edge_capacity is None:
            g, _, _ = build_graph(pos, cutoff, ptr=self.static[keys.BATCH_PTR], batch=self.static.get(keys.BATCH),
                                  cell=self.static.get(keys.CELL), pbc=self.static.get(keys.PBC))
            edge_capacity = int(g.n_edges * capacity_margin) + 1024
        self.builder = StaticGraphBuilder(pos.shape[0], self.static[keys.BATCH_PTR], cutoff, int(edge_capacity),
                                          cell=self.static.get(keys.CELL), pbc=self.static.get(keys.PBC))
        self.out: Dict[str, torch.Tensor] = {}
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        if capture:
            self._capture(warmup)

    # ---- the step itself --------------------------------------------------------------------------------
    def _body(self) -> Dict[str, torch.Tensor]:
        res: Dict[str, torch.Tensor] = {}
        d = {k: v for k, v in self.static.items() if k != keys.PBC}
        d[keys.GRAPH] = self.builder.build(self.static[keys.POSITIONS], check_overflow=False)
        out = self.model(d, compute_forces=self.forces)
        if self.train:
            loss = self.loss_fn(out, d)
            loss.backward()
            if self.flat is not None:
                self.flat.finish()
            self.opt.step()
            res["loss"] = loss.detach()
        for k, v in out.items():
            res[k] = v.detach()
        return res

    def _reset_grads(self) -> None:
        if self.train:
            if self.flat is not None:
                self.flat.zero()
            else:
                self.opt.zero_grad(set_to_none=True)
        self.static[keys.POSITIONS].grad = None

    def _capture(self, warmup: int) -> None:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                self._reset_grads()
                self._body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        self._reset_grads()
        with torch.cuda.graph(self.graph):
            self.out = self._body()  # the graph's static output tensors
        torch.cuda.synchronize()

    def load(self, batch: Dict[str, torch.Tensor]) -> None:
        """Copy a batch of the captured shape into the static input buffers (host or device tensors; pinned host
        tensors copy asynchronously)."""
        with torch.no_grad():
            for k in self.input_keys:
                if k in _STRUCTURE_KEYS:
                    continue  # the batch structure is part of the captured shape
                self.static[k].copy_(batch[k], non_blocking=True)

    def __call__(self, batch: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        if self.graph is None:
            return self.eager(batch)
        if batch is not None:
            self.load(batch)
        self.graph.replay()
        return self.out

    def eager(self, batch: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        if batch is not None:
            self.load(batch)
        self._reset_grads()
        return self._body()

    def check(self) -> None:
        """Host synchronisation: raises when a structure seen since the last check exceeded the edge capacity."""
        if int(self.builder.overflow.item()) != 0:
            self.builder.overflow.zero_()
            raise RuntimeError(f"CapturedStep: a structure had more than edge_capacity = {self.builder.cap} edges")
