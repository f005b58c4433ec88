"""torch.autograd glue over the C ABI (include/xeq_b200.h).

Each op is a `torch.autograd.Function` whose backward is itself an autograd Function backed
by a hand-written kernel, so that `torch.autograd.grad(E, pos, create_graph=True)` followed by
`loss.backward()` (nn/basic.py:150-156, utils/trainer.py:302) runs K2 -> K2b -> K2bb without
any eager fallback."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from .graph import NeighborGraph


@dataclass(frozen=True)
class Dims:
    node_dim: int
    mul0: int
    mul1: int
    mul2: int
    num_basis: int
    cutoff: float

    @property
    def M(self):
        return self.mul0 + self.mul1 + self.mul2

    @property
    def D(self):
        return self.mul0 + 3 * self.mul1 + 5 * self.mul2

    @property
    def H(self):
        return self.node_dim + 2 * self.M

    def struct(self) -> _lib.XeqDims:
        return _lib.XeqDims(self.node_dim, self.mul0, self.mul1, self.mul2, self.num_basis, self.cutoff)


def _c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        raise RuntimeError(f"xequinet_b200 kernels compute in fp32, got {t.dtype}")
    return t.contiguous()


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


from ._state import input_wanted  # noqa: E402


class KernelTimer:
    """Optional CUDA-event timing of the edge kernels on the launching stream (bench.py's
    roofline leg).  Disabled by default: no events are recorded on the product path."""

    enabled = False
    records = []  # (kind, n_nodes, n_edges, start_event, end_event)

    @classmethod
    def start(cls):
        if not cls.enabled:
            return None
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        return ev

    @classmethod
    def stop(cls, ev, kind, graph):
        if ev is None:
            return
        end = torch.cuda.Event(enable_timing=True)
        end.record()
        cls.records.append((kind, graph.n_nodes, graph.n_edges, ev, end))

    @classmethod
    def summary(cls):
        """{kind: (launches, mean ms, n_nodes, n_edges)} -- call after torch.cuda.synchronize()."""
        out = {}
        for kind, n, e, a, b in cls.records:
            cnt, tot, _, _ = out.get(kind, (0, 0.0, n, e))
            out[kind] = (cnt + 1, tot + a.elapsed_time(b), n, e)
        return {k: (c, t / c, n, e) for k, (c, t, n, e) in out.items()}


# ------------------------------------------------------------------------------------------
# raw kernel calls (no autograd)
# ------------------------------------------------------------------------------------------
def edge_message_fwd_raw(graph: NeighborGraph, dims: Dims, pos, s, v, x_in, V_in, W, b, freq):
    lib = _lib.get()
    N = graph.n_nodes
    x_out = torch.empty((N, dims.node_dim), dtype=torch.float32, device=s.device)
    V_out = torch.empty((N, dims.D), dtype=torch.float32, device=s.device)
    d = dims.struct()
    nbytes = lib.xeq_edge_message_fwd_workspace_bytes(graph.struct, d)
    ws = _workspace(nbytes, s.device)
    ev = KernelTimer.start()
    _lib.check(lib.xeq_edge_message_fwd(graph.struct, d, _lib.ptr(pos), _lib.ptr(s), _lib.ptr(v), _lib.ptr(x_in),
                                        _lib.ptr(V_in), _lib.ptr(W), _lib.ptr(b), _lib.ptr(freq), _lib.ptr(x_out),
                                        _lib.ptr(V_out), _lib.ptr(ws), nbytes, _lib.stream()), "xeq_edge_message_fwd")
    KernelTimer.stop(ev, "edge_fwd", graph)
    return x_out, V_out


def edge_message_bwd_raw(graph: NeighborGraph, dims: Dims, pos, s, v, W, b, freq, gx, gV, need_s=True, need_v=True,
                         need_pos=True, need_w=False, need_cell=False):
    """K2b.  need_cell (periodic graphs; implies need_pos): additionally returns dE/dcell [G,3,3] = -sum_e offsets_e (x)
    dE/dr_e from the per-edge d/dr records of the same launch (virial through the strain trick, nn/basic.py:93-107)."""
    lib = _lib.get()
    N, dev = graph.n_nodes, s.device
    new = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    gs = new(N, dims.H) if need_s else None
    gv = new(N, dims.D) if need_v else None
    gpos = new(N, 3) if (need_pos or need_cell) else None
    gW = new(dims.H, dims.num_basis) if need_w else None
    gb = new(dims.H) if need_w else None
    gf = new(dims.num_basis) if need_w else None
    d = dims.struct()
    nbytes = lib.xeq_edge_message_bwd_workspace_bytes(graph.struct, d, int(need_w))
    ws = _workspace(nbytes, dev)
    ev = KernelTimer.start()
    _lib.check(lib.xeq_edge_message_bwd(graph.struct, d, _lib.ptr(pos), _lib.ptr(s), _lib.ptr(v), _lib.ptr(W),
                                        _lib.ptr(b), _lib.ptr(freq), _lib.ptr(gx), _lib.ptr(gV), _lib.ptr(gs),
                                        _lib.ptr(gv), _lib.ptr(gpos), _lib.ptr(gW), _lib.ptr(gb), _lib.ptr(gf),
                                        _lib.ptr(ws), nbytes, _lib.stream()), "xeq_edge_message_bwd")
    KernelTimer.stop(ev, "edge_bwd_wgrad" if need_w else "edge_bwd", graph)
    if not need_cell:
        return gs, gv, gpos, gW, gb, gf
    return gs, gv, gpos, gW, gb, gf, _cell_grad_from_records(graph, d, ws)


def _cell_grad_from_records(graph: NeighborGraph, d, ws) -> torch.Tensor:
    """dE/dcell [G,3,3] = -sum_e offsets_e (x) dE/dr_e from the per-edge d/dr records a backward launch left in `ws`."""
    lib = _lib.get()
    if graph.offsets is None or getattr(graph, "seg_ptr", None) is None:
        raise RuntimeError("the cell gradient needs a periodic graph with its batch pointer (graph.seg_ptr)")
    N, dev = graph.n_nodes, ws.device
    rows = torch.empty((9, N), dtype=torch.float32, device=dev)
    _lib.check(lib.xeq_edge_cell_grad_rows(graph.struct, d, _lib.ptr(ws), _lib.ptr(rows), _lib.stream()), "xeq_edge_cell_grad_rows")
    G = graph.seg_ptr.numel() - 1
    sums = torch.empty((9, G), dtype=torch.float32, device=dev)
    for k in range(9):  # nine deterministic segment sums over the nodes of each graph
        _lib.check(lib.xeq_segment_sum(_lib.ptr(rows[k]), _lib.ptr(graph.seg_ptr), G, _lib.ptr(sums[k]), _lib.stream()),
                   "xeq_segment_sum")
    return -sums.t().reshape(G, 3, 3)


def edge_message_bwdbwd_raw(graph: NeighborGraph, dims: Dims, pos, s, v, W, b, freq, gx, gV, a_s, a_v, a_pos,
                            need_g=True, need_s=True, need_v=True, need_pos=True, need_w=True, a_cell=None, need_cell=False):
    """K2bb.  a_cell [G,3,3]: cotangent of the first-order cell gradient (periodic virial in a training loss); with
    need_cell the second-order cell gradient is appended to the result (it implies need_pos)."""
    lib = _lib.get()
    need_pos = need_pos or need_cell
    N, dev = graph.n_nodes, s.device
    new = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    o_gx = new(N, dims.node_dim) if need_g else None
    o_gV = new(N, dims.D) if need_g else None
    o_s = new(N, dims.H) if need_s else None
    o_v = new(N, dims.D) if need_v else None
    o_pos = new(N, 3) if need_pos else None
    o_W = new(dims.H, dims.num_basis) if need_w else None
    o_b = new(dims.H) if need_w else None
    o_f = new(dims.num_basis) if need_w else None
    d = dims.struct()
    nbytes = lib.xeq_edge_message_bwdbwd_workspace_bytes(graph.struct, d, int(need_w))
    ws = _workspace(nbytes, dev)
    ev = KernelTimer.start()
    _lib.check(lib.xeq_edge_message_bwdbwd(graph.struct, d, _lib.ptr(pos), _lib.ptr(s), _lib.ptr(v), _lib.ptr(W),
                                           _lib.ptr(b), _lib.ptr(freq), _lib.ptr(gx), _lib.ptr(gV), _lib.ptr(a_s),
                                           _lib.ptr(a_v), _lib.ptr(a_pos), _lib.ptr(a_cell), _lib.ptr(o_gx), _lib.ptr(o_gV), _lib.ptr(o_s),
                                           _lib.ptr(o_v), _lib.ptr(o_pos), _lib.ptr(o_W), _lib.ptr(o_b), _lib.ptr(o_f),
                                           _lib.ptr(ws), nbytes, _lib.stream()), "xeq_edge_message_bwdbwd")
    KernelTimer.stop(ev, "edge_bwdbwd", graph)
    if need_cell:
        return o_gx, o_gV, o_s, o_v, o_pos, o_W, o_b, o_f, _cell_grad_from_records(graph, d, ws)
    return o_gx, o_gV, o_s, o_v, o_pos, o_W, o_b, o_f


# ------------------------------------------------------------------------------------------
# autograd
# ------------------------------------------------------------------------------------------
class _EdgeMessageBwd(torch.autograd.Function):
    """K2b as a differentiable function of (gx, gV, s, v, pos, W, b, freq, cell); its backward is K2bb."""

    @staticmethod
    def forward(ctx, gx, gV, s, v, pos, W, b, freq, cell, graph, dims, needs):
        need_s, need_v, need_pos, need_w = needs[:4]
        need_cell = len(needs) > 4 and needs[4]
        gx, gV = _c(gx), _c(gV)
        res = edge_message_bwd_raw(graph, dims, pos, s, v, W, b, freq, gx, gV, need_s, need_v, need_pos, need_w, need_cell)
        gs, gv, gpos, gW, gb, gf = res[:6]
        gcell = res[6] if need_cell else None
        ctx.save_for_backward(gx, gV, s, v, pos, W, b, freq)
        ctx.graph, ctx.dims, ctx.needs = graph, dims, needs
        outs = (gs, gv, gpos, gW, gb, gf, gcell)
        ctx.mark_non_differentiable(*[o for o in outs[3:6] if o is not None])
        return outs

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, a_s, a_v, a_pos, a_W, a_b, a_f, a_cell=None):
        gx, gV, s, v, pos, W, b, freq = ctx.saved_tensors
        ni = ctx.needs_input_grad
        if a_s is None and a_v is None and a_pos is None and a_cell is None:
            return (None,) * 12
        need_g = ni[0] or ni[1]
        need_w = ni[5] or ni[6] or ni[7]
        need_cell = bool(ni[8])
        o = edge_message_bwdbwd_raw(ctx.graph, ctx.dims, pos, s, v, W, b, freq, gx, gV, _c(a_s), _c(a_v), _c(a_pos),
                                    need_g=need_g, need_s=ni[2], need_v=ni[3], need_pos=ni[4], need_w=need_w,
                                    a_cell=_c(a_cell), need_cell=need_cell)
        o_gx, o_gV, o_s, o_v, o_pos, o_W, o_b, o_f = o[:8]
        o_cell = o[8] if need_cell else None
        return (o_gx, o_gV, o_s, o_v, (o_pos if ni[4] else None), o_W, o_b, (o_f.view_as(freq) if o_f is not None else None),
                o_cell, None, None, None)


class _EdgeMessage(torch.autograd.Function):
    """K2: (x, V, s, v, pos, W_rbf, b_rbf, freq) -> (x + sum m_s, V + sum m_e)."""

    @staticmethod
    def forward(ctx, x, V, s, v, pos, W, b, freq, cell, graph, dims):
        # `cell` ([G,3,3] or None) only matters for differentiation (virial): the kernels read the lattice from the
        # graph, which holds the same values (the strain of nn/basic.py:93-107 is zero where it is evaluated)
        x, V, s, v, pos, W, b = (_c(t) for t in (x, V, s, v, pos, W, b))
        freq = _c(freq)
        x_out, V_out = edge_message_fwd_raw(graph, dims, pos, s, v, x, V, W, b, freq)
        ctx.save_for_backward(s, v, pos, W, b, freq, cell)  # cell: only to route its gradient (None otherwise)
        ctx.graph, ctx.dims = graph, dims
        return x_out, V_out

    @staticmethod
    def backward(ctx, gx, gV):
        s, v, pos, W, b, freq, cell = ctx.saved_tensors
        ni = ctx.needs_input_grad
        if gx is None:
            gx = torch.zeros((s.shape[0], ctx.dims.node_dim), dtype=s.dtype, device=s.device)
        if gV is None:
            gV = torch.zeros((s.shape[0], ctx.dims.D), dtype=s.dtype, device=s.device)
        # what this graph task asks for (the force pass of nn/basic.py:150-156 wants d/dpos only: no weight gradients)
        want = [input_wanted(ctx, i) for i in range(9)]
        need_w = want[5] or want[6] or want[7]
        need_cell = want[8]
        needs = (want[2], want[3], want[4] or need_cell, need_w, need_cell)
        gs = gv = gpos = gW = gb = gf = gcell = None
        if any(needs):
            gs, gv, gpos, gW, gb, gf, gcell = _EdgeMessageBwd.apply(gx, gV, s, v, pos, W, b, freq, cell, ctx.graph, ctx.dims, needs)
            if gf is not None:
                gf = gf.view_as(freq)
            if not want[4]:
                gpos = None
        return (gx if ni[0] else None, gV if ni[1] else None, gs, gv, gpos, gW, gb, gf, gcell, None, None)


def edge_message(x, V, s, v, pos, W_rbf, b_rbf, freq, graph: NeighborGraph, dims: Dims, cell=None):
    """Fused XPainnMessage aggregation (nn/xpainn.py:140-159); V and v in the cm layout.  `cell` is only passed
    when its gradient is wanted (virial of a periodic structure)."""
    return _EdgeMessage.apply(x, V, s, v, pos, W_rbf, b_rbf, freq, cell, graph, dims)


class _SegmentSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, seg_ptr, node_graph):
        lib = _lib.get()
        src = _c(src)
        G = seg_ptr.numel() - 1
        out = torch.empty(G, dtype=torch.float32, device=src.device)
        _lib.check(lib.xeq_segment_sum(_lib.ptr(src), _lib.ptr(seg_ptr), G, _lib.ptr(out), _lib.stream()),
                   "xeq_segment_sum")
        ctx.save_for_backward(node_graph)
        return out

    @staticmethod
    def backward(ctx, g):
        (node_graph,) = ctx.saved_tensors
        return g.index_select(0, node_graph), None, None


def segment_sum(src: torch.Tensor, seg_ptr: torch.Tensor, node_graph: torch.Tensor) -> torch.Tensor:
    """scatter_sum(src, batch) for the sorted `batch` of a collated batch (nn/output.py:124)."""
    return _SegmentSum.apply(src, seg_ptr, node_graph)


def layout_convert(V: torch.Tensor, dims: Dims, to_cm: bool) -> torch.Tensor:
    lib = _lib.get()
    V = _c(V)
    out = torch.empty_like(V)
    _lib.check(lib.xeq_layout_convert(_lib.ptr(V), _lib.ptr(out), V.shape[0], dims.struct(), 0 if to_cm else 1,
                                      _lib.stream()), "xeq_layout_convert")
    return out
