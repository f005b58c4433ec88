"""`torch.library` registration of the boundary ops ("`TORCH_LIBRARY(xeq, m)` or ctypes +
`torch.library.custom_op`"): the kernels behind the C ABI as dispatcher-visible operators

    torch.ops.xeq.radius_graph(pos, cutoff, batch)                         <- torch_cluster.radius_graph
    torch.ops.xeq.edge_message(graph tensors..., pos, s, v, x, V, W, b, freq)  <- nn/xpainn.py:140-159
    torch.ops.xeq.edge_message_bwd / edge_message_bwdbwd                   <- its first / second derivatives
    torch.ops.xeq.linear(x, weight, bias)                                  <- nn.Linear
    torch.ops.xeq.irreps_linear(V, weight, bias, muls)                     <- e3nn o3.Linear (cm layout)
    torch.ops.xeq.segment_sum(src, ptr)                                    <- torch_scatter.scatter_sum (sorted index)

with fake (meta) implementations, so that they trace under `torch.compile(fullgraph=True)` / `torch.export` and can
be called from TorchScript (`torch.jit.script` resolves `torch.ops.xeq.*` through the dispatcher), and autograd
formulas registered with `torch.library.register_autograd`, each expressed through the next op of the family --
`torch.autograd.grad(E, pos, create_graph=True)` followed by `loss.backward()` stays on hand-written kernels.

The `xequinet_b200.nn` modules call the same kernels through `torch.autograd.Function`s (ops.py / gemm.py /
nodeops.py): those can ask the autograd engine which gradients a backward pass really needs (_state.input_wanted)
and avoid the dispatcher's per-call cost in eager mode.  Both routes end in the same C entry points; tests compare
them bit for bit (tests/test_gpu_torch_ops.py).

A neighbour structure crosses the op boundary as its tensors + a short int list (`pack_graph` / `unpack_graph`)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import gemm as _gemm
from . import ops as _ops
from .graph import NeighborGraph, build_graph



# ------------------------------------------------------------------------------------------
# graph <-> tensors
# ------------------------------------------------------------------------------------------
class _GraphView:
    """A NeighborGraph rebuilt from the tensors that crossed the op boundary (no kernels are launched)."""

    def __init__(self, tensors: Sequence[Optional[Tensor]], meta: Sequence[int]):
        (self.rowptr, self.col, self.t_rowptr, self.t_row, self.t_eid, self.tile_ptr, self.t_tile_ptr,
         self.offsets, self.cell, self.node_graph, self.seg_ptr) = tensors
        self.n_nodes, self.n_edges, self.n_graphs, self.n_tiles, self.t_n_tiles, self.tile_mode, self.max_tile_nodes = (int(m) for m in meta)
        self.n_centers = self.n_nodes
        self._struct = None

    struct = NeighborGraph.struct


def pack_graph(g: NeighborGraph) -> Tuple[List[Optional[Tensor]], List[int]]:
    tensors = [g.rowptr, g.col, g.t_rowptr, g.t_row, g.t_eid, g.tile_ptr, g.t_tile_ptr, g.offsets, g.cell, g.node_graph,
               getattr(g, "seg_ptr", None)]
    return tensors, [g.n_nodes, g.n_edges, g.n_graphs, g.n_tiles, g.t_n_tiles, g.tile_mode, g.max_tile_nodes]


def _dims(d: Sequence[int], cutoff: float) -> _ops.Dims:
    return _ops.Dims(int(d[0]), int(d[1]), int(d[2]), int(d[3]), int(d[4]), float(cutoff))


def pack_dims(dims: _ops.Dims) -> List[int]:
    return [dims.node_dim, dims.mul0, dims.mul1, dims.mul2, dims.num_basis]


# ------------------------------------------------------------------------------------------
# K1
# ------------------------------------------------------------------------------------------
@torch.library.custom_op("xeq::radius_graph", mutates_args=())
def radius_graph(pos: Tensor, cutoff: float, batch: Optional[Tensor] = None) -> Tensor:
    _, ei, _ = build_graph(pos, cutoff, batch=batch, want_coo=True)
    return ei


@radius_graph.register_fake
def _(pos, cutoff, batch=None):
    n_edges = torch.library.get_ctx().new_dynamic_size()
    return pos.new_empty((2, n_edges), dtype=torch.int64)


# ------------------------------------------------------------------------------------------
# K2 / K2b / K2bb
# ------------------------------------------------------------------------------------------
@torch.library.custom_op("xeq::edge_message", mutates_args=())
def edge_message(graph: List[Optional[Tensor]], meta: List[int], dims: List[int], cutoff: float, pos: Tensor, s: Tensor,
                 v: Tensor, x: Tensor, V: Tensor, W: Tensor, b: Tensor, freq: Tensor) -> Tuple[Tensor, Tensor]:
    c = _ops._c
    return _ops.edge_message_fwd_raw(_GraphView(graph, meta), _dims(dims, cutoff), c(pos), c(s), c(v), c(x), c(V), c(W), c(b),
                                     c(freq))


@edge_message.register_fake
def _(graph, meta, dims, cutoff, pos, s, v, x, V, W, b, freq):
    return torch.empty_like(x), torch.empty_like(V)


@torch.library.custom_op("xeq::edge_message_bwd", mutates_args=())
def edge_message_bwd(graph: List[Optional[Tensor]], meta: List[int], dims: List[int], cutoff: float, pos: Tensor, s: Tensor,
                     v: Tensor, W: Tensor, b: Tensor, freq: Tensor, gx: Tensor, gV: Tensor
                     ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    c = _ops._c
    gs, gv, gpos, gW, gb, gf = _ops.edge_message_bwd_raw(_GraphView(graph, meta), _dims(dims, cutoff), c(pos), c(s), c(v), c(W),
                                                         c(b), c(freq), c(gx), c(gV), need_w=True)
    return gs, gv, gpos, gW, gb, gf.view_as(freq)


@edge_message_bwd.register_fake
def _(graph, meta, dims, cutoff, pos, s, v, W, b, freq, gx, gV):
    return torch.empty_like(s), torch.empty_like(v), torch.empty_like(pos), torch.empty_like(W), torch.empty_like(b), torch.empty_like(freq)


@torch.library.custom_op("xeq::edge_message_bwdbwd", mutates_args=())
def edge_message_bwdbwd(graph: List[Optional[Tensor]], meta: List[int], dims: List[int], cutoff: float, pos: Tensor, s: Tensor,
                        v: Tensor, W: Tensor, b: Tensor, freq: Tensor, gx: Tensor, gV: Tensor, a_s: Tensor, a_v: Tensor,
                        a_pos: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    c = _ops._c
    o = _ops.edge_message_bwdbwd_raw(_GraphView(graph, meta), _dims(dims, cutoff), c(pos), c(s), c(v), c(W), c(b), c(freq), c(gx),
                                     c(gV), c(a_s), c(a_v), c(a_pos))
    return o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7].view_as(freq)


@edge_message_bwdbwd.register_fake
def _(graph, meta, dims, cutoff, pos, s, v, W, b, freq, gx, gV, a_s, a_v, a_pos):
    return (torch.empty_like(gx), torch.empty_like(gV), torch.empty_like(s), torch.empty_like(v), torch.empty_like(pos),
            torch.empty_like(W), torch.empty_like(b), torch.empty_like(freq))


def _edge_message_setup(ctx, inputs, output):
    graph, meta, dims, cutoff, pos, s, v, x, V, W, b, freq = inputs
    ctx.save_for_backward(pos, s, v, W, b, freq)
    ctx.graph, ctx.meta, ctx.dims, ctx.cutoff = graph, meta, dims, cutoff


def _edge_message_backward(ctx, gx, gV):
    pos, s, v, W, b, freq = ctx.saved_tensors
    gs, gv, gpos, gW, gb, gf = torch.ops.xeq.edge_message_bwd(ctx.graph, ctx.meta, ctx.dims, ctx.cutoff, pos, s, v, W, b, freq,
                                                              gx.contiguous(), gV.contiguous())
    return None, None, None, None, gpos, gs, gv, gx, gV, gW, gb, gf


torch.library.register_autograd("xeq::edge_message", _edge_message_backward, setup_context=_edge_message_setup)


def _edge_message_bwd_setup(ctx, inputs, output):
    graph, meta, dims, cutoff, pos, s, v, W, b, freq, gx, gV = inputs
    ctx.save_for_backward(pos, s, v, W, b, freq, gx, gV)
    ctx.graph, ctx.meta, ctx.dims, ctx.cutoff = graph, meta, dims, cutoff


def _edge_message_bwd_backward(ctx, a_s, a_v, a_pos, a_W, a_b, a_f):
    # second derivatives through the weight gradients (a_W, a_b, a_f) are not on the XPaiNN path
    pos, s, v, W, b, freq, gx, gV = ctx.saved_tensors
    z = lambda a, like: torch.zeros_like(like) if a is None else a.contiguous()
    o = torch.ops.xeq.edge_message_bwdbwd(ctx.graph, ctx.meta, ctx.dims, ctx.cutoff, pos, s, v, W, b, freq, gx, gV, z(a_s, s),
                                          z(a_v, v), z(a_pos, pos))
    o_gx, o_gV, o_s, o_v, o_pos, o_W, o_b, o_f = o
    return None, None, None, None, o_pos, o_s, o_v, o_W, o_b, o_f, o_gx, o_gV


torch.library.register_autograd("xeq::edge_message_bwd", _edge_message_bwd_backward, setup_context=_edge_message_bwd_setup)


# ------------------------------------------------------------------------------------------
# K3
# ------------------------------------------------------------------------------------------
@torch.library.custom_op("xeq::mm", mutates_args=())
def mm(A: Tensor, B: Tensor, ta: bool, tb: bool, alpha: float = 1.0) -> Tensor:
    return _gemm.mm_raw(A, B, ta, tb, None, alpha)


@mm.register_fake
def _(A, B, ta, tb, alpha=1.0):
    m = A.shape[1] if ta else A.shape[0]
    n = B.shape[0] if tb else B.shape[1]
    return A.new_empty((m, n))


def _mm_setup(ctx, inputs, output):
    A, B, ta, tb, alpha = inputs
    ctx.save_for_backward(A, B)
    ctx.cfg = (ta, tb, alpha)


def _mm_backward(ctx, gC):
    A, B = ctx.saved_tensors
    ta, tb, alpha = ctx.cfg
    gC = gC.contiguous()
    gA = gB = None
    if ctx.needs_input_grad[0]:
        gA = torch.ops.xeq.mm(gC, B, False, not tb, alpha) if not ta else torch.ops.xeq.mm(B, gC, tb, True, alpha)
    if ctx.needs_input_grad[1]:
        gB = torch.ops.xeq.mm(A, gC, not ta, False, alpha) if not tb else torch.ops.xeq.mm(gC, A, True, ta, alpha)
    return gA, gB, None, None, None


torch.library.register_autograd("xeq::mm", _mm_backward, setup_context=_mm_setup)


def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None) -> Tensor:
    """nn.Linear through the dispatcher-visible GEMM (bias added by a torch op: its gradient is torch's)."""
    y = torch.ops.xeq.mm(x, weight, False, True, 1.0)
    return y if bias is None else y + bias


@torch.library.custom_op("xeq::irreps_linear", mutates_args=())
def irreps_linear(V: Tensor, weight: Tensor, bias: Optional[Tensor], muls: List[int], transposed: bool = False) -> Tensor:
    return _gemm.irreps_linear_raw(V, weight, bias, tuple(muls), transposed)


@irreps_linear.register_fake
def _(V, weight, bias, muls, transposed=False):
    return torch.empty_like(V)


@torch.library.custom_op("xeq::irreps_wgrad", mutates_args=())
def irreps_wgrad(A: Tensor, B: Tensor, muls: List[int]) -> Tensor:
    return _gemm.irreps_wgrad_raw(A, B, tuple(muls))


@irreps_wgrad.register_fake
def _(A, B, muls):
    return A.new_empty((sum(int(m) * int(m) for m in muls),))


def _il_setup(ctx, inputs, output):
    V, weight, bias, muls, transposed = inputs
    ctx.save_for_backward(V, weight)
    ctx.cfg = (muls, transposed, bias is not None)


def _il_backward(ctx, g):
    V, w = ctx.saved_tensors
    muls, transposed, has_bias = ctx.cfg
    g = g.contiguous()
    gV = torch.ops.xeq.irreps_linear(g, w, None, muls, not transposed) if ctx.needs_input_grad[0] else None
    gw = None
    if ctx.needs_input_grad[1]:
        gw = torch.ops.xeq.irreps_wgrad(V, g, muls) if not transposed else torch.ops.xeq.irreps_wgrad(g, V, muls)
    gb = g[:, : muls[0]].sum(0) if (has_bias and ctx.needs_input_grad[2]) else None
    return gV, gw, gb, None, None


torch.library.register_autograd("xeq::irreps_linear", _il_backward, setup_context=_il_setup)


def _iw_setup(ctx, inputs, output):
    A, B, muls = inputs
    ctx.save_for_backward(A, B)
    ctx.muls = muls


def _iw_backward(ctx, gw):
    A, B = ctx.saved_tensors
    gw = gw.contiguous()
    gA = torch.ops.xeq.irreps_linear(B, gw, None, ctx.muls, True) if ctx.needs_input_grad[0] else None
    gB = torch.ops.xeq.irreps_linear(A, gw, None, ctx.muls, False) if ctx.needs_input_grad[1] else None
    return gA, gB, None


torch.library.register_autograd("xeq::irreps_wgrad", _iw_backward, setup_context=_iw_setup)


# ------------------------------------------------------------------------------------------
# read-out
# ------------------------------------------------------------------------------------------
@torch.library.custom_op("xeq::segment_sum", mutates_args=())
def segment_sum(src: Tensor, seg_ptr: Tensor) -> Tensor:
    from . import _lib as lib

    src = _ops._c(src)
    G = seg_ptr.numel() - 1
    out = torch.empty(G, dtype=torch.float32, device=src.device)
    lib.check(lib.get().xeq_segment_sum(lib.ptr(src), lib.ptr(seg_ptr), G, lib.ptr(out), lib.stream()), "xeq_segment_sum")
    return out


@segment_sum.register_fake
def _(src, seg_ptr):
    return src.new_empty((seg_ptr.numel() - 1,))


def _ss_setup(ctx, inputs, output):
    src, seg_ptr = inputs
    ctx.save_for_backward(seg_ptr)
    ctx.n = src.shape[0]


def _ss_backward(ctx, g):
    (seg_ptr,) = ctx.saved_tensors
    counts = (seg_ptr[1:] - seg_ptr[:-1]).long()
    return torch.repeat_interleave(g, counts, output_size=ctx.n), None


torch.library.register_autograd("xeq::segment_sum", _ss_backward, setup_context=_ss_setup)
