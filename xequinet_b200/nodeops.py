"""Fused node-side kernels (csrc/node_norm.cu, csrc/node_update.cu) behind autograd Functions whose
backward is again a kernel-backed Function (forces) with a kernel for the double backward (force
training).  No eager fallback: host tensors raise."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._state import input_wanted
from .ops import _c, _workspace


def _rows(t):
    """(tensor, row stride) of a 2-D fp32 tensor whose rows are dense: column slices of a wider tensor (the halves of a
    torch.cat gradient) are passed to the kernels as they are instead of through a .contiguous() copy."""
    if t is None:
        return None, 0
    if t.dtype != torch.float32:
        raise RuntimeError(f"xequinet_b200 kernels compute in fp32, got {t.dtype}")
    if t.dim() == 2 and t.stride(1) == 1 and t.stride(0) >= t.shape[1]:
        return t, t.stride(0)
    t = t.contiguous()
    return t, t.shape[-1]


def _norm_ws(n, muls, device):
    nbytes = _lib.get().xeq_irreps_norm_workspace_bytes(n, *muls)
    return _workspace(nbytes, device), nbytes


def irreps_norm_fwd_raw(x, gamma, beta, muls, eps):
    y = torch.empty_like(x)
    _lib.check(_lib.get().xeq_irreps_norm_fwd(_lib.ptr(x), _lib.ptr(gamma), _lib.ptr(beta), x.shape[0], *muls, eps,
                                              _lib.ptr(y), _lib.stream()), "xeq_irreps_norm_fwd")
    return y


def irreps_norm_bwd_raw(x, gamma, g, muls, eps, need_x=True, need_params=True, gx_add=None):
    gx = torch.empty_like(x) if need_x else None
    gg = torch.empty(sum(muls), dtype=torch.float32, device=x.device) if need_params else None
    gb = torch.empty(muls[0], dtype=torch.float32, device=x.device) if need_params else None
    ws, nbytes = _norm_ws(x.shape[0], muls, x.device) if need_params else (None, 0)
    g, ld_g = _rows(g)
    _lib.check(_lib.get().xeq_irreps_norm_bwd(_lib.ptr(x), _lib.ptr(gamma), _lib.ptr_rows(g), ld_g, _lib.ptr(gx_add), x.shape[0], *muls, eps, _lib.ptr(gx),
                                              _lib.ptr(gg), _lib.ptr(gb), _lib.ptr(ws), nbytes, _lib.stream()),
               "xeq_irreps_norm_bwd")
    return gx, gg, gb


def irreps_norm_bwdbwd_raw(x, gamma, g, a, muls, eps, need_x=True, need_g=True, need_gamma=True):
    dx = torch.empty_like(x) if need_x else None
    dg = torch.empty_like(x) if need_g else None
    dgam = torch.empty(sum(muls), dtype=torch.float32, device=x.device) if need_gamma else None
    ws, nbytes = _norm_ws(x.shape[0], muls, x.device) if need_gamma else (None, 0)
    g, ld_g = _rows(g)
    _lib.check(_lib.get().xeq_irreps_norm_bwdbwd(_lib.ptr(x), _lib.ptr(gamma), _lib.ptr_rows(g), ld_g, _lib.ptr(a), x.shape[0], *muls, eps,
                                                 _lib.ptr(dx), _lib.ptr(dg), _lib.ptr(dgam), _lib.ptr(ws), nbytes,
                                                 _lib.stream()), "xeq_irreps_norm_bwdbwd")
    return dx, dg, dgam


class _IrrepsNormBwd(torch.autograd.Function):
    """gx = d<g, norm(x)>/dx + g_pass (the gradient that reaches x through its pass-through output, summed in-kernel)."""

    @staticmethod
    def forward(ctx, x, gamma, g, g_pass, muls, eps, need_params):
        g = _rows(g)[0]
        ctx.save_for_backward(x, gamma, g)
        ctx.cfg = (muls, eps, g_pass is not None)
        ctx.set_materialize_grads(False)
        gx, gg, gb = irreps_norm_bwd_raw(x, gamma, g, muls, eps, True, need_params, _c(g_pass) if g_pass is not None else None)
        return gx, gg, gb

    @staticmethod
    def backward(ctx, a, a_gamma, a_beta):
        if a_gamma is not None or a_beta is not None:
            raise NotImplementedError("second derivatives through the norm's parameter gradients are not on the XPaiNN path")
        x, gamma, g = ctx.saved_tensors
        muls, eps, has_pass = ctx.cfg
        if a is None:
            return None, None, None, None, None, None, None
        dx, dg, dgam = irreps_norm_bwdbwd_raw(x, gamma, g, _c(a), muls, eps, input_wanted(ctx, 0), input_wanted(ctx, 2),
                                              input_wanted(ctx, 1))
        return dx, dgam, dg, (a if has_pass else None), None, None, None


class _IrrepsNorm(torch.autograd.Function):
    """passthrough: additionally returns x itself.  A caller that feeds BOTH the norm and a residual from x uses the
    returned alias for the residual: x then has one consumer in the autograd graph, and the two gradient contributions
    are summed inside the norm's backward kernel instead of by a separate elementwise add per tensor and pass."""

    @staticmethod
    def forward(ctx, x, gamma, beta, muls, eps, passthrough):
        x, gamma, beta = _c(x), _c(gamma), _c(beta)
        ctx.save_for_backward(x, gamma)
        ctx.cfg = (muls, eps)
        ctx.set_materialize_grads(False)
        y = irreps_norm_fwd_raw(x, gamma, beta, muls, eps)
        return (y, x.view_as(x)) if passthrough else y

    @staticmethod
    def backward(ctx, g, g_pass=None):
        x, gamma = ctx.saved_tensors
        muls, eps = ctx.cfg
        if g is None:
            return g_pass, None, None, None, None, None
        need_params = input_wanted(ctx, 1) or input_wanted(ctx, 2)  # not in the force pass (d/dpos only)
        gx, gg, gb = _IrrepsNormBwd.apply(x, gamma, g, g_pass, muls, eps, need_params)
        return gx, gg, gb, None, None, None


def irreps_norm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, muls, eps: float = 1e-5) -> torch.Tensor:
    """EquivariantLayerNorm on the cm layout (nn/o3layer.py:145-171)."""
    return _IrrepsNorm.apply(x, gamma, beta, tuple(int(m) for m in muls), float(eps), False)


def layer_norm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """nn.LayerNorm(C) = the irreps norm of "Cx0e"."""
    return _IrrepsNorm.apply(x, weight, bias, (int(x.shape[1]), 0, 0), float(eps), False)


def irreps_norm_pass(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, muls, eps: float = 1e-5):
    """(norm(x), x): use the second output wherever x itself is consumed next to its norm (residuals)."""
    return _IrrepsNorm.apply(x, gamma, beta, tuple(int(m) for m in muls), float(eps), True)


def layer_norm_pass(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5):
    return _IrrepsNorm.apply(x, weight, bias, (int(x.shape[1]), 0, 0), float(eps), True)


# ------------------------------------------------------------------------------------------
# Invariant + EquivariantDot of the update block (nn/xpainn.py:214, 222)
# ------------------------------------------------------------------------------------------
def _e(*shape, like):
    return torch.empty(shape, dtype=torch.float32, device=like.device)


class _InvDotBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, U, W, gn, gt, gU_pass, muls):
        (gn, ld_gn), gt = _rows(gn), _c(gt)  # gn: usually a column slice of the update MLP's input gradient
        ctx.save_for_backward(U, W, gn, gt)
        ctx.muls = muls
        ctx.has_pass = gU_pass is not None
        ctx.set_materialize_grads(False)
        gU, gW = torch.empty_like(U), torch.empty_like(W)
        M = sum(muls)
        _lib.check(_lib.get().xeq_invariant_dot_bwd(_lib.ptr(U), _lib.ptr(W), _lib.ptr_rows(gn), ld_gn or M, _lib.ptr(gt),
                                                    _lib.ptr(_c(gU_pass) if gU_pass is not None else None), U.shape[0], *muls,
                                                    _lib.ptr(gU), _lib.ptr(gW), _lib.stream()), "xeq_invariant_dot_bwd")
        return gU, gW

    @staticmethod
    def backward(ctx, aU, aW):
        U, W, gn, gt = ctx.saved_tensors
        muls = ctx.muls
        if aU is None and aW is None:
            return None, None, None, None, None, None
        N, M = U.shape[0], sum(muls)
        d_gn, d_gt = _e(N, M, like=U), _e(N, M, like=U)
        dU, dW = torch.empty_like(U), torch.empty_like(W)
        _lib.check(_lib.get().xeq_invariant_dot_bwdbwd(_lib.ptr(U), _lib.ptr(W), _lib.ptr_rows(gn), (gn.stride(0) if gn is not None else M), _lib.ptr(gt), _lib.ptr(_c(aU)),
                                                       _lib.ptr(_c(aW)), N, *muls, _lib.ptr(d_gn), _lib.ptr(d_gt), _lib.ptr(dU),
                                                       _lib.ptr(dW), _lib.stream()), "xeq_invariant_dot_bwdbwd")
        return dU, dW, (d_gn if gn is not None else None), (d_gt if gt is not None else None), (aU if ctx.has_pass else None), None


class _InvDot(torch.autograd.Function):
    """passthrough: additionally returns U itself (see _IrrepsNorm): the gate takes U from there, and the two gradient
    contributions to U are summed inside the backward kernel."""

    @staticmethod
    def forward(ctx, U, W, muls, passthrough=False):
        U, W = _c(U), _c(W)
        ctx.save_for_backward(U, W)
        ctx.muls = muls
        ctx.set_materialize_grads(False)
        N, M = U.shape[0], sum(muls)
        nrm, t0 = _e(N, M, like=U), _e(N, M, like=U)
        _lib.check(_lib.get().xeq_invariant_dot_fwd(_lib.ptr(U), _lib.ptr(W), N, *muls, _lib.ptr(nrm), M, _lib.ptr(t0),
                                                    _lib.stream()), "xeq_invariant_dot_fwd")
        return (nrm, t0, U.view_as(U)) if passthrough else (nrm, t0)

    @staticmethod
    def backward(ctx, gn, gt, gU_pass=None):
        U, W = ctx.saved_tensors
        if gn is None and gt is None:
            return gU_pass, None, None, None
        gU, gW = _InvDotBwd.apply(U, W, gn, gt, gU_pass, ctx.muls)
        return gU, gW, None, None


def invariant_dot(U: torch.Tensor, W: torch.Tensor, muls):
    """(Invariant(W), EquivariantDot(U, W)) -> ([N, M], [N, M])."""
    return _InvDot.apply(U, W, tuple(int(m) for m in muls), False)


def invariant_dot_pass(U: torch.Tensor, W: torch.Tensor, muls):
    """(Invariant(W), EquivariantDot(U, W), U): feed U's other consumer (the gate) from the third output."""
    return _InvDot.apply(U, W, tuple(int(m) for m in muls), True)


# ------------------------------------------------------------------------------------------
# gating + residuals of the update block (nn/xpainn.py:218-229)
# ------------------------------------------------------------------------------------------
class _GateBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, U, t, gx, gV, muls):
        gx, gV = _c(gx), _c(gV)
        ctx.save_for_backward(a, U, t, gx, gV)
        ctx.muls = muls
        ctx.set_materialize_grads(False)
        ga, gU, gt = torch.empty_like(a), torch.empty_like(U), torch.empty_like(t)
        _lib.check(_lib.get().xeq_gate_residual_bwd(_lib.ptr(a), _lib.ptr(U), _lib.ptr(t), _lib.ptr(gx), _lib.ptr(gV), a.shape[0],
                                                    *muls, _lib.ptr(ga), _lib.ptr(gU), _lib.ptr(gt), _lib.stream()),
                   "xeq_gate_residual_bwd")
        return ga, gU, gt

    @staticmethod
    def backward(ctx, c_a, c_U, c_t):
        a, U, t, gx, gV = ctx.saved_tensors
        if c_a is None and c_U is None and c_t is None:
            return None, None, None, None, None, None
        d_gx, d_gV = torch.empty_like(t), torch.empty_like(U)
        d_a, d_U, d_t = torch.empty_like(a), torch.empty_like(U), torch.empty_like(t)
        _lib.check(_lib.get().xeq_gate_residual_bwdbwd(_lib.ptr(a), _lib.ptr(U), _lib.ptr(t), _lib.ptr(gx), _lib.ptr(gV),
                                                       _lib.ptr(_c(c_a)), _lib.ptr(_c(c_U)), _lib.ptr(_c(c_t)), a.shape[0],
                                                       *ctx.muls, _lib.ptr(d_gx), _lib.ptr(d_gV), _lib.ptr(d_a), _lib.ptr(d_U),
                                                       _lib.ptr(d_t), _lib.stream()), "xeq_gate_residual_bwdbwd")
        return d_a, d_U, d_t, (d_gx if gx is not None else None), (d_gV if gV is not None else None), None


class _Gate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, U, t, x, V, muls):
        a, U, t, x, V = _c(a), _c(U), _c(t), _c(x), _c(V)
        ctx.save_for_backward(a, U, t)
        ctx.muls = muls
        ctx.set_materialize_grads(False)
        x_out, V_out = torch.empty_like(x), torch.empty_like(V)
        _lib.check(_lib.get().xeq_gate_residual_fwd(_lib.ptr(a), _lib.ptr(U), _lib.ptr(t), _lib.ptr(x), _lib.ptr(V), a.shape[0],
                                                    *muls, _lib.ptr(x_out), _lib.ptr(V_out), _lib.stream()),
                   "xeq_gate_residual_fwd")
        return x_out, V_out

    @staticmethod
    def backward(ctx, gx, gV):
        a, U, t = ctx.saved_tensors
        if gx is None and gV is None:
            return None, None, None, None, None, None
        ga, gU, gt = _GateBwd.apply(a, U, t, gx, gV, ctx.muls)
        return ga, gU, gt, gx, gV, None


def gate_residual(a, U, t, x, V, muls):
    """x + a_sv * t + a_ss,  V + expand(a_vv) * U   with a = [a_vv | a_sv | a_ss]."""
    return _Gate.apply(a, U, t, x, V, tuple(int(m) for m in muls))


# ------------------------------------------------------------------------------------------
# cat([a, b], -1) whose derivatives stay one kernel per order
# ------------------------------------------------------------------------------------------
class _Split2(torch.autograd.Function):
    """(g[:, :n1], g[:, n1:]) as views; the backward of the pair is ONE cat (torch's own slice backward is a zero fill
    plus a copy per slice plus an add)."""

    @staticmethod
    def forward(ctx, g, n1):
        ctx.n = (n1, g.shape[1] - n1)
        ctx.set_materialize_grads(False)
        return g[:, :n1], g[:, n1:]

    @staticmethod
    def backward(ctx, a1, a2):
        if a1 is None and a2 is None:
            return None, None
        ref = a1 if a1 is not None else a2
        if a1 is None:
            a1 = ref.new_zeros((ref.shape[0], ctx.n[0]))
        if a2 is None:
            a2 = ref.new_zeros((ref.shape[0], ctx.n[1]))
        return _Cat2.apply(a1, a2), None


class _Cat2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.n1 = a.shape[1]
        return torch.cat([a, b], dim=1)

    @staticmethod
    def backward(ctx, g):
        g1, g2 = _Split2.apply(g, ctx.n1)
        return g1, g2


def cat2(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """torch.cat([a, b], dim=1) of two 2-D tensors."""
    return _Cat2.apply(a, b)


# ------------------------------------------------------------------------------------------
# SiLU
# ------------------------------------------------------------------------------------------
class _SiluBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, g):
        g = _c(g)
        ctx.save_for_backward(u, g)
        gu = torch.empty_like(u)
        _lib.check(_lib.get().xeq_silu_bwd(_lib.ptr(u), _lib.ptr(g), u.numel(), _lib.ptr(gu), _lib.stream()), "xeq_silu_bwd")
        return gu

    @staticmethod
    def backward(ctx, c):
        u, g = ctx.saved_tensors
        c = _c(c)
        du = torch.empty_like(u) if ctx.needs_input_grad[0] else None
        dg = torch.empty_like(u) if ctx.needs_input_grad[1] else None
        _lib.check(_lib.get().xeq_silu_bwdbwd(_lib.ptr(u), _lib.ptr(g), _lib.ptr(c), u.numel(), _lib.ptr(dg), _lib.ptr(du),
                                              _lib.stream()), "xeq_silu_bwdbwd")
        return du, dg


class _Silu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u):
        u = _c(u)
        ctx.save_for_backward(u)
        y = torch.empty_like(u)
        _lib.check(_lib.get().xeq_silu_fwd(_lib.ptr(u), u.numel(), _lib.ptr(y), _lib.stream()), "xeq_silu_fwd")
        return y

    @staticmethod
    def backward(ctx, g):
        (u,) = ctx.saved_tensors
        return _SiluBwd.apply(u, g)


def silu(u: torch.Tensor) -> torch.Tensor:
    return _Silu.apply(u)
