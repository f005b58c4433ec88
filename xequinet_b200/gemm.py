"""K3: node-side dense contractions on the tcgen05 tensor cores (csrc/node_gemm.cu, C ABI
`xeq_gemm_tf32x3`), wrapped so that they are closed under differentiation:

  mm(A, B, ta, tb)              alpha * op(A) @ op(B) (+ bias)   -> backward = two more mm calls
  linear(x, W, b)               nn.Linear (nn/xpainn.py:111-115, 190, 195-199)
  irreps_linear(V, w, b, muls)  e3nn o3.Linear on the cm layout (nn/xpainn.py:186-187): one grouped
                                launch, one problem per (l, m); backward = irreps_linear with the
                                blocks transposed + irreps_wgrad (split-K over the nodes)

so `torch.autograd.grad(E, pos, create_graph=True)` + `loss.backward()` stay on these kernels to any
order.  No eager fallback: host tensors raise."""
from __future__ import annotations

import ctypes
import math
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._state import input_wanted

_TARGET_CTAS = 148


def _operand(t: torch.Tensor) -> torch.Tensor:
    """A 2-D fp32 CUDA operand the kernel can address: unit inner stride, 16-byte aligned rows."""
    if t.dtype != torch.float32:
        raise RuntimeError(f"xequinet_b200 kernels compute in fp32, got {t.dtype}")
    if not t.is_cuda:
        raise RuntimeError("xequinet_b200 ops need CUDA tensors: there is no CPU fallback")
    ok = t.dim() == 2 and t.stride(1) == 1 and t.stride(0) % 4 == 0 and t.stride(0) >= t.shape[1] and t.data_ptr() % 16 == 0
    return t if ok else t.contiguous()


class Problem:
    """One member of a grouped launch (mirrors xeq_gemm_t)."""

    __slots__ = ("a", "b", "bias", "c", "m", "n", "k", "lda", "ldb", "ldc", "ta", "tb", "alpha", "act")

    def __init__(self, a, b, c, m, n, k, lda, ldb, ldc, ta, tb, alpha=1.0, bias=0, act=0):
        self.a, self.b, self.c, self.bias = a, b, c, bias
        self.m, self.n, self.k, self.lda, self.ldb, self.ldc = m, n, k, lda, ldb, ldc
        self.ta, self.tb, self.alpha, self.act = ta, tb, alpha, act


def launch(problems: Sequence[Problem], split_k: int = 1, device=None) -> None:
    lib = _lib.get()
    n = len(problems)
    arr = (_lib.XeqGemm * n)()
    for i, p in enumerate(problems):
        g = arr[i]
        g.a, g.b, g.c, g.bias = p.a, p.b, p.c, (p.bias or None)
        g.m, g.n, g.k, g.lda, g.ldb, g.ldc = p.m, p.n, p.k, p.lda, p.ldb, p.ldc
        g.a_trans, g.b_trans, g.alpha, g.act = int(p.ta), int(p.tb), float(p.alpha), int(p.act)
    nbytes = lib.xeq_gemm_workspace_bytes(arr, n, split_k)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=device) if nbytes else None
    _lib.check(lib.xeq_gemm_tf32x3(arr, n, split_k, _lib.ptr(ws), nbytes, _lib.stream()), "xeq_gemm_tf32x3")


def _split_for(ctas: int, k: int) -> int:
    """split-K factor for reductions over the nodes: fill the SMs, at least 4 K-blocks per split."""
    nkb = (k + 31) // 32
    return max(1, min(_TARGET_CTAS // max(ctas, 1), nkb // 4, 256))


def _ctas(m: int, n: int) -> int:
    n16 = (n + 15) // 16 * 16
    return ((m + 127) // 128) * ((n16 + 255) // 256)


def mm_raw(A: torch.Tensor, B: torch.Tensor, ta: bool, tb: bool, bias: Optional[torch.Tensor] = None,
           alpha: float = 1.0, act: int = 0) -> torch.Tensor:
    A, B = _operand(A), _operand(B)
    m, k = (A.shape[1], A.shape[0]) if ta else (A.shape[0], A.shape[1])
    kb, n = (B.shape[1], B.shape[0]) if tb else (B.shape[0], B.shape[1])
    if k != kb:
        raise RuntimeError(f"mm: inner dimensions differ ({k} vs {kb})")
    C = torch.empty((m, n), dtype=torch.float32, device=A.device)
    if m == 0:
        return C
    if k == 0:
        C.zero_()
        return C if bias is None else C + bias
    bias_c = bias.contiguous() if bias is not None else None
    split = _split_for(_ctas(m, n), k) if (ta and act == 0) else 1
    launch([Problem(A.data_ptr(), B.data_ptr(), C.data_ptr(), m, n, k, A.stride(0), B.stride(0), n, ta, tb, alpha,
                    bias_c.data_ptr() if bias_c is not None else 0, act)], split, A.device)
    return C


def colsum_raw(g: torch.Tensor) -> torch.Tensor:
    """g.sum(0) of a 2-D fp32 CUDA tensor with unit inner stride (a row-strided view is fine)."""
    if g.dtype != torch.float32 or not g.is_cuda or g.dim() != 2:
        raise RuntimeError("colsum: 2-D fp32 CUDA tensor expected (there is no CPU fallback)")
    if g.stride(1) != 1 or g.stride(0) < g.shape[1]:
        g = g.contiguous()
    out = torch.empty(g.shape[1], dtype=torch.float32, device=g.device)
    lib = _lib.get()
    src = ctypes.c_void_p(g.data_ptr()) if g.numel() else None  # row-strided views are addressed through ld
    _lib.check(lib.xeq_colsum(src, g.shape[0], g.shape[1], max(g.stride(0), g.shape[1]), _lib.ptr(out), _lib.stream()), "xeq_colsum")
    return out


class _ColSum(torch.autograd.Function):
    """Bias gradient g.sum(0) as one deterministic kernel; its own derivative is a broadcast."""

    @staticmethod
    def forward(ctx, g):
        ctx.n_rows = g.shape[0]
        return colsum_raw(g)

    @staticmethod
    def backward(ctx, gout):
        return gout.unsqueeze(0).expand(ctx.n_rows, -1)


def colsum(g: torch.Tensor) -> torch.Tensor:
    return _ColSum.apply(g)


class _MM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A, B, bias, ta, tb, alpha):
        ctx.save_for_backward(A, B)
        ctx.cfg = (ta, tb, alpha)
        return mm_raw(A, B, ta, tb, bias, alpha)

    @staticmethod
    def backward(ctx, gC):
        A, B = ctx.saved_tensors
        ta, tb, alpha = ctx.cfg
        gA = gB = gbias = None
        # only what this graph task asks for: the force pass (d/dpos) skips every weight / bias gradient
        if input_wanted(ctx, 0):
            gA = mm(gC, B, False, not tb, alpha=alpha) if not ta else mm(B, gC, tb, True, alpha=alpha)
        if input_wanted(ctx, 1):
            gB = mm(A, gC, not ta, False, alpha=alpha) if not tb else mm(gC, A, True, ta, alpha=alpha)
        if input_wanted(ctx, 2):
            gbias = colsum(gC)
        return gA, gB, gbias, None, None, None


def mm(A: torch.Tensor, B: torch.Tensor, ta: bool = False, tb: bool = False, bias: Optional[torch.Tensor] = None,
       alpha: float = 1.0) -> torch.Tensor:
    """alpha * op(A) @ op(B) + bias, differentiable to any order (every derivative is again mm)."""
    return _MM.apply(A, B, bias, bool(ta), bool(tb), float(alpha))


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """F.linear(x, weight, bias) for x [N, in], weight [out, in]."""
    return mm(x, weight, False, True, bias)


# ------------------------------------------------------------------------------------------
# nn.Linear with ONE output feature (the 64 -> 1 energy read-out, nn/output.py:107-111): a GEMM tile would be
# 127/128 padding.  Three small kernels, closed under differentiation:
#   rowdot(x, w, b)  y[r] = <x[r], w> + b        d/dx = outer(gy, w)      d/dw = wsum(gy, x)   d/db = colsum(gy)
#   outer(g, w)      o[r, c] = g[r] w[c]         d/dg = rowdot(go, w)     d/dw = wsum(g, go)
#   wsum(g, x)       o[c] = sum_r g[r] x[r, c]   d/dg = rowdot(x, go)     d/dx = outer(g, go)
# ------------------------------------------------------------------------------------------
def _rows2d(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32 or not t.is_cuda or t.dim() != 2:
        raise RuntimeError("xequinet_b200 ops need 2-D fp32 CUDA tensors here: there is no CPU fallback")
    return t if (t.stride(1) == 1 and t.stride(0) >= t.shape[1]) else t.contiguous()


def _vec(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32 or not t.is_cuda:
        raise RuntimeError("xequinet_b200 ops need fp32 CUDA tensors: there is no CPU fallback")
    return t.reshape(-1).contiguous()


def rowdot_raw(x, w, bias=None):
    x, w = _rows2d(x), _vec(w)
    y = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    b = _vec(bias) if bias is not None else None
    _lib.check(_lib.get().xeq_rowdot(x.data_ptr() if x.numel() else None, max(x.stride(0), x.shape[1]), _lib.ptr(w), _lib.ptr(b),
                                     x.shape[0], x.shape[1], _lib.ptr(y), _lib.stream()), "xeq_rowdot")
    return y


def outer_raw(g, w):
    g, w = _vec(g), _vec(w)
    out = torch.empty((g.numel(), w.numel()), dtype=torch.float32, device=g.device)
    _lib.check(_lib.get().xeq_outer(_lib.ptr(g), _lib.ptr(w), g.numel(), w.numel(), _lib.ptr(out), _lib.stream()), "xeq_outer")
    return out


def wsum_raw(g, x):
    g, x = _vec(g), _rows2d(x)
    out = torch.empty(x.shape[1], dtype=torch.float32, device=x.device)
    _lib.check(_lib.get().xeq_colsum_weighted(x.data_ptr() if x.numel() else None, _lib.ptr(g), x.shape[0], x.shape[1],
                                              max(x.stride(0), x.shape[1]), _lib.ptr(out), _lib.stream()), "xeq_colsum_weighted")
    return out


class _RowDot(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias):
        ctx.save_for_backward(x, w)
        return rowdot_raw(x, w, bias)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gx = _Outer.apply(gy, w) if input_wanted(ctx, 0) else None
        gw = _WSum.apply(gy, x).view_as(w) if input_wanted(ctx, 1) else None
        gb = colsum(gy.reshape(-1, 1)) if input_wanted(ctx, 2) else None
        return gx, gw, gb


class _Outer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g, w):
        ctx.save_for_backward(g, w)
        return outer_raw(g, w)

    @staticmethod
    def backward(ctx, go):
        g, w = ctx.saved_tensors
        gg = _RowDot.apply(go, w, None).view_as(g) if input_wanted(ctx, 0) else None
        gw = _WSum.apply(g, go).view_as(w) if input_wanted(ctx, 1) else None
        return gg, gw


class _WSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g, x):
        ctx.save_for_backward(g, x)
        return wsum_raw(g, x)

    @staticmethod
    def backward(ctx, go):
        g, x = ctx.saved_tensors
        gg = _RowDot.apply(x, go, None).view_as(g) if input_wanted(ctx, 0) else None
        gx = _Outer.apply(g, go) if input_wanted(ctx, 1) else None
        return gg, gx


def linear_to_scalar(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """F.linear(x, weight, bias) for weight [1, in] -> [N, 1], differentiable to any order."""
    return _RowDot.apply(x, weight, bias).unsqueeze(1)


# ------------------------------------------------------------------------------------------
# o3.Linear on the cm layout [mul0 | 3 x mul1 | 5 x mul2]
# ------------------------------------------------------------------------------------------
def _blocks(muls):
    """(l, mul, feature offset of the l block, flat weight offset)"""
    out, foff, woff = [], 0, 0
    for l, mul in enumerate(muls):
        if mul:
            out.append((l, mul, foff, woff))
        foff += (2 * l + 1) * mul
        woff += mul * mul
    return out


def irreps_linear_raw(V: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], muls, transposed: bool) -> torch.Tensor:
    """out[:, (l, m, w)] = sum_u V[:, (l, m, u)] W_l[u, w] / sqrt(mul_l)  (transposed: W_l[w, u]); bias on l = 0."""
    V, w = _operand(V), w.contiguous()
    N, D = V.shape
    out = torch.empty((N, D), dtype=torch.float32, device=V.device)
    if N == 0:
        return out
    bias_c = bias.contiguous() if (bias is not None and bias.numel()) else None
    ld = V.stride(0)
    probs: List[Problem] = []
    for l, mul, foff, woff in _blocks(muls):
        for m in range(2 * l + 1):
            off = foff + m * mul
            probs.append(Problem(V.data_ptr() + 4 * off, w.data_ptr() + 4 * woff, out.data_ptr() + 4 * off, N, mul, mul, ld,
                                 mul, D, False, transposed, 1.0 / math.sqrt(mul),
                                 bias_c.data_ptr() if (l == 0 and bias_c is not None) else 0))
    launch(probs, 1, V.device)
    return out


def irreps_wgrad_raw(A: torch.Tensor, B: torch.Tensor, muls) -> torch.Tensor:
    """flat o3.Linear-shaped tensor: G_l[u, w] = sum_{n, m} A[n, (l, m, u)] B[n, (l, m, w)] / sqrt(mul_l)."""
    A, B = _operand(A), _operand(B)
    N = A.shape[0]
    blocks = _blocks(muls)
    total = sum(mul * mul for _, mul, _, _ in blocks)
    if N == 0:
        return torch.zeros(total, dtype=torch.float32, device=A.device)
    # one problem per (l, m); the 2l+1 problems of an l name the same output block, so the library
    # sums their partial slabs (together with the split-K slabs) in one fixed-order reduction
    out = torch.empty(total, dtype=torch.float32, device=A.device)
    probs, ctas = [], 0
    for l, mul, foff, woff in blocks:
        for m in range(2 * l + 1):
            off = foff + m * mul
            probs.append(Problem(A.data_ptr() + 4 * off, B.data_ptr() + 4 * off, out.data_ptr() + 4 * woff, mul, mul, N,
                                 A.stride(0), B.stride(0), mul, True, False, 1.0 / math.sqrt(mul)))
            ctas += _ctas(mul, mul)
    launch(probs, _split_for(ctas, N), A.device)
    return out


class _IrrepsLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, V, w, bias, muls, transposed):
        ctx.save_for_backward(V, w)
        ctx.cfg = (muls, transposed)
        return irreps_linear_raw(V, w, bias, muls, transposed)

    @staticmethod
    def backward(ctx, g):
        V, w = ctx.saved_tensors
        muls, transposed = ctx.cfg
        gV = gw = gb = None
        if input_wanted(ctx, 0):
            gV = _IrrepsLinear.apply(g, w, None, muls, not transposed)
        if input_wanted(ctx, 1):
            gw = _IrrepsWgrad.apply(V, g, muls) if not transposed else _IrrepsWgrad.apply(g, V, muls)
        if input_wanted(ctx, 2):
            gb = colsum(g[:, : muls[0]])
        return gV, gw, gb, None, None


class _IrrepsWgrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A, B, muls):
        ctx.save_for_backward(A, B)
        ctx.muls = muls
        return irreps_wgrad_raw(A, B, muls)

    @staticmethod
    def backward(ctx, gw):
        A, B = ctx.saved_tensors
        muls = ctx.muls
        gA = _IrrepsLinear.apply(B, gw, None, muls, True) if input_wanted(ctx, 0) else None
        gB = _IrrepsLinear.apply(A, gw, None, muls, False) if input_wanted(ctx, 1) else None
        return gA, gB, None


def irreps_linear(V: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], muls) -> torch.Tensor:
    return _IrrepsLinear.apply(V, weight, bias, tuple(muls), False)
