"""GPU parity of the node-side kernels through the C ABI (ctypes): K3 tcgen05 3xTF32 GEMMs (csrc/node_gemm.cu)
and the fused norm / invariant+dot / gate+residual / SiLU kernels (csrc/node_norm.cu, csrc/node_update.cu),
values and first + second derivatives against plain torch in fp64 (the reference computes these with
nn.Linear, e3nn o3.Linear, nn.LayerNorm, EquivariantLayerNorm, Invariant, EquivariantDot: nn/xpainn.py:111-131,
186-231, nn/o3layer.py:11-171).  Tolerances are fp32-level (the 3xTF32 split keeps ~2^-22 per product)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    return float((a.detach().double().cpu() - b.detach().cpu()).abs().max() / (b.detach().abs().max() + 1e-30))


@pytest.mark.parametrize("mnk", [(128, 128, 128), (5376, 576, 128), (1000, 480, 128), (777, 128, 352), (300, 128, 224),
                                 (4608, 128, 56), (64, 16, 32), (5, 20, 8), (1, 64, 64), (3000, 256, 448)])
def test_gemm_matches_fp64(mnk):
    from xequinet_b200 import gemm

    m, n, k = mnk
    g = torch.Generator().manual_seed(m + n + k)
    for ta in (False, True):
        for tb in (False, True):
            if (ta and m % 4) or (not ta and k % 4) or (tb and k % 4) or (not tb and n % 4):
                continue
            A = torch.randn((k, m) if ta else (m, k), generator=g).to(DEV)
            B = torch.randn((n, k) if tb else (k, n), generator=g).to(DEV)
            bias = torch.randn(n, generator=g).to(DEV)
            C = gemm.mm_raw(A, B, ta, tb, bias, 0.5)
            ref = 0.5 * ((A.double().T if ta else A.double()) @ (B.double().T if tb else B.double())) + bias.double()
            assert _rel(C, ref) < 2e-6, (mnk, ta, tb)
            assert torch.equal(C, gemm.mm_raw(A, B, ta, tb, bias, 0.5)), "not deterministic"


def test_gemm_silu_epilogue_and_strided_operands():
    from xequinet_b200 import gemm

    A = torch.randn(1000, 256, device=DEV)
    W = torch.randn(128, 128, device=DEV) / 11
    b = torch.randn(128, device=DEV)
    C = gemm.mm_raw(A[:, 64:192], W, False, True, b, 1.0, act=1)  # row-strided view of A, no copy needed
    ref = torch.nn.functional.silu(A[:, 64:192].double() @ W.double().T + b.double())
    assert _rel(C, ref) < 2e-6


def test_linear_double_backward():
    from xequinet_b200 import gemm

    x = torch.randn(3000, 128, device=DEV, requires_grad=True)
    W = (torch.randn(576, 128, device=DEV) / 11).requires_grad_()
    b = torch.randn(576, device=DEV, requires_grad=True)

    def run(f, x, W, b):
        y = f(x, W, b)
        (gx,) = torch.autograd.grad((y ** 3).sum(), x, create_graph=True)
        return y, gx, torch.autograd.grad((gx ** 2).sum() + y.sum(), [x, W, b])

    y1, gx1, G1 = run(gemm.linear, x, W, b)
    y2, gx2, G2 = run(torch.nn.functional.linear, *(t.detach().double().requires_grad_() for t in (x, W, b)))
    assert _rel(y1, y2) < 2e-6 and _rel(gx1, gx2) < 1e-5
    for a, r in zip(G1, G2):
        assert _rel(a, r) < 2e-5


def test_irreps_linear_double_backward():
    from xequinet_b200 import gemm
    from xequinet_b200.nn import cm

    muls = (128, 64, 32)
    V = torch.randn(2000, 480, device=DEV, requires_grad=True)
    w = torch.randn(128 * 128 + 64 * 64 + 32 * 32, device=DEV, requires_grad=True)
    bb = torch.randn(128, device=DEV, requires_grad=True)

    def ref_lin(V, w, bb):
        m0, m1, m2 = muls
        W0, W1, W2 = w[:m0 * m0].view(m0, m0), w[m0 * m0:m0 * m0 + m1 * m1].view(m1, m1), w[m0 * m0 + m1 * m1:].view(m2, m2)
        v0, v1, v2 = cm.split(V, muls)
        return cm.join(v0 @ W0 / math.sqrt(m0) + bb, v1 @ W1 / math.sqrt(m1), v2 @ W2 / math.sqrt(m2))

    def run(f, V, w, bb):
        out = f(V, w, bb)
        (gV,) = torch.autograd.grad((out ** 3).sum(), V, create_graph=True)
        return out, gV, torch.autograd.grad((gV ** 2).sum() + out.sum(), [V, w, bb])

    o1, g1, G1 = run(lambda V, w, b: gemm.irreps_linear(V, w, b, muls), V, w, bb)
    o2, g2, G2 = run(ref_lin, *(t.detach().double().requires_grad_() for t in (V, w, bb)))
    assert _rel(o1, o2) < 2e-6 and _rel(g1, g2) < 1e-5
    for a, r in zip(G1, G2):
        assert _rel(a, r) < 2e-5


def _check(f_new, f_ref, inputs, n_diff=None, tol=5e-5):
    """outputs, first derivatives and the gradient of a loss on the first derivatives, fp32 kernel vs fp64 torch"""
    x32 = [t.detach().clone().float().to(DEV).requires_grad_() for t in inputs]
    x64 = [t.detach().clone().double().requires_grad_() for t in inputs]
    res = []
    for f, xs in ((f_new, x32), (f_ref, x64)):
        outs = f(*xs)
        outs = outs if isinstance(outs, tuple) else (outs,)
        E = sum((o ** 3).sum() + (o * o).sum() for o in outs)
        g1 = torch.autograd.grad(E, xs[:n_diff] if n_diff else xs, create_graph=True)
        L = sum((g ** 2).sum() for g in g1) + sum(o.sum() for o in outs)
        res.append((outs, g1, torch.autograd.grad(L, xs, allow_unused=True)))
    (o1, a1, b1), (o2, a2, b2) = res
    for p, q in list(zip(o1, o2)) + list(zip(a1, a2)) + [(p, q) for p, q in zip(b1, b2) if q is not None]:
        assert _rel(p, q) < tol


@pytest.mark.parametrize("muls", [(128, 64, 32), (256, 128, 64), (128, 0, 0), (32, 32, 32)])
def test_norm_invariant_gate_kernels(muls):
    from xequinet_b200 import nodeops
    from xequinet_b200.nn import cm

    m0, m1, m2 = muls
    D, M, N = m0 + 3 * m1 + 5 * m2, m0 + m1 + m2, 301
    g = torch.Generator().manual_seed(D)
    r = lambda *s: torch.randn(*s, generator=g)
    full = bool(m1 or m2)

    def ref_norm(V, gam, bet):
        scal = V[:, :m0]
        z = torch.cat([scal - scal.mean(1, keepdim=True), V[:, m0:]], 1)
        q = cm.irrep_dot(z, z, muls) if full else z * z
        rho = 1 / torch.sqrt(q.mean(1, keepdim=True) + 1e-5)
        out = z * rho * (cm.expand_gate(gam.unsqueeze(0), muls) if full else gam.unsqueeze(0))
        return torch.cat([out[:, :m0] + bet, out[:, m0:]], 1)

    _check(lambda V, ga, be: nodeops.irreps_norm(V, ga, be, muls), ref_norm, [r(N, D), r(M), r(m0)], n_diff=1)
    if not full:
        _check(lambda x, w, b: nodeops.layer_norm(x, w, b),
               lambda x, w, b: torch.nn.functional.layer_norm(x, (m0,), w, b, 1e-5), [r(N, m0), r(m0), r(m0)], n_diff=1)
        return
    _check(lambda U, W: nodeops.invariant_dot(U, W, muls),
           lambda U, W: (torch.sqrt(cm.irrep_dot(W, W, muls) + 1e-10) - 1e-5, cm.irrep_dot(U, W, muls)), [r(N, D), r(N, D)])
    _check(lambda a, U, t, x, V: nodeops.gate_residual(a, U, t, x, V, muls),
           lambda a, U, t, x, V: (x + a[:, M:M + m0] * t + a[:, M + m0:], V + U * cm.expand_gate(a[:, :M], muls)),
           [r(N, M + 2 * m0), r(N, D), r(N, m0), r(N, m0), r(N, D)])


def test_silu_kernels():
    from xequinet_b200 import nodeops

    _check(nodeops.silu, torch.nn.functional.silu, [torch.randn(1000, 128, generator=torch.Generator().manual_seed(0)) * 2])


def test_norm_parameter_gradients():
    """d/d(gamma, beta) of the fused norm (per-CTA partials + fixed-order reduction) against torch autograd."""
    from xequinet_b200 import nodeops
    from xequinet_b200.nn import cm

    muls = (128, 64, 32)
    g = torch.Generator().manual_seed(3)
    V, gam, bet = torch.randn(5000, 480, generator=g), torch.randn(224, generator=g), torch.randn(128, generator=g)
    Vc, gc, bc = (t.to(DEV).requires_grad_() for t in (V, gam, bet))
    out = nodeops.irreps_norm(Vc, gc, bc, muls)
    G1 = torch.autograd.grad((out ** 3).sum(), [gc, bc])
    Vd, gd, bd = (t.double().requires_grad_() for t in (V, gam, bet))
    scal = Vd[:, :128]
    z = torch.cat([scal - scal.mean(1, keepdim=True), Vd[:, 128:]], 1)
    rho = 1 / torch.sqrt(cm.irrep_dot(z, z, muls).mean(1, keepdim=True) + 1e-5)
    o = z * rho * cm.expand_gate(gd.unsqueeze(0), muls)
    o = torch.cat([o[:, :128] + bd, o[:, 128:]], 1)
    G2 = torch.autograd.grad((o ** 3).sum(), [gd, bd])
    for a, r in zip(G1, G2):
        assert _rel(a, r) < 2e-5
    assert torch.equal(G1[0], torch.autograd.grad((nodeops.irreps_norm(Vc, gc, bc, muls) ** 3).sum(), [gc])[0]), "not deterministic"


@pytest.mark.parametrize("shape", [(5376, 576), (1000, 130), (3, 7), (0, 64), (4097, 128)])
def test_colsum_matches_torch(shape):
    """Bias-gradient column sum (xeq_colsum): values, row-strided views, determinism, broadcast derivative."""
    from xequinet_b200 import gemm

    g = torch.randn(shape, device=DEV)
    out = gemm.colsum_raw(g)
    ref = g.double().sum(0)
    assert float((out.double() - ref).abs().max() if shape[1] else 0.0) <= 1e-5 * max(1.0, float(ref.abs().max()) if shape[0] else 1.0)
    assert torch.equal(out, gemm.colsum_raw(g))
    if shape[1] >= 8 and shape[0] > 0:
        view = g[:, 2 : shape[1] - 3]
        assert torch.allclose(gemm.colsum_raw(view), view.sum(0), rtol=1e-5, atol=1e-4)
        x = g.clone().requires_grad_(True)
        w = torch.randn(shape[1], device=DEV)
        (gemm.colsum(x) * w).sum().backward()
        assert torch.allclose(x.grad, w.expand_as(x))


def test_linear_to_scalar_kernels_match_torch():
    """xeq_rowdot / xeq_outer / xeq_colsum_weighted (the one-output Linear = energy read-out, nn/output.py:107-111)
    against fp64 torch: values, first and second derivatives through the registered formulas; row-strided input."""
    from xequinet_b200 import gemm

    gen = torch.Generator().manual_seed(11)
    for n, k in ((5376, 64), (1, 64), (333, 20), (7, 128)):
        big = torch.randn(n, k + 8, generator=gen, dtype=torch.float64)
        x0, w0, b0 = big[:, 4:4 + k], torch.randn(1, k, generator=gen, dtype=torch.float64), torch.randn(1, generator=gen, dtype=torch.float64)
        r1, r2 = torch.randn(n, 1, generator=gen, dtype=torch.float64), torch.randn(n, k, generator=gen, dtype=torch.float64)

        def run(fn, dt, dev):
            xs = big.to(dev, dt).requires_grad_(True)
            w, b = (t.to(dev, dt).clone().requires_grad_(True) for t in (w0, b0))
            x = xs[:, 4:4 + k]  # a row-strided view: the kernels take the row stride
            y = fn(x, w, b)
            (gx,) = torch.autograd.grad((y * r1.to(dev, dt)).sum(), xs, create_graph=True)
            loss = (y ** 2).sum() + (gx[:, 4:4 + k] * r2.to(dev, dt)).sum() + (gx ** 2).sum()
            return [y.detach(), gx.detach(), *torch.autograd.grad(loss, (xs, w, b))]

        ref = run(torch.nn.functional.linear, torch.float64, "cpu")
        got = run(gemm.linear_to_scalar, torch.float32, DEV)
        for name, a, r in zip(("y", "gx", "dx", "dw", "db"), got, ref):
            scale = float(r.abs().max()) + 1e-30
            assert float((a.double().cpu() - r).abs().max()) <= 2e-5 * scale, (n, k, name)
    # bitwise reproducible
    x = torch.randn(4096, 64, generator=gen).to(DEV)
    g = torch.randn(4096, generator=gen).to(DEV)
    assert torch.equal(gemm.wsum_raw(g, x), gemm.wsum_raw(g, x))
