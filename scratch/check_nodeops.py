import sys; sys.path.insert(0, ".")
import torch, math
from xequinet_b200 import nodeops
from xequinet_b200.nn import cm
torch.manual_seed(0)
dev = "cuda"
def rel(a, b): return float((a.detach().double().cpu() - b.detach().cpu()).abs().max() / (b.detach().abs().max() + 1e-30))
def check(name, f_new, f_ref, inputs, n_out, n_diff=None):
    """compare outputs, first grads of sum(out^3-ish), and grads of a second-order loss"""
    x32 = [t.detach().clone().float().to(dev).requires_grad_() for t in inputs]
    x64 = [t.detach().clone().double().requires_grad_() for t in inputs]
    res = []
    for f, xs in ((f_new, x32), (f_ref, x64)):
        outs = f(*xs)
        if not isinstance(outs, tuple): outs = (outs,)
        E = sum((o ** 3).sum() + (o * o).sum() for o in outs)
        g1 = torch.autograd.grad(E, xs[:n_diff] if n_diff else xs, create_graph=True)
        L = sum((g ** 2).sum() for g in g1) + sum(o.sum() for o in outs)
        g2 = torch.autograd.grad(L, xs, allow_unused=True)
        res.append((outs, g1, g2))
    (o1, a1, b1), (o2, a2, b2) = res
    e_o = max(rel(p, q) for p, q in zip(o1, o2)); e_1 = max(rel(p, q) for p, q in zip(a1, a2)); e_2 = max(rel(p, q) for p, q in zip(b1, b2) if q is not None)
    flag = "OK" if max(e_o, e_1, e_2) < 2e-4 else "BAD"
    print(f"{name}: out {e_o:.2e} grad {e_1:.2e} gradgrad {e_2:.2e} {flag}", flush=True)
    return flag == "OK"
ok = True
for muls in [(128, 64, 32), (256, 128, 64), (128, 0, 0), (32, 32, 32)]:
    m0, m1, m2 = muls; D = m0 + 3*m1 + 5*m2; M = m0 + m1 + m2; N = 301
    def ref_norm(V, gam, bet, muls=muls, m0=m0):
        scal = V[:, :m0]
        z = torch.cat([scal - scal.mean(1, keepdim=True), V[:, m0:]], 1)
        q = cm.irrep_dot(z, z, muls) if muls[1] or muls[2] else z * z
        rho = 1 / torch.sqrt(q.mean(1, keepdim=True) + 1e-5)
        g = cm.expand_gate(gam.unsqueeze(0), muls) if muls[1] or muls[2] else gam.unsqueeze(0)
        out = z * rho * g
        return torch.cat([out[:, :m0] + bet, out[:, m0:]], 1)
    ok &= check(f"norm{muls}", lambda V, g, b: nodeops.irreps_norm(V, g, b, muls), ref_norm,
                [torch.randn(N, D), torch.randn(M), torch.randn(m0)], 1, n_diff=1)
    if m1 == 0: continue
    def ref_invdot(U, W, muls=muls):
        return torch.sqrt(cm.irrep_dot(W, W, muls) + 1e-10) - 1e-5, cm.irrep_dot(U, W, muls)
    ok &= check(f"invdot{muls}", lambda U, W: nodeops.invariant_dot(U, W, muls), ref_invdot, [torch.randn(N, D), torch.randn(N, D)], 2)
    def ref_gate(a, U, t, x, V, muls=muls, M=M, C=m0):
        return x + a[:, M:M+C] * t + a[:, M+C:], V + U * cm.expand_gate(a[:, :M], muls)
    ok &= check(f"gate{muls}", lambda a, U, t, x, V: nodeops.gate_residual(a, U, t, x, V, muls), ref_gate,
                [torch.randn(N, M + 2*m0), torch.randn(N, D), torch.randn(N, m0), torch.randn(N, m0), torch.randn(N, D)], 2)
ok &= check("layer_norm", lambda x, w, b: nodeops.layer_norm(x, w, b), lambda x, w, b: torch.nn.functional.layer_norm(x, (128,), w, b, 1e-5),
            [torch.randn(500, 128), torch.randn(128), torch.randn(128)], 1, n_diff=1)
ok &= check("silu", nodeops.silu, torch.nn.functional.silu, [torch.randn(1000, 128) * 2], 1)
print("ALL OK" if ok else "FAILED")
