import sys; sys.path.insert(0, ".")
import torch, math
from xequinet_b200 import gemm
torch.manual_seed(0)
dev = "cuda"
def rel(a, b): return float((a.double() - b).abs().max() / b.abs().max())
ok = True
for (m, n, k) in [(128, 128, 128), (5376, 576, 128), (1000, 480, 128), (777, 128, 352), (300, 128, 224), (4608, 128, 56), (64, 16, 32), (5, 20, 8), (5376, 64, 64)]:
    for ta in (False, True):
        for tb in (False, True):
            if ta and m % 4: continue
            if not ta and k % 4: continue
            if tb and k % 4: continue
            if not tb and n % 4: continue
            A = torch.randn((k, m) if ta else (m, k), device=dev)
            B = torch.randn((n, k) if tb else (k, n), device=dev)
            bias = torch.randn(n, device=dev)
            C = gemm.mm_raw(A, B, ta, tb, bias, 0.5)
            ref = 0.5 * ((A.double().T if ta else A.double()) @ (B.double().T if tb else B.double())) + bias.double()
            e = rel(C, ref)
            e32 = rel(0.5 * ((A.T if ta else A) @ (B.T if tb else B)) + bias, ref)
            flag = "OK" if e < 2e-6 else "BAD"
            if flag == "BAD": ok = False
            print(f"m{m} n{n} k{k} ta{int(ta)} tb{int(tb)}: rel err {e:.2e} (torch fp32 {e32:.2e}) {flag}", flush=True)
# SiLU epilogue
A = torch.randn(1000, 128, device=dev); W = torch.randn(128, 128, device=dev) / 11; b = torch.randn(128, device=dev)
C = gemm.mm_raw(A, W, False, True, b, 1.0, act=1)
ref = torch.nn.functional.silu(A.double() @ W.double().T + b.double())
print("silu", rel(C, ref))
# irreps linear + grads (double backward) vs torch
from xequinet_b200.nn import cm
muls = (128, 64, 32); D = 480; N = 2000
V = torch.randn(N, D, device=dev, requires_grad=True); w = torch.randn(128*128+64*64+32*32, device=dev, requires_grad=True); bb = torch.randn(128, device=dev, requires_grad=True)
def ref_lin(V, w, bb):
    m0, m1, m2 = muls
    W0, W1, W2 = w[:m0*m0].view(m0, m0), w[m0*m0:m0*m0+m1*m1].view(m1, m1), w[m0*m0+m1*m1:].view(m2, m2)
    v0, v1, v2 = cm.split(V, muls)
    return cm.join(v0 @ W0 / math.sqrt(m0) + bb, v1 @ W1 / math.sqrt(m1), v2 @ W2 / math.sqrt(m2))
def loss2(f, V, w, bb):
    out = f(V, w, bb)
    E = (out ** 3).sum()
    gV, = torch.autograd.grad(E, V, create_graph=True)
    L = (gV ** 2).sum() + out.sum()
    return out, gV, torch.autograd.grad(L, [V, w, bb])
o1, g1, gg1 = loss2(lambda V, w, b: gemm.irreps_linear(V, w, b, muls), V, w, bb)
Vd, wd, bd = (t.detach().double().requires_grad_() for t in (V, w, bb))
o2, g2, gg2 = loss2(ref_lin, Vd, wd, bd)
print("irreps fwd", rel(o1, o2), "gV", rel(g1, g2), "ggV", rel(gg1[0], gg2[0]), "ggw", rel(gg1[1], gg2[1]), "ggb", rel(gg1[2], gg2[2]))
# linear double backward
x = torch.randn(3000, 128, device=dev, requires_grad=True); W = (torch.randn(576, 128, device=dev) / 11).requires_grad_(); b = torch.randn(576, device=dev, requires_grad=True)
def loss3(f, x, W, b):
    y = f(x, W, b); E = (y ** 3).sum()
    gx, = torch.autograd.grad(E, x, create_graph=True)
    L = (gx ** 2).sum() + y.sum()
    return y, gx, torch.autograd.grad(L, [x, W, b])
y1, gx1, G1 = loss3(gemm.linear, x, W, b)
xd, Wd, bd = (t.detach().double().requires_grad_() for t in (x, W, b))
y2, gx2, G2 = loss3(torch.nn.functional.linear, xd, Wd, bd)
print("linear fwd", rel(y1, y2), "gx", rel(gx1, gx2), "ggx", rel(G1[0], G2[0]), "ggW", rel(G1[1], G2[1]), "ggb", rel(G1[2], G2[2]))
# timing vs torch fp32
def tm(f, n=50):
    f(); torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b_ = torch.cuda.Event(enable_timing=True); a.record()
    for _ in range(n): f()
    b_.record(); torch.cuda.synchronize(); return a.elapsed_time(b_) / n * 1e3
for (m, n, k) in [(5376, 576, 128), (5376, 128, 128), (5376, 480, 128), (5376, 128, 352), (147456, 576, 128)]:
    A = torch.randn(m, k, device=dev); W = torch.randn(n, k, device=dev); b = torch.randn(n, device=dev)
    t1 = tm(lambda: gemm.mm_raw(A, W, False, True, b)); t2 = tm(lambda: torch.addmm(b, A, W.T))
    print(f"m{m} n{n} k{k}: tcgen05 3xTF32 {t1:.1f} us ({2*m*n*k/t1*1e-6:.1f} TFLOP/s)  torch fp32 {t2:.1f} us")
print("ALL OK" if ok else "FAILED")
