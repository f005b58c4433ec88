"""GEMM kernel timings on the shapes of a c3 step (M = 5376 nodes) vs torch.matmul fp32 (cuBLAS, allow_tf32 = False).
Each candidate is captured 20x back to back into a CUDA graph and replayed, so the numbers are device times per
launch (launch-to-launch, L2-warm operands) and not host launch overhead."""
import sys; sys.path.insert(0, ".")
import torch
from xequinet_b200 import gemm
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
REP = 20
def tm(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        f()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        for _ in range(REP): f()
    g.replay(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); a.record()
    for _ in range(n): g.replay()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n / REP * 1e3
M = int(sys.argv[1]) if len(sys.argv) > 1 else 5376
print(f"M = {M}")
for (n, k, ta, tb) in [(128, 128, False, True), (576, 128, False, True), (128, 352, False, True), (480, 128, False, True), (128, 224, False, True),
                       (128, 576, False, False), (128, 56, False, True), (64, 128, False, True), (128, M, True, False), (576, M, True, False), (352, M, True, False)]:
    m = M if not ta else 128
    A = torch.randn((k, m) if ta else (m, k), device=dev); B = torch.randn((n, k) if tb else (k, n), device=dev)
    out = torch.empty(m, n, device=dev)
    t1 = tm(lambda: gemm.mm_raw(A, B, ta, tb))
    t2 = tm(lambda: torch.matmul(A.T if ta else A, B.T if tb else B, out=out))
    fl = 2.0 * m * n * k
    print(f"m={m} n={n} k={k} ta={ta} tb={tb}: xeq {t1:.1f} us ({fl/t1*1e-6:.1f} TF/s fp32-equivalent)   cuBLAS fp32 {t2:.1f} us")
V = torch.randn(M, 480, device=dev); w = torch.randn(128*128+64*64+32*32, device=dev)
print("irreps_linear %.1f us" % tm(lambda: gemm.irreps_linear_raw(V, w, None, (128, 64, 32), False)))
print("irreps_wgrad %.1f us" % tm(lambda: gemm.irreps_wgrad_raw(V, V, (128, 64, 32))))
