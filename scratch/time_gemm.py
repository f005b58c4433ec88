"""GEMM kernel timings on the shapes of a c3 step (M = 5376 nodes) vs torch.matmul fp32 (cuBLAS)."""
import sys; sys.path.insert(0, ".")
import os, torch
from pathlib import Path
from xequinet_b200 import _lib
if os.environ.get("XEQ_LIB"): _lib.LIB_PATH = Path(os.environ["XEQ_LIB"]).resolve()
from xequinet_b200 import gemm
dev = "cuda"
def tm(f, n=50):
    for _ in range(5): f()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n * 1e3
M = 5376
for (n, k, ta, tb) in [(128, 128, False, True), (576, 128, False, True), (128, 352, False, True), (480, 128, False, True), (128, 224, False, True),
                       (128, 576, False, False), (128, 5376, True, False), (576, 5376, True, False)]:
    m = M if not ta else 128
    A = torch.randn((k, m) if ta else (m, k), device=dev); B = torch.randn((n, k) if tb else (k, n), device=dev)
    t1 = tm(lambda: gemm.mm_raw(A, B, ta, tb))
    t2 = tm(lambda: torch.matmul(A.T if ta else A, B.T if tb else B))
    print(f"m={m} n={n} k={k} ta={ta} tb={tb}: xeq {t1:.1f} us   torch fp32 {t2:.1f} us")
V = torch.randn(M, 480, device=dev); w = torch.randn(128*128+64*64+32*32, device=dev)
print("irreps_linear %.1f us" % tm(lambda: gemm.irreps_linear_raw(V, w, None, (128, 64, 32), False)))
print("irreps_wgrad %.1f us" % tm(lambda: gemm.irreps_wgrad_raw(V, V, (128, 64, 32))))
