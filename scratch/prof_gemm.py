"""ncu target: the K3 GEMM on the c3 shapes (M = 5376): forward Linear shapes, the grouped o3.Linear, a split-K weight gradient."""
import sys; sys.path.insert(0, ".")
import torch
from xequinet_b200 import gemm
M = 5376
dev = "cuda"
shapes = [(128, 128), (576, 128), (128, 352), (480, 128), (128, 224)]
ops = []
for n, k in shapes:
    A = torch.randn(M, k, device=dev); B = torch.randn(n, k, device=dev)
    ops.append(lambda A=A, B=B: gemm.mm_raw(A, B, False, True))
V = torch.randn(M, 480, device=dev); w = torch.randn(128 * 128 + 64 * 64 + 32 * 32, device=dev)
ops.append(lambda: gemm.irreps_linear_raw(V, w, None, (128, 64, 32), False))
G = torch.randn(M, 576, device=dev); X = torch.randn(M, 128, device=dev)
ops.append(lambda: gemm.mm_raw(X, G, True, False))
for _ in range(3):
    for f in ops: f()
torch.cuda.synchronize()
