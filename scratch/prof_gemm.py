import sys; sys.path.insert(0, ".")
import torch
from xequinet_b200 import gemm
dev = "cuda"
for (m, n, k) in [(5376, 128, 128), (5376, 576, 128)]:
    A = torch.randn(m, k, device=dev); W = torch.randn(n, k, device=dev); b = torch.randn(n, device=dev)
    for _ in range(3):
        gemm.mm_raw(A, W, False, True, b)
    torch.cuda.synchronize()
G = torch.randn(5376, 576, device=dev); X = torch.randn(5376, 128, device=dev)
for _ in range(3):
    gemm.mm_raw(G, X, True, False)
torch.cuda.synchronize()
