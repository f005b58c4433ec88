import sys; sys.path.insert(0, ".")
import torch
from xequinet_b200 import gemm
dev = "cuda"
M = 5376
A = torch.randn(M, 128, device=dev); W1 = torch.randn(128, 128, device=dev); W2 = torch.randn(576, 128, device=dev); b = torch.randn(576, device=dev)
G = torch.randn(M, 576, device=dev)
for _ in range(3):
    gemm.mm_raw(A, W1, False, True, b[:128])          # (42,1,1)
    gemm.mm_raw(A, W2, False, True, b)                # (42,3,1)
    gemm.mm_raw(G, A, True, False)                    # grad weight: [576,128] = G^T A, split-K
torch.cuda.synchronize()
