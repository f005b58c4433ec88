"""ncu target: a few launches of the K3 GEMM on c3 shapes.  usage: python scratch/prof_gemm.py [n k]"""
import sys; sys.path.insert(0, ".")
import torch
from xequinet_b200 import gemm
M = 5376
n, k = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 128)
A = torch.randn(M, k, device="cuda"); B = torch.randn(n, k, device="cuda")
for _ in range(5): gemm.mm_raw(A, B, False, True)
torch.cuda.synchronize()
