import sys; sys.path.insert(0, ".")
import torch
from xequinet_b200 import gemm
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
for (m, n, k, pos) in [(4096, 128, 704, True), (4096, 128, 704, False), (4096, 576, 128, True), (4096, 128, 128, False)]:
    A = torch.rand(m, k, device="cuda") if pos else torch.randn(m, k, device="cuda")
    B = torch.rand(n, k, device="cuda") if pos else torch.randn(n, k, device="cuda")
    C64 = A.double() @ B.double().T
    for name, C in (("xeq", gemm.mm_raw(A, B, False, True)), ("cublas", A @ B.T)):
        e = (C.double() - C64)
        print(f"m{m} n{n} k{k} pos={pos} {name}: rel rms {float(e.pow(2).mean().sqrt() / C64.pow(2).mean().sqrt()):.2e}  rel bias {float(e.mean() / C64.abs().mean()):.2e}  max {float(e.abs().max() / C64.abs().max()):.2e}")
