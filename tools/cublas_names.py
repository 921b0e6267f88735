"""Not a test: torch.matmul on the tower's GEMM shapes, for an ncu launch list that names the
cuBLAS kernels (tile / stage / cluster choices) the library picks on this B200."""
import torch

M = 50100
for Nn, K in ((2304, 768), (3072, 768), (768, 3072), (768, 768)):
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = torch.randn(Nn, K, device="cuda").bfloat16()
    for _ in range(2):
        c = torch.matmul(a, w.t())
    torch.cuda.synchronize()
print("done")
