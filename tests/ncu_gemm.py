"""Not a test: a short GEMM-only run for `ncu --set full` captures (one tower chunk's shapes)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from clip_ppo_b200 import _native as N

L = N.lib()
st = torch.cuda.current_stream().cuda_stream
M = 25050                                            # 501 images x 50 tokens
gen = torch.Generator(device="cuda").manual_seed(0)
for (Nn, K, epi) in ((2304, 768, 0), (3072, 768, 1), (768, 3072, 2), (768, 768, 2)):
    a = (torch.randn(M, K, device="cuda", generator=gen) * 0.5).bfloat16()
    w = (torch.randn(Nn, K, device="cuda", generator=gen) * (K ** -0.5)).bfloat16()
    bias = torch.randn(Nn, device="cuda", generator=gen) * 0.1
    out = torch.zeros(M, Nn, device="cuda", dtype=torch.bfloat16 if epi in (0, 1) else torch.float32)
    for _ in range(2):
        N.check(L.clipppo_gemm_bf16(a.data_ptr(), w.data_ptr(), M, Nn, K, epi, bias.data_ptr(), None, 0, out.data_ptr(), Nn, st))
    torch.cuda.synchronize()
print("done")
