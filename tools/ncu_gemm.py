"""Not a test: a short GEMM-only run for `ncu --set full` captures - the four GEMM shapes of one
tower block at the bench's chunk size, with the epilogues the tower runs them with."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from clip_ppo_b200 import _native as N

L = N.lib()
st = torch.cuda.current_stream().cuda_stream
M = int(sys.argv[1]) if len(sys.argv) > 1 else 51200          # 1024 images x 50 tokens
gen = torch.Generator(device="cuda").manual_seed(0)
SHAPES = ((2304, 768, 6), (3072, 768, 7), (768, 3072, 8), (768, 768, 8), (768, 3072, 9), (768, 768, 9))
if len(sys.argv) > 2:                                        # e.g. "8,9": only these epilogues
    SHAPES = tuple(s for s in SHAPES if str(s[2]) in sys.argv[2].split(","))
for (Nn, K, epi) in SHAPES:
    a = (torch.randn(M, K, device="cuda", generator=gen) * 0.5).bfloat16()
    w = (torch.randn(Nn, K, device="cuda", generator=gen) * (K ** -0.5)).bfloat16()
    bias = torch.randn(Nn, device="cuda", generator=gen) * 0.1
    out = torch.zeros(M, Nn, device="cuda", dtype=torch.bfloat16)
    stats = torch.stack([torch.zeros(M, device="cuda"), torch.ones(M, device="cuda")], 1).contiguous()
    colsum = w.float().sum(1).contiguous()
    parts = torch.empty(M, (Nn + 127) // 128, 2, device="cuda")
    for _ in range(2):
        if epi == 9:
            N.check(L.clipppo_gemm_bf16_resid_stats(a.data_ptr(), w.data_ptr(), M, Nn, K, bias.data_ptr(), out.data_ptr(), Nn,
                                                    parts.data_ptr(), st))
            continue
        N.check(L.clipppo_gemm_bf16_fused(a.data_ptr(), w.data_ptr(), M, Nn, K, epi, bias.data_ptr(),
                                          stats.data_ptr() if epi < 8 else None, colsum.data_ptr() if epi < 8 else None,
                                          out.data_ptr(), Nn, st))
    torch.cuda.synchronize()
print("done")
