import os, sys
sys.path.insert(0, '/root/repo')
import torch
from clip_ppo_b200 import _native as N
L = N.lib(); st = torch.cuda.current_stream().cuda_stream
for T in (256, 257):
    n, H = 1024, 16
    qkv = torch.randn(n * T, 3 * H * 64, device="cuda").bfloat16()
    out = torch.empty(n * T, H * 64, device="cuda", dtype=torch.bfloat16)
    print("T", T, "items per CTA", n * H / 148, flush=True)
    N.check(L.clipppo_attention_bf16(qkv.data_ptr(), n, T, H, 64, out.data_ptr(), st))
    torch.cuda.synchronize()
    N.check(L.clipppo_attention_bf16(qkv.data_ptr(), n, T, H, 64, out.data_ptr(), st))
    torch.cuda.synchronize()
