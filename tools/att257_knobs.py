"""Not a test: CUDA-event timing of the long-sequence attention kernel at T = 257 (ViT-L/14) and T = 256 (no tail token)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_ppo_b200 import _native as N
L = N.lib(); st = torch.cuda.current_stream().cuda_stream
for (n, T, H) in ((1024, 257, 16), (1024, 256, 16), (256, 257, 16)):
    qkv = torch.randn(n * T, 3 * H * 64, device="cuda").bfloat16()
    out = torch.empty(n * T, H * 64, device="cuda", dtype=torch.bfloat16)
    for _ in range(3): N.check(L.clipppo_attention_bf16(qkv.data_ptr(), n, T, H, 64, out.data_ptr(), st))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): N.check(L.clipppo_attention_bf16(qkv.data_ptr(), n, T, H, 64, out.data_ptr(), st))
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(f"n={n} T={T} H={H}: {us:.1f} us  {4.0 * n * H * T * T * 64 / us / 1e6:.0f} TFLOP/s  {n * T * 4 * H * 64 * 2 / us / 1e3:.0f} GB/s")
