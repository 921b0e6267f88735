"""Not a test: CUDA-event timing of the attention kernels.  python tools/bench_attention.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from clip_ppo_b200 import _native as N

L = N.lib()
st = torch.cuda.current_stream().cuda_stream
for (n, T, H, causal) in ((4096, 50, 12, False), (1024, 50, 12, False), (1024, 257, 16, False), (256, 257, 16, False),
                          (4096, 77, 8, True), (4096, 77, 8, False), (4096, 77, 12, True)):
    dh = 64
    fn = L.clipppo_attention_causal_bf16 if causal else L.clipppo_attention_bf16
    qkv = torch.randn(n * T, 3 * H * dh, device="cuda").bfloat16()
    out = torch.empty(n * T, H * dh, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        N.check(fn(qkv.data_ptr(), n, T, H, dh, out.data_ptr(), st))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        N.check(fn(qkv.data_ptr(), n, T, H, dh, out.data_ptr(), st))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = n * T * 4 * H * dh * 2 / 1e9
    fl = 4.0 * n * H * T * T * dh
    print(f"attention{' causal' if causal else ''} n={n} T={T} H={H}: {ms * 1e3:8.1f} us  {gb / ms * 1e3:6.0f} GB/s (QKV in + O out)  {fl / ms / 1e9:6.1f} TFLOP/s")
