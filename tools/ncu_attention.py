"""Not a test: an attention-only run for `ncu --set full` (ViT-B/32: T = 50, 12 heads, d = 64)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from clip_ppo_b200 import _native as N

L = N.lib()
st = torch.cuda.current_stream().cuda_stream
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
T = int(sys.argv[2]) if len(sys.argv) > 2 else 50            # 257 with H = 16: the ViT-L/14 kernel
H = int(sys.argv[3]) if len(sys.argv) > 3 else 12
dh = 64
qkv = torch.randn(n * T, 3 * H * dh, device="cuda").bfloat16()
out = torch.empty(n * T, H * dh, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    N.check(L.clipppo_attention_bf16(qkv.data_ptr(), n, T, H, dh, out.data_ptr(), st))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    N.check(L.clipppo_attention_bf16(qkv.data_ptr(), n, T, H, dh, out.data_ptr(), st))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
gb = n * T * (3 + 1) * H * dh * 2 / 1e9
print(f"attention n={n}: {ms * 1e3:.1f} us  {gb / ms * 1e3:.0f} GB/s")
