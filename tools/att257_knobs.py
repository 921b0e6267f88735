import os, sys
sys.path.insert(0, '/root/repo')
import torch
from clip_ppo_b200 import _native as N
L = N.lib(); st = torch.cuda.current_stream().cuda_stream
n, T, H = 1024, 257, 16
qkv = torch.randn(n * T, 3 * H * 64, device="cuda").bfloat16()
out = torch.empty(n * T, H * 64, device="cuda", dtype=torch.bfloat16)
for _ in range(3): N.check(L.clipppo_attention_bf16(qkv.data_ptr(), n, T, H, 64, out.data_ptr(), st))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): N.check(L.clipppo_attention_bf16(qkv.data_ptr(), n, T, H, 64, out.data_ptr(), st))
e1.record(); torch.cuda.synchronize()
print(os.environ.get("CLIPPPO_ATC_WG1_DELAY"), f"{e0.elapsed_time(e1) * 100:.1f} us")
