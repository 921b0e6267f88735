"""Not a test: CUDA-event timing of the row-statistics kernel (what is left of ln_1 / ln_2)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from clip_ppo_b200 import _native as N

L = N.lib()
st = torch.cuda.current_stream().cuda_stream
for rows in (51200, 204800):
    x = torch.randn(rows, 768, device="cuda").bfloat16()
    out = torch.empty(rows, 2, device="cuda")
    for _ in range(3):
        N.check(L.clipppo_rowstats_bf16(x.data_ptr(), rows, 768, 768, out.data_ptr(), st))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        N.check(L.clipppo_rowstats_bf16(x.data_ptr(), rows, 768, 768, out.data_ptr(), st))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"rowstats rows={rows}: {ms * 1e3:.1f} us  {rows * 768 * 2 / ms / 1e6:.0f} GB/s")
