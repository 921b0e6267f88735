"""Not a test: a short up-sampling preprocess run for `ncu --set full` captures (84x84x3 fp32 -> 224x224 patch matrix)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from clip_ppo_b200 import _native as N

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
img = torch.randint(0, 256, (n, 3, 84, 84), device="cuda").float()
out = torch.empty(n * 49, 3072, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    N.check(N.lib().clipppo_preprocess_bf16(img.data_ptr(), 0, N.strides4(img), n, 3, 84, 84, 1 / 255.0, 1, 32, 224, out.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print("done")
