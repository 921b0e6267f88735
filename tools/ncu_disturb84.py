"""Not a test: 84x84x3 disturb run for ncu.  python tools/ncu_disturb84.py SEVERE|MODERATE [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_ppo_b200 import disturb as D
from shared.disturbance_types import DisturbanceSeverity, SEVERITY_CONFIGS
sev = sys.argv[1] if len(sys.argv) > 1 else "SEVERE"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
row = SEVERITY_CONFIGS[DisturbanceSeverity[sev]]
x = torch.rand(B, 3, 84, 84, device="cuda"); n = torch.randn(B, 3, 84, 84, device="cuda")
k = D.blur_kernel_size(row["gaussian_blur_sigma"]); taps = D.gaussian_taps(k, row["gaussian_blur_sigma"])
ph, pw = D.cutout_patch(84, 84, row["cutout_ratio"])
for _ in range(3):
    D.fused_disturb(x, stages=15, noise=n, noise_sigma=row["gaussian_noise_sigma"], contrast=1.1, taps=taps, window=(3, 5, ph, pw))
torch.cuda.synchronize()
print("done")
