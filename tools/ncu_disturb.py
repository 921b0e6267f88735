"""Not a test: a short disturb-only run for `ncu --set full` captures (BASELINE configs[2] frame shape)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from clip_ppo_b200 import disturb as D
from shared.disturbance_types import DisturbanceSeverity, SEVERITY_CONFIGS

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
SEV = sys.argv[2] if len(sys.argv) > 2 else "MODERATE"
_row = SEVERITY_CONFIGS[DisturbanceSeverity[SEV]]
cfg = {"blur_sigma": _row["gaussian_blur_sigma"], "noise_sigma": _row["gaussian_noise_sigma"], "cutout": _row["cutout_ratio"]}
x = torch.rand(B, 3, 224, 224, device="cuda")
n = torch.randn(B, 3, 224, 224, device="cuda")
k = D.blur_kernel_size(cfg["blur_sigma"])
taps = D.gaussian_taps(k, cfg["blur_sigma"])
ph, pw = D.cutout_patch(224, 224, cfg["cutout"])
for _ in range(3):
    out = D.fused_disturb(x, stages=15, noise=n, noise_sigma=cfg["noise_sigma"], contrast=1.1, taps=taps, window=(3, 5, ph, pw))
torch.cuda.synchronize()
print("done")
