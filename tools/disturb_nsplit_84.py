"""Not a test: blur-task split sweep for 84x84 frames.  CLIPPPO_DISTURB_NSPLIT=n python tools/disturb_nsplit_84.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_ppo_b200 import disturb as D
from shared.disturbance_types import DisturbanceSeverity, SEVERITY_CONFIGS
for sev in ("MILD", "MODERATE", "SEVERE"):
    for (B, C) in ((16384, 3), (65536, 1)):
        row = SEVERITY_CONFIGS[DisturbanceSeverity[sev]]
        x = torch.rand(B, C, 84, 84, device="cuda"); n = torch.randn(B, C, 84, 84, device="cuda")
        k = D.blur_kernel_size(row["gaussian_blur_sigma"]); taps = D.gaussian_taps(k, row["gaussian_blur_sigma"])
        ph, pw = D.cutout_patch(84, 84, row["cutout_ratio"])
        fn = lambda: D.fused_disturb(x, stages=15, noise=n, noise_sigma=row["gaussian_noise_sigma"], contrast=1.1, taps=taps, window=(3, 5, ph, pw))
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"NSPLIT={os.environ.get('CLIPPPO_DISTURB_NSPLIT', 'auto'):4s} {sev:8s} k={k} B={B} C={C}: {ms*1e3:8.1f} us  {12.0*B*C*84*84/ms/1e6/6537.6*100:5.1f}% HBM")
