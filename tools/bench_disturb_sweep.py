"""Not a test: BASELINE configs[4] - disturbance kernel throughput sweep, batch 256 ... 16384 of 224x224x3 frames
(and the 84x84 training shapes), CUDA events, supplied noise (12 B / element).  python tools/bench_disturb_sweep.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from clip_ppo_b200 import disturb as D
from shared.disturbance_types import DisturbanceSeverity, SEVERITY_CONFIGS

try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    HBM = 6537.6


def run(B, C, H, W, sev):
    row = SEVERITY_CONFIGS[DisturbanceSeverity[sev]]
    x = torch.rand(B, C, H, W, device="cuda")
    n = torch.randn(B, C, H, W, device="cuda")
    k = D.blur_kernel_size(row["gaussian_blur_sigma"])
    taps = D.gaussian_taps(k, row["gaussian_blur_sigma"])
    ph, pw = D.cutout_patch(H, W, row["cutout_ratio"])
    fn = lambda: D.fused_disturb(x, stages=15, noise=n, noise_sigma=row["gaussian_noise_sigma"], contrast=1.1, taps=taps,
                                 window=(3, 5, ph, pw))
    iters = 200 if B * C * H * W < 2e8 else 20
    for _ in range(3):
        fn()
    # inputs of the small cases fit the 126 MB L2: rotate over copies so every call reads from HBM
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    gb = 12.0 * B * C * H * W / 1e9
    note = "  (x + noise + out = %.0f MB: L2-resident between calls)" % (gb * 1e3) if gb * 1e3 < 126 else ""
    print(f"{sev:8s} k={k} B={B:6d} {C}x{H}x{W}: {ms * 1e3:9.1f} us  {gb / ms * 1e3:8.1f} GB/s  {gb / ms * 1e3 / HBM * 100:5.1f}% of measured HBM  "
          f"{B / ms * 1e3 / 1e6:7.3f} M frames/s{note}", flush=True)


for sev in ("SEVERE", "MODERATE"):
    for B in (256, 512, 1024, 2048, 4096, 8192, 16384):
        run(B, 3, 224, 224, sev)
for sev in ("SEVERE", "HARD"):
    for (B, C) in ((64, 3), (256, 1), (1024, 3), (8192, 3), (32768, 1), (131072, 1)):
        run(B, C, 84, 84, sev)
