"""Not a test: CUDA-event micro-benchmarks of the memory-bound kernels (disturb, preprocess, LN,
attention) against the measured HBM peak.  python tools/bench_kernels.py [disturb|pre|ln|attn ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from clip_ppo_b200 import _native as N
from clip_ppo_b200 import disturb as D

L = N.lib()
HBM = 6537.6
try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_disturb():
    from shared.disturbance_types import DisturbanceSeverity, SEVERITY_CONFIGS
    SEVERITY_TABLE = {s.name: {"noise_sigma": r["gaussian_noise_sigma"], "blur_sigma": r["gaussian_blur_sigma"],
                               "cutout": r["cutout_ratio"]} for s, r in SEVERITY_CONFIGS.items()}
    for (B, C, H, W, sev) in [(4096, 3, 224, 224, "MODERATE"), (4096, 3, 224, 224, "SEVERE"), (501, 3, 224, 224, "MODERATE"),
                              (16384, 3, 84, 84, "MODERATE"), (16384, 1, 84, 84, "HARD"), (64, 3, 84, 84, "MODERATE"),
                              (256, 1, 84, 84, "HARD")]:
        cfg = SEVERITY_TABLE[sev]
        x = torch.rand(B, C, H, W, device="cuda")
        n = torch.randn(B, C, H, W, device="cuda")
        k = D.blur_kernel_size(cfg["blur_sigma"])
        taps = D.gaussian_taps(k, cfg["blur_sigma"])
        ph, pw = D.cutout_patch(H, W, cfg["cutout"])
        stages = int(os.environ.get("DISTURB_STAGES", "15"))
        ms = timeit(lambda: D.fused_disturb(x, stages=stages, noise=n, noise_sigma=cfg["noise_sigma"], contrast=1.1, taps=taps,
                                            window=(3, 5, ph, pw)), iters=10 if B >= 4096 else 50)
        gb = 12.0 * B * C * H * W / 1e9
        print(f"disturb B={B} C={C} {H}x{W} {sev:8s} k={k}: {ms*1e3:9.1f} us  {gb/ms*1e3:8.1f} GB/s  "
              f"{gb/ms*1e3/HBM*100:5.1f}% of measured HBM  {B/ms*1e3:12.0f} frames/s", flush=True)
        del x, n


def bench_pre():
    for (n, C, h, dt) in [(501, 3, 224, torch.float32), (501, 3, 84, torch.float32), (501, 3, 84, torch.uint8), (501, 1, 84, torch.float32)]:
        img = (torch.rand(n, C, h, h, device="cuda") * 255).to(dt)
        out = torch.empty(n * 49, 3072, device="cuda", dtype=torch.bfloat16)
        st = torch.cuda.current_stream().cuda_stream
        ms = timeit(lambda: N.check(L.clipppo_preprocess_bf16(img.data_ptr(), 1 if dt == torch.uint8 else 0, N.strides4(img), n, C, h, h,
                                                               1 / 255.0, 1, 32, 224, out.data_ptr(), st)), iters=50)
        gb = (img.numel() * img.element_size() + out.numel() * 2) / 1e9
        print(f"preprocess n={n} C={C} {h}x{h} {dt}: {ms*1e3:8.1f} us  {gb/ms*1e3:8.1f} GB/s ({gb/ms*1e3/HBM*100:5.1f}% HBM)", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["disturb", "pre"]
    if "disturb" in which:
        bench_disturb()
    if "pre" in which:
        bench_pre()
