"""Not a test (pytest does not collect it): the GPU half of the reference's own disturbance benchmark
(shared/benchmark_disturbances.py:39-93: 84x84x3 frames, HARD, batch 1..64, 5 warm-up + 50 calls, wall clock
with a synchronize per call) run on this repo's DisturbanceWrapperGPU, next to the PyTorch-eager op sequence the
reference dispatches on a GPU (oracle/disturb.py on cuda:0 - torchvision's ops restated, pinned against the
reference's outputs).  These are the shapes of the env-step call sites (clip_ppo_minigrid.py:381-388: E = 64;
clip_ppo_atari.py:568-584: 4 x [256,1,84,84]), where a call is launch-latency bound, not HBM bound.

    python tests/disturb_latency_baseline.py

Lives under tests/ because only tests/, smoke() and bench.py's CPU legs may import oracle/.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from oracle import disturb as od
from shared.disturbance_types import DisturbanceSeverity
from shared.disturbances_gpu import DisturbanceWrapperGPU

ITERS = 50
dev = torch.device("cuda", 0)
sev = od.SEVERITY_TABLE["HARD"]


def eager_call(x):
    """One apply_disturbances of the reference on a GPU: device randn_like, CPU scalar draws, ~35 eager launches."""
    noise = torch.randn_like(x)
    torch.randperm(4)
    c = float(torch.empty(1).uniform_(*sev["contrast"]))
    sigma = torch.empty(1).uniform_(sev["blur_sigma"], sev["blur_sigma"]).item()
    k1d = od.gaussian_kernel1d(od.blur_kernel_size(sev["blur_sigma"]), sigma).to(x.device)
    H, W = x.shape[-2:]
    ph, pw = od.cutout_patch(H, W, sev["cutout"])
    sh = torch.randint(0, max(1, H - ph + 1), (1,)).item()
    sw = torch.randint(0, max(1, W - pw + 1), (1,)).item()
    return od.disturb(x, noise, sev["noise_sigma"], c, k1d, sh, sw, ph, pw)


def wall(fn, x):
    for _ in range(5):
        fn(x)
    torch.cuda.synchronize()
    ts = []
    for _ in range(ITERS):
        t0 = time.perf_counter()
        fn(x)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return float(np.mean(ts)) * 1e6, float(np.std(ts)) * 1e6


def main():
    w = DisturbanceWrapperGPU(device="cuda", severity=DisturbanceSeverity.HARD, seed=42)
    print(f"GPU: {torch.cuda.get_device_name()}   84x84 frames, HARD, {ITERS} calls after 5 warm-ups, wall clock incl. synchronize")
    print(f"{'shape':>18s} {'this repo us/call':>20s} {'PyTorch eager us/call':>24s} {'ratio':>7s}")
    rng = np.random.default_rng(0)
    for B, C in [(1, 3), (4, 3), (8, 3), (16, 3), (32, 3), (64, 3), (256, 3), (256, 1), (1024, 1)]:
        frames = torch.from_numpy(rng.integers(0, 256, (B, 84, 84, C), dtype=np.uint8)).to(dev).float() / 255.0
        x = frames.permute(0, 3, 1, 2)                     # NHWC-strided NCHW view, like the reference's call sites
        ours, so = wall(w.apply_disturbances, x)
        eager, se = wall(eager_call, x)
        print(f"{str((B, C, 84, 84)):>18s} {ours:12.1f} +- {so:5.1f} {eager:16.1f} +- {se:5.1f} {eager / ours:7.1f}x")


if __name__ == "__main__":
    main()
