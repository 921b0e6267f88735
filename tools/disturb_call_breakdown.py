"""Where one env-step apply_disturbances call spends its time (host segments by perf_counter, device by CUDA events).
    python tools/disturb_call_breakdown.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shared.disturbance_types import DisturbanceSeverity
from shared.disturbances_gpu import DisturbanceWrapperGPU
from clip_ppo_b200 import disturb as D, _native as N

w = DisturbanceWrapperGPU(device="cuda", severity=DisturbanceSeverity.HARD, seed=1)
for shape, nhwc in [((64, 3, 84, 84), True), ((64, 3, 84, 84), False), ((256, 1, 84, 84), False)]:
    B, C, H, W = shape
    x = (torch.rand(B, H, W, C, device="cuda").permute(0, 3, 1, 2) if nhwc else torch.rand(shape, device="cuda"))
    seg = {"randn_like": 0.0, "cpu draws": 0.0, "fused_disturb (host)": 0.0}
    dev_ms = 0.0
    R = 200
    for it in range(R + 20):
        if it == 20:
            seg = {k: 0.0 for k in seg}; dev_ms = 0.0
        t0 = time.perf_counter()
        noise = torch.randn_like(x)
        t1 = time.perf_counter()
        c = w._draw_contrast(); taps = w._draw_blur_taps(); win = w._draw_cutout(H, W, None)
        t2 = time.perf_counter()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        t3 = time.perf_counter()
        out = D.fused_disturb(x, stages=N.STAGE_ALL, noise=noise, noise_sigma=w.gaussian_noise_sigma, contrast=c, taps=taps, window=win)
        t4 = time.perf_counter()
        e1.record()
        torch.cuda.synchronize()
        seg["randn_like"] += t1 - t0; seg["cpu draws"] += t2 - t1; seg["fused_disturb (host)"] += t4 - t3
        dev_ms += e0.elapsed_time(e1)
    print(shape, "NHWC view" if nhwc else "contiguous", " ".join(f"{k}: {v / R * 1e6:.1f} us" for k, v in seg.items()),
          f"device (alloc + kernel, events): {dev_ms / R * 1e3:.1f} us")

# pure device time of the kernel: launches queued behind a 1 ms spin so the events see no host gaps
print("kernel alone (events, launches queued behind a busy GPU):")
for shape, nhwc in [((8, 3, 84, 84), True), ((64, 3, 84, 84), True), ((8, 3, 84, 84), False), ((64, 3, 84, 84), False),
                    ((256, 1, 84, 84), False), ((256, 3, 84, 84), True), ((1024, 3, 84, 84), True), ((8192, 3, 84, 84), True)]:
    B, C, H, W = shape
    x = (torch.rand(B, H, W, C, device="cuda").permute(0, 3, 1, 2) if nhwc else torch.rand(shape, device="cuda"))
    noise = torch.randn_like(x)
    taps = w._draw_blur_taps(); win = w._draw_cutout(H, W, None)
    tot = 0.0
    for it in range(30):
        torch.cuda._sleep(2_000_000)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        D.fused_disturb(x, stages=N.STAGE_ALL, noise=noise, noise_sigma=0.13, contrast=1.1, taps=taps, window=win)
        e1.record()
        torch.cuda.synchronize()
        if it >= 10: tot += e0.elapsed_time(e1)
    print(f"  {shape} {'NHWC view' if nhwc else 'contiguous'}: {tot / 20 * 1e3:.1f} us")

# Atari call site: one frame of a 4-frame stack, [E,1,84,84] with the images 4 frames apart
for E in (64, 256):
    stack = torch.rand(E, 4, 84, 84, device="cuda")
    x = stack[:, 1:2]
    noise = torch.randn(E, 1, 84, 84, device="cuda")
    taps = w._draw_blur_taps(); win = w._draw_cutout(84, 84, None)
    tot = 0.0
    for it in range(30):
        torch.cuda._sleep(2_000_000)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        D.fused_disturb(x, stages=N.STAGE_ALL, noise=noise, noise_sigma=0.13, contrast=1.1, taps=taps, window=win)
        e1.record()
        torch.cuda.synchronize()
        if it >= 10: tot += e0.elapsed_time(e1)
    print(f"  ({E}, 1, 84, 84) frame of an Atari stack (batch stride 4 frames): {tot / 20 * 1e3:.1f} us")
