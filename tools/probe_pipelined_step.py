"""Not a test: the bench step with the NEXT batch's disturbances running on a second stream under the current batch's tower
(software pipelining across steps; the GEMM grids optionally capped with CLIPPPO_GEMM_MAX_SMS so that the second stream finds
free SMs).  Same public calls, same work per step; sustained timing, interleaved with the sequential step.
    CLIPPPO_GEMM_MAX_SMS=132 python tools/probe_pipelined_step.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CLIPPPO_ALLOW_RANDOM_WEIGHTS", "1")
import torch

import shared.clip_ppo_utils as U
from shared.disturbances_gpu import DisturbanceWrapperGPU
from shared.disturbance_types import DisturbanceSeverity

dev = torch.device("cuda", 0)
B = 4096
model = U.load_clip_model("ViT-B/32", device=dev)
w = DisturbanceWrapperGPU(device=dev, seed=1, severity=DisturbanceSeverity.MODERATE)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.rand(B, 3, 224, 224, device=dev, generator=g)
noise = torch.randn(B, 3, 224, 224, device=dev, generator=g)
z = torch.relu(torch.randn(B, 512, device=dev, generator=g))
side = torch.cuda.Stream(device=dev)
main = torch.cuda.current_stream(dev)


def disturb():
    return w.apply_disturbances(x, noise=noise, contrast_factor=1.1, cutout_start=(44, 56), out_scale=255.0)


def tower(d):
    emb = U.generate_clip_embeddings(U.AblationMode.NONE, model, "image", B, dev, images=d)
    return U.compute_cosine_embedding_loss(z, emb)


def sequential(steps):
    for _ in range(steps):
        loss = tower(disturb())
    return loss


def pipelined(steps):
    """disturb(i + 1) on `side` while tower(i) runs on `main`; exactly `steps` disturb calls and `steps` tower calls."""
    ready = torch.cuda.Event()
    d = disturb()                      # the first batch (belongs to this call's work)
    for i in range(steps):
        nxt = None
        if i + 1 < steps:
            side.wait_stream(main)     # frames / noise are ready; also orders buffer reuse
            with torch.cuda.stream(side):
                nxt = disturb()
                ready.record(side)
        loss = tower(d)
        if nxt is not None:
            main.wait_event(ready)
            nxt.record_stream(main)
            d = nxt
    return loss


def timed(fn, steps=10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    loss = fn(steps)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps, float(loss)


sequential(3); pipelined(3)
print("max SMs per GEMM:", os.environ.get("CLIPPPO_GEMM_MAX_SMS", "148"))
for rep in range(3):
    t1, l1 = timed(sequential)
    t2, l2 = timed(pipelined)
    print(f"rep {rep}: sequential {t1:7.2f} ms/step   pipelined {t2:7.2f} ms/step   (loss {l1:.6f} / {l2:.6f})")
