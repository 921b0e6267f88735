"""Not a test: in-situ kernel time breakdown of the bench step (torch.profiler / CUPTI, real clocks,
kernels overlapping as they do in the step; a kernel launched with PDL counts the time it waits for its predecessor).
python tools/step_profile.py [frames] [ViT-B/32 | ViT-L/14]"""
import os
import re
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CLIPPPO_ALLOW_RANDOM_WEIGHTS", "1")
import torch
from torch.profiler import ProfilerActivity, profile

import shared.clip_ppo_utils as U
from clip_ppo_b200 import disturb as D
from shared.disturbance_types import DisturbanceSeverity
from shared.disturbances_gpu import DisturbanceWrapperGPU

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
NAME = sys.argv[2] if len(sys.argv) > 2 else "ViT-B/32"
dev = torch.device("cuda", 0)
model = U.load_clip_model(NAME, device=dev)
engine = U._engine_for(model)
w = DisturbanceWrapperGPU(device=dev, seed=1, severity=DisturbanceSeverity.MODERATE)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.rand(B, 3, 224, 224, device=dev, generator=g)
noise = torch.randn(B, 3, 224, 224, device=dev, generator=g)
z = torch.relu(torch.randn(B, engine.cfg.out_dim, device=dev, generator=g))


def step():
    d = w.apply_disturbances(x, noise=noise, contrast_factor=1.1, cutout_start=(44, 56), out_scale=255.0)
    emb = U.generate_clip_embeddings(U.AblationMode.NONE, model, "image", B, dev, images=d)
    return U.compute_cosine_embedding_loss(z, emb)


for _ in range(3):
    step()
torch.cuda.synchronize()
STEPS = 3
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    e0.record()
    for _ in range(STEPS):
        step()
    e1.record()
    torch.cuda.synchronize()
wall = e0.elapsed_time(e1) / STEPS
agg = defaultdict(lambda: [0, 0.0])
first, last = None, None
for ev in prof.events():
    if ev.device_type.name != "CUDA" or ev.device_time_total <= 0:
        continue
    m = re.search(r"([A-Za-z_0-9]+(?:<[^()]*>)?)\(", ev.name.replace("(anonymous namespace)", "").replace("<unnamed>", ""))
    name = m.group(1) if m else ev.name[:60]
    agg[name][0] += 1
    agg[name][1] += ev.device_time_total
tot = sum(v[1] for v in agg.values()) / STEPS / 1e3
print(f"{B} frames/step: wall {wall:.2f} ms/step (under the profiler), sum of kernel durations {tot:.2f} ms/step")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
    print(f"  {t / STEPS / 1e3:8.3f} ms/step  {n // STEPS:5d} x {t / n:9.1f} us   {name}")
