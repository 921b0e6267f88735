"""Kernel-time breakdown of one small-batch tower pass (FROZEN_CLIP policy forward).  python tools/smallbatch_profile.py [n]"""
import os, re, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from clip_ppo_b200 import clip_compat
from clip_ppo_b200.vit import VitEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
eng = VitEngine(clip_compat.random_visual_state_dict("ViT-B/32", 0), device="cuda")
x = torch.rand(n, 3, 84, 84, device="cuda")
for _ in range(3):
    eng.encode(x, pre_scale=1.0, l2norm=False)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        eng.encode(x, pre_scale=1.0, l2norm=False)
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for ev in prof.events():
    if ev.device_type is not None and "cuda" in str(ev.device_type).lower() and ev.name:
        m = re.search(r"(\w+_kernel)(<[^>]*>)?", ev.name)
        key = m.group(0) if m else ev.name[:50]
        a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print(f"n = {n}: {tot / 5:.1f} us of kernel time per pass (CUPTI durations include the PDL wait on the previous kernel)")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"  {k:50s} {c // 5:3d} x {v / c:7.2f} us = {v / 5:8.1f} us {100 * v / tot:5.1f}%")
