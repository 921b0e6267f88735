import os, sys
sys.path.insert(0, '.')
import torch, torch.nn as nn
from torch.profiler import profile, ProfilerActivity
from clip_ppo_b200.policy import NatureCNN
torch.manual_seed(0)
seq = nn.Sequential(nn.Conv2d(3, 32, 8, stride=4), nn.ReLU(), nn.Conv2d(32, 64, 4, stride=2), nn.ReLU(), nn.Conv2d(64, 64, 3, stride=1), nn.ReLU(), nn.Flatten(), nn.Linear(64 * 7 * 7, 512), nn.ReLU()).cuda()
net = NatureCNN.from_sequential(seq)
for mb in (8, 64, 256):
    x = torch.rand(mb, 84, 84, 3, device="cuda").permute(0, 3, 1, 2)
    def t(fn, n=50):
        for _ in range(5): fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(n): fn()
        b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n * 1e3
    with torch.no_grad():
        print(f"mb={mb}: eager {t(lambda: seq(x)):.1f} us   native {t(lambda: net(x)):.1f} us per forward")
x = torch.rand(64, 84, 84, 3, device="cuda").permute(0, 3, 1, 2)
with torch.no_grad():
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5): net(x)
        torch.cuda.synchronize()
import collections
agg = collections.OrderedDict()
for ev in prof.events():
    if "cuda" in str(ev.device_type).lower():
        a = agg.setdefault(ev.name[:70], [0, 0.0]); a[0] += 1; a[1] += ev.device_time
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:10]:
    print(f"  {k:70s} {c // 5:3d} x {v / c:8.1f} us")
