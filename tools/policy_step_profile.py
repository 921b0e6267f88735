"""Not a test: kernel-time breakdown of one NatureCNN forward + backward (csrc/policy.cu) at a minibatch size.
python tools/policy_step_profile.py [mb]"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from torch.profiler import profile, ProfilerActivity

from clip_ppo_b200.policy import NatureCNN

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
torch.manual_seed(0)
seq = nn.Sequential(nn.Conv2d(3, 32, 8, stride=4), nn.ReLU(), nn.Conv2d(32, 64, 4, stride=2), nn.ReLU(), nn.Conv2d(64, 64, 3, stride=1), nn.ReLU(),
                    nn.Flatten(), nn.Linear(64 * 7 * 7, 512), nn.ReLU()).cuda()
net = NatureCNN.from_sequential(seq)
x = torch.rand(mb, 84, 84, 3, device="cuda").permute(0, 3, 1, 2)
gh = torch.randn(mb, 512, device="cuda")


def run():
    net.zero_grad(set_to_none=True)
    net(x).backward(gh)


for _ in range(3):
    run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        run()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for ev in prof.events():
    if "cuda" in str(ev.device_type).lower():
        a = agg.setdefault(ev.name.replace("clipppo::(anonymous namespace)::", "").replace("void ", "")[:64], [0, 0.0])
        a[0] += 1; a[1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print(f"mb = {mb}: {tot / 3 / 1e3:.3f} ms of kernel time per forward + backward")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"  {k:64s} {c // 3:3d} x {v / c:8.1f} us = {v / 3:8.1f} us {100 * v / tot:5.1f}%")
