"""Kernel-time breakdown of one text-tower pass (torch.profiler / CUPTI).  python tools/text_step_profile.py [texts]"""
import os, re, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from clip_ppo_b200 import clip_compat
from clip_ppo_b200.text import TextEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sd = clip_compat.random_text_state_dict("ViT-B/32", 0)
eng = TextEngine(sd, device="cuda")
g = torch.Generator().manual_seed(0)
tok = torch.randint(1, 40000, (n, 77), generator=g)
tok[:, 0] = 49406
tok[torch.arange(n), torch.randint(2, 77, (n,), generator=g)] = 49407
tok = tok.cuda()
for _ in range(2):
    eng.encode(tok)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    eng.encode(tok)
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for ev in prof.events():
    if ev.device_type is not None and "cuda" in str(ev.device_type).lower() and ev.name:
        m = re.search(r"(\w+_kernel)(<[^>]*>)?", ev.name)
        key = m.group(0) if m else ev.name[:60]
        a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print(f"{n} texts: {tot / 1e3:.2f} ms of kernel time")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:58s} {c:4d} {v / 1e3:9.3f} ms {100 * v / tot:5.1f}%")
