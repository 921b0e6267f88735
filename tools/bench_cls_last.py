"""Not a test: the tower pass with and without CLIPPPO_VIT_CLS_LAST_BLOCK (class-token rows only after the last block's attention)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_ppo_b200.clip_compat.model import random_visual_state_dict
from clip_ppo_b200.vit import VitEngine
for name, n in (("ViT-B/32", 4096), ("ViT-L/14", 1024), ("ViT-B/32", 64)):
    eng = VitEngine(random_visual_state_dict(name, 0), device="cuda")
    x = torch.rand(n, 3, 224, 224, device="cuda") * 255.0
    res = {}
    for rep in range(2):
        for flag in ("0", "1"):
            os.environ["CLIPPPO_VIT_CLS_LAST_BLOCK"] = flag
            for _ in range(3): out = eng.encode(x)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): out = eng.encode(x)
            e1.record(); torch.cuda.synchronize()
            res.setdefault(flag, []).append(e0.elapsed_time(e1) / 10)
    print(f"{name} n={n}: full {min(res['0']):.3f} ms ({n / min(res['0']):.1f} k frames/s)   class-token rows only in the last block {min(res['1']):.3f} ms ({n / min(res['1']):.1f} k frames/s)")
