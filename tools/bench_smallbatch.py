"""Not a test: latency of the FROZEN_CLIP policy path (get_frozen_clip_features inside every policy
forward: E frames per call, reference clip_ppo_minigrid.py:249-254 / clip_ppo_atari.py:213-228) at small
batches, eager launches vs one CUDA-graph replay.  python tools/bench_smallbatch.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import shared.clip_ppo_utils as U

model = U.load_clip_model("ViT-B/32", device="cuda")
eng = U._engine_for(model)
for n in (8, 64, 256, 1024):
    x = torch.rand(n, 3, 84, 84, device="cuda")
    for _ in range(5):
        f = U.get_frozen_clip_features(x, model.visual)
    torch.cuda.synchronize()
    iters = 50
    t0 = time.perf_counter()
    for _ in range(iters):
        f = U.get_frozen_clip_features(x, model.visual)
    t_cpu = (time.perf_counter() - t0) / iters * 1e6          # host time to enqueue (no sync)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        f = U.get_frozen_clip_features(x, model.visual)
    torch.cuda.synchronize()
    t_wall = (time.perf_counter() - t0) / iters * 1e6
    line = f"n={n:5d}  eager: host enqueue {t_cpu:8.1f} us/call   wall {t_wall:8.1f} us/call  ({n / t_wall * 1e6:9.0f} frames/s)"
    if hasattr(eng, "encode_graphed"):
        for _ in range(3):
            g = eng.encode_graphed(x, pre_scale=1.0, l2norm=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            g = eng.encode_graphed(x, pre_scale=1.0, l2norm=False)
        torch.cuda.synchronize()
        t_g = (time.perf_counter() - t0) / iters * 1e6
        line += f"   graph replay: wall {t_g:8.1f} us/call  ({n / t_g * 1e6:9.0f} frames/s)  equal={torch.equal(f, g)}"
    print(line, flush=True)
