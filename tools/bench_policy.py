"""Not a test: the native NatureCNN encoder (csrc/policy.cu) against PyTorch eager on the same GPU, forward + backward.
python tools/bench_policy.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn

from clip_ppo_b200.policy import NatureCNN


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for C, mb in ((3, 256), (3, 2048), (4, 8192)):
    torch.manual_seed(0)
    seq = nn.Sequential(nn.Conv2d(C, 32, 8, stride=4), nn.ReLU(), nn.Conv2d(32, 64, 4, stride=2), nn.ReLU(),
                        nn.Conv2d(64, 64, 3, stride=1), nn.ReLU(), nn.Flatten(), nn.Linear(64 * 7 * 7, 512), nn.ReLU()).cuda()
    net = NatureCNN.from_sequential(seq)
    x = torch.rand(mb, C, 84, 84, device="cuda")
    gh = torch.randn(mb, 512, device="cuda")

    def run(m):
        m.zero_grad(set_to_none=True)
        h = m(x)
        h.backward(gh)

    flops = mb * (2.0 * (400 * 32 * C * 64 + 81 * 64 * 512 + 49 * 64 * 576 + 512 * 3136)) * 3      # fwd + wgrad + dgrad (upper bound)
    rows = []
    for name, tf32 in (("eager fp32 (TF32 off)", False), ("eager default (cuDNN TF32 convs)", True)):
        torch.backends.cudnn.allow_tf32 = tf32
        rows.append((name, timeit(lambda: run(seq))))
    rows.append(("native fp32 (csrc/policy.cu)", timeit(lambda: run(net))))
    with torch.no_grad():
        fwd_e = timeit(lambda: seq(x)); fwd_n = timeit(lambda: net(x))
    print(f"C={C} mb={mb}: " + "  ".join(f"{n}: {t:.3f} ms ({mb / t * 1e3:.0f} frames/s)" for n, t in rows) +
          f"   forward only: eager {fwd_e:.3f} ms, native {fwd_n:.3f} ms   [{flops / rows[-1][1] / 1e9:.1f} TFLOP/s native]")
