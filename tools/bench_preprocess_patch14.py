import sys, torch
sys.path.insert(0, "/root/repo")
from clip_ppo_b200 import _native as N
n = 1024
img = torch.rand(n, 3, 224, 224, device="cuda")
for P, kpad in ((32, 3072), (14, 640)):
    G = 224 // P
    out = torch.empty(n * G * G, kpad, device="cuda", dtype=torch.bfloat16)
    f = lambda: N.check(N.lib().clipppo_preprocess_bf16(img.data_ptr(), 0, N.strides4(img), n, 3, 224, 224, 1.0, 1, P, 224, out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    print(f"preprocess 224x224 identity, patch {P}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us for {n} frames")
