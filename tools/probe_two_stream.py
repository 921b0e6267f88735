"""Not a test: does running the two halves of a 4096-frame tower pass on two streams (one half's attention / preprocess /
LayerNorm under the other half's GEMMs) beat one pass over the whole batch?  VERDICT r01 item 1(a).
    CLIPPPO_GEMM_MAX_SMS=132 python tools/probe_two_stream.py      # GEMM grids capped, 16 SMs left to the other stream
Sustained timing (the part is at its power cap after ~1 s): 3 x 10 passes each way, interleaved."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CLIPPPO_ALLOW_RANDOM_WEIGHTS", "1")
import torch

import shared.clip_ppo_utils as U

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
model = U.load_clip_model("ViT-B/32", device=dev)
eng = U._engine_for(model)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.rand(B, 3, 224, 224, device=dev, generator=g) * 255.0
halves = (x[: B // 2], x[B // 2:])
s = [torch.cuda.Stream(device=dev) for _ in range(2)]


def one_stream():
    return eng.encode(x, pre_scale=1 / 255.0, l2norm=True)


def two_streams():
    cur = torch.cuda.current_stream(dev)
    outs = []
    for st, h in zip(s, halves):
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            outs.append(eng.encode(h, pre_scale=1 / 255.0, l2norm=True))
    for st in s:
        cur.wait_stream(st)
    return torch.cat(outs)


def timed(fn, iters=10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


ref = one_stream()
two = two_streams()
torch.cuda.synchronize()
print("max SMs per GEMM:", os.environ.get("CLIPPPO_GEMM_MAX_SMS", "148"), " two-stream result equals one-stream:", torch.equal(ref, two))
for _ in range(5):
    one_stream(); two_streams()
for rep in range(3):
    t1 = timed(one_stream)
    t2 = timed(two_streams)
    print(f"rep {rep}: one stream {t1:7.2f} ms   two streams (2 x {B // 2}) {t2:7.2f} ms")
