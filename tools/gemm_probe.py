"""Not a test: attributes GEMM time to TMA feed / MMA issue / epilogue.  Needs a probe build of the library:
CLIPPPO_BUILD_PROBES=1 python clip-ppo_b200/build.py --force; python tools/gemm_probe.py [M]

For each tower GEMM shape: the product kernel, the same kernel with the epilogue reduced to a TMEM
drain (dbg 1), without TMA loads (dbg 2), with neither (dbg 3 = pure tcgen05 issue rate), and cuBLAS
(torch.matmul, no epilogue) for reference.  CUDA events, 20 launches after 3 warm-ups.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from clip_ppo_b200 import _native as N

L = N.lib()
st = torch.cuda.current_stream().cuda_stream
M = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 50100
gen = torch.Generator(device="cuda").manual_seed(0)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


print(f"M={M}  (us per launch | TFLOP/s)")
LEGACY = "--legacy" in sys.argv
SHAPES = ((("qkv", 2304, 768, 0), ("fc+gelu", 3072, 768, 1), ("proj+res", 768, 3072, 2), ("out+res", 768, 768, 2)) if LEGACY else
          (("qkv+ln", 2304, 768, 6), ("fc+ln+gelu", 3072, 768, 7), ("proj+res16", 768, 3072, 8), ("out+res16", 768, 768, 8)))
for name, Nn, K, epi in SHAPES:
    a = (torch.randn(M, K, device="cuda", generator=gen) * 0.5).bfloat16()
    w = (torch.randn(Nn, K, device="cuda", generator=gen) * (K ** -0.5)).bfloat16()
    bias = torch.randn(Nn, device="cuda", generator=gen) * 0.1
    out = torch.zeros(M, Nn, device="cuda", dtype=torch.float32 if epi == 2 else torch.bfloat16)
    stats = torch.stack([torch.zeros(M, device="cuda"), torch.ones(M, device="cuda")], 1).contiguous()
    colsum = w.float().sum(1).contiguous()
    fl = 2.0 * M * Nn * K

    def prod():
        if epi >= 6:
            N.check(L.clipppo_gemm_bf16_fused(a.data_ptr(), w.data_ptr(), M, Nn, K, epi, bias.data_ptr(),
                                              stats.data_ptr() if epi < 8 else None, colsum.data_ptr() if epi < 8 else None,
                                              out.data_ptr(), Nn, st))
        else:
            N.check(L.clipppo_gemm_bf16(a.data_ptr(), w.data_ptr(), M, Nn, K, epi, bias.data_ptr(), None, 0, out.data_ptr(), Nn, st))

    def probe(d):
        return lambda: N.check(L.clipppo_gemm_bf16_probe(a.data_ptr(), w.data_ptr(), M, Nn, K, epi, bias.data_ptr(),
                                                         out.data_ptr(), Nn, d, st))

    cols = [("product", timeit(prod))]
    for d, label in ((1, "no-epilogue"), (2, "no-TMA"), (3, "MMA-only")) + (((4, "no-output"), (8, "math-only")) if epi >= 6 else ()):
        cols.append((label, timeit(probe(d))))
    cols.append(("cuBLAS", timeit(lambda: torch.matmul(a, w.t()))))
    print(f"{name:11s} N={Nn:4d} K={K:4d}  " + "  ".join(f"{lb} {us:7.1f} | {fl / us / 1e6:6.0f}" for lb, us in cols), flush=True)
