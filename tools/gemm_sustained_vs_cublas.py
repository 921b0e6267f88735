"""Not a test: the tower's four block GEMM shapes, SUSTAINED (each looped for ~1.5 s, i.e. under the 1 kW power cap the bench
step runs in), this library's fused-epilogue tcgen05 kernel against cuBLAS (torch.matmul, bf16, no epilogue at all) and
against the 8192^3 cuBLAS figure MEASURED_PEAKS.json calls the sustained peak.  python tools/gemm_sustained_vs_cublas.py"""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from clip_ppo_b200 import _native as N

L = N.lib()
st = torch.cuda.current_stream().cuda_stream
M = 204800                       # 4096 images x 50 tokens: the bench's launch size
dev = "cuda"


def sustained(fn, flops, seconds=1.5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    n, t0 = 0, time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # find a batch of launches worth ~50 ms, then loop batches until `seconds` have passed; time the LAST second
    fn(); torch.cuda.synchronize()
    t1 = time.perf_counter(); fn(); torch.cuda.synchronize(); per = max(time.perf_counter() - t1, 1e-5)
    batch = max(1, int(0.05 / per))
    total, tstart = 0.0, time.perf_counter()
    last = []
    while time.perf_counter() - tstart < seconds:
        a.record()
        for _ in range(batch):
            fn()
        b.record(); torch.cuda.synchronize()
        last.append(a.elapsed_time(b) / batch)
    tail = last[len(last) // 2:]                 # second half: the cap has settled
    ms = sum(tail) / len(tail)
    clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.strip()
    return flops / ms / 1e9, ms, clk


print(f"M = {M}; TFLOP/s sustained (second half of a {1.5} s loop)   [SM MHz, W right after]")
x8 = torch.randn(8192, 8192, device=dev).bfloat16(); y8 = torch.randn(8192, 8192, device=dev).bfloat16()
tf, ms, clk = sustained(lambda: torch.matmul(x8, y8.t()), 2.0 * 8192 ** 3)
print(f"cuBLAS 8192^3 (the 'sustained peak' shape): {tf:7.1f}   [{clk}]")
del x8, y8
for name, Nn, K, epi in (("QKV  + ln_1 folded", 2304, 768, N.EPI_ROWAFFINE_BF16), ("c_fc + ln_2 + QuickGELU", 3072, 768, N.EPI_ROWAFFINE_GELU_BF16),
                         ("c_proj += residual", 768, 3072, N.EPI_RESID_BF16), ("out_proj += residual", 768, 768, N.EPI_RESID_BF16)):
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(Nn, K, device=dev) * 0.03).bfloat16()
    bias = torch.randn(Nn, device=dev)
    stats = torch.stack([torch.zeros(M, device=dev), torch.ones(M, device=dev)], 1).contiguous()
    colsum = torch.randn(Nn, device=dev)
    out = torch.zeros(M, Nn, device=dev, dtype=torch.bfloat16)
    flops = 2.0 * M * Nn * K
    if epi == N.EPI_RESID_BF16:
        mine = lambda: N.check(L.clipppo_gemm_bf16_fused(a.data_ptr(), w.data_ptr(), M, Nn, K, epi, bias.data_ptr(), None, None, out.data_ptr(), Nn, st))
    else:
        mine = lambda: N.check(L.clipppo_gemm_bf16_fused(a.data_ptr(), w.data_ptr(), M, Nn, K, epi, bias.data_ptr(), stats.data_ptr(), colsum.data_ptr(), out.data_ptr(), Nn, st))
    tf_m, _, clk_m = sustained(mine, flops)
    wt = w.t()
    tf_c, _, clk_c = sustained(lambda: torch.matmul(a, wt, out=out), flops)
    print(f"{name:26s} N={Nn:5d} K={K:5d}: this kernel (fused epilogue) {tf_m:7.1f} [{clk_m}]   cuBLAS (plain GEMM) {tf_c:7.1f} [{clk_c}]   ratio {tf_m / tf_c:.3f}")
