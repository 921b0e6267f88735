"""Not a test: CUDA-event timing of the tower's GEMM shapes with the epilogues the tower runs (6 QKV + ln_1, 7 c_fc + ln_2 +
QuickGELU, 9 residual + row statistics, 8 residual as a TMA reduce-add), M = 204 800 (4096 images) by default.
    python tools/bench_gemm_shapes.py [M] [iters] [width]
Short bursts (boost clock) - the in-step figures are in bench.py's roofline.by_shape."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from clip_ppo_b200 import _native as N

L = N.lib()
st = torch.cuda.current_stream().cuda_stream
M = int(sys.argv[1]) if len(sys.argv) > 1 else 204800
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
D = int(sys.argv[3]) if len(sys.argv) > 3 else 768
gen = torch.Generator(device="cuda").manual_seed(0)
print(f"M={M} D={D} env: L2PF={os.environ.get('CLIPPPO_GEMM_L2PF', '1')}")
for (Nn, K, epi) in ((3 * D, D, 6), (4 * D, D, 7), (D, 4 * D, 9), (D, D, 9), (D, 4 * D, 8), (D, D, 8)):
    a = (torch.randn(M, K, device="cuda", generator=gen) * 0.5).bfloat16()
    w = (torch.randn(Nn, K, device="cuda", generator=gen) * (K ** -0.5)).bfloat16()
    bias = torch.randn(Nn, device="cuda", generator=gen) * 0.1
    out = torch.zeros(M, Nn, device="cuda", dtype=torch.bfloat16)
    stats = torch.stack([torch.zeros(M, device="cuda"), torch.ones(M, device="cuda")], 1).contiguous()
    colsum = w.float().sum(1).contiguous()
    parts = torch.empty(M, (Nn + 127) // 128, 2, device="cuda")

    def run():
        if epi == 9:
            N.check(L.clipppo_gemm_bf16_resid_stats(a.data_ptr(), w.data_ptr(), M, Nn, K, bias.data_ptr(), out.data_ptr(), Nn,
                                                    parts.data_ptr(), st))
        else:
            N.check(L.clipppo_gemm_bf16_fused(a.data_ptr(), w.data_ptr(), M, Nn, K, epi, bias.data_ptr(),
                                              stats.data_ptr() if epi < 8 else None, colsum.data_ptr() if epi < 8 else None,
                                              out.data_ptr(), Nn, st))
    for _ in range(3):
        run()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    med = ts[len(ts) // 2]
    print(f"epi {epi} N={Nn:5d} K={K:5d}: median {med:8.1f} us  best {ts[0]:8.1f} us  {2.0 * M * Nn * K / med / 1e6:7.1f} TFLOP/s (median)")
    del a, w, out
    torch.cuda.synchronize()
    import time; time.sleep(0.5)
