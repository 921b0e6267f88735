"""Not a test: a diagnostic run for the GPU box.  python tools/gpu_probe.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from clip_ppo_b200 import _native as N

L = N.lib()
print(torch.cuda.get_device_name(0), torch.version.cuda)
st = torch.cuda.current_stream().cuda_stream


def gemm(M, Nn, K, epi=4, check=True, iters=0):
    gen = torch.Generator(device="cuda").manual_seed(1)
    a = (torch.randn(M, K, device="cuda", generator=gen) * 0.5).bfloat16()
    w = (torch.randn(Nn, K, device="cuda", generator=gen) * (K ** -0.5)).bfloat16()
    bias = torch.randn(Nn, device="cuda", generator=gen) * 0.1
    out = torch.zeros(M, Nn, device="cuda", dtype=torch.bfloat16 if epi in (0, 1) else torch.float32)
    rc = L.clipppo_gemm_bf16(a.data_ptr(), w.data_ptr(), M, Nn, K, epi, bias.data_ptr(), None, 0, out.data_ptr(), Nn, st)
    torch.cuda.synchronize()
    msg = f"gemm M={M} N={Nn} K={K} epi={epi} rc={rc}"
    if check:
        ref = a.float() @ w.float().t()
        if epi in (0, 1, 2):
            ref = ref + bias
        if epi == 1:
            ref = ref * torch.sigmoid(1.702 * ref)
        diff = (out.float() - ref).abs()
        msg += f" max_err={diff.max().item():.4e} mean_err={diff.mean().item():.3e} ref_absmean={ref.abs().mean().item():.3f}"
        if diff.max().item() > 0.05:
            bad = (diff > 0.05).nonzero()
            msg += f" BAD n={len(bad)} first={bad[:4].tolist()} rows_bad={(diff > 0.05).any(1).sum().item()} cols_bad={(diff > 0.05).any(0).sum().item()}"
    if iters:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            L.clipppo_gemm_bf16(a.data_ptr(), w.data_ptr(), M, Nn, K, epi, bias.data_ptr(), None, 0, out.data_ptr(), Nn, st)
        e0.record()
        for _ in range(iters):
            L.clipppo_gemm_bf16(a.data_ptr(), w.data_ptr(), M, Nn, K, epi, bias.data_ptr(), None, 0, out.data_ptr(), Nn, st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        msg += f" | {ms:.3f} ms  {2.0 * M * Nn * K / ms / 1e9:.1f} TFLOP/s"
        t0 = time.time()
        for _ in range(iters):
            torch.matmul(a, w.t())
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            torch.matmul(a, w.t())
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / iters
        msg += f" | cuBLAS {ms2:.3f} ms {2.0 * M * Nn * K / ms2 / 1e9:.1f} TFLOP/s"
    print(msg, flush=True)


if __name__ == "__main__":
    gemm(128, 256, 64)
    gemm(128, 256, 256)
    gemm(256, 512, 768)
    gemm(1000, 768, 768, epi=0)
    for epi in (0, 1, 2, 4):
        gemm(25600, 2304 if epi == 0 else (3072 if epi == 1 else 768), 768, epi=epi, iters=20)
    gemm(25600, 768, 3072, epi=2, iters=20)
    gemm(204800, 3072, 768, epi=1, check=False, iters=5)
    gemm(204800, 768, 3072, epi=2, check=False, iters=5)
