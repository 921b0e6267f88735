"""Randomized parity sweep of the round-2 kernels - tcgen05 short-sequence attention (pairs / causal single tile), the RESID_STATS
GEMM epilogue and its partial-sum consumer, the disturbance kernel's out_scale and in-kernel noise - against torch fp32 and the
numpy Philox oracle on random shapes, with guard regions around every output.  Two fixed seeds; also runnable by hand for a
longer sweep:  python tests/test_fuzz_gpu.py [cases per kernel] [seed]"""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def sweep(cases: int, seed: int) -> list:
    from clip_ppo_b200 import _native as N
    from oracle import philox as PH
    from shared.disturbances_gpu import DisturbanceWrapperGPU
    from shared.disturbance_types import DisturbanceSeverity
    CASES = cases
    rng = np.random.RandomState(seed)
    L = N.lib()
    st = torch.cuda.current_stream().cuda_stream
    failures = []

    def report(kind, desc, ok, detail=""):
        if not ok:
            failures.append(f"{kind} {desc} {detail}")

    # ---- attention: T <= 64 pairs, causal T <= 128 ----
    for _ in range(CASES):
        causal = bool(rng.randint(0, 2))
        T = int(rng.randint(16, 129 if causal else 65))
        H = int(rng.choice([8, 12, 16]))
        n = int(rng.choice([1, 2, 3, 5, 8, 31, 64, 150, 301]))
        Dm = H * 64
        g = torch.Generator(device="cuda").manual_seed(int(rng.randint(1 << 30)))
        qkv = (torch.randn(n * T, 3 * Dm, device="cuda", generator=g) * float(rng.uniform(0.3, 2.0))).bfloat16()
        buf = torch.full((n * T + 64, Dm), 9.0, device="cuda", dtype=torch.bfloat16)
        out = buf[32:32 + n * T]
        fn = L.clipppo_attention_causal_bf16 if causal else L.clipppo_attention_bf16
        N.check(fn(qkv.data_ptr(), n, T, H, 64, out.data_ptr(), st))
        q, k, v = qkv.float().reshape(n, T, 3, H, 64).permute(2, 0, 3, 1, 4)
        s = q @ k.transpose(-1, -2) / 8.0
        if causal:
            s = s + torch.full((T, T), float("-inf"), device="cuda").triu(1)
        ref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(n * T, Dm)
        err = (out.float() - ref).abs().max().item()
        guard = bool((buf[:32] == 9.0).all() and (buf[32 + n * T:] == 9.0).all())
        report("attention", f"n={n} T={T} H={H} causal={causal}", err <= 4e-2 * max(1.0, ref.abs().max().item()) and guard, f"err {err:.4f} guard {guard}")

    # ---- RESID_STATS GEMM + partial-sum consumer ----
    for _ in range(CASES):
        M = int(rng.choice([1, 7, 50, 127, 128, 129, 255, 256, 257, 1000, 3200, 6401, 20000]))
        Nn = int(rng.choice([64, 128, 192, 256, 512, 768, 1024]))
        K = int(rng.choice([64, 128, 512, 768, 1024, 3072]))
        g = torch.Generator(device="cuda").manual_seed(int(rng.randint(1 << 30)))
        a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
        w = (torch.randn(Nn, K, device="cuda", generator=g) * (K ** -0.5)).bfloat16()
        bias = torch.randn(Nn, device="cuda", generator=g) * 0.1
        x0 = (torch.randn(M, Nn, device="cuda", generator=g) * 2 + 0.3).bfloat16()
        P = (Nn + 127) // 128
        xbuf = torch.full((M + 160, Nn), 5.5, device="cuda", dtype=torch.bfloat16)
        xbuf[:M] = x0
        parts = torch.full((M + 160, P, 2), -3.0, device="cuda")
        N.check(L.clipppo_gemm_bf16_resid_stats(a.data_ptr(), w.data_ptr(), M, Nn, K, bias.data_ptr(), xbuf.data_ptr(), Nn, parts.data_ptr(), st))
        ref = x0.float() + a.float() @ w.float().t() + bias
        X = xbuf[:M].float()
        err = (X - ref).abs().max().item()
        pad = P * 128 - Nn
        xp = torch.nn.functional.pad(X, (0, pad)).view(M, P, 128)
        ok = err <= 2 ** -6 * max(1.0, ref.abs().max().item())
        ok &= bool(torch.allclose(parts[:M, :, 0], xp.sum(2), atol=3e-3, rtol=1e-5) and torch.allclose(parts[:M, :, 1], (xp * xp).sum(2), atol=3e-3, rtol=1e-5))
        ok &= bool((xbuf[M:] == 5.5).all() and (parts[M:] == -3.0).all())
        report("resid_stats", f"M={M} N={Nn} K={K}", ok, f"err {err:.4f}")
        if Nn % 128 == 0 and M >= 2:                      # feed the partial sums to the folded-LayerNorm epilogue of a second GEMM
            N2 = int(rng.choice([256, 768]))
            W2 = (torch.randn(N2, Nn, device="cuda", generator=g) * (Nn ** -0.5))
            gamma = 1 + 0.1 * torch.randn(Nn, device="cuda", generator=g)
            beta = 0.1 * torch.randn(Nn, device="cuda", generator=g)
            b2 = torch.randn(N2, device="cuda", generator=g) * 0.1
            Wf = (W2 * gamma).bfloat16()
            colsum = Wf.float().sum(1).contiguous()
            bias2 = (b2 + W2 @ beta).contiguous()
            o = torch.empty(M, N2, device="cuda", dtype=torch.bfloat16)
            xs = xbuf[:M].contiguous()
            N.check(L.clipppo_gemm_bf16_fused_parts(xs.data_ptr(), Wf.data_ptr(), M, N2, Nn, 6, bias2.data_ptr(), parts[:M].contiguous().data_ptr(), P,
                                                    colsum.data_ptr(), o.data_ptr(), N2, st))
            ref2 = torch.nn.functional.layer_norm(xs.float(), (Nn,), gamma, beta, 1e-5) @ W2.t() + b2
            e2 = (o.float() - ref2).abs()
            report("rowaffine_parts", f"M={M} K={Nn} N={N2}", e2.max().item() <= 6e-2 * max(1.0, ref2.abs().max().item() / 4) and e2.mean().item() <= 6e-3,
                   f"max {e2.max().item():.4f} mean {e2.mean().item():.5f}")

    # ---- disturbance: out_scale bitwise, in-kernel noise vs its oracle ----
    for _ in range(max(8, CASES // 3)):
        sev = str(rng.choice(["MILD", "MODERATE", "HARD", "SEVERE"]))
        hw = int(rng.choice([84, 224]))
        C = int(rng.choice([1, 3]))
        B = int(rng.choice([1, 2, 9, 33, 150])) if hw == 84 else int(rng.choice([1, 2, 5]))
        u8 = bool(rng.randint(0, 2))
        g = torch.Generator().manual_seed(int(rng.randint(1 << 30)))
        x = torch.randint(0, 256, (B, C, hw, hw), generator=g, dtype=torch.uint8).cuda()
        if not u8:
            x = x.float() / 255.0
        w = DisturbanceWrapperGPU(device="cuda", severity=DisturbanceSeverity[sev])
        noise = torch.randn(B, C, hw, hw, generator=g).cuda()
        kw = dict(contrast_factor=float(rng.uniform(0.7, 1.3)), cutout_start=(int(rng.randint(0, 20)), int(rng.randint(0, 20))))
        a = w.apply_disturbances(x, noise=noise, **kw) * 255.0
        b = w.apply_disturbances(x, noise=noise, out_scale=255.0, **kw)
        report("out_scale", f"{sev} {B}x{C}x{hw} u8={u8}", torch.equal(a, b))
        seed = int(rng.randint(1 << 62))
        o1 = w.apply_disturbances(x, noise_seed=seed, **kw)
        nz = torch.from_numpy(PH.normal_noise(seed, w._philox_calls - 1, (B, C, hw, hw))).cuda()
        o2 = w.apply_disturbances(x, noise=nz, **kw)
        e = (o1 - o2).abs().max().item()
        report("philox", f"{sev} {B}x{C}x{hw} u8={u8}", e <= 1e-5, f"err {e:.2e}")

    torch.cuda.synchronize()
    return failures


@pytest.mark.parametrize("seed", [1, 2])
def test_random_shapes_of_the_round2_kernels(native, seed):
    failures = sweep(24, seed)
    assert not failures, failures


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    f = sweep(int(sys.argv[1]) if len(sys.argv) > 1 else 60, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    print("\n".join(f))
    print(f"fuzz done: {len(f)} failures")
