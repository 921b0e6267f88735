"""Parity of the tower and its building blocks.  bf16 operands / fp32 accumulate vs the fp32
oracle: embeddings must reach cosine >= 0.999 (north_star); building blocks are checked against
plain torch fp32 references of the same op on bf16-rounded inputs."""
import ctypes
import math
import os

import numpy as np
import pytest
import torch

from oracle import vit as ov

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _stream():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 128), (1000, 768, 768), (6400, 2304, 768),
                                   (6400, 768, 3072), (50, 512, 768), (12800, 3072, 768), (19000, 768, 768)])
@pytest.mark.parametrize("epi", [0, 1, 2, 4])
def test_gemm_epilogues_vs_torch(native, M, N, K, epi):
    from clip_ppo_b200 import _native as Nn
    gen = torch.Generator(device="cuda").manual_seed(M + N + K + epi)
    a = (torch.randn(M, K, device="cuda", generator=gen) * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=gen) * (K ** -0.5)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=gen) * 0.1
    acc = a.float() @ w.float().t()
    if epi in (0, 1):
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ref = acc + bias
        if epi == 1:
            ref = ref * torch.sigmoid(1.702 * ref)
    elif epi == 2:
        x0 = torch.randn(M, N, device="cuda", generator=gen)
        out = x0.clone()
        ref = x0 + acc + bias
    else:
        out = torch.empty(M, N, device="cuda", dtype=torch.float32)
        ref = acc
    st = native.clipppo_gemm_bf16(a.data_ptr(), w.data_ptr(), M, N, K, epi, bias.data_ptr(), None, 0, out.data_ptr(), N, _stream())
    Nn.check(st)
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    tol = 2e-2 if epi in (0, 1) else 2e-3            # bf16 output rounding vs fp32 accumulate-order noise
    assert err <= tol, err


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (1000, 768, 768), (6400, 2304, 768), (12800, 3072, 768), (77, 64, 128)])
@pytest.mark.parametrize("epi,with_stats", [(6, True), (6, False), (7, True)])
def test_gemm_rowaffine_epilogue_is_folded_layernorm(native, M, N, K, epi, with_stats):
    """ROWAFFINE (ln_1 / ln_2 folded through the GEMM, [clip] ResidualAttentionBlock): the kernel output must
    equal LayerNorm(x) @ W^T + b computed the ordinary way in fp32 on the same bf16-rounded x."""
    from clip_ppo_b200 import _native as Nn
    gen = torch.Generator(device="cuda").manual_seed(M + N + K + epi)
    x = (torch.randn(M, K, device="cuda", generator=gen) * 1.5 + 0.7).bfloat16()          # un-normalised rows, non-zero mean
    W = torch.randn(N, K, device="cuda", generator=gen) * (K ** -0.5)
    b = torch.randn(N, device="cuda", generator=gen) * 0.1
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    if with_stats:
        gamma = 1 + 0.1 * torch.randn(K, device="cuda", generator=gen)
        beta = 0.1 * torch.randn(K, device="cuda", generator=gen)
        Wf = (W * gamma).bfloat16()
        colsum = Wf.float().sum(1).contiguous()
        bias2 = (b + W @ beta).contiguous()
        xf = x.float()
        mean = xf.mean(1)
        rstd = torch.rsqrt(xf.var(1, unbiased=False) + 1e-5)
        stats = torch.stack([mean, rstd], 1).contiguous()
        ref = torch.nn.functional.layer_norm(xf, (K,), gamma, beta, 1e-5) @ W.t() + b
        st = native.clipppo_gemm_bf16_fused(x.data_ptr(), Wf.data_ptr(), M, N, K, epi, bias2.data_ptr(), stats.data_ptr(),
                                            colsum.data_ptr(), out.data_ptr(), N, _stream())
        tol = 4e-2               # bf16 rounding of W*gamma (2^-9 relative per weight) + bf16 output
    else:
        Wf = W.bfloat16()
        ref = x.float() @ Wf.float().t() + b
        st = native.clipppo_gemm_bf16_fused(x.data_ptr(), Wf.data_ptr(), M, N, K, epi, b.data_ptr(), None, None,
                                            out.data_ptr(), N, _stream())
        tol = 3e-2
    Nn.check(st)
    if epi == 7:
        ref = ref * torch.sigmoid(1.702 * ref)
    torch.cuda.synchronize()
    err = (out.float() - ref).abs()
    assert err.max().item() <= tol * max(1.0, ref.abs().max().item() / 4), err.max().item()
    assert err.mean().item() <= 4e-3


@pytest.mark.parametrize("M,N,K,split", [(128, 256, 64, False), (1000, 768, 768, True), (6400, 768, 3072, False),
                                         (19000, 768, 768, False), (50, 64, 64, False),
                                         (400, 768, 3072, True), (3200, 768, 3072, True), (3200, 1024, 4096, True)])
def test_gemm_bf16_residual_reduce_add(native, monkeypatch, M, N, K, split):
    """RESID_BF16: X (bf16, in place) += bf16(acc + bias), added by the TMA unit in L2.  split=True: with
    CLIPPPO_GEMM_KSPLIT=auto the K range of a small-M tile is cut into slices that each add a bf16-rounded
    partial sum (opt-in latency schedule for FROZEN_CLIP policy batches)."""
    if split:
        monkeypatch.setenv("CLIPPPO_GEMM_KSPLIT", "auto")
    from clip_ppo_b200 import _native as Nn
    gen = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = (torch.randn(M, K, device="cuda", generator=gen) * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=gen) * (K ** -0.5)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=gen) * 0.1
    x0 = torch.randn(M, N, device="cuda", generator=gen).bfloat16()
    X = x0.clone()
    Nn.check(native.clipppo_gemm_bf16_fused(a.data_ptr(), w.data_ptr(), M, N, K, 8, bias.data_ptr(), None, None,
                                            X.data_ptr(), N, _stream()))
    torch.cuda.synchronize()
    delta = (a.float() @ w.float().t() + bias).bfloat16()
    ref = (x0.float() + delta.float()).bfloat16()
    # two bf16 roundings (the delta, then the sum); allow one bf16 ulp of the result for accumulate-order noise
    err = (X.float() - ref.float()).abs()
    if split:     # one more bf16 rounding per K slice: bounded by a few ulps, and unbiased
        ref32 = x0.float() + a.float() @ w.float().t() + bias
        err32 = (X.float() - ref32).abs()
        assert err32.max().item() <= 2 ** -6 * max(1.0, ref32.abs().max().item()), err32.max().item()
        assert err32.mean().item() <= 6e-3
        assert abs((X.float() - ref32).mean().item()) <= 2e-4
    else:
        assert err.max().item() <= 2 ** -7 * max(1.0, ref.float().abs().max().item()), err.max().item()
        assert (err > 0).float().mean().item() < 0.05


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (1000, 768, 768), (6400, 768, 3072), (19000, 768, 768), (50, 64, 64),
                                   (3200, 1024, 4096), (77 * 9, 512, 2048), (40000, 768, 768)])
def test_gemm_resid_stats_epilogue(native, M, N, K):
    """RESID_STATS: X (bf16, in place) = bf16(X + A W^T + bias) with one rounding, and the (sum, sum of squares) of every
    updated row over each 128-column slice - what is left of the LayerNorm that follows the residual add
    ([clip] ResidualAttentionBlock: ln_2 after the attention branch, the next block's ln_1 after the MLP)."""
    from clip_ppo_b200 import _native as Nn
    gen = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = (torch.randn(M, K, device="cuda", generator=gen) * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=gen) * (K ** -0.5)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=gen) * 0.1
    x0 = (torch.randn(M, N, device="cuda", generator=gen) * 2 + 0.3).bfloat16()
    X = x0.clone()
    P = (N + 127) // 128
    parts = torch.full((M, P, 2), float("nan"), device="cuda")
    Nn.check(native.clipppo_gemm_bf16_resid_stats(a.data_ptr(), w.data_ptr(), M, N, K, bias.data_ptr(), X.data_ptr(), N,
                                                  parts.data_ptr(), _stream()))
    torch.cuda.synchronize()
    ref32 = x0.float() + a.float() @ w.float().t() + bias
    ref = ref32.bfloat16()
    err = (X.float() - ref.float()).abs()
    # one rounding of the fp32 sum: differences only where accumulation order moves a value across a rounding boundary
    assert err.max().item() <= 2 ** -7 * max(1.0, ref32.abs().max().item()), err.max().item()
    assert (err > 0).float().mean().item() < 0.02
    # the statistics are those of the rows as STORED
    xs = X.float()
    assert not torch.isnan(parts).any()
    pad = P * 128 - N
    xp = torch.nn.functional.pad(xs, (0, pad)).view(M, P, 128)
    assert torch.allclose(parts[:, :, 0], xp.sum(2), atol=2e-3, rtol=1e-5)
    assert torch.allclose(parts[:, :, 1], (xp * xp).sum(2), atol=2e-3, rtol=1e-5)


@pytest.mark.parametrize("M,N,K,epi", [(1000, 768, 768, 6), (6400, 2304, 768, 6), (12800, 3072, 768, 7), (77 * 5, 1536, 512, 6)])
def test_gemm_rowaffine_from_partial_sums(native, M, N, K, epi):
    """The folded-LayerNorm epilogue fed with the partial sums of RESID_STATS must equal the same epilogue fed with
    two-pass (mean, rstd) up to fp32 rounding of the variance."""
    from clip_ppo_b200 import _native as Nn
    gen = torch.Generator(device="cuda").manual_seed(M + N + K + epi)
    x = (torch.randn(M, K, device="cuda", generator=gen) * 1.5 + 0.7).bfloat16()
    x[:, 3] *= 40.0                                                     # an outlier channel, as real CLIP streams have
    W = torch.randn(N, K, device="cuda", generator=gen) * (K ** -0.5)
    b = torch.randn(N, device="cuda", generator=gen) * 0.1
    gamma = 1 + 0.1 * torch.randn(K, device="cuda", generator=gen)
    beta = 0.1 * torch.randn(K, device="cuda", generator=gen)
    Wf = (W * gamma).bfloat16()
    colsum = Wf.float().sum(1).contiguous()
    bias2 = (b + W @ beta).contiguous()
    xf = x.float()
    P = K // 128
    xp = xf.view(M, P, 128)
    parts = torch.stack([xp.sum(2), (xp * xp).sum(2)], 2).contiguous()
    stats = torch.stack([xf.mean(1), torch.rsqrt(xf.var(1, unbiased=False) + 1e-5)], 1).contiguous()
    o1 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    o2 = torch.empty_like(o1)
    Nn.check(native.clipppo_gemm_bf16_fused(x.data_ptr(), Wf.data_ptr(), M, N, K, epi, bias2.data_ptr(), stats.data_ptr(),
                                            colsum.data_ptr(), o1.data_ptr(), N, _stream()))
    Nn.check(native.clipppo_gemm_bf16_fused_parts(x.data_ptr(), Wf.data_ptr(), M, N, K, epi, bias2.data_ptr(), parts.data_ptr(), P,
                                                  colsum.data_ptr(), o2.data_ptr(), N, _stream()))
    torch.cuda.synchronize()
    d = (o1.float() - o2.float()).abs()
    assert d.max().item() <= 2 ** -7 * max(1.0, o1.float().abs().max().item()), d.max().item()
    assert (d > 0).float().mean().item() < 0.01
    ref = torch.nn.functional.layer_norm(xf, (K,), gamma, beta, 1e-5) @ W.t() + b
    if epi == 7:
        ref = ref * torch.sigmoid(1.702 * ref)
    assert (o2.float() - ref).abs().mean().item() <= 4e-3


@pytest.mark.parametrize("rows,width", [(50, 768), (6401, 768), (257, 1024)])
def test_rowstats_vs_torch(native, rows, width):
    from clip_ppo_b200 import _native as Nn
    gen = torch.Generator(device="cuda").manual_seed(rows)
    x = (torch.randn(rows, width, device="cuda", generator=gen) * 3 + 0.5).bfloat16()
    stats = torch.empty(rows, 2, device="cuda")
    Nn.check(native.clipppo_rowstats_bf16(x.data_ptr(), rows, width, width, stats.data_ptr(), _stream()))
    xf = x.float()
    assert torch.allclose(stats[:, 0], xf.mean(1), atol=1e-5, rtol=1e-5)
    assert torch.allclose(stats[:, 1], torch.rsqrt(xf.var(1, unbiased=False) + 1e-5), atol=0, rtol=1e-5)


def test_gemm_patch_epilogue(native):
    from clip_ppo_b200 import _native as Nn
    n, G2, T, D, K = 5, 49, 50, 768, 3072
    gen = torch.Generator(device="cuda").manual_seed(1)
    a = (torch.randn(n * G2, K, device="cuda", generator=gen) * 0.5).bfloat16()
    w = (torch.randn(D, K, device="cuda", generator=gen) * (K ** -0.5)).bfloat16()
    pos = torch.randn(T, D, device="cuda", generator=gen)
    X = torch.full((n * T, D), 7.0, device="cuda")
    Nn.check(native.clipppo_gemm_bf16(a.data_ptr(), w.data_ptr(), n * G2, D, K, 3, None, pos.data_ptr(), T, X.data_ptr(), D, _stream()))
    torch.cuda.synchronize()
    ref = (a.float() @ w.float().t()).reshape(n, G2, D) + pos[1:]
    Xv = X.reshape(n, T, D)
    assert (Xv[:, 1:] - ref).abs().max().item() <= 2e-3
    assert torch.all(Xv[:, 0] == 7.0)                # CLS rows untouched by the patch GEMM


@pytest.mark.parametrize("rows,width", [(50, 768), (6401, 768), (257, 1024)])
def test_layernorm_vs_torch(native, rows, width):
    from clip_ppo_b200 import _native as Nn
    gen = torch.Generator(device="cuda").manual_seed(rows)
    x = torch.randn(rows, width, device="cuda", generator=gen) * 3 + 0.5
    g = 1 + 0.1 * torch.randn(width, device="cuda", generator=gen)
    b = 0.1 * torch.randn(width, device="cuda", generator=gen)
    y = torch.empty(rows, width, device="cuda", dtype=torch.bfloat16)
    Nn.check(native.clipppo_layernorm_bf16(x.data_ptr(), g.data_ptr(), b.data_ptr(), rows, width, width, y.data_ptr(), _stream()))
    ref = torch.nn.functional.layer_norm(x, (width,), g, b, 1e-5)
    assert (y.float() - ref).abs().max().item() <= 2e-2
    assert (y.float() - ref.bfloat16().float()).abs().mean().item() <= 1e-4


@pytest.mark.parametrize("n,T,H", [(1, 50, 12), (37, 50, 12), (3, 17, 12), (2, 64, 12),
                                   (2, 257, 16), (3, 65, 12), (1, 128, 16), (2, 200, 12)])      # T > 64: key-block kernel (ViT-L/14)
def test_attention_vs_torch(native, n, T, H):
    from clip_ppo_b200 import _native as Nn
    dh = 64
    D = H * dh
    gen = torch.Generator(device="cuda").manual_seed(n * 100 + T)
    qkv = torch.randn(n * T, 3 * D, device="cuda", generator=gen).bfloat16()
    out = torch.empty(n * T, D, device="cuda", dtype=torch.bfloat16)
    Nn.check(native.clipppo_attention_bf16(qkv.data_ptr(), n, T, H, dh, out.data_ptr(), _stream()))
    q, k, v = qkv.float().reshape(n, T, 3, H, dh).permute(2, 0, 3, 1, 4)
    s = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh), dim=-1)
    ref = (s @ v).permute(0, 2, 1, 3).reshape(n * T, D)
    assert (out.float() - ref).abs().max().item() <= 3e-2


@pytest.mark.parametrize("n,T,H", [(1, 50, 12), (2, 50, 12), (301, 50, 12), (7, 33, 12), (5, 64, 16), (4, 16, 12), (1200, 50, 12)])
def test_attention_pair_kernel_vs_torch_and_vs_the_mma_sync_kernel(native, monkeypatch, n, T, H):
    """T <= 64 (ViT-B/32: T = 50): the tcgen05 kernel that packs two images into one 128-row tile, against torch fp32 and
    against the mma.sync kernel it replaces - odd image counts (the second half of the last tile is zero-filled by the TMA
    unit and its stores are clipped), more items than CTAs x stages (barrier phase wrap), every token count class."""
    from clip_ppo_b200 import _native as Nn
    dh, D = 64, H * 64
    gen = torch.Generator(device="cuda").manual_seed(n * 131 + T)
    qkv = (torch.randn(n * T, 3 * D, device="cuda", generator=gen) * 1.5).bfloat16()
    out = torch.full((n * T, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    Nn.check(native.clipppo_attention_bf16(qkv.data_ptr(), n, T, H, dh, out.data_ptr(), _stream()))
    monkeypatch.setenv("CLIPPPO_ATT_TC50", "0")
    old = torch.empty_like(out)
    Nn.check(native.clipppo_attention_bf16(qkv.data_ptr(), n, T, H, dh, old.data_ptr(), _stream()))
    torch.cuda.synchronize()
    assert not torch.isnan(out.float()).any()
    q, k, v = qkv.float().reshape(n, T, 3, H, dh).permute(2, 0, 3, 1, 4)
    sm = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh), dim=-1)
    ref = (sm @ v).permute(0, 2, 1, 3).reshape(n * T, D)
    assert (out.float() - ref).abs().max().item() <= 3e-2
    assert (out.float() - ref).abs().mean().item() <= 2e-3
    # two bf16 kernels, each within 3e-2 of fp32 (P is rounded to bf16 before normalisation here, after it there)
    assert (out.float() - old.float()).abs().max().item() <= 6e-2
    assert (out.float() - old.float()).abs().mean().item() <= 2e-3


@pytest.mark.parametrize("kind,n,T,H", [("pair", 5, 50, 12), ("pair", 1, 50, 12), ("pair", 3, 17, 12), ("causal", 5, 77, 8), ("causal", 2, 128, 8),
                                         ("long", 3, 257, 16)])
def test_attention_kernels_write_nothing_outside_their_output(native, kind, n, T, H):
    """The tcgen05 attention kernels store through TMA boxes that are larger than what they own (32-row slabs, a second image
    that does not exist when n is odd): the TMA unit must clip them.  Guard rows around the output stay untouched."""
    from clip_ppo_b200 import _native as Nn
    D, G = H * 64, 40
    gen = torch.Generator(device="cuda").manual_seed(n + T)
    qkv = torch.randn(n * T, 3 * D, device="cuda", generator=gen).bfloat16()
    buf = torch.full(((n * T + 2 * G), D), 7.5, device="cuda", dtype=torch.bfloat16)
    out = buf[G:G + n * T]
    fn = native.clipppo_attention_causal_bf16 if kind == "causal" else native.clipppo_attention_bf16
    Nn.check(fn(qkv.data_ptr(), n, T, H, 64, out.data_ptr(), _stream()))
    torch.cuda.synchronize()
    assert (buf[:G] == 7.5).all() and (buf[G + n * T:] == 7.5).all()
    assert torch.isfinite(out.float()).all() and not (out == 7.5).all(dim=1).any()       # every row was written


@pytest.mark.parametrize("M,N,K", [(130, 768, 768), (50, 64, 64), (1000, 768, 3072), (257, 1024, 1024)])
def test_gemm_resid_stats_writes_nothing_outside_its_rows(native, M, N, K):
    """RESID_STATS loads and stores 32 x 64 boxes that overhang M: loads beyond M are zero-filled, stores clipped, and the partial
    statistics are written for rows < M only."""
    from clip_ppo_b200 import _native as Nn
    gen = torch.Generator(device="cuda").manual_seed(M + N)
    a = (torch.randn(M, K, device="cuda", generator=gen) * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=gen) * (K ** -0.5)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=gen) * 0.1
    G, P = 200, (N + 127) // 128
    xbuf = torch.full((M + G, N), 3.25, device="cuda", dtype=torch.bfloat16)
    pbuf = torch.full((M + G, P, 2), -77.0, device="cuda")
    x0 = torch.randn(M, N, device="cuda", generator=gen).bfloat16()
    xbuf[:M] = x0
    Nn.check(native.clipppo_gemm_bf16_resid_stats(a.data_ptr(), w.data_ptr(), M, N, K, bias.data_ptr(), xbuf.data_ptr(), N,
                                                  pbuf.data_ptr(), _stream()))
    torch.cuda.synchronize()
    assert (xbuf[M:] == 3.25).all() and (pbuf[M:] == -77.0).all()
    ref = (x0.float() + a.float() @ w.float().t() + bias)
    assert (xbuf[:M].float() - ref).abs().max().item() <= 2 ** -6 * max(1.0, ref.abs().max().item())
    assert (pbuf[:M] != -77.0).all()


@pytest.mark.parametrize("n,T,H,skew", [(2, 257, 16, 1), (2, 257, 16, 2), (3, 200, 12, 1), (2, 129, 8, 2), (150, 257, 16, 0)])
def test_attention_tc_running_shift(native, n, T, H, skew):
    """The tcgen05 kernel (64 < T <= 257) keeps ONE running softmax shift per row and only moves it when a later 32-key chunk
    beats it by more than 2^8: keys whose scores grow (skew 1) or shrink (skew 2) by whole factors along the sequence drive
    every path of that logic - shift moves inside a block, across blocks (accumulator row rescaled in TMEM), and on the
    tail key.  n = 150 with skew 0 runs more items than there are CTAs (both shared-memory stages, barrier phase wrap)."""
    from clip_ppo_b200 import _native as Nn
    dh, D = 64, H * 64
    gen = torch.Generator(device="cuda").manual_seed(7 * n + T + skew)
    qkv = torch.randn(n * T, 3 * D, device="cuda", generator=gen)
    if skew:
        t = torch.arange(T, device="cuda").repeat(n)
        f = (1.0 + 2.0 * (t // 64).float()) if skew == 1 else (1.0 + 2.0 * ((T - 1 - t) // 64).float())
        qkv[:, D:2 * D] *= f[:, None]
    qkv = qkv.bfloat16()
    out = torch.full((n * T, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    Nn.check(native.clipppo_attention_bf16(qkv.data_ptr(), n, T, H, dh, out.data_ptr(), _stream()))
    q, k, v = qkv.float().reshape(n, T, 3, H, dh).permute(2, 0, 3, 1, 4)
    s = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh), dim=-1)
    ref = (s @ v).permute(0, 2, 1, 3).reshape(n * T, D)
    assert not out.float().isnan().any()
    assert (out.float() - ref).abs().max().item() <= 3e-2


@pytest.mark.parametrize("shape,dtype,scale", [((3, 3, 84, 84), torch.float32, 1 / 255.0), ((2, 3, 224, 224), torch.float32, 1 / 255.0),
                                               ((2, 1, 84, 84), torch.float32, 1 / 255.0 / 255.0), ((3, 3, 84, 84), torch.uint8, 1 / 255.0),
                                               ((2, 3, 60, 100), torch.float32, 1.0)])
def test_preprocess_vs_oracle(native, shape, dtype, scale):
    from clip_ppo_b200 import _native as Nn
    gen = torch.Generator().manual_seed(sum(shape))
    raw = torch.randint(0, 256, shape, generator=gen)
    img = raw.to(dtype) if dtype == torch.uint8 else (raw.float() if scale != 1.0 else raw.float() / 255.0)
    n, C, h, w = shape
    dev = img.cuda()
    out = torch.empty(n * 49, 3072, device="cuda", dtype=torch.bfloat16)
    Nn.check(native.clipppo_preprocess_bf16(dev.data_ptr(), 1 if dtype == torch.uint8 else 0, Nn.strides4(dev), n, C, h, w,
                                            scale, 1, 32, 224, out.data_ptr(), _stream()))
    x = img.float() * scale
    if C == 1:
        x = x.repeat(1, 3, 1, 1)
    ref = ov.preprocess(x, False)                                        # antialias bilinear + normalise, fp32
    ref = ref.reshape(n, 3, 7, 32, 7, 32).permute(0, 2, 4, 1, 3, 5).reshape(n * 49, 3072)
    err = (out.float().cpu() - ref).abs()
    assert err.max().item() <= 2e-2 and err.mean().item() <= 3e-3        # bf16 rounding of values up to ~2.2


def _engine(seed=0):
    from clip_ppo_b200.clip_compat.model import random_visual_state_dict
    from clip_ppo_b200.vit import VitEngine
    return VitEngine(random_visual_state_dict("ViT-B/32", seed), device="cuda")


def test_embeddings_match_reference_golden(native):
    import shared.clip_ppo_utils as U
    g = np.load(os.path.join(GOLDEN, "vit_b32_seed0.npz"))
    model = U.load_clip_model("ViT-B/32", device="cuda")               # compat shim, seed-0 random weights
    for tag in ("84", "224"):
        img = torch.from_numpy(g[f"img{tag}"]).cuda().float()
        emb = U.generate_clip_embeddings(U.AblationMode.NONE, model, "image", img.shape[0], "cuda", images=img)
        ref = torch.from_numpy(g[f"emb{tag}"])
        assert emb.shape == ref.shape and emb.dtype == torch.float32
        cos = torch.sum(emb.cpu() * ref, dim=-1)
        assert cos.min().item() >= 0.999, cos
        assert torch.allclose(emb.norm(dim=-1).cpu(), torch.ones(len(ref)), atol=1e-5)
    # Atari path: gray frames already divided by 255 once by the caller (clip_ppo_atari.py:661)
    from clip_ppo_b200 import rollout
    gray = torch.from_numpy(g["gray_atari"]).cuda().float()          # [3,1,84,84]
    stack = gray.reshape(1, 3, 84, 84)                                 # pretend a 3-frame stack of one env
    rgb = rollout.convert_atari_frames_for_clip(stack) / 255.0
    e = rollout.process_multiframe_clip_embeddings(rgb, model, U.AblationMode.NONE, "image", 1, "cuda")
    assert e.shape == (1, 3 * 512)
    cos = torch.sum(e.reshape(3, 512).cpu() * torch.from_numpy(g["emb_atari"]), dim=-1)
    assert cos.min().item() >= 0.999, cos


@pytest.mark.parametrize("n,hw", [(1, 84), (37, 84), (130, 84), (5, 224)])
def test_embeddings_vs_oracle(native, n, hw):
    eng = _engine(0)
    sd = ov.random_state_dict(ov.VIT_B32, 0)
    gen = torch.Generator().manual_seed(n + hw)
    img = torch.randint(0, 256, (n, 3, hw, hw), generator=gen).float()
    emb = eng.encode(img.cuda(), pre_scale=1 / 255.0, l2norm=True).cpu()
    m = min(n, 6)                                                     # the CPU oracle is slow: check a subset
    idx = torch.linspace(0, n - 1, m).long()
    ref = ov.image_embeddings(sd, img[idx])
    cos = torch.sum(emb[idx] * ref, dim=-1)
    assert cos.min().item() >= 0.999, cos
    # batch-size invariance: an image encoded alone == encoded inside the batch (bitwise)
    alone = eng.encode(img[idx[-1:]].cuda(), pre_scale=1 / 255.0, l2norm=True).cpu()
    assert torch.equal(alone[0], emb[idx[-1]])


def test_vit_l14_embeddings_vs_oracle(native):
    """BASELINE configs[4] variant: ViT-L/14 (width 1024, 24 blocks, 16 heads, patch 14, 257 tokens, out 768)."""
    from clip_ppo_b200.clip_compat.model import random_visual_state_dict
    from clip_ppo_b200.vit import VitEngine
    eng = VitEngine(random_visual_state_dict("ViT-L/14", 0), device="cuda")
    sd = ov.random_state_dict(ov.VIT_L14, 0)
    gen = torch.Generator().manual_seed(14)
    img = torch.randint(0, 256, (3, 3, 84, 84), generator=gen).float()
    emb = eng.encode(img.cuda(), pre_scale=1 / 255.0, l2norm=True).cpu()
    assert emb.shape == (3, 768)
    ref = ov.image_embeddings(sd, img[:2])
    cos = torch.sum(emb[:2] * ref, dim=-1)
    assert cos.min().item() >= 0.999, cos
    alone = eng.encode(img[2:].cuda(), pre_scale=1 / 255.0, l2norm=True).cpu()
    assert torch.equal(alone[0], emb[2])


def test_small_batch_graph_replay_is_bitwise_eager(native):
    """FROZEN_CLIP policy path: E frames per call through a captured CUDA graph (one per call shape)."""
    eng = _engine(0)
    gen = torch.Generator().manual_seed(21)
    for n in (8, 64):
        outs = []
        for rep in range(3):
            x = torch.rand(n, 3, 84, 84, generator=gen).cuda()
            a = eng.encode(x, pre_scale=1.0, l2norm=False)
            b = eng.encode_graphed(x, pre_scale=1.0, l2norm=False)           # first call captures, later calls replay
            assert torch.equal(a, b)
            outs.append(b)
        assert not torch.equal(outs[0], outs[1])                             # fresh tensors, not views of the static output
    big = eng.encode(torch.rand(700, 3, 84, 84, generator=gen).cuda())       # a bigger eager call reallocates the shared workspace
    x = torch.rand(8, 3, 84, 84, generator=gen).cuda()
    assert torch.equal(eng.encode(x, pre_scale=1.0, l2norm=False), eng.encode_graphed(x, pre_scale=1.0, l2norm=False))
    assert len(eng._graphs) == 2 and torch.isfinite(big).all()


def test_chunked_batch_and_uint8_input(native, monkeypatch):
    monkeypatch.setenv("CLIPPPO_VIT_CHUNK", "500")      # force several tower passes: chunk boundaries at 500 / 1000
    eng = _engine(0)
    gen = torch.Generator().manual_seed(4)
    u8 = torch.randint(0, 256, (1500, 3, 84, 84), generator=gen, dtype=torch.uint8).cuda()    # three chunks of 500
    a = eng.encode(u8, pre_scale=1 / 255.0, l2norm=True)
    b = eng.encode(u8.float(), pre_scale=1 / 255.0, l2norm=True)
    assert torch.equal(a, b)
    for lo in (490, 740, 990):                          # slices that straddle / sit inside the chunk boundaries
        c = eng.encode(u8[lo:lo + 25], pre_scale=1 / 255.0, l2norm=True)
        assert torch.equal(a[lo:lo + 25], c)
    monkeypatch.delenv("CLIPPPO_VIT_CHUNK")
    whole = eng.encode(u8, pre_scale=1 / 255.0, l2norm=True)                  # one pass over all 1500 images
    assert torch.equal(whole, a)
    assert torch.isfinite(a).all()


def test_fused_row_statistics_schedule_matches_the_rowstats_schedule(native, monkeypatch):
    """The default schedule (residual GEMM epilogues leave the next LayerNorm's row statistics) against the round-1
    schedule (TMA reduce-add + a rowstats pass per folded GEMM): same embeddings up to the bf16 rounding of the stream."""
    eng = _engine(0)
    gen = torch.Generator().manual_seed(11)
    x = torch.rand(300, 3, 224, 224, generator=gen).cuda()
    a = eng.encode(x, pre_scale=1.0, l2norm=True)
    monkeypatch.setenv("CLIPPPO_GEMM_RESID", "reduce")
    b = eng.encode(x, pre_scale=1.0, l2norm=True)
    monkeypatch.delenv("CLIPPPO_GEMM_RESID")
    c = eng.encode(x, pre_scale=1.0, l2norm=True)
    assert torch.equal(a, c)                                  # deterministic
    cos = torch.nn.functional.cosine_similarity(a, b, dim=1)
    assert cos.min().item() >= 0.9995, cos.min().item()


def test_two_streams_encode_concurrently_without_sharing_a_workspace(native):
    """Two tower passes enqueued on two streams may overlap on the device: each stream gets its own cached workspace, and
    the halves equal the whole batch bit for bit (batch invariance)."""
    eng = _engine(0)
    gen = torch.Generator().manual_seed(12)
    x = torch.rand(600, 3, 224, 224, generator=gen).cuda()
    whole = eng.encode(x, pre_scale=1.0, l2norm=True)
    cur = torch.cuda.current_stream()
    streams = [torch.cuda.Stream() for _ in range(2)]
    outs = []
    for st, h in zip(streams, (x[:300], x[300:])):
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            for _ in range(3):                                  # several passes in flight per stream
                o = eng.encode(h, pre_scale=1.0, l2norm=True)
            outs.append(o)
    for st in streams:
        cur.wait_stream(st)
    torch.cuda.synchronize()
    assert torch.equal(torch.cat(outs), whole)


def test_frozen_features_and_encode_image(native):
    import shared.clip_ppo_utils as U
    model = U.load_clip_model("ViT-B/32", device="cuda")
    sd = ov.random_state_dict(ov.VIT_B32, 0)
    gen = torch.Generator().manual_seed(8)
    x = torch.rand(4, 3, 84, 84, generator=gen)
    f = U.get_frozen_clip_features(x.cuda(), model.visual).cpu()       # bare VisionTransformer dispatch (:212-213)
    f2 = U.get_frozen_clip_features(x.cuda(), model).cpu()             # full CLIP object
    ref = ov.frozen_features(sd, x)
    assert torch.equal(f, f2)
    cos = torch.nn.functional.cosine_similarity(f, ref, dim=-1)
    assert cos.min().item() >= 0.999
    assert abs(f.norm(dim=-1).mean().item() / ref.norm(dim=-1).mean().item() - 1) < 2e-2       # not L2-normalised
    # upstream-style encode_image on an already normalised batch
    pre = ov.preprocess(x, False)
    e = model.encode_image(pre.cuda()).cpu()
    cos = torch.nn.functional.cosine_similarity(e, ov.vision_tower(sd, pre), dim=-1)
    assert cos.min().item() >= 0.999
    assert all(not p.requires_grad for p in model.parameters())
    assert "conv1.weight" in model.visual.state_dict()


def test_uint8_obs_store_matches_fp32_rollout_storage(native):
    """§8f-4: MiniGrid observations stored as uint8 give bitwise the embeddings of the script's fp32 0..255 store."""
    import shared.clip_ppo_utils as U
    from clip_ppo_b200 import rollout
    model = U.load_clip_model("ViT-B/32", "cuda")
    T, E = 4, 6
    gen = torch.Generator().manual_seed(5)
    frames = torch.randint(0, 256, (T, E, 84, 84, 3), generator=gen, dtype=torch.uint8).cuda()
    obs_f32 = torch.zeros((T, E, 84, 84, 3), device="cuda")                      # clip_ppo_minigrid.py:346
    store = rollout.ObsStoreU8(T, E, (84, 84, 3))
    for t in range(T):
        obs_f32[t] = frames[t].float()
        store[t] = frames[t].float() if t % 2 else frames[t]                       # fp32-valued and uint8 writes
    assert store.data.element_size() * 4 == obs_f32.element_size()
    mb = torch.tensor([0, 5, 7, 13, 23], device="cuda")
    ref = U.generate_clip_embeddings(U.AblationMode.NONE, model, "image", len(mb), "cuda",
                                     images=obs_f32.reshape(-1, 84, 84, 3)[mb].permute(0, 3, 1, 2))
    got = U.generate_clip_embeddings(U.AblationMode.NONE, model, "image", len(mb), "cuda", images=store.clip_images(mb))
    assert torch.equal(ref, got)
    assert torch.equal(store.policy_input(mb), obs_f32.reshape(-1, 84, 84, 3)[mb])
    with pytest.raises(ValueError):
        store[0] = torch.full((E, 84, 84, 3), 0.5, device="cuda")


def test_frames_larger_than_the_tower_resolution(native):
    """h, w > 224: antialiased down-sampling (torch's op, as in the reference) in front of the fused path."""
    eng = _engine(0)
    sd = ov.random_state_dict(ov.VIT_B32, 0)
    gen = torch.Generator().manual_seed(9)
    img = torch.randint(0, 256, (3, 3, 300, 260), generator=gen).float()
    emb = eng.encode(img.cuda(), pre_scale=1 / 255.0, l2norm=True).cpu()
    ref = ov.image_embeddings(sd, img)
    assert torch.sum(emb * ref, dim=-1).min().item() >= 0.999
    mixed = torch.randint(0, 256, (2, 3, 240, 100), generator=gen).float()          # one side down, one side up
    emb = eng.encode(mixed.cuda(), pre_scale=1 / 255.0, l2norm=True).cpu()
    assert torch.sum(emb * ov.image_embeddings(sd, mixed), dim=-1).min().item() >= 0.999


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_second_device_in_the_same_process(native):
    """Kernel attributes (large dynamic smem, cluster sizes) are opted in per device: a process that used cuda:0 can
    run the same path on cuda:1 and gets the same bits."""
    from shared.disturbances_gpu import DisturbanceWrapperGPU
    from shared.disturbance_types import DisturbanceSeverity
    from clip_ppo_b200.vit import VitEngine
    sd = ov.random_state_dict(ov.VIT_B32, 0)
    gen = torch.Generator().manual_seed(3)
    x = torch.rand(5, 3, 224, 224, generator=gen)
    noise = torch.randn(5, 3, 224, 224, generator=gen)
    small = torch.rand(9, 3, 84, 84, generator=gen)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        w = DisturbanceWrapperGPU(device=dev, severity=DisturbanceSeverity.SEVERE)
        d = w.apply_disturbances(x.to(dev), noise=noise.to(dev), contrast_factor=1.2, cutout_start=(5, 9))
        eng = VitEngine(sd, device=dev)
        e = eng.encode(d * 255.0, pre_scale=1 / 255.0, l2norm=True)
        e2 = eng.encode(small.to(dev) * 255.0, pre_scale=1 / 255.0, l2norm=True)
        outs.append((d.cpu(), e.cpu(), e2.cpu()))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)


def test_full_size_batches_of_the_baseline_configs(native):
    """BASELINE configs at their full sizes, through size-independent properties: unit-norm rows, and an image's
    embedding is bitwise the same alone, inside a small batch and inside the full batch (several tower chunks)."""
    import shared.clip_ppo_utils as U
    from clip_ppo_b200 import rollout
    model = U.load_clip_model("ViT-B/32", "cuda")
    gen = torch.Generator(device="cuda").manual_seed(1)
    # configs[2]: 4096 frames of 224 x 224 x 3 (uint8-valued)
    frames = torch.randint(0, 256, (4096, 3, 224, 224), device="cuda", generator=gen, dtype=torch.uint8)
    emb = U.generate_clip_embeddings(U.AblationMode.NONE, model, "image", 4096, "cuda", images=frames)
    assert emb.shape == (4096, 512) and torch.isfinite(emb).all()
    assert (emb.norm(dim=-1) - 1).abs().max().item() <= 1e-5
    idx = torch.tensor([0, 1337, 4095], device="cuda")
    sub = U.generate_clip_embeddings(U.AblationMode.NONE, model, "image", 3, "cuda", images=frames[idx].float())
    assert torch.equal(sub, emb[idx])
    del frames, emb
    # configs[3]: 256 envs x 128 steps of 4 stacked 84 x 84 gray frames = 131 072 CLIP frames -> [32768, 2048]
    stacks = torch.randint(0, 256, (32768, 4, 84, 84), device="cuda", generator=gen, dtype=torch.uint8).float()
    rgb = rollout.convert_atari_frames_for_clip(stacks) / 255.0               # the Atari call site's double /255 (clip_ppo_atari.py:661)
    e = rollout.process_multiframe_clip_embeddings(rgb, model, U.AblationMode.NONE, "image", 32768, "cuda")
    assert e.shape == (32768, 4 * 512) and torch.isfinite(e).all()
    assert (e.reshape(-1, 512).norm(dim=-1) - 1).abs().max().item() <= 1e-5
    pick = torch.tensor([0, 20000, 32767], device="cuda")
    e_sub = rollout.process_multiframe_clip_embeddings(rgb[pick], model, U.AblationMode.NONE, "image", 3, "cuda")
    assert torch.equal(e_sub, e[pick])


def test_graph_cache_is_bounded(native):
    """encode_graphed keeps at most GRAPH_CACHE_ENTRIES captured call shapes (least recently used out first): varying batch
    sizes (eval, tail minibatches) no longer grow device memory without bound; evicted shapes are simply captured again."""
    from clip_ppo_b200.vit import VitEngine
    eng = VitEngine(ov.random_state_dict(ov.VIT_B32, 0), device="cuda")
    g = torch.Generator(device="cuda").manual_seed(1)
    outs = {}
    for n in (3, 5, 7, 9, 11, 13, 3, 5):
        x = torch.rand(n, 3, 84, 84, device="cuda", generator=g)
        e = eng.encode_graphed(x, pre_scale=1.0, l2norm=False)
        assert torch.equal(e, eng.encode(x, pre_scale=1.0, l2norm=False))
        assert len(eng._graphs) <= eng.GRAPH_CACHE_ENTRIES
        outs[n] = e
    assert [k[0][0] for k in eng._graphs] == [11, 13, 3, 5]          # LRU order, oldest first
