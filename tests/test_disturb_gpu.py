"""Parity of the fused disturbance kernel (through the C ABI) with the reference goldens and the
oracle.  Tolerance: 1e-5 absolute (BASELINE.json north_star)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import disturb as od

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
TOL = 1e-5
FILES = sorted(glob.glob(os.path.join(GOLDEN, "disturb_*.npz")))


def _wrapper(sev):
    from shared.disturbances_gpu import DisturbanceWrapperGPU
    from shared.disturbance_types import DisturbanceSeverity
    return DisturbanceWrapperGPU(device="cuda", severity=DisturbanceSeverity[sev])


def _case(path):
    g = np.load(path)
    u8 = torch.from_numpy(g["u8"]).cuda()
    x = (u8.float() / 255.0).permute(0, 3, 1, 2)
    if str(g["layout"]) != "nhwc_view":
        x = x.contiguous()
    return g, x


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p)[:-4] for p in FILES])
def test_chain_matches_reference_golden(native, path):
    g, x = _case(path)
    w = _wrapper(str(g["severity"]))
    noise = torch.from_numpy(g["noise"]).cuda()
    if str(g["layout"]) == "nhwc_view":                   # randn_like of the view keeps its strides
        noise = noise.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    out = w.apply_disturbances(x, noise=noise, contrast_factor=float(g["c"]), cutout_start=(int(g["sh"]), int(g["sw"])))
    assert out.shape == x.shape and out.is_contiguous() and out.dtype == torch.float32
    err = (out.cpu() - torch.from_numpy(g["out"])).abs().max().item()
    assert err <= TOL, err
    sh, sw, ph, pw = int(g["sh"]), int(g["sw"]), int(g["ph"]), int(g["pw"])
    assert out[:, :, sh:sh + ph, sw:sw + pw].abs().max().item() == 0.0


@pytest.mark.parametrize("path", FILES[2:4] + FILES[8:], ids=["rgb", "gray", "nhwc", "224", "ragged"])
def test_each_stage_matches_oracle(native, path):
    g, x = _case(path)
    sev = str(g["severity"])
    cfg = od.SEVERITY_TABLE[sev]
    w = _wrapper(sev)
    xc = x.cpu()
    noise = torch.from_numpy(g["noise"])
    k1d = torch.from_numpy(g["k1d"])
    a = w.apply_gaussian_noise(x, noise=noise.cuda())
    assert torch.equal(a.cpu(), od.add_noise(xc, noise, cfg["noise_sigma"]))          # bit-exact stage
    b = w.apply_contrast_jitter(x, contrast_factor=float(g["c"]))
    assert (b.cpu() - od.contrast(xc, float(g["c"]))).abs().max() <= TOL
    c = w.apply_gaussian_blur(x)
    assert (c.cpu() - od.blur(xc, k1d)).abs().max() <= TOL
    d = w.apply_cutout(x, cutout_start=(int(g["sh"]), int(g["sw"])))
    assert torch.equal(d.cpu(), od.cutout(xc, int(g["sh"]), int(g["sw"]), int(g["ph"]), int(g["pw"])))


def test_seeded_call_consumes_rng_like_the_reference(native):
    """Same seed => same contrast factor / window as the CPU reference stream, and the noise is
    exactly torch.randn_like(obs) from the device generator."""
    from shared.disturbances_gpu import DisturbanceWrapperGPU
    from shared.disturbance_types import DisturbanceSeverity
    x = torch.rand(4, 3, 84, 84, device="cuda")
    w = DisturbanceWrapperGPU(device="cuda", seed=77, severity=DisturbanceSeverity.MODERATE)
    out = w.apply_disturbances(x)
    torch.manual_seed(77)
    r = od.draw_call_randomness(x, od.SEVERITY_TABLE["MODERATE"])       # randn_like on the CUDA generator
    k1d = od.gaussian_kernel1d(r["k"], r["sigma_b"])
    ref = od.disturb(x.cpu(), r["noise"].cpu(), 0.12, r["c"], k1d, r["sh"], r["sw"], r["ph"], r["pw"])
    assert (out.cpu() - ref).abs().max() <= TOL
    # and twice the same seed => identical result
    w2 = DisturbanceWrapperGPU(device="cuda", seed=77, severity=DisturbanceSeverity.MODERATE)
    assert torch.equal(out, w2.apply_disturbances(x))


@pytest.mark.parametrize("shape,sev", [((64, 3, 84, 84), "MODERATE"), ((256, 1, 84, 84), "HARD"),
                                       ((9, 3, 224, 224), "SEVERE"), ((5, 3, 50, 70), "MILD"),
                                       ((3, 1, 17, 23), "SEVERE"), ((2, 3, 224, 224), "MILD")])
def test_random_shapes_against_oracle(native, shape, sev):
    g = torch.Generator().manual_seed(hash((shape, sev)) % 2**31)
    x = torch.rand(shape, generator=g)
    noise = torch.randn(shape, generator=g)
    cfg = od.SEVERITY_TABLE[sev]
    H, W = shape[-2:]
    ph, pw = od.cutout_patch(H, W, cfg["cutout"])
    sh, sw = (H - ph) // 3, (W - pw) // 2
    c = 0.5 * (cfg["contrast"][0] + cfg["contrast"][1]) + 0.07
    w = _wrapper(sev)
    out = w.apply_disturbances(x.cuda(), noise=noise.cuda(), contrast_factor=c, cutout_start=(sh, sw))
    k = od.blur_kernel_size(cfg["blur_sigma"])
    sigma32 = float(torch.tensor(cfg["blur_sigma"], dtype=torch.float32))
    ref = od.disturb(x, noise, cfg["noise_sigma"], c, od.gaussian_kernel1d(k, sigma32), sh, sw, ph, pw)
    assert (out.cpu() - ref).abs().max() <= TOL


# Shapes that steer the dispatcher through every variant of the fast kernel (one CTA per frame / 8-CTA
# clusters, template widths 84 and 224 / run-time widths, short last stripe, a single frame) and into the
# general kernel (odd width, width not a multiple of 4, strided input).
@pytest.mark.parametrize("shape,sev", [((3, 3, 224, 224), "MODERATE"), ((3, 3, 224, 224), "HARD"), ((4, 1, 224, 224), "MODERATE"),
                                       ((1, 3, 224, 224), "SEVERE"), ((3, 3, 250, 224), "MODERATE"), ((2, 3, 300, 200), "HARD"),
                                       ((6, 3, 96, 128), "MILD"), ((2, 3, 64, 512), "SEVERE"), ((2, 1, 90, 88), "HARD"),
                                       ((3, 3, 84, 86), "MODERATE"), ((2, 3, 61, 61), "SEVERE"), ((1, 1, 8, 8), "MILD")])
def test_dispatch_variants_against_oracle(native, shape, sev):
    g = torch.Generator().manual_seed(sum(shape) * 7 + len(sev))
    x = torch.rand(shape, generator=g)
    noise = torch.randn(shape, generator=g)
    cfg = od.SEVERITY_TABLE[sev]
    H, W = shape[-2:]
    ph, pw = od.cutout_patch(H, W, cfg["cutout"])
    sh, sw = max(0, (H - ph) // 2), max(0, (W - pw) // 3 - 1)           # a patch taller than the frame is clipped, as in the reference
    c = cfg["contrast"][1] - 0.02                          # > 1: the clamp of the blend is active for dark pixels
    w = _wrapper(sev)
    k = od.blur_kernel_size(cfg["blur_sigma"])
    k1d = od.gaussian_kernel1d(k, float(torch.tensor(cfg["blur_sigma"], dtype=torch.float32)))
    out = w.apply_disturbances(x.cuda(), noise=noise.cuda(), contrast_factor=c, cutout_start=(sh, sw))
    ref = od.disturb(x, noise, cfg["noise_sigma"], c, k1d, sh, sw, ph, pw)
    assert (out.cpu() - ref).abs().max() <= TOL
    assert out[:, :, sh:sh + ph, sw:sw + pw].abs().max().item() == 0.0
    # each stage on its own (stage mask): without the contrast stage a multi-stripe image runs as plain CTAs
    xg = x.cuda()
    assert torch.equal(w.apply_gaussian_noise(xg, noise=noise.cuda()).cpu(), od.add_noise(x, noise, cfg["noise_sigma"]))
    assert (w.apply_contrast_jitter(xg, contrast_factor=c).cpu() - od.contrast(x, c)).abs().max() <= TOL
    assert (w.apply_gaussian_blur(xg).cpu() - od.blur(x, k1d)).abs().max() <= TOL
    assert torch.equal(w.apply_cutout(xg, cutout_start=(sh, sw)).cpu(), od.cutout(x, sh, sw, ph, pw))
    # a strided view of the same data (general kernel) gives the same chain
    xs = torch.empty(shape[0], H, W, shape[1]).cuda()
    xs.copy_(x.permute(0, 2, 3, 1))
    out_v = w.apply_disturbances(xs.permute(0, 3, 1, 2), noise=noise.cuda(), contrast_factor=c, cutout_start=(sh, sw))
    assert (out_v.cpu() - ref).abs().max() <= TOL


@pytest.mark.parametrize("B,C", [(8, 3), (64, 3), (40, 1), (150, 3)])
@pytest.mark.parametrize("sev", ["MODERATE", "SEVERE"])
def test_env_step_batches_run_as_stripe_clusters(native, B, C, sev):
    """Env-step call shapes (clip_ppo_minigrid.py:381-388: E x 84 x 84 x 3 NHWC; clip_ppo_atari.py:568-584:
    [E,1,84,84]): below 148 frames a frame is cut into a cluster of stripes so the grid covers the SMs.  Same
    chain as the oracle in both layouts; a frame alone agrees with the frame in its batch to the last bit or two
    (the gray mean's summation order follows the stripe count)."""
    g = torch.Generator().manual_seed(B * 3 + C)
    frames = torch.rand(B, 84, 84, C, generator=g)
    x = frames.permute(0, 3, 1, 2)                                  # NHWC-strided view, like the reference's call site
    noise = torch.randn(B, 84, 84, C, generator=g).permute(0, 3, 1, 2)
    cfg = od.SEVERITY_TABLE[sev]
    ph, pw = od.cutout_patch(84, 84, cfg["cutout"])
    k = od.blur_kernel_size(cfg["blur_sigma"])
    k1d = od.gaussian_kernel1d(k, float(torch.tensor(cfg["blur_sigma"], dtype=torch.float32)))
    idx = torch.linspace(0, B - 1, 5).long()
    ref = od.disturb(x[idx].contiguous(), noise[idx].contiguous(), cfg["noise_sigma"], 1.21, k1d, 11, 30, ph, pw)
    w = _wrapper(sev)
    out_v = w.apply_disturbances(x.cuda(), noise=noise.cuda(), contrast_factor=1.21, cutout_start=(11, 30))
    out_c = w.apply_disturbances(x.contiguous().cuda(), noise=noise.contiguous().cuda(), contrast_factor=1.21, cutout_start=(11, 30))
    assert (out_v.cpu()[idx] - ref).abs().max() <= TOL
    assert (out_c.cpu()[idx] - ref).abs().max() <= TOL
    big = w.apply_disturbances(x.contiguous().repeat(4, 1, 1, 1).cuda(), noise=noise.contiguous().repeat(4, 1, 1, 1).cuda(),
                               contrast_factor=1.21, cutout_start=(11, 30))
    assert (big[:B] - out_c).abs().max().item() <= 1e-6


def test_uint8_frames_equal_the_float_path_bitwise(native):
    """Additive: uint8 [B,C,H,W] input is converted inside the kernel - bit-identical to handing over
    `obs.float() / 255` computed by PyTorch on the device (the reference benchmark's conversion; ATen multiplies by
    fl(1/255) there), for every byte value and every dispatch variant."""
    from clip_ppo_b200 import disturb as D, _native as Nn
    # all 256 byte values through the load path alone (cutout stage with an empty window = copy)
    ramp = torch.arange(256, dtype=torch.uint8, device="cuda").reshape(1, 1, 8, 32)
    out = D.fused_disturb(ramp, stages=Nn.STAGE_CUTOUT, window=(0, 0, 0, 0))
    assert torch.equal(out, ramp.float() / 255.0)
    gen = torch.Generator(device="cuda").manual_seed(12)
    for shape, sev in (((5, 3, 224, 224), "MODERATE"), ((2, 3, 224, 224), "SEVERE"), ((200, 3, 84, 84), "HARD"),
                       ((8, 3, 84, 84), "MODERATE"), ((160, 1, 84, 84), "SEVERE"), ((3, 3, 84, 86), "MILD"), ((2, 3, 61, 61), "HARD")):
        u8 = torch.randint(0, 256, shape, device="cuda", generator=gen, dtype=torch.uint8)
        noise = torch.randn(shape, device="cuda", generator=gen)
        w = _wrapper(sev)
        a = w.apply_disturbances(u8, noise=noise, contrast_factor=1.17, cutout_start=(4, 9))
        b = w.apply_disturbances(u8.float() / 255.0, noise=noise, contrast_factor=1.17, cutout_start=(4, 9))
        assert a.dtype == torch.float32 and torch.equal(a, b), (shape, sev, (a - b).abs().max().item())
    # default randomness: same seed, same draws as the float call
    u8 = torch.randint(0, 256, (6, 3, 84, 84), device="cuda", generator=gen, dtype=torch.uint8)
    w = _wrapper("MODERATE")
    torch.manual_seed(5)
    a = w.apply_disturbances(u8)
    torch.manual_seed(5)
    b = w.apply_disturbances(u8.float() / 255.0)
    assert torch.equal(a, b)


@pytest.mark.parametrize("E", [8, 256])
def test_atari_stack_frames_take_the_fast_kernel(native, E):
    """clip_ppo_atari.py:568-584 disturbs each of the 4 stacked frames as `x[:, f:f+1]`: [E,1,84,84] views whose images
    are contiguous but 4 frames apart.  They run through the same kernel as a contiguous copy (bitwise the same result)."""
    from clip_ppo_b200 import rollout
    gen = torch.Generator(device="cuda").manual_seed(E)
    stack = torch.rand(E, 4, 84, 84, device="cuda", generator=gen)
    noise = torch.randn(E, 1, 84, 84, device="cuda", generator=gen)
    w = _wrapper("HARD")
    for f in (0, 3):
        view = stack[:, f:f + 1]
        assert not view.is_contiguous() or E == 1
        a = w.apply_disturbances(view, noise=noise, contrast_factor=0.8, cutout_start=(20, 7))
        b = w.apply_disturbances(view.contiguous(), noise=noise, contrast_factor=0.8, cutout_start=(20, 7))
        assert torch.equal(a, b)
    cfg = od.SEVERITY_TABLE["HARD"]
    ph, pw = od.cutout_patch(84, 84, cfg["cutout"])
    k1d = od.gaussian_kernel1d(od.blur_kernel_size(cfg["blur_sigma"]), float(torch.tensor(cfg["blur_sigma"], dtype=torch.float32)))
    ref = od.disturb(stack[:4, 3:4].cpu().contiguous(), noise[:4].cpu(), cfg["noise_sigma"], 0.8, k1d, 20, 7, ph, pw)
    assert (a[:4].cpu() - ref).abs().max() <= TOL
    # the whole call site: same draws, frame by frame, as the reference loop
    torch.manual_seed(9)
    got = rollout.disturb_atari_stack(w, stack * 255.0)
    torch.manual_seed(9)
    want = torch.cat([w.apply_disturbances((stack * 255.0 / 255.0)[:, f:f + 1].contiguous()) for f in range(4)], dim=1) * 255.0
    assert torch.equal(got, want)


def test_wide_blur_kernels_use_the_general_kernel(native):
    """Custom sigma: k = 9 .. 15 taps are outside the fast path's register ring."""
    from shared.disturbances_gpu import DisturbanceWrapperGPU
    g = torch.Generator().manual_seed(99)
    x = torch.rand(3, 3, 84, 84, generator=g)
    noise = torch.randn(3, 3, 84, 84, generator=g)
    for sigma_b in (4.0, 5.5, 7.0):
        w = DisturbanceWrapperGPU(device="cuda", severity=None, gaussian_noise_sigma=0.1, gaussian_blur_sigma=sigma_b,
                                  contrast_range=(0.8, 1.2), cutout_ratio=0.2)
        k = od.blur_kernel_size(sigma_b)
        assert k >= 9
        k1d = od.gaussian_kernel1d(k, float(torch.tensor(sigma_b, dtype=torch.float32)))
        ph, pw = od.cutout_patch(84, 84, 0.2)
        out = w.apply_disturbances(x.cuda(), noise=noise.cuda(), contrast_factor=0.9, cutout_start=(5, 7))
        ref = od.disturb(x, noise, 0.1, 0.9, k1d, 5, 7, ph, pw)
        assert (out.cpu() - ref).abs().max() <= TOL


def test_full_size_batch_properties(native):
    """BASELINE config sizes (4096 x 3 x 224 x 224 is 2.4 GB): size-independent properties plus an
    oracle comparison on a slice of the batch."""
    B = 1024
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.rand(B, 3, 224, 224, device="cuda", generator=g)
    noise = torch.randn(B, 3, 224, 224, device="cuda", generator=g)
    w = _wrapper("SEVERE")
    out = w.apply_disturbances(x, noise=noise, contrast_factor=1.23, cutout_start=(30, 60))
    assert out.min().item() >= 0.0 and out.max().item() <= 1.0 + 1e-6
    assert out[:, :, 30:30 + 112, 60:60 + 112].abs().max().item() == 0.0
    # per-image independence: a slice processed alone gives the same answer
    idx = torch.tensor([0, 511, 1023], device="cuda")
    sub = w.apply_disturbances(x[idx], noise=noise[idx], contrast_factor=1.23, cutout_start=(30, 60))
    assert torch.equal(sub, out[idx])
    k1d = od.gaussian_kernel1d(7, 3.0)
    ref = od.disturb(x[idx].cpu(), noise[idx].cpu(), 0.26, 1.23, k1d, 30, 60, 112, 112)
    assert (sub.cpu() - ref).abs().max() <= TOL
    # linearity of the blur stage: blur(a + b) == blur(a) + blur(b)
    a, b = x[:8] * 0.5, noise[:8].abs().clamp(0, 0.5)
    lhs = w.apply_gaussian_blur(a + b)
    rhs = w.apply_gaussian_blur(a) + w.apply_gaussian_blur(b)
    assert (lhs - rhs).abs().max() <= 1e-6


def test_numpy_shims_and_minigrid_call_site(native):
    from clip_ppo_b200 import rollout
    w = _wrapper("MODERATE")
    rng = np.random.RandomState(0)
    img = rng.randint(0, 256, (84, 84, 3), dtype=np.uint8)
    batch = rng.randint(0, 256, (6, 84, 84, 3), dtype=np.uint8)
    for arr in (img, batch):
        out = w.apply_cutout_numpy(arr)
        assert out.shape == arr.shape and out.dtype == np.uint8
        # cutout of uint8 frames is exact: every pixel is either untouched or zero
        assert np.all((out == arr) | (out == 0))
        for fn in (w.apply_disturbances_numpy, w.apply_gaussian_noise_numpy, w.apply_contrast_jitter_numpy, w.apply_gaussian_blur_numpy):
            o = fn(arr)
            assert o.shape == arr.shape and o.dtype == np.uint8
    # MiniGrid env-step call site: fp32 0..255 NHWC in -> uint8 NHWC out == reference expression
    obs = torch.from_numpy(batch).cuda().float()
    torch.manual_seed(3)
    got = rollout.disturb_minigrid_obs(w, obs)
    torch.manual_seed(3)
    chw = (obs / 255.0).permute(0, 3, 1, 2)
    r = od.draw_call_randomness(chw, od.SEVERITY_TABLE["MODERATE"])
    ref = od.disturb(chw.cpu(), r["noise"].cpu(), 0.12, r["c"], od.gaussian_kernel1d(r["k"], r["sigma_b"]),
                     r["sh"], r["sw"], r["ph"], r["pw"])
    ref_u8 = (ref.permute(0, 2, 3, 1) * 255.0).byte()
    assert got.dtype == torch.uint8 and got.shape == obs.shape
    diff = (got.cpu().int() - ref_u8.int()).abs()
    assert diff.max().item() <= 1 and (diff > 0).float().mean().item() < 1e-3      # floor-boundary flips only


@pytest.mark.parametrize("batched", [False, True])
def test_numpy_shims_values_vs_oracle(native, batched):
    """a8: every `*_numpy` shim (reference shared/disturbances_gpu.py:75-95, 101-115, 121-135, 141-155, 174-194) returns the
    VALUES of the reference expression  (stage(x / 255) * 255).byte()  for the randomness a seeded reference call would draw:
    each stage's own draws are replayed from the same global-generator state and pushed through the oracle.
    Bytes may differ by one where the float result sits on a floor boundary (< 1e-3 of the pixels)."""
    cfg = od.SEVERITY_TABLE["HARD"]
    rng = np.random.RandomState(5)
    arr = rng.randint(0, 256, (5, 84, 84, 3) if batched else (84, 84, 3), dtype=np.uint8)
    x = torch.from_numpy(arr).float().div(255.0)
    x = (x if batched else x.unsqueeze(0)).permute(0, 3, 1, 2)             # the reference's NHWC-strided NCHW view
    H, W = 84, 84

    def ref_noise():
        n = torch.randn_like(x.cuda()).cpu()
        return od.add_noise(x, n, cfg["noise_sigma"])

    def ref_contrast():
        torch.randperm(4)
        c = float(torch.empty(1).uniform_(*cfg["contrast"]))
        return od.contrast(x, c)

    def ref_blur():
        sigma = torch.empty(1).uniform_(cfg["blur_sigma"], cfg["blur_sigma"]).item()
        return od.blur(x, od.gaussian_kernel1d(od.blur_kernel_size(cfg["blur_sigma"]), sigma))

    def ref_cutout():
        ph, pw = od.cutout_patch(H, W, cfg["cutout"])
        sh = torch.randint(0, max(1, H - ph + 1), (1,)).item()
        sw = torch.randint(0, max(1, W - pw + 1), (1,)).item()
        return od.cutout(x, sh, sw, ph, pw)

    def ref_all():
        r = od.draw_call_randomness(x.cuda(), cfg)           # device randn_like on the same strided view, then the CPU draws
        return od.disturb(x, r["noise"].cpu(), cfg["noise_sigma"], r["c"], od.gaussian_kernel1d(r["k"], r["sigma_b"]),
                          r["sh"], r["sw"], r["ph"], r["pw"])

    w = _wrapper("HARD")
    for name, ref_fn in (("apply_gaussian_noise_numpy", ref_noise), ("apply_contrast_jitter_numpy", ref_contrast),
                         ("apply_gaussian_blur_numpy", ref_blur), ("apply_cutout_numpy", ref_cutout),
                         ("apply_disturbances_numpy", ref_all)):
        torch.manual_seed(11)
        got = getattr(w, name)(arr)
        torch.manual_seed(11)
        want = (ref_fn().permute(0, 2, 3, 1) * 255.0).byte().numpy()
        want = want if batched else want[0]
        assert got.shape == arr.shape and got.dtype == np.uint8, name
        diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
        assert diff.max() <= 1 and (diff > 0).mean() < 1e-3, (name, int(diff.max()), float((diff > 0).mean()))


@pytest.mark.parametrize("shape,sev,u8", [((9, 3, 224, 224), "MODERATE", False), ((5, 3, 224, 224), "SEVERE", True),
                                          ((300, 1, 84, 84), "HARD", False), ((40, 3, 84, 84), "MILD", True)])
def test_out_scale_equals_a_separate_multiply_bitwise(native, shape, sev, u8):
    """apply_disturbances(out_scale=255) == apply_disturbances(...) * 255 (the `* 255` of the call sites,
    clip_ppo_atari.py:584) bit for bit - one fp32 multiply either way."""
    gen = torch.Generator().manual_seed(sum(shape))
    w = _wrapper(sev)
    x = torch.randint(0, 256, shape, generator=gen, dtype=torch.uint8).cuda()
    if not u8:
        x = x.float() / 255.0
    noise = torch.randn(shape, generator=gen).cuda()
    kw = dict(noise=noise, contrast_factor=1.17, cutout_start=(7, 12))
    a = w.apply_disturbances(x, **kw) * 255.0
    b = w.apply_disturbances(x, out_scale=255.0, **kw)
    assert torch.equal(a, b)
    assert b.max().item() > 100.0


def test_out_scale_falls_back_for_shapes_outside_the_fast_kernel(native):
    gen = torch.Generator().manual_seed(3)
    w = _wrapper("MODERATE")
    x = torch.rand(4, 3, 40, 56, generator=gen).cuda()
    noise = torch.randn(4, 3, 40, 56, generator=gen).cuda()
    kw = dict(noise=noise, contrast_factor=0.9, cutout_start=(3, 4))
    assert torch.equal(w.apply_disturbances(x, **kw) * 255.0, w.apply_disturbances(x, out_scale=255.0, **kw))


@pytest.mark.parametrize("shape,sev,u8", [((9, 3, 224, 224), "MODERATE", False), ((6, 3, 224, 224), "SEVERE", True),
                                          ((300, 1, 84, 84), "HARD", False), ((150, 3, 84, 84), "MILD", False),
                                          ((7, 3, 84, 84), "SEVERE", True)])
def test_in_kernel_philox_noise_matches_its_oracle(native, shape, sev, u8):
    """apply_disturbances(noise_seed=s) (opt-in: N(0,1) draws generated inside the kernel) against the same call fed with
    the noise tensor oracle/philox.py computes for (seed, call counter): <= 1e-5 on the disturbed frames, for whole
    images, stripe clusters (224 x 224: halo rows re-generate the neighbours' draws) and env-step batches."""
    from oracle import philox as P
    gen = torch.Generator().manual_seed(sum(shape) + 1)
    x = torch.randint(0, 256, shape, generator=gen, dtype=torch.uint8).cuda()
    if not u8:
        x = x.float() / 255.0
    seed = 0x1234_5678_9ABC_DEF0 + shape[0]
    w = _wrapper(sev)
    kw = dict(contrast_factor=0.83, cutout_start=(5, 9))
    state = torch.cuda.get_rng_state()
    outs = [w.apply_disturbances(x, noise_seed=seed, **kw) for _ in range(2)]       # call counter 0, 1
    assert torch.equal(state, torch.cuda.get_rng_state())                             # the device generator is not consumed
    assert not torch.equal(outs[0], outs[1])
    for call, out in enumerate(outs):
        noise = torch.from_numpy(P.normal_noise(seed, call, shape)).cuda()
        ref = w.apply_disturbances(x, noise=noise, **kw)
        err = (out - ref).abs().max().item()
        assert err <= TOL, (call, err)
    # a shard of the batch draws the whole batch's noise (G-invariance)
    w2 = _wrapper(sev)
    lo = shape[0] // 3
    part = w2.apply_disturbances(x[lo:].contiguous(), noise_seed=seed, first_image=lo, **kw)
    d = (part - outs[0][lo:]).abs().max().item()
    # the noise is a function of the GLOBAL element index, so a shard sees exactly the draws of the whole batch; the only thing
    # that may move is the last bit of an 84 x 84 image's gray mean when the shard falls below 148 images and its frames are cut
    # into stripe clusters (run_disturb: widen_for_small_batches) - 224 x 224 frames always use the same 8 stripes
    assert d == 0.0 if shape[-1] == 224 else d <= 5e-7, d


def test_in_kernel_noise_refuses_shapes_it_does_not_serve(native):
    w = _wrapper("MODERATE")
    x = torch.rand(2, 3, 40, 56).cuda()
    with pytest.raises(NotImplementedError):
        w.apply_disturbances(x, noise_seed=1)
    with pytest.raises(ValueError):
        from clip_ppo_b200 import disturb as D
        D.fused_disturb(torch.rand(2, 3, 84, 84).cuda(), stages=15, noise=torch.randn(2, 3, 84, 84).cuda(), noise_sigma=0.1,
                        taps=(0.25, 0.5, 0.25), window=(0, 0, 4, 4), philox=(1, 0, 0))


def test_atari_call_site(native):
    from clip_ppo_b200 import rollout
    w = _wrapper("HARD")
    obs = torch.randint(0, 256, (16, 4, 84, 84), device="cuda").float()
    torch.manual_seed(9)
    got = rollout.disturb_atari_stack(w, obs)
    torch.manual_seed(9)
    frames = []
    for f in range(4):
        fr = (obs / 255.0)[:, f:f + 1]
        r = od.draw_call_randomness(fr, od.SEVERITY_TABLE["HARD"])
        frames.append(od.disturb(fr.cpu(), r["noise"].cpu(), 0.13, r["c"], od.gaussian_kernel1d(r["k"], r["sigma_b"]),
                                 r["sh"], r["sw"], r["ph"], r["pw"]))
    ref = torch.cat(frames, dim=1) * 255.0
    assert got.shape == obs.shape and (got.cpu() - ref).abs().max() <= 255 * TOL


def test_error_mapping(native):
    w = _wrapper("MILD")
    with pytest.raises(TypeError):
        w.apply_contrast_jitter(torch.rand(2, 2, 16, 16, device="cuda"))
    with pytest.raises(RuntimeError):
        _wrapper("SEVERE").apply_gaussian_blur(torch.rand(1, 3, 3, 40, device="cuda"))       # pad 3 >= H 3
    with pytest.raises(ValueError):
        w.apply_disturbances(torch.rand(3, 16, 16, device="cuda"))
    assert w.apply_disturbances(torch.rand(0, 3, 16, 16, device="cuda")).shape == (0, 3, 16, 16)
