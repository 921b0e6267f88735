"""Oracle (CPU, fp32 torch) for the visual disturbance chain.  TEST INFRASTRUCTURE ONLY.

Restates, stage by stage and with *supplied* randomness, what
``DisturbanceWrapperGPU.apply_disturbances`` does (reference
``shared/disturbances_gpu.py:66-73``): noise -> contrast -> blur -> cutout.  The arithmetic
of three of the four stages lives in torchvision (un-vendored, ``requirements.txt:2``); the
op order below follows torchvision 0.26 so that the restatement is bit-exact against the
reference module on CPU (checked by ``oracle/make_goldens.py`` and ``tests/test_oracle.py``).

Pinned by: tests/golden/disturb_*.npz (outputs of the reference module itself).
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

# reference shared/disturbance_types.py:18-43 (also README.md:102-107)
SEVERITY_TABLE: Dict[str, Dict[str, object]] = {
    "MILD": dict(noise_sigma=0.08, blur_sigma=1.0, contrast=(0.75, 1.25), cutout=0.10),
    "MODERATE": dict(noise_sigma=0.12, blur_sigma=2.0, contrast=(0.7, 1.3), cutout=0.17),
    "HARD": dict(noise_sigma=0.13, blur_sigma=2.1, contrast=(0.69, 1.31), cutout=0.18),
    "SEVERE": dict(noise_sigma=0.26, blur_sigma=3.0, contrast=(0.6, 1.4), cutout=0.25),
}


def blur_kernel_size(blur_sigma: float) -> int:
    """reference shared/disturbances_gpu.py:58-60."""
    k = max(3, int(2 * blur_sigma) + 1)
    return k + 1 if k % 2 == 0 else k


def gaussian_kernel1d(k: int, sigma: float) -> torch.Tensor:
    """[tv] _functional_tensor.py:727-734 (fp32 linspace / exp / normalise)."""
    half = (k - 1) * 0.5
    t = torch.linspace(-half, half, steps=k, dtype=torch.float32)
    pdf = torch.exp(-0.5 * (t / sigma).pow(2))
    return pdf / pdf.sum()


def cutout_patch(H: int, W: int, ratio: float) -> Tuple[int, int]:
    """reference shared/disturbances_gpu.py:163-165."""
    area = int(H * W * ratio)
    ph = int(math.sqrt(area))
    return ph, area // ph


def add_noise(x: torch.Tensor, noise: torch.Tensor, sigma_n: float) -> torch.Tensor:
    """[tv] v2/functional/_misc.py:197-206 with the randn tensor supplied."""
    return torch.clamp(x + (0.0 + noise * sigma_n), 0, 1)


def contrast(x: torch.Tensor, c: float) -> torch.Tensor:
    """[tv] _functional_tensor.py:180-194,258-261.  One factor per call, one mean per image."""
    if x.shape[-3] == 3:
        r, g, b = x.unbind(dim=-3)
        gray = (0.2989 * r + 0.587 * g + 0.114 * b).unsqueeze(-3)
    elif x.shape[-3] == 1:
        gray = x
    else:
        raise TypeError("contrast needs 1 or 3 channels")
    m = torch.mean(gray, dim=(-3, -2, -1), keepdim=True)
    c = float(c)
    return (c * x + (1.0 - c) * m).clamp(0, 1)


def blur(x: torch.Tensor, k1d: torch.Tensor) -> torch.Tensor:
    """[tv] _functional_tensor.py:737-764: outer-product kernel, reflect pad, depthwise conv."""
    k = k1d.numel()
    k2d = torch.mm(k1d[:, None], k1d[None, :])
    w = k2d.expand(x.shape[-3], 1, k, k)
    p = k // 2
    xp = F.pad(x, [p, p, p, p], mode="reflect")
    return F.conv2d(xp, w, groups=x.shape[-3])


def cutout(x: torch.Tensor, sh: int, sw: int, ph: int, pw: int) -> torch.Tensor:
    """reference shared/disturbances_gpu.py:157-172: one window for the whole batch."""
    out = x.clone()
    out[:, :, sh:sh + ph, sw:sw + pw] = 0.0
    return out


def disturb(x: torch.Tensor, noise: torch.Tensor, sigma_n: float, c: float,
            k1d: torch.Tensor, sh: int, sw: int, ph: int, pw: int) -> torch.Tensor:
    """The full chain, reference shared/disturbances_gpu.py:66-73."""
    y = add_noise(x, noise, sigma_n)
    y = contrast(y, c)
    y = blur(y, k1d)
    return cutout(y, sh, sw, ph, pw)


def draw_call_randomness(x: torch.Tensor, cfg: Dict[str, object]) -> Dict[str, object]:
    """Consume the global torch generators exactly as one reference ``apply_disturbances``
    call does: randn_like(x) ; randperm(4) ; uniform_(lo,hi) ;
    uniform_(sigma,sigma) ; randint ; randint.  Returns the drawn values."""
    H, W = x.shape[-2:]
    noise = torch.randn_like(x)
    torch.randperm(4)
    lo, hi = cfg["contrast"]
    c = float(torch.empty(1).uniform_(lo, hi))
    sigma_b = torch.empty(1).uniform_(cfg["blur_sigma"], cfg["blur_sigma"]).item()
    ph, pw = cutout_patch(H, W, cfg["cutout"])
    sh = torch.randint(0, max(1, H - ph + 1), (1,)).item()
    sw = torch.randint(0, max(1, W - pw + 1), (1,)).item()
    k = blur_kernel_size(cfg["blur_sigma"])
    return dict(noise=noise, c=c, sigma_b=sigma_b, k=k, sh=sh, sw=sw, ph=ph, pw=pw)


def disturb_seeded(x: torch.Tensor, severity: str) -> torch.Tensor:
    """What the reference returns for the current global RNG state."""
    cfg = SEVERITY_TABLE[severity]
    r = draw_call_randomness(x, cfg)
    k1d = gaussian_kernel1d(r["k"], r["sigma_b"])
    return disturb(x, r["noise"], cfg["noise_sigma"], r["c"], k1d, r["sh"], r["sw"], r["ph"], r["pw"])
