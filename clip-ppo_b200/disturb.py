"""Host side of the fused disturbance kernel (D1): parameter derivation + one C-ABI launch.

Mirrors the arithmetic that reference ``shared/disturbances_gpu.py`` delegates to torchvision
(SURVEY.md Appendix A); the launch itself is ``clipppo_disturb_f32`` / ``clipppo_disturb_nhwc_u8``.
"""
from __future__ import annotations

import ctypes as C
import math
from functools import lru_cache
from typing import Optional, Sequence, Tuple

import torch

from . import _native as N


def blur_kernel_size(blur_sigma: float) -> int:
    """kernel_size the reference derives from sigma (shared/disturbances_gpu.py:58-60)."""
    k = max(3, int(2 * blur_sigma) + 1)
    return k + 1 if k % 2 == 0 else k


@lru_cache(maxsize=64)
def gaussian_taps(k: int, sigma: float) -> Tuple[float, ...]:
    """Normalised 1-D Gaussian taps, evaluated with the same fp32 op sequence torchvision
    uses ([tv] _functional_tensor.py:727-734) so the values are bit-identical; k scalars on
    the host, once per (k, sigma)."""
    half = (k - 1) * 0.5
    t = torch.linspace(-half, half, steps=k, dtype=torch.float32)
    pdf = torch.exp(-0.5 * (t / sigma).pow(2))
    return tuple((pdf / pdf.sum()).tolist())


def cutout_patch(H: int, W: int, ratio: float) -> Tuple[int, int]:
    """(patch_h, patch_w) as in shared/disturbances_gpu.py:163-165."""
    area = int(H * W * ratio)
    ph = int(math.sqrt(area))
    return ph, area // ph


def _stream_ptr(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor - the B200 path has no CPU fallback "
                           f"(got device {t.device})")


def _fused_disturb_ex(x, stages, noise, noise_sigma, contrast, taps, window, out_scale, philox) -> Optional[torch.Tensor]:
    """clipppo_disturb_ex: the fast kernel with an output scale and / or in-kernel Philox noise.  None when the kernel
    does not serve the shape (the caller falls back)."""
    B, Cc, H, W = x.shape
    out = torch.empty((B, Cc, H, W), dtype=torch.float32, device=x.device)
    if B == 0:
        return out
    d = N.DisturbDesc()
    d.x, d.x_dtype = x.data_ptr(), (N.IMG_U8 if x.dtype == torch.uint8 else N.IMG_F32)
    xs = N.strides4(x)
    d.x_strides_host = C.cast(xs, C.POINTER(C.c_int64))
    ns = None
    if (stages & N.STAGE_NOISE) and philox is None:
        if noise is None:
            raise ValueError("noise stage needs a noise tensor")
        if noise.shape != x.shape or noise.dtype != torch.float32 or noise.device != x.device:
            raise ValueError("noise must match x in shape, dtype (fp32) and device")
        ns = N.strides4(noise)
        d.noise, d.noise_strides_host = noise.data_ptr(), C.cast(ns, C.POINTER(C.c_int64))
    d.out = out.data_ptr()
    d.B, d.C, d.H, d.W, d.stages = B, Cc, H, W, stages
    d.noise_sigma, d.contrast = float(noise_sigma), float(contrast)
    k = len(taps) if (stages & N.STAGE_BLUR) else 0
    taps_arr = (C.c_float * max(k, 1))(*(taps if k else (1.0,)))
    d.k1d_host, d.k = C.cast(taps_arr, C.POINTER(C.c_float)), k
    d.sh, d.sw, d.ph, d.pw = (int(v) for v in window)
    d.out_scale = float(out_scale)
    if philox is not None:
        seed, offset, first_image = philox
        d.flags = N.DISTURB_PHILOX
        d.philox_seed, d.philox_offset, d.first_image = int(seed) & (2 ** 64 - 1), int(offset) & (2 ** 64 - 1), int(first_image)
    with N.device_ctx(x.device):
        st = N.lib().clipppo_disturb_ex(C.byref(d), _stream_ptr(x))
    if st == N.ERR_UNSUPPORTED:
        return None
    N.check(st, "clipppo_disturb_ex")
    return out


def fused_disturb(x: torch.Tensor, *, stages: int, noise: Optional[torch.Tensor] = None,
                  noise_sigma: float = 0.0, contrast: float = 1.0,
                  taps: Optional[Sequence[float]] = None,
                  window: Tuple[int, int, int, int] = (0, 0, 0, 0),
                  out_scale: float = 1.0, philox: Optional[Tuple[int, int, int]] = None) -> torch.Tensor:
    """out[B,C,H,W] fp32 (contiguous) = cutout(blur(contrast(noise(x)))) restricted to `stages`.
    x may have arbitrary strides (e.g. the NHWC view of clip_ppo_minigrid.py:385).  uint8 x (additive): pixels
    0..255 are read as float(v) * fl(1/255), bit-identical to `x.float() / 255` evaluated by PyTorch on the device,
    without materialising the fp32 batch.
    Additive: `out_scale` multiplies the result inside the kernel (== `out * out_scale`, bit for bit; a separate
    in-place multiply for the shapes the fast kernel does not serve); `philox = (seed, offset, first_image)` draws the
    noise inside the kernel instead of reading a tensor (opt-in, its own stream - see include/clipppo_b200.h)."""
    _require_cuda(x, "fused_disturb")
    if x.dim() != 4:
        raise ValueError(f"expected [B,C,H,W], got shape {tuple(x.shape)}")
    if philox is not None and noise is not None:
        raise ValueError("either a noise tensor or in-kernel noise, not both")
    if out_scale != 1.0 or philox is not None:
        xe = x if x.dtype in (torch.uint8, torch.float32) else x.float()
        out = _fused_disturb_ex(xe, stages, noise, noise_sigma, contrast, taps, window, out_scale, philox)
        if out is not None:
            return out
        if philox is not None:
            raise NotImplementedError("in-kernel noise is served for contiguous NCHW frames of width 84 / 224 with a blur stage "
                                      "(k = 3, 5, 7); draw the noise with torch for other shapes")
        out = fused_disturb(x, stages=stages, noise=noise, noise_sigma=noise_sigma, contrast=contrast, taps=taps, window=window)
        return out.mul_(out_scale)
    if x.dtype == torch.uint8:
        out = _fused_disturb_u8(x, stages, noise, noise_sigma, contrast, taps, window)
        if out is not None:
            return out
        x = x.float() / 255.0                       # shapes outside the fast kernel: convert, then the general path
    if x.dtype != torch.float32:
        x = x.float()
    B, Cc, H, W = x.shape
    out = torch.empty((B, Cc, H, W), dtype=torch.float32, device=x.device)
    if B == 0:
        return out
    nptr, nstr = None, None
    if stages & N.STAGE_NOISE:
        if noise is None:
            raise ValueError("noise stage needs a noise tensor")
        if noise.shape != x.shape or noise.dtype != torch.float32 or noise.device != x.device:
            raise ValueError("noise must match x in shape, dtype (fp32) and device")
        nptr, nstr = noise.data_ptr(), N.strides4(noise)
    k = len(taps) if (stages & N.STAGE_BLUR) else 0
    taps_arr = (C.c_float * max(k, 1))(*(taps if k else (1.0,)))
    sh, sw, ph, pw = window
    with N.device_ctx(x.device):
        st = N.lib().clipppo_disturb_f32(
            x.data_ptr(), N.strides4(x), nptr, nstr, out.data_ptr(), B, Cc, H, W, stages,
            float(noise_sigma), float(contrast), taps_arr, k, int(sh), int(sw), int(ph), int(pw),
            _stream_ptr(x))
    N.check(st, "clipppo_disturb_f32")
    return out


def _fused_disturb_u8(x, stages, noise, noise_sigma, contrast, taps, window) -> Optional[torch.Tensor]:
    """uint8 contiguous NCHW frames through clipppo_disturb_u8_f32; None when that entry point does not serve the shape."""
    B, Cc, H, W = x.shape
    if not x.is_contiguous() or W % 4:
        return None
    nptr = None
    if stages & N.STAGE_NOISE:
        if noise is None:
            raise ValueError("noise stage needs a noise tensor")
        if noise.shape != x.shape or noise.dtype != torch.float32 or noise.device != x.device:
            raise ValueError("noise must match x in shape and device, dtype fp32")
        if not noise.is_contiguous():
            return None
        nptr = noise.data_ptr()
    out = torch.empty((B, Cc, H, W), dtype=torch.float32, device=x.device)
    if B == 0:
        return out
    k = len(taps) if (stages & N.STAGE_BLUR) else 0
    taps_arr = (C.c_float * max(k, 1))(*(taps if k else (1.0,)))
    sh, sw, ph, pw = window
    with N.device_ctx(x.device):
        st = N.lib().clipppo_disturb_u8_f32(x.data_ptr(), nptr, out.data_ptr(), B, Cc, H, W, stages, float(noise_sigma),
                                            float(contrast), taps_arr, k, int(sh), int(sw), int(ph), int(pw), _stream_ptr(x))
    if st == N.ERR_UNSUPPORTED:
        return None
    N.check(st, "clipppo_disturb_u8_f32")
    return out


def fused_disturb_nhwc_u8(obs: torch.Tensor, *, stages: int, noise: Optional[torch.Tensor] = None,
                          noise_sigma: float = 0.0, contrast: float = 1.0,
                          taps: Optional[Sequence[float]] = None,
                          window: Tuple[int, int, int, int] = (0, 0, 0, 0)) -> torch.Tensor:
    """uint8 [B,H,W,C] = trunc(255 * chain(obs / 255)); obs is contiguous NHWC, uint8 or fp32
    holding 0..255 (the MiniGrid call site, clip_ppo_minigrid.py:381-388).  `noise` is fp32,
    logical [B,C,H,W] with any strides."""
    _require_cuda(obs, "fused_disturb_nhwc_u8")
    if obs.dim() != 4:
        raise ValueError(f"expected [B,H,W,C], got shape {tuple(obs.shape)}")
    if obs.dtype not in (torch.uint8, torch.float32):
        obs = obs.float()
    obs = obs.contiguous()
    B, H, W, Cc = obs.shape
    out = torch.empty((B, H, W, Cc), dtype=torch.uint8, device=obs.device)
    if B == 0:
        return out
    nptr, nstr = None, None
    if stages & N.STAGE_NOISE:
        if noise is None or tuple(noise.shape) != (B, Cc, H, W) or noise.dtype != torch.float32:
            raise ValueError("noise must be fp32 with logical shape [B,C,H,W]")
        nptr, nstr = noise.data_ptr(), N.strides4(noise)
    k = len(taps) if (stages & N.STAGE_BLUR) else 0
    taps_arr = (C.c_float * max(k, 1))(*(taps if k else (1.0,)))
    sh, sw, ph, pw = window
    with N.device_ctx(obs.device):
        st = N.lib().clipppo_disturb_nhwc_u8(
            obs.data_ptr(), int(obs.dtype == torch.float32), nptr, nstr, out.data_ptr(), B, H, W, Cc, stages,
            float(noise_sigma), float(contrast), taps_arr, k, int(sh), int(sw), int(ph), int(pw),
            _stream_ptr(obs))
    N.check(st, "clipppo_disturb_nhwc_u8")
    return out
