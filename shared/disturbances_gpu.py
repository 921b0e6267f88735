"""Drop-in for reference ``shared/disturbances_gpu.py``: same class, constructor, attributes and
methods, but every call is ONE launch of the fused sm_100a kernel (csrc/disturb.cu) through the
C ABI ``clipppo_disturb_f32`` instead of ~35 eager torchvision launches.

Randomness is consumed exactly as the reference consumes it, so a run seeded
like the reference draws the same noise tensor, contrast factor and cutout window:

    device generator : randn_like(obs)                                  (noise)
    CPU generator    : randperm(4) ; uniform_(lo, hi)                   (contrast, [tv] ColorJitter)
                       uniform_(sigma, sigma)                           (blur, [tv] GaussianBlur)
                       randint(0, H-ph+1) ; randint(0, W-pw+1)          (cutout)

Additive, non-breaking keyword arguments (``noise=``, ``contrast_factor=``, ``cutout_start=``)
let a caller supply the randomness instead; parity tests use them.  There is no CPU fallback.
"""
from __future__ import annotations

from typing import Optional, Tuple, Union

import numpy as np
import torch

from clip_ppo_b200 import _native as _N
from clip_ppo_b200 import disturb as _D
from shared.disturbance_types import DisturbanceSeverity, SEVERITY_CONFIGS


class _NoiseStage:
    """Stands where the reference keeps ``T2.GaussianNoise(mean=0.0, sigma=...)``."""

    def __init__(self, owner: "DisturbanceWrapperGPU"):
        self._owner, self.mean, self.sigma, self.clip = owner, 0.0, owner.gaussian_noise_sigma, True

    def __call__(self, obs):
        return self._owner.apply_gaussian_noise(obs)


class _ContrastStage:
    """Stands where the reference keeps ``T.ColorJitter(contrast=range)``."""

    def __init__(self, owner: "DisturbanceWrapperGPU"):
        self._owner, self.contrast = owner, tuple(float(v) for v in owner.contrast_range)
        self.brightness = self.saturation = self.hue = None

    def __call__(self, obs):
        return self._owner.apply_contrast_jitter(obs)


class _BlurStage:
    """Stands where the reference keeps ``T.GaussianBlur(kernel_size, sigma)``."""

    def __init__(self, owner: "DisturbanceWrapperGPU", kernel_size: int):
        self._owner = owner
        self.kernel_size = (kernel_size, kernel_size)
        self.sigma = (float(owner.gaussian_blur_sigma), float(owner.gaussian_blur_sigma))

    def __call__(self, obs):
        return self._owner.apply_gaussian_blur(obs)


class DisturbanceWrapperGPU:
    """Batched visual disturbances on [B,C,H,W] float tensors in [0,1]
    (reference shared/disturbances_gpu.py:14-194)."""

    def __init__(
        self,
        device: Union[str, torch.device] = "cuda",
        seed: Optional[int] = None,
        severity: Optional[DisturbanceSeverity] = DisturbanceSeverity.MILD,
        gaussian_noise_sigma: Optional[float] = None,
        gaussian_blur_sigma: Optional[float] = None,
        contrast_range: Optional[tuple] = None,
        cutout_ratio: Optional[float] = None,
    ):
        self.device = torch.device(device) if isinstance(device, str) else device
        if severity is not None:
            row = SEVERITY_CONFIGS[severity]
            gaussian_noise_sigma, gaussian_blur_sigma = row["gaussian_noise_sigma"], row["gaussian_blur_sigma"]
            contrast_range, cutout_ratio = row["contrast_range"], row["cutout_ratio"]
        elif None in (gaussian_noise_sigma, gaussian_blur_sigma, contrast_range, cutout_ratio):
            raise ValueError("All custom parameters must not be None if not setting a severity.")
        self.gaussian_noise_sigma = gaussian_noise_sigma
        self.gaussian_blur_sigma = gaussian_blur_sigma
        self.contrast_range = contrast_range
        self.cutout_ratio = cutout_ratio
        if seed is not None:
            torch.manual_seed(seed)          # global, like the reference (:54-55)
        self._philox_calls = 0               # counter offset of the opt-in in-kernel noise (apply_disturbances(noise_seed=))
        self._kernel_size = _D.blur_kernel_size(self.gaussian_blur_sigma)
        self.blur_transform = _BlurStage(self, self._kernel_size)
        self.noise_transform = _NoiseStage(self)
        self.contrast_transform = _ContrastStage(self)

    # ---- randomness, drawn in the reference's order ------------------------------------------
    def _draw_contrast(self) -> float:
        torch.randperm(4)                                                   # [tv] transforms.py:1259
        lo, hi = self.contrast_range
        return float(torch.empty(1).uniform_(lo, hi))                        # [tv] transforms.py:1262

    def _draw_blur_taps(self):
        s = self.gaussian_blur_sigma
        sigma = torch.empty(1).uniform_(s, s).item()                         # [tv] transforms.py:1809
        return _D.gaussian_taps(self._kernel_size, sigma)

    def _draw_cutout(self, H: int, W: int, start: Optional[Tuple[int, int]]):
        ph, pw = _D.cutout_patch(H, W, self.cutout_ratio)
        if start is None:
            sh = torch.randint(0, max(1, H - ph + 1), (1,)).item()           # reference :168-169
            sw = torch.randint(0, max(1, W - pw + 1), (1,)).item()
        else:
            sh, sw = start
        return int(sh), int(sw), ph, pw

    @staticmethod
    def _noise_like(obs: torch.Tensor) -> torch.Tensor:
        """`torch.randn_like(obs)` (reference: [tv] gaussian_noise_image); for the additive uint8 input, the noise
        `randn_like(obs.float() / 255)` would draw - same generator consumption, same values."""
        if obs.dtype == torch.uint8:
            return torch.randn(obs.shape, dtype=torch.float32, device=obs.device)
        return torch.randn_like(obs)

    # ---- tensor API --------------------------------------------------------------------------
    def apply_disturbances(self, obs: torch.Tensor, *, noise: Optional[torch.Tensor] = None,
                           contrast_factor: Optional[float] = None,
                           cutout_start: Optional[Tuple[int, int]] = None,
                           out_scale: float = 1.0, noise_seed: Optional[int] = None,
                           first_image: int = 0) -> torch.Tensor:
        """noise -> contrast -> blur -> cutout (reference :66-73), one fused launch.  Additive: uint8 `obs` holds
        0..255 pixels and equals `apply_disturbances(obs.float() / 255)` bit for bit (same RNG consumption).
        Additive keywords (defaults = the reference's behaviour):
          out_scale   `apply_disturbances(obs, out_scale=255.0)` == `apply_disturbances(obs) * 255` bit for bit, without
                      the extra pass over the batch (the call sites store 0..255: clip_ppo_atari.py:584);
          noise_seed  opt-in: the Gaussian noise is generated inside the kernel (Philox4x32-10 keyed by this seed, one
                      counter offset per call) instead of `torch.randn_like(obs)` - N(0,1) draws, but NOT torch's stream
                      and the device generator is not consumed; `first_image` is the global index of obs[0] when obs is a
                      shard of a larger batch (the shards then draw the whole batch's noise)."""
        philox = None
        if noise is None and noise_seed is not None:
            philox = (noise_seed, self._philox_calls, first_image)
            self._philox_calls += 1
        elif noise is None:
            noise = self._noise_like(obs)
        c = self._draw_contrast() if contrast_factor is None else float(contrast_factor)
        taps = self._draw_blur_taps()
        window = self._draw_cutout(obs.shape[-2], obs.shape[-1], cutout_start)
        return _D.fused_disturb(obs, stages=_N.STAGE_ALL, noise=noise, noise_sigma=self.gaussian_noise_sigma,
                                contrast=c, taps=taps, window=window, out_scale=out_scale, philox=philox)

    def apply_gaussian_noise(self, obs: torch.Tensor, *, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        if noise is None:
            noise = self._noise_like(obs)
        return _D.fused_disturb(obs, stages=_N.STAGE_NOISE, noise=noise, noise_sigma=self.gaussian_noise_sigma)

    def apply_contrast_jitter(self, obs: torch.Tensor, *, contrast_factor: Optional[float] = None) -> torch.Tensor:
        c = self._draw_contrast() if contrast_factor is None else float(contrast_factor)
        return _D.fused_disturb(obs, stages=_N.STAGE_CONTRAST, contrast=c)

    def apply_gaussian_blur(self, obs: torch.Tensor) -> torch.Tensor:
        return _D.fused_disturb(obs, stages=_N.STAGE_BLUR, taps=self._draw_blur_taps())

    def apply_cutout(self, obs: torch.Tensor, *, cutout_start: Optional[Tuple[int, int]] = None) -> torch.Tensor:
        window = self._draw_cutout(obs.shape[-2], obs.shape[-1], cutout_start)
        return _D.fused_disturb(obs, stages=_N.STAGE_CUTOUT, window=window)

    # ---- numpy API (uint8 HWC / BHWC in and out; reference :75-95, 101-115, 121-135, 141-155, 174-194)
    def _numpy_call(self, obs: np.ndarray, stages: int) -> np.ndarray:
        single = obs.ndim == 3
        u8 = torch.from_numpy(np.ascontiguousarray(obs)).to(self.device)
        if single:
            u8 = u8.unsqueeze(0)
        B, H, W, Cc = u8.shape
        kw = {}
        if stages & _N.STAGE_NOISE:
            # same strides as the reference's permuted float view => same randn_like stream
            like = torch.empty((B, H, W, Cc), dtype=torch.float32, device=self.device).permute(0, 3, 1, 2)
            kw.update(noise=torch.randn_like(like), noise_sigma=self.gaussian_noise_sigma)
        if stages & _N.STAGE_CONTRAST:
            kw.update(contrast=self._draw_contrast())
        if stages & _N.STAGE_BLUR:
            kw.update(taps=self._draw_blur_taps())
        if stages & _N.STAGE_CUTOUT:
            kw.update(window=self._draw_cutout(H, W, None))
        out = _D.fused_disturb_nhwc_u8(u8, stages=stages, **kw)
        if single:
            out = out.squeeze(0)
        return out.cpu().numpy()

    def apply_disturbances_numpy(self, obs: np.ndarray) -> np.ndarray:
        return self._numpy_call(obs, _N.STAGE_ALL)

    def apply_gaussian_noise_numpy(self, obs: np.ndarray) -> np.ndarray:
        return self._numpy_call(obs, _N.STAGE_NOISE)

    def apply_contrast_jitter_numpy(self, obs: np.ndarray) -> np.ndarray:
        return self._numpy_call(obs, _N.STAGE_CONTRAST)

    def apply_gaussian_blur_numpy(self, obs: np.ndarray) -> np.ndarray:
        return self._numpy_call(obs, _N.STAGE_BLUR)

    def apply_cutout_numpy(self, obs: np.ndarray) -> np.ndarray:
        return self._numpy_call(obs, _N.STAGE_CUTOUT)


def create_disturbance_wrapper(use_gpu=True, **kwargs):
    """Factory with the reference's signature (:198-214).  The reference falls back to its cv2
    CPU twin without CUDA; this path has no CPU fallback and raises instead."""
    if use_gpu and torch.cuda.is_available():
        return DisturbanceWrapperGPU(device="cuda", **kwargs)
    raise RuntimeError("create_disturbance_wrapper: CUDA is required (use_gpu=True on a GPU box); "
                       "the B200 implementation ships no CPU fallback")
