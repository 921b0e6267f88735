"""Host side of the frozen CLIP image tower: weight repacking + the C-ABI encode call.

``VitEngine`` takes a state dict in openai/CLIP ``visual.*`` key layout (what
``clip.load(...).state_dict()`` holds), repacks it once into the bf16 / fp32 device tensors the
kernels want, creates the native handle (``clipppo_vit_create``) and then serves
``encode(images, pre_scale, l2norm)`` = preprocessing + tower + optional L2 normalise in one
C-ABI call (``clipppo_vit_encode``).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import _native as N


@dataclass(frozen=True)
class TowerConfig:
    width: int
    layers: int
    heads: int
    patch: int
    image: int
    out_dim: int

    @property
    def grid(self) -> int:
        return self.image // self.patch

    @property
    def tokens(self) -> int:
        return self.grid * self.grid + 1


def config_from_state_dict(sd: Dict[str, torch.Tensor], prefix: str = "visual.") -> TowerConfig:
    w = sd[prefix + "conv1.weight"]
    D, P = int(w.shape[0]), int(w.shape[-1])
    T = int(sd[prefix + "positional_embedding"].shape[0])
    G = int(round(math.sqrt(T - 1)))
    blocks = {k[len(prefix):].split(".")[2] for k in sd if k.startswith(prefix + "transformer.resblocks.")}
    return TowerConfig(width=D, layers=len(blocks), heads=D // 64, patch=P, image=G * P,
                       out_dim=int(sd[prefix + "proj"].shape[1]))


class VitEngine:
    """Frozen tower on one GPU.  Owns the repacked weights; the native handle only borrows them."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device: torch.device | str = "cuda",
                 prefix: str = "visual."):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("VitEngine needs a CUDA device - the B200 path has no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.cfg = config_from_state_dict(state_dict, prefix)
        cfg = self.cfg
        D, P = cfg.width, cfg.patch
        self._keep = []          # every tensor the native handle points into

        def dev(t, dtype):
            t = t.detach().to(device=self.device, dtype=dtype).contiguous()
            self._keep.append(t)
            return t

        def g(name):
            return state_dict[prefix + name]

        kreal = 3 * P * P
        kpad = (kreal + 63) // 64 * 64
        wp = g("conv1.weight").detach().float().reshape(D, kreal)
        if kpad != kreal:
            wp = torch.nn.functional.pad(wp, (0, kpad - kreal))
        W = N.VitWeights()
        W.w_patch = dev(wp, torch.bfloat16).data_ptr()
        pos = g("positional_embedding").detach().float()
        W.cls_pos0 = dev(g("class_embedding").detach().float() + pos[0], torch.float32).data_ptr()
        W.pos = dev(pos, torch.float32).data_ptr()
        W.ln_pre_g = dev(g("ln_pre.weight"), torch.float32).data_ptr()
        W.ln_pre_b = dev(g("ln_pre.bias"), torch.float32).data_ptr()
        W.ln_post_g = dev(g("ln_post.weight"), torch.float32).data_ptr()
        W.ln_post_b = dev(g("ln_post.bias"), torch.float32).data_ptr()
        W.w_head = dev(g("proj").detach().float().t(), torch.bfloat16).data_ptr()      # [out, D], K-major
        layers = (N.VitLayer * cfg.layers)()
        for i in range(cfg.layers):
            p = f"transformer.resblocks.{i}."
            L = layers[i]
            L.w_qkv = dev(g(p + "attn.in_proj_weight"), torch.bfloat16).data_ptr()
            L.b_qkv = dev(g(p + "attn.in_proj_bias"), torch.float32).data_ptr()
            L.w_out = dev(g(p + "attn.out_proj.weight"), torch.bfloat16).data_ptr()
            L.b_out = dev(g(p + "attn.out_proj.bias"), torch.float32).data_ptr()
            L.w_fc = dev(g(p + "mlp.c_fc.weight"), torch.bfloat16).data_ptr()
            L.b_fc = dev(g(p + "mlp.c_fc.bias"), torch.float32).data_ptr()
            L.w_proj = dev(g(p + "mlp.c_proj.weight"), torch.bfloat16).data_ptr()
            L.b_proj = dev(g(p + "mlp.c_proj.bias"), torch.float32).data_ptr()
            L.ln1_g = dev(g(p + "ln_1.weight"), torch.float32).data_ptr()
            L.ln1_b = dev(g(p + "ln_1.bias"), torch.float32).data_ptr()
            L.ln2_g = dev(g(p + "ln_2.weight"), torch.float32).data_ptr()
            L.ln2_b = dev(g(p + "ln_2.bias"), torch.float32).data_ptr()
        W.layers_host = C.cast(layers, C.POINTER(N.VitLayer))
        ncfg = N.VitConfig(cfg.width, cfg.layers, cfg.heads, cfg.patch, cfg.image, cfg.out_dim)
        self._handle = C.c_void_p()
        with torch.cuda.device(self.device):
            N.check(N.lib().clipppo_vit_create(C.byref(self._handle), C.byref(ncfg), C.byref(W)), "clipppo_vit_create")
        self._workspace: Optional[torch.Tensor] = None

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h:
            try:
                N.lib().clipppo_vit_destroy(h)
            except Exception:
                pass
            self._handle = None

    def _workspace_for(self, n: int) -> torch.Tensor:
        need = C.c_size_t()
        N.check(N.lib().clipppo_vit_workspace_bytes(self._handle, n, C.byref(need)), "clipppo_vit_workspace_bytes")
        if self._workspace is None or self._workspace.numel() < need.value:
            self._workspace = None
            self._workspace = torch.empty(need.value, dtype=torch.uint8, device=self.device)
        return self._workspace

    @torch.no_grad()
    def encode_normalized(self, x: torch.Tensor) -> torch.Tensor:
        """Tower only, on an already resized-and-normalised batch (what upstream's
        ``encode_image`` / ``VisionTransformer.forward`` receive)."""
        return self.encode(x, pre_scale=1.0, l2norm=False, prenormalized=True)

    @torch.no_grad()
    def encode(self, images: torch.Tensor, pre_scale: float = 1.0 / 255.0, l2norm: bool = True,
               prenormalized: bool = False) -> torch.Tensor:
        """images [N,C,h,w] (fp32 or uint8, any strides, C in {1,3}, h,w <= 224) -> fp32 [N,out_dim]."""
        if images.device != self.device:
            raise RuntimeError(f"images on {images.device}, tower on {self.device}")
        if images.dim() != 4:
            raise ValueError(f"expected [N,C,h,w], got {tuple(images.shape)}")
        if images.dtype not in (torch.float32, torch.uint8):
            images = images.float()
        n, c, h, w = images.shape
        out = torch.empty((n, self.cfg.out_dim), dtype=torch.float32, device=self.device)
        if n == 0:
            return out
        ws = self._workspace_for(n)
        with torch.cuda.device(self.device):
            st = N.lib().clipppo_vit_encode(
                self._handle, images.data_ptr(), N.IMG_U8 if images.dtype == torch.uint8 else N.IMG_F32,
                N.strides4(images), n, c, h, w, float(pre_scale),
                (N.VIT_L2NORM if l2norm else 0) | (N.VIT_PRENORMALIZED if prenormalized else 0), out.data_ptr(),
                ws.data_ptr(), ws.numel(), torch.cuda.current_stream(self.device).cuda_stream)
        N.check(st, "clipppo_vit_encode")
        return out
