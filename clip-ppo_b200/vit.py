"""Host side of the frozen CLIP image tower: weight repacking + the C-ABI encode call.

``VitEngine`` takes a state dict in openai/CLIP ``visual.*`` key layout (what
``clip.load(...).state_dict()`` holds), stages it on the device as fp32, lets the native side repack
it into handle-owned bf16 operands (``clipppo_vit_create``: K-major GEMM weights, ln_1 / ln_2 folded
into in_proj / c_fc) and then serves
``encode(images, pre_scale, l2norm)`` = preprocessing + tower + optional L2 normalise in one
C-ABI call (``clipppo_vit_encode``).
"""
from __future__ import annotations

import ctypes as C
import math
import os
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


def _cls_last_flag() -> int:
    """CLIPPPO_VIT_CLS_LAST_BLOCK=1 (opt-in, read per call): after the last block's attention only the class-token rows go
    through out_proj / c_fc / c_proj.  Exact - ``VisionTransformer.forward`` reads ``x[:, 0, :]`` only, embeddings are bitwise
    equal - and 6 % fewer FLOPs on ViT-B/32.  Off by default: the default pass runs every token through every block."""
    return N.VIT_CLS_LAST_BLOCK if os.environ.get("CLIPPPO_VIT_CLS_LAST_BLOCK", "") not in ("", "0") else 0


class VitEngine:
    """Frozen tower on one GPU.  The native handle owns the repacked weights."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device: torch.device | str = "cuda",
                 prefix: str = "visual."):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("VitEngine needs a CUDA device - the B200 path has no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.cfg = config_from_state_dict(state_dict, prefix)
        cfg = self.cfg
        keep = []                # fp32 staging copies: only needed until clipppo_vit_create returns

        def dev(name):
            t = state_dict[prefix + name].detach().to(device=self.device, dtype=torch.float32).contiguous()
            keep.append(t)
            return t.data_ptr()

        W = N.VitWeights()
        W.conv1 = dev("conv1.weight")
        W.class_embedding = dev("class_embedding")
        W.positional_embedding = dev("positional_embedding")
        W.ln_pre_g, W.ln_pre_b = dev("ln_pre.weight"), dev("ln_pre.bias")
        W.ln_post_g, W.ln_post_b = dev("ln_post.weight"), dev("ln_post.bias")
        W.proj = dev("proj")
        layers = (N.VitLayer * cfg.layers)()
        for i in range(cfg.layers):
            p = f"transformer.resblocks.{i}."
            L = layers[i]
            L.ln1_g, L.ln1_b = dev(p + "ln_1.weight"), dev(p + "ln_1.bias")
            L.w_qkv, L.b_qkv = dev(p + "attn.in_proj_weight"), dev(p + "attn.in_proj_bias")
            L.w_out, L.b_out = dev(p + "attn.out_proj.weight"), dev(p + "attn.out_proj.bias")
            L.ln2_g, L.ln2_b = dev(p + "ln_2.weight"), dev(p + "ln_2.bias")
            L.w_fc, L.b_fc = dev(p + "mlp.c_fc.weight"), dev(p + "mlp.c_fc.bias")
            L.w_proj, L.b_proj = dev(p + "mlp.c_proj.weight"), dev(p + "mlp.c_proj.bias")
        W.layers_host = C.cast(layers, C.POINTER(N.VitLayer))
        ncfg = N.VitConfig(cfg.width, cfg.layers, cfg.heads, cfg.patch, cfg.image, cfg.out_dim)
        self._handle = C.c_void_p()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)          # staging copies above ran on torch's stream
            N.check(N.lib().clipppo_vit_create(C.byref(self._handle), C.byref(ncfg), C.byref(W)), "clipppo_vit_create")
        del keep                                         # create() has repacked into handle-owned memory
        self._workspaces: Dict[int, torch.Tensor] = {}   # one cached workspace per CUDA stream that has called encode()
        self._graphs: Dict[tuple, tuple] = {}            # small-batch schedule: one captured CUDA graph per call shape

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h:
            try:
                N.lib().clipppo_vit_destroy(h)
            except Exception:
                pass
            self._handle = None

    WORKSPACE_STREAMS = 4             # cached workspaces kept (least recently used stream evicted)

    def _workspace_for(self, n: int) -> torch.Tensor:
        """The activation workspace of a tower pass, cached PER STREAM: two passes enqueued on different streams may run
        concurrently and must not share one (a pass owns its workspace from its first kernel to its last)."""
        need = C.c_size_t()
        N.check(N.lib().clipppo_vit_workspace_bytes(self._handle, n, C.byref(need)), "clipppo_vit_workspace_bytes")
        key = torch.cuda.current_stream(self.device).cuda_stream
        ws = self._workspaces.pop(key, None)
        if ws is None or ws.numel() < need.value:
            ws = None
            while len(self._workspaces) >= self.WORKSPACE_STREAMS:
                self._workspaces.pop(next(iter(self._workspaces)))
            ws = torch.empty(need.value, dtype=torch.uint8, device=self.device)
        self._workspaces[key] = ws
        return ws

    # A tower pass is ~90 launches; below a few hundred images the host cannot enqueue them as fast as the
    # GPU runs them (measured on B200: 64 frames = 1.0 ms eager, of which 0.8 ms is host enqueue).  That is
    # the FROZEN_CLIP policy path (reference clip_ppo_minigrid.py:249-254, clip_ppo_atari.py:213-228: E
    # frames inside every policy forward), so small batches replay one captured CUDA graph instead.
    GRAPH_MAX_IMAGES = 512
    GRAPH_CACHE_ENTRIES = 4           # call shapes kept captured at a time (LRU): env-step batch, eval batch, a tail minibatch ...

    @torch.no_grad()
    def encode_graphed(self, images: torch.Tensor, pre_scale: float = 1.0 / 255.0, l2norm: bool = True,
                       prenormalized: bool = False) -> torch.Tensor:
        """Same result as :meth:`encode` (bitwise), one graph launch per call.  The graph owns a private
        input / output / workspace set, so later, larger ``encode`` calls cannot invalidate it."""
        if images.device != self.device:
            raise RuntimeError(f"images on {images.device}, tower on {self.device}")
        if images.dim() != 4:
            raise ValueError(f"expected [N,C,h,w], got {tuple(images.shape)}")
        if images.dtype not in (torch.float32, torch.uint8):
            images = images.float()
        key = (tuple(images.shape), images.dtype, float(pre_scale), bool(l2norm), bool(prenormalized), _cls_last_flag())
        ent = self._graphs.pop(key, None)                 # re-inserted below: the dict is kept in least-recently-used order
        if ent is None:
            while len(self._graphs) >= self.GRAPH_CACHE_ENTRIES:       # each entry owns an input, an output and a full workspace
                self._graphs.pop(next(iter(self._graphs)))
            n = images.shape[0]
            static_in = torch.empty(images.shape, dtype=images.dtype, device=self.device)
            static_out = torch.empty((n, self.cfg.out_dim), dtype=torch.float32, device=self.device)
            need = C.c_size_t()
            N.check(N.lib().clipppo_vit_workspace_bytes(self._handle, n, C.byref(need)), "clipppo_vit_workspace_bytes")
            ws = torch.empty(need.value, dtype=torch.uint8, device=self.device)
            static_in.copy_(images)
            cur = torch.cuda.current_stream(self.device)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):                       # warm-up outside capture (lazy kernel attributes)
                for _ in range(2):
                    self._encode_into(static_in, static_out, ws, pre_scale, l2norm, prenormalized)
            cur.wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._encode_into(static_in, static_out, ws, pre_scale, l2norm, prenormalized)
            ent = (graph, static_in, static_out, ws)
        self._graphs[key] = ent
        graph, static_in, static_out, _ = ent
        static_in.copy_(images)
        graph.replay()
        return static_out.clone()

    def _encode_into(self, images, out, ws, pre_scale, l2norm, prenormalized):
        n, c, h, w = images.shape
        with torch.cuda.device(self.device):
            st = N.lib().clipppo_vit_encode(
                self._handle, images.data_ptr(), N.IMG_U8 if images.dtype == torch.uint8 else N.IMG_F32,
                N.strides4(images), n, c, h, w, float(pre_scale),
                (N.VIT_L2NORM if l2norm else 0) | (N.VIT_PRENORMALIZED if prenormalized else 0) | _cls_last_flag(), out.data_ptr(),
                ws.data_ptr(), ws.numel(), torch.cuda.current_stream(self.device).cuda_stream)
        N.check(st, "clipppo_vit_encode")

    @torch.no_grad()
    def encode_normalized(self, x: torch.Tensor) -> torch.Tensor:
        """Tower only, on an already resized-and-normalised batch (what upstream's
        ``encode_image`` / ``VisionTransformer.forward`` receive)."""
        return self.encode(x, pre_scale=1.0, l2norm=False, prenormalized=True)

    @torch.no_grad()
    def encode(self, images: torch.Tensor, pre_scale: float = 1.0 / 255.0, l2norm: bool = True,
               prenormalized: bool = False) -> torch.Tensor:
        """images [N,C,h,w] (fp32 or uint8, any strides, C in {1,3}) -> fp32 [N,out_dim].  Frames up to the tower's
        resolution take the fused resize (up-sampling: antialias is a no-op); larger ones - never produced by the
        reference's environments - are first reduced by torch's antialiased bilinear resize, the op the reference
        itself calls (shared/clip_ppo_utils.py:151-157), in chunks so the fp32 intermediate stays bounded."""
        if images.device != self.device:
            raise RuntimeError(f"images on {images.device}, tower on {self.device}")
        if images.dim() != 4:
            raise ValueError(f"expected [N,C,h,w], got {tuple(images.shape)}")
        if images.dtype not in (torch.float32, torch.uint8):
            images = images.float()
        n, c, h, w = images.shape
        R = self.cfg.image
        if (h > R or w > R) and not prenormalized and n > 0:
            outs = []
            for i in range(0, n, 256):
                x = torch.nn.functional.interpolate(images[i:i + 256].float() * pre_scale, size=(R, R), mode="bilinear",
                                                    align_corners=False, antialias=True)
                outs.append(self.encode(x, pre_scale=1.0, l2norm=l2norm))
            return torch.cat(outs)
        out = torch.empty((n, self.cfg.out_dim), dtype=torch.float32, device=self.device)
        if n == 0:
            return out
        ws = self._workspace_for(n)
        with torch.cuda.device(self.device):
            st = N.lib().clipppo_vit_encode(
                self._handle, images.data_ptr(), N.IMG_U8 if images.dtype == torch.uint8 else N.IMG_F32,
                N.strides4(images), n, c, h, w, float(pre_scale),
                (N.VIT_L2NORM if l2norm else 0) | (N.VIT_PRENORMALIZED if prenormalized else 0) | _cls_last_flag(), out.data_ptr(),
                ws.data_ptr(), ws.numel(), torch.cuda.current_stream(self.device).cuda_stream)
        N.check(st, "clipppo_vit_encode")
        return out
