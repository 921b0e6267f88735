"""Host side of the frozen CLIP text tower: weight hand-over and the encode call.

Replaces ``clip_model.encode_text`` as the reference calls it at shared/clip_ppo_utils.py:136-139
(``clip.tokenize(descriptions)`` -> ``encode_text`` -> ``.float()`` -> ``F.normalize``), the default
MiniGrid modality (``ClipPPOConfig.clip_modality = "text"``).  Token ids are an input: the BPE tokenizer
is host code of the openai package (its merges file is not reproducible offline).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import _native as N


@dataclass(frozen=True)
class TextTowerConfig:
    width: int = 512
    layers: int = 12
    heads: int = 8
    context: int = 77
    vocab: int = 49408
    out_dim: int = 512


def text_config_from_state_dict(sd: Dict[str, torch.Tensor], prefix: str = "") -> TextTowerConfig:
    V, D = sd[prefix + "token_embedding.weight"].shape
    blocks = {k[len(prefix):].split(".")[2] for k in sd if k.startswith(prefix + "transformer.resblocks.")}
    return TextTowerConfig(width=D, layers=len(blocks), heads=D // 64, context=sd[prefix + "positional_embedding"].shape[0],
                           vocab=V, out_dim=sd[prefix + "text_projection"].shape[1])


class TextEngine:
    """Frozen text tower on one GPU.  The native handle owns the repacked weights."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device: torch.device | str = "cuda", prefix: str = ""):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("TextEngine needs a CUDA device - the B200 path has no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.cfg = cfg = text_config_from_state_dict(state_dict, prefix)
        keep = []                # fp32 staging copies: only needed until clipppo_text_create returns

        def dev(name):
            t = state_dict[prefix + name].detach().to(device=self.device, dtype=torch.float32).contiguous()
            keep.append(t)
            return t.data_ptr()

        W = N.TextWeights()
        W.token_embedding = dev("token_embedding.weight")
        W.positional_embedding = dev("positional_embedding")
        W.ln_final_g, W.ln_final_b = dev("ln_final.weight"), dev("ln_final.bias")
        W.text_projection = dev("text_projection")
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
        ncfg = N.TextConfig(cfg.width, cfg.layers, cfg.heads, cfg.context, cfg.vocab, cfg.out_dim)
        self._handle = C.c_void_p()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)          # staging copies above ran on torch's stream
            N.check(N.lib().clipppo_text_create(C.byref(self._handle), C.byref(ncfg), C.byref(W)), "clipppo_text_create")
        del keep
        self._workspaces = {}                # one cached workspace per CUDA stream (concurrent passes on two streams must not share one)

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h:
            try:
                N.lib().clipppo_text_destroy(h)
            except Exception:
                pass
            self._handle = None

    def _workspace_for(self, n: int) -> torch.Tensor:
        need = C.c_size_t()
        N.check(N.lib().clipppo_text_workspace_bytes(self._handle, n, C.byref(need)), "clipppo_text_workspace_bytes")
        key = torch.cuda.current_stream(self.device).cuda_stream
        ws = self._workspaces.pop(key, None)
        if ws is None or ws.numel() < need.value:
            ws = None
            while len(self._workspaces) >= 4:
                self._workspaces.pop(next(iter(self._workspaces)))
            ws = torch.empty(need.value, dtype=torch.uint8, device=self.device)
        self._workspaces[key] = ws
        return ws

    @torch.no_grad()
    def encode(self, tokens: torch.Tensor, l2norm: bool = True) -> torch.Tensor:
        """tokens [N, context] integer ids (``clip.tokenize`` output) -> fp32 [N, out_dim]."""
        if tokens.device != self.device:
            raise RuntimeError(f"tokens on {tokens.device}, tower on {self.device}")
        if tokens.dim() != 2 or tokens.shape[1] != self.cfg.context:
            raise ValueError(f"expected [N, {self.cfg.context}] token ids, got {tuple(tokens.shape)}")
        if tokens.dtype.is_floating_point or tokens.dtype == torch.bool:
            raise TypeError(f"token ids must be integers, got {tokens.dtype}")
        n = tokens.shape[0]
        out = torch.empty((n, self.cfg.out_dim), dtype=torch.float32, device=self.device)
        if n == 0:
            return out
        lo, hi = (int(v) for v in torch.stack(torch.aminmax(tokens)).tolist())      # one read-back for both bounds
        if lo < 0 or hi >= self.cfg.vocab:                                          # nn.Embedding's IndexError, raised on the host
            raise IndexError(f"token id out of range [0, {self.cfg.vocab})")
        tok = tokens.to(torch.int32).contiguous()
        ws = self._workspace_for(n)
        with N.device_ctx(self.device):
            st = N.lib().clipppo_text_encode(self._handle, tok.data_ptr(), n, N.VIT_L2NORM if l2norm else 0, out.data_ptr(),
                                             ws.data_ptr(), ws.numel(), torch.cuda.current_stream(self.device).cuda_stream)
        N.check(st, "clipppo_text_encode")
        return out
