"""``clip.model`` surface: ``VisionTransformer`` and ``CLIP`` classes backed by ``VitEngine``.

The modules hold the frozen weights as fp32 ``nn.Parameter``s under the upstream names (so
``state_dict()`` / ``parameters()`` / checkpointing keep working, reference
clip_ppo_minigrid.py:223-226) and lazily build the native engine on the device they live on.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn

from ..vit import TowerConfig, VitEngine, config_from_state_dict

# name -> (width, layers, patch, image, out_dim)
ARCHS = {
    "ViT-B/32": (768, 12, 32, 224, 512),
    "ViT-B/16": (768, 12, 16, 224, 512),
    "ViT-L/14": (1024, 24, 14, 224, 768),
}


def random_visual_state_dict(name: str = "ViT-B/32", seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded random weights in upstream key layout, for benches and tests (no real CLIP weights
    exist on these machines).  LayerNorm gains ~1±0.1 and biases ~±0.1 so every affine path is
    exercised.  Generated on the CPU generator so every rank / box gets identical weights."""
    D, L, P, I, O = ARCHS[name]
    T = (I // P) ** 2 + 1
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    sd: Dict[str, torch.Tensor] = {}
    scale = D ** -0.5
    sd["visual.conv1.weight"] = rn(D, 3, P, P, std=(3 * P * P) ** -0.5)
    sd["visual.class_embedding"] = rn(D, std=scale)
    sd["visual.positional_embedding"] = rn(T, D, std=scale)
    for n in ("ln_pre", "ln_post"):
        sd[f"visual.{n}.weight"] = 1.0 + rn(D, std=0.1)
        sd[f"visual.{n}.bias"] = rn(D, std=0.1)
    attn_std, proj_std, fc_std = D ** -0.5, (D ** -0.5) * ((2 * L) ** -0.5), (2 * D) ** -0.5
    for i in range(L):
        p = f"visual.transformer.resblocks.{i}."
        for n in ("ln_1", "ln_2"):
            sd[p + f"{n}.weight"] = 1.0 + rn(D, std=0.1)
            sd[p + f"{n}.bias"] = rn(D, std=0.1)
        sd[p + "attn.in_proj_weight"] = rn(3 * D, D, std=attn_std)
        sd[p + "attn.in_proj_bias"] = rn(3 * D, std=0.02)
        sd[p + "attn.out_proj.weight"] = rn(D, D, std=proj_std)
        sd[p + "attn.out_proj.bias"] = rn(D, std=0.02)
        sd[p + "mlp.c_fc.weight"] = rn(4 * D, D, std=fc_std)
        sd[p + "mlp.c_fc.bias"] = rn(4 * D, std=0.02)
        sd[p + "mlp.c_proj.weight"] = rn(D, 4 * D, std=proj_std)
        sd[p + "mlp.c_proj.bias"] = rn(D, std=0.02)
    sd["visual.proj"] = rn(D, O, std=scale)
    return sd


class VisionTransformer(nn.Module):
    """Frozen image tower; ``forward(x)`` takes an already resized + normalised [N,3,224,224] batch
    like upstream's.  The fused resize/normalise path is ``encode_raw``."""

    def __init__(self, visual_state_dict: Dict[str, torch.Tensor], device="cuda"):
        super().__init__()
        self.config: TowerConfig = config_from_state_dict(visual_state_dict, prefix="")
        self._names = {}
        for k, v in visual_state_dict.items():
            pname = k.replace(".", "__")
            self._names[pname] = k
            self.register_parameter(pname, nn.Parameter(v.detach().clone().float(), requires_grad=False))
        self.output_dim = self.config.out_dim
        self.input_resolution = self.config.image
        self._engine = None
        self._engine_device = None
        self.to(device)

    # upstream-style key names in checkpoints
    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        out = {} if destination is None else destination
        for pname, k in self._names.items():
            p = getattr(self, pname)
            out[prefix + k] = p if keep_vars else p.detach()
        return out

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        for pname, k in self._names.items():
            if prefix + k in state_dict:
                getattr(self, pname).data.copy_(state_dict[prefix + k])
            elif strict:
                missing_keys.append(prefix + k)
        self._engine = None

    def engine(self) -> VitEngine:
        dev = next(self.parameters()).device
        if self._engine is None or self._engine_device != dev:
            sd = {k: getattr(self, pname).detach() for pname, k in self._names.items()}
            self._engine = VitEngine(sd, device=dev, prefix="")
            self._engine_device = dev
        return self._engine

    @property
    def conv1(self):          # upstream exposes conv1.weight.dtype via CLIP.dtype
        class _W:             # noqa: N801
            pass
        w = _W()
        w.weight = getattr(self, "conv1__weight")
        return w

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        # x is normalised already: undo nothing, just run the tower on it.  The native
        # preprocessing would normalise again, so feed it through the "identity" entry:
        return self.engine().encode_normalized(x)

    @torch.no_grad()
    def encode_raw(self, images: torch.Tensor, pre_scale: float, l2norm: bool) -> torch.Tensor:
        return self.engine().encode(images, pre_scale=pre_scale, l2norm=l2norm)


class CLIP(nn.Module):
    """``clip.model.CLIP`` surface: ``.visual``, ``.encode_image``, ``.dtype``; the text side raises."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda"):
        super().__init__()
        vis = {k[len("visual."):]: v for k, v in state_dict.items() if k.startswith("visual.")}
        if not vis:
            raise ValueError("state dict holds no 'visual.*' entries")
        self.visual = VisionTransformer(vis, device=device)

    @property
    def dtype(self):
        return torch.float32

    def encode_image(self, image: torch.Tensor) -> torch.Tensor:
        return self.visual(image)

    def encode_text(self, text):
        raise NotImplementedError("text tower is outside the B200 hot path (SURVEY.md §8f)")

    def forward(self, image, text):
        raise NotImplementedError("joint image/text forward is outside the B200 hot path")
