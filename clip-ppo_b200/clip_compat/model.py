"""``clip.model`` surface: ``VisionTransformer`` and ``CLIP`` classes backed by ``VitEngine``.

The modules hold the frozen weights as fp32 ``nn.Parameter``s under the upstream names (so
``state_dict()`` / ``parameters()`` / checkpointing keep working, reference
clip_ppo_minigrid.py:223-226) and lazily build the native engine on the device they live on.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn

from ..text import TextEngine, TextTowerConfig, text_config_from_state_dict
from ..vit import TowerConfig, VitEngine, config_from_state_dict

# name -> (width, layers, patch, image, out_dim)
ARCHS = {
    "ViT-B/32": (768, 12, 32, 224, 512),
    "ViT-B/16": (768, 12, 16, 224, 512),
    "ViT-L/14": (1024, 24, 14, 224, 768),
}


# name -> text tower (width, layers, heads, out_dim); context 77 and vocabulary 49408 in every release
TEXT_ARCHS = {
    "ViT-B/32": (512, 12, 8, 512),
    "ViT-B/16": (512, 12, 8, 512),
    "ViT-L/14": (768, 12, 12, 768),
}
TEXT_CONTEXT, TEXT_VOCAB = 77, 49408


def random_text_state_dict(name: str = "ViT-B/32", seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded random text-tower weights in upstream key layout (``token_embedding.weight``,
    ``positional_embedding``, ``transformer.resblocks.{i}.*``, ``ln_final.*``, ``text_projection``)."""
    D, L, _, O = TEXT_ARCHS[name]
    g = torch.Generator().manual_seed(seed + 1000)

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    sd: Dict[str, torch.Tensor] = {}
    sd["token_embedding.weight"] = rn(TEXT_VOCAB, D, std=0.02)
    sd["positional_embedding"] = rn(TEXT_CONTEXT, D, std=0.01)
    attn_std, proj_std, fc_std = D ** -0.5, (D ** -0.5) * ((2 * L) ** -0.5), (2 * D) ** -0.5
    for i in range(L):
        p = f"transformer.resblocks.{i}."
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
    sd["ln_final.weight"] = 1.0 + rn(D, std=0.1)
    sd["ln_final.bias"] = rn(D, std=0.1)
    sd["text_projection"] = rn(D, O, std=D ** -0.5)
    return sd


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


class _TextTower(nn.Module):
    """Holder of the text-side parameters, which live at the TOP level of upstream's ``CLIP`` state dict
    (``token_embedding.weight`` ... ``text_projection``): this module keeps them under mangled names and
    writes / reads them without its own name in the key."""

    _OWN = "_text."

    def __init__(self, text_state_dict: Dict[str, torch.Tensor], device="cuda"):
        super().__init__()
        self.config: TextTowerConfig = text_config_from_state_dict(text_state_dict)
        self._names = {}
        for k, v in text_state_dict.items():
            pname = k.replace(".", "__")
            self._names[pname] = k
            self.register_parameter(pname, nn.Parameter(v.detach().clone().float(), requires_grad=False))
        self._engine = None
        self._engine_device = None
        self.to(device)

    def _strip(self, prefix: str) -> str:
        return prefix[:-len(self._OWN)] if prefix.endswith(self._OWN) else prefix

    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        out = {} if destination is None else destination
        for pname, k in self._names.items():
            p = getattr(self, pname)
            out[self._strip(prefix) + k] = p if keep_vars else p.detach()
        return out

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        pass           # a child only sees keys under its own name; CLIP._load_from_state_dict feeds load_top_level

    def load_top_level(self, state_dict, prefix, strict, missing_keys):
        for pname, k in self._names.items():
            if prefix + k in state_dict:
                getattr(self, pname).data.copy_(state_dict[prefix + k])
            elif strict:
                missing_keys.append(prefix + k)
        self._engine = None

    def engine(self) -> TextEngine:
        dev = next(self.parameters()).device
        if self._engine is None or self._engine_device != dev:
            sd = {k: getattr(self, pname).detach() for pname, k in self._names.items()}
            self._engine = TextEngine(sd, device=dev, prefix="")
            self._engine_device = dev
        return self._engine


class CLIP(nn.Module):
    """``clip.model.CLIP`` surface: ``.visual``, ``.encode_image``, ``.encode_text``, ``.dtype``."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda"):
        super().__init__()
        vis = {k[len("visual."):]: v for k, v in state_dict.items() if k.startswith("visual.")}
        if not vis:
            raise ValueError("state dict holds no 'visual.*' entries")
        self.visual = VisionTransformer(vis, device=device)
        text = {k: v for k, v in state_dict.items()
                if k.startswith(("token_embedding.", "transformer.resblocks.", "ln_final.")) or k in ("positional_embedding", "text_projection")}
        self._text = _TextTower(text, device=device) if "token_embedding.weight" in text else None

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        if self._text is not None:
            self._text.load_top_level(state_dict, prefix, strict, missing_keys)

    @property
    def dtype(self):
        return torch.float32

    def encode_image(self, image: torch.Tensor) -> torch.Tensor:
        return self.visual(image)

    def text_engine(self) -> TextEngine:
        if self._text is None:
            raise NotImplementedError("this CLIP was built from a state dict without text-tower weights")
        return self._text.engine()

    @torch.no_grad()
    def encode_text(self, text: torch.Tensor) -> torch.Tensor:
        """[N, 77] token ids -> [N, out_dim] (not normalised), like upstream's."""
        return self.text_engine().encode(text, l2norm=False)

    def forward(self, image, text):
        raise NotImplementedError("joint image/text forward (logits) is outside the B200 hot path")
