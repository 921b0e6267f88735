"""Oracle (CPU, fp32 torch) for the frozen CLIP image tower.  TEST INFRASTRUCTURE ONLY.

The tower's arithmetic lives in openai/CLIP (``clip/model.py``; reference
``requirements.txt:12``, un-vendored, unpinned HEAD, absent from /root/reference and from
this image).  This restates its published algorithm as plain
functions over a state dict in openai key layout (``visual.conv1.weight`` ...), and the
reference-side preprocessing around it (``shared/clip_ppo_utils.py:141-164`` and
``:185-217``).

Pinned by: ``tests/test_oracle.py::test_vit_oracle_matches_hf_small`` (transformers
CLIPVisionModelWithProjection with the same weights) and tests/golden/vit_*.npz.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict

import torch
import torch.nn.functional as F

# reference shared/clip_ppo_utils.py:21-22
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


@dataclass(frozen=True)
class VitConfig:
    width: int = 768
    layers: int = 12
    heads: int = 12
    patch: int = 32
    image: int = 224
    out_dim: int = 512

    @property
    def grid(self) -> int:
        return self.image // self.patch

    @property
    def tokens(self) -> int:
        return self.grid * self.grid + 1


VIT_B32 = VitConfig()
VIT_L14 = VitConfig(width=1024, layers=24, heads=16, patch=14, image=224, out_dim=768)


def random_state_dict(cfg: VitConfig = VIT_B32, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded random weights in openai/CLIP ``visual.*`` key layout.
    Scales follow upstream's initialisation orders of magnitude; LN gamma ~ 1±0.1 and
    beta ~ ±0.1 so the affine paths are exercised."""
    g = torch.Generator().manual_seed(seed)
    D, L, P, T, O = cfg.width, cfg.layers, cfg.patch, cfg.tokens, cfg.out_dim

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    sd: Dict[str, torch.Tensor] = {}
    scale = D ** -0.5
    sd["visual.conv1.weight"] = rn(D, 3, P, P, std=(3 * P * P) ** -0.5)
    sd["visual.class_embedding"] = rn(D, std=scale)
    sd["visual.positional_embedding"] = rn(T, D, std=scale)
    for name in ("ln_pre", "ln_post"):
        sd[f"visual.{name}.weight"] = 1.0 + rn(D, std=0.1)
        sd[f"visual.{name}.bias"] = rn(D, std=0.1)
    attn_std = D ** -0.5
    proj_std = (D ** -0.5) * ((2 * L) ** -0.5)
    fc_std = (2 * D) ** -0.5
    for i in range(L):
        p = f"visual.transformer.resblocks.{i}."
        for name in ("ln_1", "ln_2"):
            sd[p + f"{name}.weight"] = 1.0 + rn(D, std=0.1)
            sd[p + f"{name}.bias"] = rn(D, std=0.1)
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


def config_from_state_dict(sd: Dict[str, torch.Tensor]) -> VitConfig:
    w = sd["visual.conv1.weight"]
    D, P = w.shape[0], w.shape[-1]
    T = sd["visual.positional_embedding"].shape[0]
    G = int(round(math.sqrt(T - 1)))
    L = len({k.split(".")[3] for k in sd if k.startswith("visual.transformer.resblocks.")})
    return VitConfig(width=D, layers=L, heads=D // 64, patch=P, image=G * P, out_dim=sd["visual.proj"].shape[1])


def preprocess(images: torch.Tensor, div255: bool) -> torch.Tensor:
    """reference shared/clip_ppo_utils.py:146-159 (div255=True) and :201-208 (div255=False):
    bilinear antialias resize to 224x224, then (u - mean) / std."""
    x = images.float()
    if div255:
        x = x / 255.0
    x = F.interpolate(x, size=(224, 224), mode="bilinear", align_corners=False, antialias=True)
    mean = torch.tensor(CLIP_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(CLIP_STD).view(1, 3, 1, 1)
    return (x - mean) / std


def _ln(x, w, b):
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def vision_tower(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """[clip] VisionTransformer.forward on a normalised [N,3,224,224] fp32 batch -> [N,out]."""
    cfg = config_from_state_dict(sd)
    D, H, P, G = cfg.width, cfg.heads, cfg.patch, cfg.grid
    dh = D // H
    N = x.shape[0]
    # conv1 (stride = kernel = P, no bias) as im2col + matmul
    a = x.reshape(N, 3, G, P, G, P).permute(0, 2, 4, 1, 3, 5).reshape(N, G * G, 3 * P * P)
    tok = a @ sd["visual.conv1.weight"].reshape(D, -1).t()
    cls = sd["visual.class_embedding"].expand(N, 1, D)
    X = torch.cat([cls, tok], dim=1) + sd["visual.positional_embedding"]
    X = _ln(X, sd["visual.ln_pre.weight"], sd["visual.ln_pre.bias"])
    T = X.shape[1]
    for i in range(cfg.layers):
        p = f"visual.transformer.resblocks.{i}."
        Y = _ln(X, sd[p + "ln_1.weight"], sd[p + "ln_1.bias"])
        qkv = Y @ sd[p + "attn.in_proj_weight"].t() + sd[p + "attn.in_proj_bias"]
        q, k, v = qkv.split(D, dim=-1)
        q = q.reshape(N, T, H, dh).transpose(1, 2)
        k = k.reshape(N, T, H, dh).transpose(1, 2)
        v = v.reshape(N, T, H, dh).transpose(1, 2)
        s = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(dh), dim=-1)
        o = (s @ v).transpose(1, 2).reshape(N, T, D)
        X = X + o @ sd[p + "attn.out_proj.weight"].t() + sd[p + "attn.out_proj.bias"]
        Y = _ln(X, sd[p + "ln_2.weight"], sd[p + "ln_2.bias"])
        h = Y @ sd[p + "mlp.c_fc.weight"].t() + sd[p + "mlp.c_fc.bias"]
        h = h * torch.sigmoid(1.702 * h)
        X = X + h @ sd[p + "mlp.c_proj.weight"].t() + sd[p + "mlp.c_proj.bias"]
    e = _ln(X[:, 0, :], sd["visual.ln_post.weight"], sd["visual.ln_post.bias"])
    return e @ sd["visual.proj"]


def image_embeddings(sd: Dict[str, torch.Tensor], images: torch.Tensor) -> torch.Tensor:
    """reference generate_clip_embeddings(modality="image"), shared/clip_ppo_utils.py:141-164."""
    e = vision_tower(sd, preprocess(images, True)).float()
    return F.normalize(e, dim=-1)


def frozen_features(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """reference get_frozen_clip_features, shared/clip_ppo_utils.py:185-217 (no /255, no L2
    norm; the reference runs the tower in fp16 there, the oracle stays fp32)."""
    return vision_tower(sd, preprocess(x, False)).float()
