"""Oracle (CPU, fp32 torch) for the frozen CLIP TEXT tower.  TEST INFRASTRUCTURE ONLY.

The arithmetic lives in openai/CLIP (``clip/model.py``: ``CLIP.encode_text``, ``Transformer``,
``ResidualAttentionBlock``, ``build_attention_mask``; reference ``requirements.txt:12``, un-vendored,
unpinned HEAD, absent from /root/reference and from this image).  This restates its published
algorithm over a state dict in openai key layout, as called by the reference at
``shared/clip_ppo_utils.py:132-139`` (``clip.tokenize`` -> ``encode_text`` -> ``.float()`` ->
``F.normalize``):

    x = token_embedding[tokens] + positional_embedding            [N, 77, W]
    L x [ x += out_proj(causal_attention(ln_1(x)));  x += c_proj(QuickGELU(c_fc(ln_2(x)))) ]
    x = ln_final(x);  e = x[n, argmax_t tokens[n, t]] @ text_projection        (the EOT token has the largest id)

The BPE tokenizer itself (``clip/simple_tokenizer.py`` + its 1.3 MB merges file) is not reproducible
offline: token ids are an INPUT here.

Pinned by: ``tests/test_oracle.py::test_text_oracle_matches_hf`` (transformers
CLIPTextModelWithProjection with the same weights) and tests/golden/text_b32_seed0.npz.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class TextConfig:
    width: int = 512
    layers: int = 12
    heads: int = 8
    context: int = 77
    vocab: int = 49408
    out_dim: int = 512


TEXT_B32 = TextConfig()
SOT, EOT = 49406, 49407


def random_state_dict(cfg: TextConfig = TEXT_B32, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded random weights in openai/CLIP key layout (``token_embedding.weight``, ``positional_embedding``,
    ``transformer.resblocks.{i}.*``, ``ln_final.*``, ``text_projection``); upstream's initialisation orders of
    magnitude, LN gamma ~ 1 +- 0.1 and beta ~ +-0.1 so the affine paths are exercised."""
    g = torch.Generator().manual_seed(seed + 1000)
    D, L, O = cfg.width, cfg.layers, cfg.out_dim

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    sd: Dict[str, torch.Tensor] = {}
    sd["token_embedding.weight"] = rn(cfg.vocab, D, std=0.02)
    sd["positional_embedding"] = rn(cfg.context, D, std=0.01)
    attn_std, proj_std, fc_std = D ** -0.5, (D ** -0.5) * ((2 * L) ** -0.5), (2 * D) ** -0.5
    for i in range(L):
        p = f"transformer.resblocks.{i}."
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
    sd["ln_final.weight"] = 1.0 + rn(D, std=0.1)
    sd["ln_final.bias"] = rn(D, std=0.1)
    sd["text_projection"] = rn(D, O, std=D ** -0.5)
    return sd


def config_from_state_dict(sd: Dict[str, torch.Tensor]) -> TextConfig:
    V, D = sd["token_embedding.weight"].shape
    L = len({k.split(".")[2] for k in sd if k.startswith("transformer.resblocks.")})
    return TextConfig(width=D, layers=L, heads=D // 64, context=sd["positional_embedding"].shape[0], vocab=V,
                      out_dim=sd["text_projection"].shape[1])


def random_tokens(n: int, cfg: TextConfig = TEXT_B32, seed: int = 0) -> torch.Tensor:
    """[n, context] int64 shaped like ``clip.tokenize`` output: SOT, 1 .. context-2 word ids, EOT, zero padding."""
    g = torch.Generator().manual_seed(seed)
    tok = torch.zeros(n, cfg.context, dtype=torch.int64)
    hi = min(cfg.vocab - 2, SOT) if cfg.vocab > 3 else 1
    for i in range(n):
        words = int(torch.randint(1, cfg.context - 1, (1,), generator=g))
        tok[i, 0] = cfg.vocab - 2
        tok[i, 1:1 + words] = torch.randint(1, max(2, hi), (words,), generator=g)
        tok[i, 1 + words] = cfg.vocab - 1
    return tok


def _ln(x, w, b):
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def text_tower(sd: Dict[str, torch.Tensor], tokens: torch.Tensor) -> torch.Tensor:
    """[clip] CLIP.encode_text on [N, context] token ids -> [N, out_dim] (not normalised)."""
    cfg = config_from_state_dict(sd)
    D, H = cfg.width, cfg.heads
    dh = D // H
    N, T = tokens.shape
    X = sd["token_embedding.weight"][tokens] + sd["positional_embedding"][:T]
    mask = torch.full((T, T), float("-inf")).triu_(1)                 # [clip] build_attention_mask
    for i in range(cfg.layers):
        p = f"transformer.resblocks.{i}."
        Y = _ln(X, sd[p + "ln_1.weight"], sd[p + "ln_1.bias"])
        qkv = Y @ sd[p + "attn.in_proj_weight"].t() + sd[p + "attn.in_proj_bias"]
        q, k, v = (t.reshape(N, T, H, dh).transpose(1, 2) for t in qkv.split(D, dim=-1))
        s = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(dh) + mask, dim=-1)
        o = (s @ v).transpose(1, 2).reshape(N, T, D)
        X = X + o @ sd[p + "attn.out_proj.weight"].t() + sd[p + "attn.out_proj.bias"]
        Y = _ln(X, sd[p + "ln_2.weight"], sd[p + "ln_2.bias"])
        h = Y @ sd[p + "mlp.c_fc.weight"].t() + sd[p + "mlp.c_fc.bias"]
        h = h * torch.sigmoid(1.702 * h)
        X = X + h @ sd[p + "mlp.c_proj.weight"].t() + sd[p + "mlp.c_proj.bias"]
    X = _ln(X, sd["ln_final.weight"], sd["ln_final.bias"])
    return X[torch.arange(N), tokens.argmax(dim=-1)] @ sd["text_projection"]


def text_embeddings(sd: Dict[str, torch.Tensor], tokens: torch.Tensor) -> torch.Tensor:
    """reference generate_clip_embeddings(modality="text") after tokenisation, shared/clip_ppo_utils.py:136-139."""
    return F.normalize(text_tower(sd, tokens).float(), dim=-1)
