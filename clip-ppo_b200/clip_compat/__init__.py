"""A stand-in for the un-vendored openai ``clip`` package (reference requirements.txt:12), used
when the real one is not importable.  It exposes exactly the surface the reference touches:
``clip.load`` (shared/clip_ppo_utils.py:90), ``clip.tokenize`` (:136), ``clip.model.CLIP`` and
``clip.model.VisionTransformer`` (:187, :212).  The image tower it returns runs on the sm_100a
kernels and so does the text tower (``encode_text`` on token ids); only the BPE tokenizer is missing -
its 1.3 MB merges file ships with the openai package and cannot be reproduced offline - so ``tokenize`` raises.

Weight sources, in order: an explicit state dict / checkpoint file (``CLIPPPO_CLIP_WEIGHTS`` env
var or ``load(..., state_dict=...)``), else seeded random weights of the named architecture
(there are no CLIP weights and no network on the build or GPU boxes).
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import torch

from . import model
from .model import CLIP, VisionTransformer, random_visual_state_dict, random_text_state_dict, ARCHS

__all__ = ["load", "tokenize", "available_models", "model", "CLIP", "VisionTransformer"]


def available_models():
    return list(ARCHS)


def load(name: str = "ViT-B/32", device: str | torch.device = "cuda", jit: bool = False,
         download_root: Optional[str] = None, *, state_dict: Optional[Dict[str, torch.Tensor]] = None,
         seed: int = 0) -> Tuple[CLIP, None]:
    """Mirror of ``clip.load``: returns (model, preprocess).  ``preprocess`` (a PIL transform
    upstream) is unused by the reference and is None here."""
    if state_dict is None:
        path = os.environ.get("CLIPPPO_CLIP_WEIGHTS")
        if path:
            obj = torch.load(path, map_location="cpu")
            state_dict = obj.state_dict() if hasattr(obj, "state_dict") else obj
    if state_dict is None:
        if name not in ARCHS:
            raise RuntimeError(f"Model {name} not found; available models = {available_models()}")
        state_dict = random_visual_state_dict(name, seed)
        state_dict.update(random_text_state_dict(name, seed))
    return CLIP(state_dict, device=device), None


def tokenize(texts, context_length: int = 77, truncate: bool = False):
    raise NotImplementedError("clip_compat has no BPE tokenizer (the merges file ships with openai/CLIP): install it for "
                              "string descriptions, or pass pre-tokenised [N, 77] ids to generate_clip_embeddings / encode_text")
