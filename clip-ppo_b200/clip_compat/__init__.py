"""A stand-in for the un-vendored openai ``clip`` package (reference requirements.txt:12), used
when the real one is not importable.  It exposes exactly the surface the reference touches:
``clip.load`` (shared/clip_ppo_utils.py:90), ``clip.tokenize`` (:136), ``clip.model.CLIP`` and
``clip.model.VisionTransformer`` (:187, :212).  The image tower it returns runs on the sm_100a
kernels and so does the text tower (``encode_text`` on token ids).  ``tokenize`` is the byte-level BPE of
clip_compat/tokenizer.py; its 1.3 MB merge list is DATA of the openai package that cannot be reproduced offline, so
``CLIPPPO_BPE_PATH`` has to name it (``tokenize`` raises FileNotFoundError with that remedy otherwise).

Weight sources, in order: an explicit state dict (``load(..., state_dict=...)``), a checkpoint file
(``CLIPPPO_CLIP_WEIGHTS``: a ``torch.save``d openai state dict, module or TorchScript archive).  With neither,
``load`` RAISES - the reference's ``clip.load`` returns pretrained weights, and silently handing back an
untrained tower would turn every run into the RANDOM_ENCODER ablation.  Seeded random weights of the named
architecture (all the build and GPU boxes can have: no CLIP checkpoint, no network) are an explicit opt-in:
``load(..., random_init=True)`` or ``CLIPPPO_ALLOW_RANDOM_WEIGHTS=1`` (tests, bench.py and smoke() set it).
The source is recorded on the model as ``model.weight_source``.
"""
from __future__ import annotations

import os
import warnings
from typing import Dict, Optional, Tuple

import torch

from . import model
from .model import CLIP, VisionTransformer, random_visual_state_dict, random_text_state_dict, ARCHS

__all__ = ["load", "tokenize", "tokenizer_available", "available_models", "model", "CLIP", "VisionTransformer"]


def available_models():
    return list(ARCHS)


def load(name: str = "ViT-B/32", device: str | torch.device = "cuda", jit: bool = False,
         download_root: Optional[str] = None, *, state_dict: Optional[Dict[str, torch.Tensor]] = None,
         seed: int = 0, random_init: Optional[bool] = None) -> Tuple[CLIP, None]:
    """Mirror of ``clip.load``: returns (model, preprocess).  ``preprocess`` (a PIL transform
    upstream) is unused by the reference and is None here."""
    source = "state_dict argument"
    if state_dict is None:
        path = os.environ.get("CLIPPPO_CLIP_WEIGHTS")
        if path:
            try:
                obj = torch.load(path, map_location="cpu", weights_only=True)
            except Exception:                       # a pickled module or a TorchScript archive (what openai ships)
                try:
                    obj = torch.jit.load(path, map_location="cpu")
                except Exception:
                    obj = torch.load(path, map_location="cpu", weights_only=False)
            state_dict = obj.state_dict() if hasattr(obj, "state_dict") else obj
            source = f"checkpoint {path}"
    if state_dict is None:
        if name not in ARCHS:
            raise RuntimeError(f"Model {name} not found; available models = {available_models()}")
        if random_init is None:
            random_init = os.environ.get("CLIPPPO_ALLOW_RANDOM_WEIGHTS", "") not in ("", "0")
        if not random_init:
            raise RuntimeError(
                "clip_compat.load: no CLIP weights available - the openai `clip` package is not installed and "
                "CLIPPPO_CLIP_WEIGHTS does not name a checkpoint.  Install openai/CLIP, point CLIPPPO_CLIP_WEIGHTS at a "
                "saved state dict, or opt in to an UNTRAINED tower with load(..., random_init=True) / "
                "CLIPPPO_ALLOW_RANDOM_WEIGHTS=1 (throughput measurements and parity tests only).")
        warnings.warn(f"clip_compat.load({name!r}): seeded RANDOM weights (seed {seed}) - not a pretrained CLIP", stacklevel=2)
        state_dict = random_visual_state_dict(name, seed)
        state_dict.update(random_text_state_dict(name, seed))
        source = f"random(seed={seed})"
    clip_model = CLIP(state_dict, device=device)
    clip_model.weight_source = source
    return clip_model, None


def tokenize(texts, context_length: int = 77, truncate: bool = False):
    """``clip.tokenize`` on the byte-level BPE of clip_compat/tokenizer.py.  The merge list is data of the openai package:
    ``CLIPPPO_BPE_PATH`` must name it (FileNotFoundError with the remedy otherwise)."""
    from .tokenizer import tokenize as _tokenize
    return _tokenize(texts, context_length, truncate)


def tokenizer_available() -> bool:
    p = os.environ.get("CLIPPPO_BPE_PATH", "")
    return bool(p) and os.path.exists(p)
