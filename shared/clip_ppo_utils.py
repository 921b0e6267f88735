"""Drop-in for reference ``shared/clip_ppo_utils.py``: same names and signatures, with the image
branch, the frozen-feature path and the alignment loss running on the sm_100a kernels.

* ``generate_clip_embeddings(modality="image")`` / ``get_frozen_clip_features`` make ONE C-ABI call
  (``clipppo_vit_encode``): resize + normalise + bf16 im2col, the tcgen05 tower, projection and the
  L2 normalise - instead of eager interpolate / normalise / fp16 tower / normalize.
* ``compute_cosine_embedding_loss`` is a fused forward (+ custom backward wrt both arguments).
* The host-side scalars (lambda warm-up, gating, config) are plain Python, as in the reference.

``clip`` (openai/CLIP) is an un-vendored dependency of the reference; if it is importable it is used
for loading weights, otherwise ``clip_ppo_b200.clip_compat`` stands in under the same name.
"""
from __future__ import annotations

import sys
from dataclasses import dataclass
from enum import Enum
from typing import List, Optional

import torch

try:                                    # real openai/CLIP, if the user has it
    import clip  # type: ignore
    import clip.model  # type: ignore  # noqa: F401
except ImportError:                     # stand-in with the same surface (image tower only)
    from clip_ppo_b200 import clip_compat as clip
    sys.modules.setdefault("clip", clip)
    sys.modules.setdefault("clip.model", clip.model)

from clip_ppo_b200 import losses as _L
from clip_ppo_b200.text import TextEngine
from clip_ppo_b200.vit import VitEngine


class AblationMode(Enum):
    """Ablation study modes (reference :13-17)."""
    NONE = "NONE"
    FROZEN_CLIP = "FROZEN_CLIP"
    RANDOM_ENCODER = "RANDOM_ENCODER"


# reference :21-22; the kernels carry the same constants (csrc/preprocess.cu)
_CLIP_MEAN = torch.tensor([0.48145466, 0.4578275, 0.40821073])
_CLIP_STD = torch.tensor([0.26862954, 0.26130258, 0.27577711])

CLIP_LOSS_FREQUENCY = 4


def get_clip_lambda_with_warmup(target_lambda: float, current_iteration: int, total_iterations: int,
                                warmup_fraction: float = 0.2) -> float:
    """Linear 0 -> target over the first ``warmup_fraction`` of training (reference :26-46)."""
    ramp = int(total_iterations * warmup_fraction)
    if current_iteration >= ramp:
        return target_lambda
    return target_lambda * (current_iteration / ramp)


def compute_cosine_embedding_loss(z: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
    """mean_b(1 - cos(z_b, c_b)) (reference :48-76); raises ValueError on a width mismatch."""
    return _L.cosine_embedding_loss(z, c)


def load_clip_model(model_name: str = "ViT-B/32", device: str = "cuda") -> torch.nn.Module:
    """``clip.load`` -> eval -> freeze (reference :79-97)."""
    model, _ = clip.load(model_name, device=device)
    model.eval()
    for p in model.parameters():
        p.requires_grad = False
    return model


def _engine_for(clip_model) -> VitEngine:
    """The native tower for whatever the caller holds: our compat ``CLIP`` / ``VisionTransformer``
    (they build their engine lazily) or a real openai module, whose ``visual.*`` weights are
    repacked once and cached on the module."""
    visual = getattr(clip_model, "visual", clip_model)
    if hasattr(visual, "engine"):
        return visual.engine()
    eng = getattr(visual, "_clipppo_engine", None)
    dev = next(visual.parameters()).device
    if eng is None or eng.device != dev:
        eng = VitEngine(visual.state_dict(), device=dev, prefix="")
        object.__setattr__(visual, "_clipppo_engine", eng)
    return eng


def _text_engine_for(clip_model) -> TextEngine:
    """The native text tower: our compat ``CLIP`` builds it lazily; a real openai module's top-level text weights
    (``token_embedding.weight`` ... ``text_projection``) are repacked once and cached on the module."""
    if hasattr(clip_model, "text_engine"):
        return clip_model.text_engine()
    eng = getattr(clip_model, "_clipppo_text_engine", None)
    dev = next(clip_model.parameters()).device
    if eng is None or eng.device != dev:
        eng = TextEngine(clip_model.state_dict(), device=dev, prefix="")
        object.__setattr__(clip_model, "_clipppo_text_engine", eng)
    return eng


def generate_clip_embeddings(
    ablation_mode: AblationMode,
    clip_model: torch.nn.Module,
    modality: str,
    batch_size: int,
    device: str,
    descriptions: Optional[List[str]] = None,
    images: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """Unit-norm [N,512] embeddings (reference :100-167)."""
    if ablation_mode == AblationMode.RANDOM_ENCODER:
        e = torch.randn(batch_size, 512, device=device)
        return torch.nn.functional.normalize(e, dim=-1)
    if modality == "text":
        if descriptions is None:
            raise ValueError("descriptions required for text modality")
        eng = _text_engine_for(clip_model)
        if isinstance(descriptions, torch.Tensor):          # additive: pre-tokenised [N, 77] ids, taken as they are
            return eng.encode(descriptions.to(device), l2norm=True)
        # A rollout batch repeats a handful of state descriptions thousands of times (clip_ppo_minigrid.py:459-462):
        # tokenise and encode each distinct string once and gather.  The tower is batch-invariant bit for bit, so the
        # result equals encoding every row (tokenize -> encode_text -> float -> normalize, :136-139).
        index = {}
        inverse = [index.setdefault(d, len(index)) for d in descriptions]
        tokens = clip.tokenize(list(index))
        e = eng.encode(tokens.to(device), l2norm=True)
        if len(index) == len(inverse):
            return e
        return e[torch.tensor(inverse, device=e.device)]
    if modality == "image":
        if images is None:
            raise ValueError("images required for image modality")
        # images / 255 -> bilinear 224 -> (u - mean) / std -> tower -> float -> normalize, one call
        return _engine_for(clip_model).encode(images, pre_scale=1.0 / 255.0, l2norm=True)
    raise ValueError(f"Invalid modality: {modality}. Must be 'image' or 'text'")


def should_compute_clip_loss(ablation_mode: AblationMode, clip_lambda: float) -> bool:
    """reference :170-182."""
    return clip_lambda > 0.0 and ablation_mode != AblationMode.FROZEN_CLIP


def get_frozen_clip_features(
    x: torch.Tensor,
    clip_model: "clip.model.VisionTransformer | clip.model.CLIP",
) -> torch.Tensor:
    """Frozen-encoder features (reference :185-217): no /255, no L2 normalise, fp32 out.
    This runs inside every policy forward in FROZEN_CLIP mode (E frames per call), where a tower pass is
    bound by host launch latency: small batches replay a captured CUDA graph (same result, bitwise)."""
    eng = _engine_for(clip_model)
    if x.dim() == 4 and 0 < x.shape[0] <= eng.GRAPH_MAX_IMAGES and not torch.cuda.is_current_stream_capturing():
        return eng.encode_graphed(x, pre_scale=1.0, l2norm=False)
    return eng.encode(x, pre_scale=1.0, l2norm=False)


@dataclass
class ClipPPOConfig:
    """CLIP-PPO specific parameters shared by the training scripts (reference :220-240)."""

    clip_lambda: float = 0.00001
    """coefficient for CLIP alignment loss"""
    clip_model: str = "ViT-B/32"
    """CLIP model variant to use"""
    clip_modality: str = "text"
    """CLIP modality to use for alignment: 'image' or 'text'"""
    ablation_mode: AblationMode = AblationMode.NONE
    """ablation mode for controlled experiments"""
    apply_disturbances: bool = False
    """whether to apply visual disturbances during training"""
    disturbance_severity: str = "MODERATE"
    """disturbance severity level: MILD, MODERATE, HARD, SEVERE"""
