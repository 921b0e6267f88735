"""Host side of L1 (cosine alignment loss), G1 (GAE) and P1 (PPO minibatch loss)."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import _native as N


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected CUDA tensors - the B200 path has no CPU fallback")


class _CosineLoss(torch.autograd.Function):
    """mean(1 - cos(z, c)); differentiable wrt both arguments like the reference expression
    (shared/clip_ppo_utils.py:66-74)."""

    @staticmethod
    def forward(ctx, z: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
        zc, cc = _f32c(z), _f32c(c)
        rows, dim = zc.shape
        loss = torch.empty((), dtype=torch.float32, device=zc.device)
        stats = torch.empty((rows, 3), dtype=torch.float32, device=zc.device)
        with N.device_ctx(zc.device):
            N.check(N.lib().clipppo_cosine_loss_fwd(zc.data_ptr(), cc.data_ptr(), rows, dim, loss.data_ptr(),
                                                    stats.data_ptr(), _stream(zc)), "clipppo_cosine_loss_fwd")
        ctx.save_for_backward(zc, cc, stats)
        ctx.in_dtypes = (z.dtype, c.dtype)
        return loss

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        zc, cc, stats = ctx.saved_tensors
        need_z, need_c = ctx.needs_input_grad
        rows, dim = zc.shape
        gz = torch.empty_like(zc) if need_z else None
        gc = torch.empty_like(cc) if need_c else None
        g = grad_out.detach().to(torch.float32).contiguous()
        with N.device_ctx(zc.device):
            N.check(N.lib().clipppo_cosine_loss_bwd(zc.data_ptr(), cc.data_ptr(), stats.data_ptr(), g.data_ptr(),
                                                    rows, dim, gz.data_ptr() if need_z else None,
                                                    gc.data_ptr() if need_c else None, _stream(zc)),
                    "clipppo_cosine_loss_bwd")
        zt, ct = ctx.in_dtypes
        return (gz.to(zt) if need_z else None), (gc.to(ct) if need_c else None)


def cosine_embedding_loss(z: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
    if z.shape[-1] != c.shape[-1]:
        raise ValueError(f"Dimension mismatch: PPO latents ({z.shape[-1]}) vs CLIP embeddings ({c.shape[-1]}). "
                         f"Both should be 512-dim for ViT-B/32. Check agent architecture.")
    _require_cuda(z, "cosine_embedding_loss")
    z2, c2 = z.reshape(-1, z.shape[-1]), c.reshape(-1, c.shape[-1])
    if z2.shape[0] != c2.shape[0]:
        raise ValueError(f"row count mismatch: {z2.shape[0]} vs {c2.shape[0]}")
    return _CosineLoss.apply(z2, c2)


def gae(rewards: torch.Tensor, values: torch.Tensor, dones: torch.Tensor, next_value: torch.Tensor,
        next_done: torch.Tensor, gamma: float = 0.99, gae_lambda: float = 0.95) -> Tuple[torch.Tensor, torch.Tensor]:
    """(advantages, returns), both [T,E]; one launch instead of the reference's T-step Python
    loop (clip_ppo_minigrid.py:437-450)."""
    _require_cuda(rewards, "gae")
    r, v, d = _f32c(rewards), _f32c(values), _f32c(dones)
    T, E = r.shape
    nv, nd = _f32c(next_value).reshape(-1), _f32c(next_done).reshape(-1)
    if nv.numel() != E or nd.numel() != E or v.shape != r.shape or d.shape != r.shape:
        raise ValueError("gae: shape mismatch")
    adv, ret = torch.empty_like(r), torch.empty_like(r)
    with N.device_ctx(r.device):
        N.check(N.lib().clipppo_gae_f32(r.data_ptr(), v.data_ptr(), d.data_ptr(), nv.data_ptr(), nd.data_ptr(), T, E,
                                        float(gamma), float(gae_lambda), adv.data_ptr(), ret.data_ptr(), _stream(r)),
                "clipppo_gae_f32")
    return adv, ret


_STAT_NAMES = ("loss", "pg_loss", "v_loss", "entropy", "old_approx_kl", "approx_kl", "clipfrac", "adv_std")


class _PpoLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, newlogprob, entropy, newvalue, old_logprob, advantages, returns, old_values, clip_loss,
                clip_coef, ent_coef, vf_coef, clip_lambda, norm_adv, clip_vloss):
        nlp, ent, nv = _f32c(newlogprob).reshape(-1), _f32c(entropy).reshape(-1), _f32c(newvalue).reshape(-1)
        olp, adv, ret = _f32c(old_logprob).reshape(-1), _f32c(advantages).reshape(-1), _f32c(returns).reshape(-1)
        ov = _f32c(old_values).reshape(-1)
        n = nlp.numel()
        stats = torch.empty(8, dtype=torch.float32, device=nlp.device)
        need = any(ctx.needs_input_grad[:3]) or (clip_loss is not None and ctx.needs_input_grad[7])
        g = [torch.empty_like(nlp) for _ in range(3)] if need else [None] * 3
        cl = _f32c(clip_loss).reshape(()) if clip_loss is not None else None
        with N.device_ctx(nlp.device):
            N.check(N.lib().clipppo_ppo_loss_f32(
                nlp.data_ptr(), ent.data_ptr(), nv.data_ptr(), olp.data_ptr(), adv.data_ptr(), ret.data_ptr(),
                ov.data_ptr(), cl.data_ptr() if cl is not None else None, n, float(clip_coef), float(ent_coef),
                float(vf_coef), float(clip_lambda), int(bool(norm_adv)), int(bool(clip_vloss)), stats.data_ptr(),
                *(t.data_ptr() if t is not None else None for t in g), _stream(nlp)), "clipppo_ppo_loss_f32")
        if need:
            ctx.save_for_backward(*g)
        ctx.shapes = (newlogprob.shape, entropy.shape, newvalue.shape)
        ctx.clip_lambda = float(clip_lambda)
        ctx.mark_non_differentiable(stats)
        return stats[0].clone(), stats

    @staticmethod
    def backward(ctx, grad_loss, _grad_stats):
        g_nlp, g_ent, g_nv = ctx.saved_tensors
        s0, s1, s2 = ctx.shapes
        gl = grad_loss
        out = [None] * 14
        if ctx.needs_input_grad[0]:
            out[0] = (g_nlp * gl).reshape(s0)
        if ctx.needs_input_grad[1]:
            out[1] = (g_ent * gl).reshape(s1)
        if ctx.needs_input_grad[2]:
            out[2] = (g_nv * gl).reshape(s2)
        if ctx.needs_input_grad[7]:
            out[7] = gl * ctx.clip_lambda
        return tuple(out)


def ppo_loss(newlogprob: torch.Tensor, entropy: torch.Tensor, newvalue: torch.Tensor, old_logprob: torch.Tensor,
             advantages: torch.Tensor, returns: torch.Tensor, old_values: torch.Tensor,
             clip_loss: Optional[torch.Tensor] = None, clip_lambda: float = 0.0, clip_coef: float = 0.1,
             ent_coef: float = 0.01, vf_coef: float = 0.5, norm_adv: bool = True,
             clip_vloss: bool = True, adv_stats: Optional[tuple] = None) -> Dict[str, torch.Tensor]:
    """The reference's minibatch loss (clip_ppo_minigrid.py:498-531,559) as one launch.  Returns
    a dict of 0-d tensors: the differentiable total `loss` (gradients flow to newlogprob, entropy,
    newvalue and clip_loss) plus the detached diagnostics - no `.item()` sync is forced.

    ``adv_stats=(mean, std)`` (additive; 0-d tensors or floats): normalise the advantages with THESE statistics instead of
    the minibatch's own (reference :509).  Data-parallel runs pass ``distributed.global_advantage_stats(adv)`` so that every
    rank's shard is normalised exactly as the single-GPU run normalises the whole minibatch."""
    _require_cuda(newlogprob, "ppo_loss")
    if adv_stats is not None and norm_adv:
        mean, std = adv_stats
        advantages = (advantages - mean) / (std + 1e-8)
        norm_adv = False
    if clip_loss is not None and not torch.is_tensor(clip_loss):
        clip_loss = torch.tensor(float(clip_loss), dtype=torch.float32, device=newlogprob.device)
    loss, stats = _PpoLoss.apply(newlogprob, entropy, newvalue, old_logprob, advantages, returns, old_values,
                                 clip_loss, clip_coef, ent_coef, vf_coef, clip_lambda, norm_adv, clip_vloss)
    out = {name: stats[i] for i, name in enumerate(_STAT_NAMES)}
    out["loss"] = loss
    return out
