"""Oracle (CPU, fp32 torch) for the alignment loss, GAE and the PPO minibatch loss.
TEST INFRASTRUCTURE ONLY.

Each function restates reference code that is importable or quotable line by line; autograd
of these torch expressions is the gradient oracle.

Pinned by: tests/golden/losses.npz (cosine loss and lambda warm-up produced by the
reference's own ``shared/clip_ppo_utils.py`` functions; GAE / PPO terms produced by executing
the reference script's statements on fixed inputs, see oracle/make_goldens.py).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F


def cosine_embedding_loss(z: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
    """reference shared/clip_ppo_utils.py:48-76."""
    if z.shape[-1] != c.shape[-1]:
        raise ValueError("Dimension mismatch")
    zn = F.normalize(z, dim=-1)
    cn = F.normalize(c, dim=-1)
    return torch.mean(1 - torch.sum(zn * cn, dim=-1))


def clip_lambda_with_warmup(target: float, it: int, total: int, frac: float = 0.2) -> float:
    """reference shared/clip_ppo_utils.py:26-46."""
    warm = int(total * frac)
    return target * (it / warm) if it < warm else target


def gae(rewards: torch.Tensor, values: torch.Tensor, dones: torch.Tensor,
        next_value: torch.Tensor, next_done: torch.Tensor,
        gamma: float = 0.99, lam: float = 0.95):
    """reference minigrid_experiments/clip_ppo/clip_ppo_minigrid.py:437-450
    (= atari_experiments/clip_ppo/clip_ppo_atari.py:619-632).  [T,E] inputs."""
    T = rewards.shape[0]
    adv = torch.zeros_like(rewards)
    last = torch.zeros_like(rewards[0])
    for t in reversed(range(T)):
        if t == T - 1:
            nnt = 1.0 - next_done
            nv = next_value.reshape(-1)
        else:
            nnt = 1.0 - dones[t + 1]
            nv = values[t + 1]
        delta = rewards[t] + gamma * nv * nnt - values[t]
        adv[t] = last = delta + gamma * lam * nnt * last
    return adv, adv + values


def ppo_loss(newlogprob: torch.Tensor, entropy: torch.Tensor, newvalue: torch.Tensor,
             old_logprob: torch.Tensor, advantages: torch.Tensor, returns: torch.Tensor,
             old_values: torch.Tensor, clip_loss: torch.Tensor | float = 0.0,
             clip_lambda: float = 0.0, clip_coef: float = 0.1, ent_coef: float = 0.01,
             vf_coef: float = 0.5, norm_adv: bool = True, clip_vloss: bool = True) -> Dict[str, torch.Tensor]:
    """reference clip_ppo_minigrid.py:498-531,559 (same code at clip_ppo_atari.py:691-720,747)."""
    logratio = newlogprob - old_logprob
    ratio = logratio.exp()
    with torch.no_grad():
        old_approx_kl = (-logratio).mean()
        approx_kl = ((ratio - 1) - logratio).mean()
        clipfrac = ((ratio - 1.0).abs() > clip_coef).float().mean()
    adv = advantages
    if norm_adv:
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
    pg1 = -adv * ratio
    pg2 = -adv * torch.clamp(ratio, 1 - clip_coef, 1 + clip_coef)
    pg_loss = torch.max(pg1, pg2).mean()
    newvalue = newvalue.view(-1)
    if clip_vloss:
        vu = (newvalue - returns) ** 2
        vc = old_values + torch.clamp(newvalue - old_values, -clip_coef, clip_coef)
        v_loss = 0.5 * torch.max(vu, (vc - returns) ** 2).mean()
    else:
        v_loss = 0.5 * ((newvalue - returns) ** 2).mean()
    ent = entropy.mean()
    loss = pg_loss - ent_coef * ent + v_loss * vf_coef + clip_lambda * clip_loss
    return dict(loss=loss, pg_loss=pg_loss, v_loss=v_loss, entropy=ent,
                old_approx_kl=old_approx_kl, approx_kl=approx_kl, clipfrac=clipfrac)
