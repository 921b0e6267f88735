"""Data parallelism over observations: one process per GPU, rollout rows / frames
sharded by environment, the frozen tower replicated, and exactly ONE exchange per optimizer step -
an all-reduce of the trainable (PPO encoder / actor / critic / temporal_projection) gradients,
inserted between ``loss.backward()`` and ``clip_grad_norm_`` (reference
clip_ppo_minigrid.py:562-563).  No activation or embedding ever crosses a GPU boundary.

``torch.distributed`` is the plumbing (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of n units (envs / frames) owned by `rank`; sizes differ by <= 1."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class GradBucket:
    """Flat fp32 bucket over the trainable parameters' gradients.  ``all_reduce_mean()`` gathers the
    grads with one multi-tensor copy, all-reduces once (sum), divides by the world size and scatters them back with a
    second multi-tensor copy - four launches and one
    collective of ~6.7 MB (MiniGrid agent) / ~11 MB (Atari image agent) per optimizer step."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group: Optional[dist.ProcessGroup] = None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.group = group
        # one view of the bucket per parameter, in parameter order: gather / scatter are single multi-tensor copies
        self.views: List[torch.Tensor] = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view(p.shape))
            off += p.numel()

    def all_reduce_mean(self) -> None:
        if not (dist.is_available() and dist.is_initialized()):
            return
        world = dist.get_world_size(self.group)
        if world == 1:
            return
        for p, v in zip(self.params, self.views):           # parameters the backward never reached
            if p.grad is None:
                p.grad = torch.zeros_like(v)
        grads = [p.grad for p in self.params]
        torch._foreach_copy_(self.views, grads)             # gather: one launch
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.div_(world)
        torch._foreach_copy_(grads, self.views)             # scatter back: one launch


def global_advantage_stats(adv: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(mean, unbiased std) of the advantages over ALL ranks' minibatch rows with one 3-float
    all-reduce - what ``mb_advantages.mean()`` / ``.std()`` (reference :509) see on one GPU."""
    s = torch.stack([adv.sum(dtype=torch.float64), (adv.double() ** 2).sum(), torch.tensor(float(adv.numel()), device=adv.device, dtype=torch.float64)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(s, group=group)
    n = s[2]
    mean = s[0] / n
    var = (s[1] - n * mean * mean) / (n - 1)
    return mean.float(), var.clamp_min(0).sqrt().float()
