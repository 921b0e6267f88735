"""The PPO encoder of the reference's ``Agent`` on the sm_100a kernels of csrc/policy.cu.

``NatureCNN`` is a drop-in for the ``nn.Sequential`` the scripts build (reference
minigrid_experiments/clip_ppo/clip_ppo_minigrid.py:229-242, atari_experiments/clip_ppo/clip_ppo_atari.py:196-209):

    nn.Sequential(Conv2d(C, 32, 8, 4), ReLU, Conv2d(32, 64, 4, 2), ReLU, Conv2d(64, 64, 3, 1), ReLU, Flatten, Linear(3136, 512), ReLU)

with the SAME parameter names (``0.weight`` ... ``7.bias``), shapes and initialisation interface, so ``state_dict()``,
checkpoints, ``optim.Adam(agent.parameters())`` and ``clip_grad_norm_`` keep working: ``agent.network = NatureCNN.from_sequential(agent.network)``.
Forward and backward are one C-ABI call each (fp32 FMA arithmetic throughout, deterministic); autograd sees a single
``torch.autograd.Function``.  No gradient flows to the observations (the reference never asks for one).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.nn as nn

from . import _native as N


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


class _NatureFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, obs, in_scale, w1, b1, w2, b2, w3, b3, wfc, bfc):
        if not obs.is_cuda:
            raise RuntimeError("NatureCNN runs on the CUDA kernels only (no CPU fallback)")
        if obs.dim() != 4 or obs.shape[2] != 84 or obs.shape[3] != 84:
            raise ValueError(f"expected [mb, C, 84, 84] observations, got {tuple(obs.shape)}")
        if obs.dtype != torch.float32:
            obs = obs.float()
        mb, ch = obs.shape[0], obs.shape[1]
        if w1.shape != (32, ch, 8, 8):
            raise ValueError(f"first convolution expects {w1.shape[1]} channels, observations have {ch}")
        params = [p.contiguous() for p in (w1, b1, w2, b2, w3, b3, wfc, bfc)]
        need = C.c_size_t()
        L = N.lib()
        N.check(L.clipppo_nature_workspace_bytes(mb, ch, C.byref(need)), "clipppo_nature_workspace_bytes")
        ws = torch.empty(need.value, dtype=torch.uint8, device=obs.device)
        hidden = torch.empty((mb, 512), dtype=torch.float32, device=obs.device)
        def run(o):
            with N.device_ctx(o.device):
                return L.clipppo_nature_forward(o.data_ptr(), N.strides4(o), float(in_scale), mb, ch, *[p.data_ptr() for p in params],
                                                hidden.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream(o.device).cuda_stream)
        st = run(obs)
        if st in (N.ERR_UNSUPPORTED, N.ERR_ALIGN):          # exotic strides: one NCHW copy, then the same kernels
            obs = obs.contiguous()
            st = run(obs)
        N.check(st, "clipppo_nature_forward")
        ctx.save_for_backward(hidden, ws, obs, params[2], params[4])
        ctx.shape = (mb, ch)
        ctx.in_scale = float(in_scale)
        ctx.param_shapes = [p.shape for p in params]
        return hidden

    @staticmethod
    def backward(ctx, grad_hidden):
        hidden, ws, obs, w2, w3 = ctx.saved_tensors
        mb, ch = ctx.shape
        grad_hidden = grad_hidden.contiguous().float()
        grads = [torch.empty(s, dtype=torch.float32, device=hidden.device) for s in ctx.param_shapes]
        with N.device_ctx(hidden.device):
            st = N.lib().clipppo_nature_backward(grad_hidden.data_ptr(), hidden.data_ptr(), obs.data_ptr(), N.strides4(obs), ctx.in_scale,
                                                 mb, ch, w2.data_ptr(), w3.data_ptr(), *[g.data_ptr() for g in grads],
                                                 ws.data_ptr(), ws.numel(), torch.cuda.current_stream(hidden.device).cuda_stream)
        N.check(st, "clipppo_nature_backward")
        return (None, None, *grads)


class NatureCNN(nn.Module):
    """``network(x)`` of the reference's Agent: x [mb, C, 84, 84] fp32 (already divided by 255, like ``self.network(x / 255.0)`` /
    ``self.network(self._pre(x))``) -> post-ReLU features [mb, 512].  ``forward(x, in_scale=1/255)`` folds that division
    into the first kernel; ``x`` may be any strided view (e.g. ``obs.permute(0, 3, 1, 2)`` of NHWC MiniGrid frames)."""

    _KEYS = (("0", (32, None, 8, 8)), ("2", (64, 32, 4, 4)), ("4", (64, 64, 3, 3)), ("7", (512, 3136)))

    def __init__(self, in_channels: int = 3):
        super().__init__()
        ref = nn.Sequential(nn.Conv2d(in_channels, 32, 8, stride=4), nn.ReLU(), nn.Conv2d(32, 64, 4, stride=2), nn.ReLU(),
                            nn.Conv2d(64, 64, 3, stride=1), nn.ReLU(), nn.Flatten(), nn.Linear(64 * 7 * 7, 512), nn.ReLU())
        self._adopt(ref)

    def _adopt(self, seq: nn.Sequential) -> None:
        # parameters registered under the Sequential's own names ("0.weight", ...) via per-index holder modules
        for idx in ("0", "2", "4", "7"):
            holder = nn.Module()
            holder.weight = nn.Parameter(seq[int(idx)].weight.detach().clone())
            holder.bias = nn.Parameter(seq[int(idx)].bias.detach().clone())
            self.add_module(idx, holder)
        self.in_channels = seq[0].weight.shape[1]

    @classmethod
    def from_sequential(cls, seq: nn.Sequential) -> "NatureCNN":
        """Adopt the weights of the scripts' ``nn.Sequential`` (same device, same values, same state-dict keys)."""
        m = cls.__new__(cls)
        nn.Module.__init__(m)
        m._adopt(seq)
        return m.to(seq[0].weight.device)

    def forward(self, x: torch.Tensor, in_scale: float = 1.0) -> torch.Tensor:
        g = lambda i: getattr(self, i)
        return _NatureFn.apply(x, in_scale, g("0").weight, g("0").bias, g("2").weight, g("2").bias, g("4").weight, g("4").bias,
                               g("7").weight, g("7").bias)
