"""The hot-path fragments of the reference TRAINING SCRIPTS, as functions over the B200 kernels.

The scripts themselves (env stepping, logging, checkpoints) are out of scope; these
are the statements of theirs that sit on the observation path,
each with the reference line range it stands for, so the scripts can call them in place.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import _native as N
from . import disturb as D
from . import losses as L


# ---- a10: MiniGrid env-step disturbance (clip_ppo_minigrid.py:381-388) -----------------------------
def disturb_minigrid_obs(disturber, next_obs: torch.Tensor) -> torch.Tensor:
    """next_obs [E,H,W,C] (fp32 holding 0..255, or uint8) -> disturbed uint8 [E,H,W,C].
    Equals ``(disturber.apply_disturbances((next_obs.float()/255).permute(0,3,1,2)).permute(0,2,3,1)*255).byte()``
    with the same RNG consumption, in one launch (no NCHW round trip, no fp32 intermediate)."""
    if next_obs.dim() != 4:
        raise ValueError(f"expected [E,H,W,C], got {tuple(next_obs.shape)}")
    E, H, W, C = next_obs.shape
    like = next_obs if next_obs.dtype == torch.float32 and next_obs.is_contiguous() else \
        torch.empty((E, H, W, C), dtype=torch.float32, device=next_obs.device)
    noise = torch.randn_like(like.permute(0, 3, 1, 2))           # NHWC strides, like the reference's view
    c = disturber._draw_contrast()
    taps = disturber._draw_blur_taps()
    window = disturber._draw_cutout(H, W, None)
    return D.fused_disturb_nhwc_u8(next_obs, stages=N.STAGE_ALL, noise=noise, noise_sigma=disturber.gaussian_noise_sigma,
                                   contrast=c, taps=taps, window=window)


# ---- a11: Atari env-step disturbance (clip_ppo_atari.py:568-584) -----------------------------------
def disturb_atari_stack(disturber, next_obs: torch.Tensor) -> torch.Tensor:
    """next_obs [E,4,H,W] fp32 0..255 -> disturbed fp32 [E,4,H,W] (non-integer, x255), each stacked
    frame disturbed as its own [E,1,H,W] call with its own draws, in the reference's order."""
    x = next_obs.float() / 255.0
    frames = [disturber.apply_disturbances(x[:, f:f + 1]) for f in range(x.shape[1])]
    return torch.cat(frames, dim=1) * 255.0


# ---- §8f-4: rollout observation storage (clip_ppo_minigrid.py:346, 467-483) --------------------------
class ObsStoreU8:
    """MiniGrid rollout observations kept as uint8 [T,E,H,W,C] instead of the script's fp32 0..255
    (``obs = torch.zeros((num_steps, num_envs) + obs_shape)``, clip_ppo_minigrid.py:346): the values are
    integral already (the env renders uint8; the disturbed frames go through ``.byte()``, :388), so the
    store is lossless at a quarter of the HBM, and the tower's preprocess reads the bytes directly
    (``CLIPPPO_IMG_U8``, any strides) - no fp32 copy of a minibatch is ever materialised.
    Atari stacks are NOT integral after ``*255`` (clip_ppo_atari.py:584) and stay fp32."""

    def __init__(self, num_steps: int, num_envs: int, obs_shape: Tuple[int, int, int], device="cuda"):
        self.data = torch.zeros((num_steps, num_envs) + tuple(obs_shape), dtype=torch.uint8, device=device)

    def __setitem__(self, step: int, next_obs: torch.Tensor) -> None:
        """``obs[step] = next_obs``; fp32 input must hold integers in 0..255 (raises otherwise)."""
        if next_obs.dtype != torch.uint8:
            q = next_obs.to(torch.uint8)
            if not torch.equal(q.to(next_obs.dtype), next_obs):
                raise ValueError("ObsStoreU8 holds integral 0..255 observations only")
            next_obs = q
        self.data[step] = next_obs

    def flat(self) -> torch.Tensor:
        """``b_obs = obs.reshape((-1,) + obs_shape)`` (clip_ppo_minigrid.py:453), uint8 [T*E,H,W,C]."""
        return self.data.reshape((-1,) + tuple(self.data.shape[2:]))

    def clip_images(self, mb_inds: torch.Tensor) -> torch.Tensor:
        """The ``images=`` argument of generate_clip_embeddings for a minibatch: uint8 [mb,C,H,W] as an
        NHWC-strided view of the gathered rows (what ``b_obs[mb_inds].permute(0, 3, 1, 2)`` is in the script)."""
        return self.flat()[mb_inds].permute(0, 3, 1, 2)

    def policy_input(self, mb_inds: torch.Tensor) -> torch.Tensor:
        """fp32 0..255 [mb,H,W,C] for the PPO encoder (it divides by 255 itself, :262)."""
        return self.flat()[mb_inds].float()


# ---- a17: Atari frame stack -> CLIP embeddings (clip_ppo_atari.py:249-299, 661) --------------------
def convert_atari_frames_for_clip(obs_batch: torch.Tensor) -> torch.Tensor:
    """[B,4,84,84] gray -> [B,4,3,84,84] by channel repeat (reference :249-269).  Returned as an
    expanded VIEW (no copy): the tower broadcasts a gray plane to RGB in registers."""
    B, F, H, W = obs_batch.shape
    return obs_batch.unsqueeze(2).expand(B, F, 3, H, W)


def process_multiframe_clip_embeddings(rgb_frames: torch.Tensor, clip_model, ablation_mode, modality: str,
                                       batch_size: int, device) -> torch.Tensor:
    """[B,F,3,h,w] -> [B, F*512]: all B*F frames through the tower in one call (reference :272-299)."""
    import shared.clip_ppo_utils as U
    B, F = rgb_frames.shape[0], rgb_frames.shape[1]
    if rgb_frames.stride(2) == 0:                       # expanded gray view: feed the single plane, C = 1
        frames = rgb_frames[:, :, 0].reshape(B * F, 1, *rgb_frames.shape[-2:])
    else:
        frames = rgb_frames.reshape(B * F, *rgb_frames.shape[2:])
    e = U.generate_clip_embeddings(ablation_mode, clip_model, modality=modality, batch_size=B * F, device=device,
                                   images=frames)
    return e.reshape(B, F * e.shape[-1])


# ---- a20: latents for the alignment loss without the second encoder forward ---------------------------
def action_value_and_latents(agent, obs: torch.Tensor, action: torch.Tensor):
    """``agent.get_action_and_value(obs, action)`` plus the detached encoder output that
    ``agent.get_latent_representation(obs)`` would recompute (clip_ppo_minigrid.py:262-271, 534): one forward
    of the PPO encoder per minibatch instead of two, bitwise the same latents.  Returns
    (action, logprob, entropy, value, latents); works with any agent exposing ``_pre`` / ``_get_features`` /
    ``actor`` / ``critic`` like the reference's ``Agent``."""
    from torch.distributions.categorical import Categorical
    hidden = agent._get_features(agent._pre(obs))
    probs = Categorical(logits=agent.actor(hidden))
    if action is None:
        action = probs.sample()
    return action, probs.log_prob(action), probs.entropy(), agent.critic(hidden), hidden.detach()


class PolicyStepGraph:
    """The policy forward of one env step - ``agent.get_action_and_value(next_obs)`` under ``torch.no_grad()``
    (clip_ppo_minigrid.py:395-399, clip_ppo_atari.py:590-594) - captured once as a CUDA graph and replayed.

    At env-step batches (E = 8 ... 256 frames) the encoder is ~0.1 ms of device time while the ~25 small launches of the
    heads, ``Categorical`` and the sampling cost ~0.5 ms of host time per step; a replay is one launch.  The agent's
    parameters are read at replay time (optimizers update them in place), sampling draws from the device generator through
    torch's graph-safe Philox offsets, so every replay samples fresh actions.  Returns ``(action, logprob, entropy, value)``
    as views of the graph's static outputs: valid until the next call - copy what must survive (``actions[step] = action``
    does).  ``obs`` must keep the shape / dtype of the example."""

    def __init__(self, agent, example_obs: torch.Tensor):
        self.agent = agent
        self.static_obs = torch.empty_like(example_obs)
        self.static_obs.copy_(example_obs)
        dev = example_obs.device
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():          # warm-up outside capture (lazy kernel attributes, cuBLAS handles)
            for _ in range(2):
                self._forward()
        cur.wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.out = self._forward()

    def _forward(self):
        # the statements of Agent.get_action_and_value; argument validation off: its `.all()` is a host sync, illegal in a capture
        from torch.distributions.categorical import Categorical
        agent = self.agent
        hidden = agent._get_features(agent._pre(self.static_obs.float()))
        probs = Categorical(logits=agent.actor(hidden), validate_args=False)
        action = probs.sample()
        return action, probs.log_prob(action), probs.entropy(), agent.critic(hidden)

    def __call__(self, obs: torch.Tensor):
        self.static_obs.copy_(obs)
        self.graph.replay()
        return self.out


# ---- §8f-1: the PPO encoder on the native kernels ----------------------------------------------------
def use_native_encoder(agent):
    """Swap the scripts' ``agent.network`` (``nn.Sequential`` NatureCNN, clip_ppo_minigrid.py:229-242 / clip_ppo_atari.py:196-209)
    for :class:`clip_ppo_b200.policy.NatureCNN` in place: same parameter names, values, state-dict keys and call signature
    (``agent.network(x)``), forward + backward on csrc/policy.cu.  Build the optimizer AFTER the swap (the parameters are new
    objects).  FROZEN_CLIP agents (``network`` is the CLIP tower / an Identity) are left alone."""
    import torch.nn as nn
    from .policy import NatureCNN
    net = getattr(agent, "network", None)
    if isinstance(net, nn.Sequential) and len(net) == 9 and isinstance(net[0], nn.Conv2d) and isinstance(net[7], nn.Linear):
        agent.network = NatureCNN.from_sequential(net)
    return agent


# ---- a21: GAE (clip_ppo_minigrid.py:437-450 = clip_ppo_atari.py:619-632) ---------------------------
def compute_gae(rewards: torch.Tensor, values: torch.Tensor, dones: torch.Tensor, next_value: torch.Tensor,
                next_done: torch.Tensor, gamma: float = 0.99, gae_lambda: float = 0.95) -> Tuple[torch.Tensor, torch.Tensor]:
    return L.gae(rewards, values, dones, next_value, next_done, gamma, gae_lambda)


# ---- a22: PPO minibatch loss (clip_ppo_minigrid.py:498-531,559) ------------------------------------
def ppo_minibatch_loss(newlogprob, entropy, newvalue, old_logprob, advantages, returns, old_values,
                       clip_loss: Optional[torch.Tensor] = None, clip_lambda: float = 0.0, clip_coef: float = 0.1,
                       ent_coef: float = 0.01, vf_coef: float = 0.5, norm_adv: bool = True,
                       clip_vloss: bool = True, adv_stats: Optional[tuple] = None) -> Dict[str, torch.Tensor]:
    return L.ppo_loss(newlogprob, entropy, newvalue, old_logprob, advantages, returns, old_values, clip_loss,
                      clip_lambda, clip_coef, ent_coef, vf_coef, norm_adv, clip_vloss, adv_stats)
