"""One CLIP-PPO minibatch update, end to end: the statements of clip_ppo_minigrid.py:453-564 (embeddings once per
iteration, policy forward, PPO terms, alignment loss, lambda warm-up, backward, all the way to the gradients
that get clipped and all-reduced) run twice on the same inputs - through this repository's kernels on the
GPU, and as the script's own torch expressions on the CPU in fp32 with the oracle tower.
north_star tolerances: embeddings cosine >= 0.999, loss within 1e-3 relative."""
import numpy as np
import pytest
import torch
import torch.nn as nn
from torch.distributions.categorical import Categorical

from oracle import losses as ol, vit as ov

pytestmark = pytest.mark.gpu


class Agent(nn.Module):
    """The script's NatureCNN actor-critic (clip_ppo_minigrid.py:213-271 restated; 7 actions), seeded."""

    def __init__(self, n_actions=7, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.network = nn.Sequential(nn.Conv2d(3, 32, 8, stride=4), nn.ReLU(), nn.Conv2d(32, 64, 4, stride=2), nn.ReLU(),
                                     nn.Conv2d(64, 64, 3, stride=1), nn.ReLU(), nn.Flatten(), nn.Linear(64 * 7 * 7, 512), nn.ReLU())
        self.actor = nn.Linear(512, n_actions)
        self.critic = nn.Linear(512, 1)
        self.temporal_projection = nn.Linear(512, 512)        # the "c with grad" case of compute_cosine_embedding_loss
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * (0.5 / max(1.0, float(np.sqrt(p[0].numel() if p.dim() > 1 else 1)))))

    def hidden(self, x):                                       # x [B,H,W,C] fp32 0..255
        return self.network(x.permute(0, 3, 1, 2).contiguous() / 255.0)

    def get_action_and_value(self, x, action):
        h = self.hidden(x)
        probs = Categorical(logits=self.actor(h))
        return probs.log_prob(action), probs.entropy(), self.critic(h)


def _rollout(T, E, seed):
    g = torch.Generator().manual_seed(seed)
    obs = torch.randint(0, 256, (T * E, 84, 84, 3), generator=g).float()
    return dict(obs=obs, actions=torch.randint(0, 7, (T * E,), generator=g),
                logprobs=-torch.rand(T * E, generator=g) * 2.0, rewards=(torch.rand(T, E, generator=g) < 0.1).float(),
                dones=(torch.rand(T, E, generator=g) < 0.05).float(), values=torch.randn(T, E, generator=g) * 0.5,
                next_value=torch.randn(1, E, generator=g) * 0.5, next_done=(torch.rand(E, generator=g) < 0.05).float())


def _update(agent, r, mb, clip_embeddings, lam, detach_latents, gae_fn, cos_fn, ppo_fn):
    """clip_ppo_minigrid.py:437-450 (GAE) and :494-562 (one minibatch) with the three kernels injected."""
    T, E = r["rewards"].shape
    adv, ret = gae_fn(r["rewards"], r["values"], r["dones"], r["next_value"], r["next_done"], 0.99, 0.95)
    b_adv, b_ret, b_val = adv.reshape(-1), ret.reshape(-1), r["values"].reshape(-1)
    newlogprob, entropy, newvalue = agent.get_action_and_value(r["obs"][mb], r["actions"][mb])
    z = agent.hidden(r["obs"][mb])
    if detach_latents:                                         # MiniGrid script: get_latent_representation detaches
        clip_loss = cos_fn(z.detach(), clip_embeddings[mb])
    else:                                                      # both arguments carry gradient
        clip_loss = cos_fn(z, agent.temporal_projection(clip_embeddings[mb]))
    out = ppo_fn(newlogprob, entropy, newvalue, r["logprobs"][mb], b_adv[mb], b_ret[mb], b_val[mb], clip_loss, lam)
    agent.zero_grad()
    out["loss"].backward()
    return out, clip_loss.detach(), {n: p.grad.detach().clone() for n, p in agent.named_parameters() if p.grad is not None}


@pytest.mark.parametrize("native_encoder", [False, True], ids=["torch-encoder", "native-encoder"])
@pytest.mark.parametrize("detach_latents", [True, False])
def test_minibatch_update_matches_the_script_on_cpu(native, detach_latents, native_encoder):
    import shared.clip_ppo_utils as U
    from clip_ppo_b200 import rollout
    T, E, mbsz = 8, 6, 24
    r = _rollout(T, E, seed=11)
    mb = torch.from_numpy(np.random.RandomState(3).permutation(T * E)[:mbsz])
    lam = U.get_clip_lambda_with_warmup(0.5, 7, 16)             # past the warm-up: 0.5 (large, so the term matters)
    assert lam == ol.clip_lambda_with_warmup(0.5, 7, 16)

    # ---- CPU: the script's expressions, fp32, oracle tower ----
    sd = ov.random_state_dict(ov.VIT_B32, 0)
    emb_cpu = ov.image_embeddings(sd, r["obs"].permute(0, 3, 1, 2).contiguous())
    agent_cpu = Agent(seed=5)
    ppo_cpu = lambda nlp, ent, nv, olp, a, R, V, cl, l: ol.ppo_loss(nlp, ent, nv, olp, a, R, V, cl, l)
    out_c, cl_c, g_c = _update(agent_cpu, r, mb, emb_cpu, lam, detach_latents, ol.gae, ol.cosine_embedding_loss, ppo_cpu)

    # ---- GPU: this repository ----
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False     # the stock-PyTorch agent (and the heads) in true fp32
    try:
        model = U.load_clip_model("ViT-B/32", "cuda")          # seeded-random weights == ov.random_state_dict(VIT_B32, 0)
        rg = {k: v.cuda() for k, v in r.items()}
        emb_gpu = U.generate_clip_embeddings(U.AblationMode.NONE, model, "image", T * E, "cuda",
                                             images=rg["obs"].permute(0, 3, 1, 2).contiguous())
        agent_gpu = Agent(seed=5).cuda()
        if native_encoder:                                      # the NatureCNN on csrc/policy.cu, same parameter names
            from clip_ppo_b200.policy import NatureCNN
            rollout.use_native_encoder(agent_gpu)
            assert isinstance(agent_gpu.network, NatureCNN)
        ppo_gpu = lambda nlp, ent, nv, olp, a, R, V, cl, l: rollout.ppo_minibatch_loss(nlp, ent, nv, olp, a, R, V, cl, l)
        out_g, cl_g, g_g = _update(agent_gpu, rg, mb.cuda(), emb_gpu, lam, detach_latents, rollout.compute_gae,
                                   U.compute_cosine_embedding_loss, ppo_gpu)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32

    cos = torch.sum(emb_gpu.cpu() * emb_cpu, dim=-1)
    assert cos.min().item() >= 0.999, cos.min()
    assert abs(cl_g.item() - cl_c.item()) <= 1e-3 * abs(cl_c.item()), (cl_g.item(), cl_c.item())
    lc, lg = out_c["loss"].item(), out_g["loss"].item()
    assert abs(lg - lc) <= 1e-3 * abs(lc), (lg, lc)
    assert set(g_g) == set(g_c)
    if not detach_latents:
        assert "temporal_projection.weight" in g_g and g_c["network.0.weight"].abs().max() > 0
    for name, gc in g_c.items():
        gg = g_g[name].cpu()
        denom = gc.norm().item()
        # gradients that are linear in the bf16 tower's embeddings inherit their error (cosine 0.9999 = 1.4e-2
        # relative in the vector); everything else is fp32 arithmetic on both sides
        rel = 2e-2 if name.startswith("temporal_projection") else 2e-3
        assert (gg - gc).norm().item() <= rel * denom + 1e-7, (name, (gg - gc).norm().item(), denom)
