"""§8f-1: the native NatureCNN encoder (csrc/policy.cu) against torch.nn in fp32 (TF32 off) - outputs and ALL gradients."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _reference_network(C, seed):
    """The scripts' encoder (clip_ppo_minigrid.py:229-242 / clip_ppo_atari.py:196-209), orthogonal init like layer_init."""
    torch.manual_seed(seed)
    seq = nn.Sequential(nn.Conv2d(C, 32, 8, stride=4), nn.ReLU(), nn.Conv2d(32, 64, 4, stride=2), nn.ReLU(),
                        nn.Conv2d(64, 64, 3, stride=1), nn.ReLU(), nn.Flatten(), nn.Linear(64 * 7 * 7, 512), nn.ReLU())
    for m in seq:
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            nn.init.orthogonal_(m.weight, 2 ** 0.5)
            nn.init.normal_(m.bias, std=0.1)          # non-zero biases so their path is exercised
    return seq.cuda()


@pytest.fixture(autouse=True)
def _fp32_math():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def _native_activations(hidden, mb):
    """The three post-ReLU activations the native forward kept for its backward (NHWC, first in its workspace), as NCHW."""
    ws = [t for t in hidden.grad_fn.saved_tensors if t.dtype == torch.uint8][0].view(torch.float32)
    r64 = lambda n: (n + 63) // 64 * 64
    n1, n2, n3 = mb * 400 * 32, mb * 81 * 64, mb * 49 * 64
    a1 = ws[:n1].view(mb, 20, 20, 32).permute(0, 3, 1, 2)
    a2 = ws[r64(n1):r64(n1) + n2].view(mb, 9, 9, 64).permute(0, 3, 1, 2)
    a3 = ws[r64(n1) + r64(n2):r64(n1) + r64(n2) + n3].view(mb, 7, 7, 64).permute(0, 3, 1, 2)
    return a1, a2, a3


@pytest.mark.parametrize("mb,C,layout", [(256, 3, "nhwc"), (37, 4, "nchw"), (2048, 3, "nhwc"), (300, 4, "nchw"), (33, 3, "nchw"),
                                         (8, 3, "nhwc"), (64, 3, "nhwc"), (64, 4, "nchw"), (1, 3, "nchw")])      # env-step batches: split-K forward
def test_nature_cnn_forward_backward_vs_torch(native, mb, C, layout):
    """Outputs and all eight gradients against torch.nn in fp32.  A ReLU whose pre-activation sits within rounding noise of zero
    is open in one fp32 implementation and closed in the other; the gradients then differ by that unit's whole contribution
    (1e-3 .. 1e-2 relative at small batches) although both are right.  So the torch reference applies the NATIVE forward's
    gates (checked to differ from torch's own only where |pre-activation| < 1e-5) - everything else is plain torch autograd."""
    from clip_ppo_b200.policy import NatureCNN
    seq = _reference_network(C, seed=mb + C)
    net = NatureCNN.from_sequential(seq)
    assert list(net.state_dict().keys()) == list(seq.state_dict().keys())
    g = torch.Generator(device="cuda").manual_seed(mb)
    if layout == "nhwc":            # MiniGrid: [mb, 84, 84, 3] fp32 0..255, `_pre` = permute + /255 (clip_ppo_minigrid.py:244-247)
        obs = torch.randint(0, 256, (mb, 84, 84, C), device="cuda", generator=g).float()
        x_ref = obs.permute(0, 3, 1, 2).contiguous() / 255.0
        x_view, scale = obs.permute(0, 3, 1, 2), 1.0 / 255.0
    else:                           # Atari: [mb, 4, 84, 84] fp32 0..255, `self.network(x / 255.0)` (clip_ppo_atari.py:229)
        obs = torch.randint(0, 256, (mb, C, 84, 84), device="cuda", generator=g).float()
        x_ref, x_view, scale = obs / 255.0, obs, 1.0 / 255.0
    gh = torch.randn(mb, 512, device="cuda", generator=g)
    h = net(x_view, in_scale=scale)
    acts = [a.clone() for a in _native_activations(h, mb)]      # before backward() frees the saved workspace
    (h * gh).sum().backward()
    # torch forward with the native gates
    t = x_ref
    for conv, act in zip((seq[0], seq[2], seq[4]), acts):
        pre = conv(t)
        gate = act > 0
        disagree = gate != (pre > 0)
        assert not disagree.any() or pre[disagree].abs().max().item() < 1e-5
        assert _rel(act, torch.relu(pre)) <= 1e-4
        t = pre * gate
    pre = seq[7](t.flatten(1))
    gate = h > 0
    disagree = gate != (pre > 0)
    assert not disagree.any() or pre[disagree].abs().max().item() < 1e-5
    h_ref = pre * gate
    (h_ref * gh).sum().backward()
    assert _rel(h, h_ref) <= 1e-4
    for (k, p), (_, q) in zip(net.named_parameters(), seq.named_parameters()):
        assert p.grad is not None and _rel(p.grad, q.grad) <= 1e-4, k


def test_nature_cnn_is_deterministic_and_batch_invariant(native):
    from clip_ppo_b200.policy import NatureCNN
    net = NatureCNN.from_sequential(_reference_network(3, seed=1))
    x = torch.rand(96, 3, 84, 84, device="cuda")
    a, b = net(x), net(x)
    assert torch.equal(a, b)
    assert torch.equal(net(x[:32]), a[:32])             # a row's features do not depend on the rest of the minibatch
    # ... nor on whether the batch is small enough for the split-K forward (env-step batches) or not (minibatches): both add
    # the same canonical K chunks in the same order
    big = torch.rand(1500, 3, 84, 84, device="cuda")
    big[100:164] = x[:64]
    with torch.no_grad():
        assert torch.equal(net(big)[100:164], net(x[:64]))
        assert torch.equal(net(big[:600])[100:164], net(x[:64]))
    a.sum().backward()
    g1 = [p.grad.clone() for p in net.parameters()]
    net.zero_grad()
    net(x).sum().backward()
    assert all(torch.equal(u, p.grad) for u, p in zip(g1, net.parameters()))


def test_agent_update_with_native_encoder_matches_torch(native):
    """One optimiser step of the reference Agent (encoder + actor + critic, Adam, clip_grad_norm_) with the encoder swapped
    for the native one: same loss, same updated parameters up to fp32 noise."""
    from clip_ppo_b200.policy import NatureCNN

    class Agent(nn.Module):
        def __init__(self, network):
            super().__init__()
            self.network, self.actor, self.critic = network, nn.Linear(512, 7), nn.Linear(512, 1)

        def forward(self, x):
            h = self.network(x)
            return h, self.actor(h), self.critic(h)

    seq = _reference_network(3, seed=5)
    torch.manual_seed(9)
    ref = Agent(seq).cuda()
    mine = Agent(NatureCNN.from_sequential(seq)).cuda()
    mine.actor.load_state_dict(ref.actor.state_dict()); mine.critic.load_state_dict(ref.critic.state_dict())
    x = torch.rand(128, 3, 84, 84, device="cuda")
    tgt = torch.randn(128, 512, device="cuda")
    losses = []
    for ag in (ref, mine):
        opt = torch.optim.Adam(ag.parameters(), lr=2.5e-4, eps=1e-5)
        h, logits, v = ag(x)
        loss = (h - tgt).square().mean() + logits.logsumexp(-1).mean() + 0.5 * v.square().mean()
        opt.zero_grad(); loss.backward()
        nn.utils.clip_grad_norm_(ag.parameters(), 0.5)
        opt.step()
        losses.append(loss.item())
    assert abs(losses[0] - losses[1]) <= 1e-5 * abs(losses[0])
    for (k, p), (_, q) in zip(mine.named_parameters(), ref.named_parameters()):
        assert (p - q).abs().max().item() <= 1e-5, k         # lr-sized steps; Adam normalises the gradient magnitude away


def test_use_native_encoder_swaps_in_place(native):
    from clip_ppo_b200 import rollout
    from clip_ppo_b200.policy import NatureCNN

    class Agent(nn.Module):
        def __init__(self):
            super().__init__()
            self.network = _reference_network(4, seed=2)
            self.actor, self.critic = nn.Linear(512, 4), nn.Linear(512, 1)

        def _pre(self, x):
            return x / 255.0

        def _get_features(self, x):
            return self.network(x)

    a = Agent().cuda()
    sd = {k: v.clone() for k, v in a.state_dict().items()}
    x = torch.randint(0, 256, (16, 4, 84, 84), device="cuda").float()
    want = a.network(x / 255.0)
    rollout.use_native_encoder(a)
    assert isinstance(a.network, NatureCNN) and list(a.state_dict().keys()) == list(sd.keys())
    assert all(torch.equal(a.state_dict()[k], sd[k]) for k in sd)
    act, lp, ent, val, lat = rollout.action_value_and_latents(a, x, None)
    assert _rel(lat, want) <= 1e-4 and not lat.requires_grad


def test_policy_step_graph_matches_eager_and_tracks_the_parameters(native):
    """rollout.PolicyStepGraph: the env step's policy forward as one CUDA-graph replay - same logits / values as the eager
    statements (given the sampled action), fresh samples on every replay, and parameter updates are seen without re-capture."""
    import torch.nn as nn
    from torch.distributions.categorical import Categorical
    from clip_ppo_b200 import rollout
    from clip_ppo_b200.policy import NatureCNN

    class Agent(nn.Module):
        def __init__(self):
            super().__init__()
            self.network = NatureCNN.from_sequential(_reference_network(3, seed=2))
            self.actor, self.critic = nn.Linear(512, 7), nn.Linear(512, 1)

        def _pre(self, x):
            return x.permute(0, 3, 1, 2)

        def _get_features(self, x):
            return self.network(x, in_scale=1.0 / 255.0)

    torch.manual_seed(0)
    agent = Agent().cuda()
    obs = torch.randint(0, 256, (64, 84, 84, 3), device="cuda", dtype=torch.uint8)
    step = rollout.PolicyStepGraph(agent, obs)
    acts = []
    for trial in range(3):
        a, lp, ent, v = step(obs)
        with torch.no_grad():
            h = agent._get_features(agent._pre(obs.float()))
            probs = Categorical(logits=agent.actor(h))
            assert torch.allclose(lp, probs.log_prob(a), atol=1e-6) and torch.allclose(ent, probs.entropy(), atol=1e-6)
            assert torch.allclose(v, agent.critic(h), atol=1e-6)
        acts.append(a.clone())
        if trial == 1:                                  # an optimizer-style in-place update must be visible to the next replay
            with torch.no_grad():
                for p in agent.parameters():
                    p.add_(0.01 * torch.randn_like(p))
    assert not torch.equal(acts[0], acts[1])            # the device generator advances between replays
    obs2 = torch.randint(0, 256, (64, 84, 84, 3), device="cuda", dtype=torch.uint8)
    a2, lp2, _, v2 = step(obs2)
    with torch.no_grad():
        h2 = agent._get_features(agent._pre(obs2.float()))
        assert torch.allclose(v2, agent.critic(h2), atol=1e-6)

