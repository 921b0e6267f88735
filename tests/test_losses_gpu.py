"""Parity of the alignment loss / GAE / PPO loss kernels with the reference goldens and the oracle.
Tolerance: 1e-3 relative on losses (north_star); GAE is bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import losses as ol

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _g():
    return np.load(os.path.join(GOLDEN, "losses.npz"))


def test_cosine_loss_golden_forward_backward(native):
    import shared.clip_ppo_utils as U
    g = _g()
    z = torch.from_numpy(g["cos_z"]).cuda().requires_grad_(True)
    c = torch.from_numpy(g["cos_c"]).cuda().requires_grad_(True)
    loss = U.compute_cosine_embedding_loss(z, c)
    assert loss.shape == () and loss.dtype == torch.float32
    assert abs(loss.item() - float(g["cos_loss"])) <= 1e-3 * abs(float(g["cos_loss"]))
    assert abs(loss.item() - float(g["cos_loss"])) <= 1e-6
    loss.backward()
    assert torch.allclose(c.grad.cpu(), torch.from_numpy(g["cos_gc"]), rtol=1e-4, atol=1e-8)
    assert torch.allclose(z.grad.cpu(), torch.from_numpy(g["cos_gz"]), rtol=1e-4, atol=1e-8)


@pytest.mark.parametrize("rows,dim", [(256, 512), (2048, 512), (8192, 512), (7, 768), (33, 100)])
def test_cosine_loss_random_vs_oracle(native, rows, dim):
    import shared.clip_ppo_utils as U
    gen = torch.Generator().manual_seed(rows * 31 + dim)
    z = torch.relu(torch.randn(rows, dim, generator=gen))
    c = torch.randn(rows, dim, generator=gen)
    zr, cr = z.clone().requires_grad_(True), c.clone().requires_grad_(True)
    ref = ol.cosine_embedding_loss(zr, cr)
    (ref * 2.5).backward()
    zg, cg = z.cuda().requires_grad_(True), c.cuda().requires_grad_(True)
    out = U.compute_cosine_embedding_loss(zg, cg)
    (out * 2.5).backward()
    assert abs(out.item() - ref.item()) <= 1e-5
    assert torch.allclose(zg.grad.cpu(), zr.grad, rtol=1e-4, atol=1e-9)
    assert torch.allclose(cg.grad.cpu(), cr.grad, rtol=1e-4, atol=1e-9)
    # gradient only through c (the Atari temporal_projection case, clip_ppo_atari.py:730)
    c2 = c.cuda().requires_grad_(True)
    U.compute_cosine_embedding_loss(z.cuda(), c2).backward()
    assert torch.allclose(c2.grad.cpu() * 2.5, cr.grad, rtol=1e-4, atol=1e-9)
    # self-similarity => 0
    assert abs(U.compute_cosine_embedding_loss(c.cuda(), c.cuda()).item()) < 1e-6


def test_gae_golden_bit_exact(native):
    from clip_ppo_b200 import rollout
    g = _g()
    t = lambda k: torch.from_numpy(g[k]).cuda()
    adv, ret = rollout.compute_gae(t("gae_rewards"), t("gae_values"), t("gae_dones"), t("gae_next_value"), t("gae_next_done"))
    assert torch.equal(adv.cpu(), torch.from_numpy(g["gae_advantages"]))
    assert torch.equal(ret.cpu(), torch.from_numpy(g["gae_returns"]))


@pytest.mark.parametrize("T,E", [(128, 8), (128, 64), (128, 256), (5, 3), (1, 1000)])
def test_gae_random_vs_oracle(native, T, E):
    from clip_ppo_b200 import rollout
    gen = torch.Generator().manual_seed(T * 1000 + E)
    r = (torch.rand(T, E, generator=gen) < 0.1).float() * torch.rand(T, E, generator=gen)
    v = torch.randn(T, E, generator=gen)
    d = (torch.rand(T, E, generator=gen) < 0.02).float()
    nv, nd = torch.randn(1, E, generator=gen), (torch.rand(E, generator=gen) < 0.3).float()
    ra, rr = ol.gae(r, v, d, nv, nd, 0.99, 0.95)
    adv, ret = rollout.compute_gae(r.cuda(), v.cuda(), d.cuda(), nv.cuda(), nd.cuda(), 0.99, 0.95)
    assert torch.equal(adv.cpu(), ra) and torch.equal(ret.cpu(), rr)


def test_ppo_loss_golden(native):
    from clip_ppo_b200 import rollout
    g = _g()
    t = lambda k: torch.from_numpy(g[k]).cuda()
    nlp, ent, nv = t("ppo_newlogprob").requires_grad_(True), t("ppo_entropy").requires_grad_(True), t("ppo_newvalue").requires_grad_(True)
    r = rollout.ppo_minibatch_loss(nlp, ent, nv, t("ppo_b_logprobs"), t("ppo_b_advantages"), t("ppo_b_returns"), t("ppo_b_values"),
                                   clip_loss=torch.tensor(float(g["ppo_clip_loss"]), device="cuda"), clip_lambda=float(g["ppo_clip_lambda"]))
    for k, gk in (("loss", "ppo_loss"), ("pg_loss", "ppo_pg_loss"), ("v_loss", "ppo_v_loss"), ("entropy", "ppo_entropy_loss"),
                  ("old_approx_kl", "ppo_old_approx_kl"), ("approx_kl", "ppo_approx_kl"), ("clipfrac", "ppo_clipfrac")):
        want = float(g[gk])
        assert abs(r[k].item() - want) <= 1e-3 * abs(want) + 1e-7, (k, r[k].item(), want)
    r["loss"].backward()
    assert nv.grad.shape == nv.shape
    assert torch.allclose(nlp.grad.cpu(), torch.from_numpy(g["ppo_g_newlogprob"]), rtol=1e-3, atol=1e-8)
    assert torch.allclose(ent.grad.cpu(), torch.from_numpy(g["ppo_g_entropy"]), rtol=1e-3, atol=1e-9)
    assert torch.allclose(nv.grad.cpu(), torch.from_numpy(g["ppo_g_newvalue"]), rtol=1e-3, atol=1e-8)


@pytest.mark.parametrize("n,norm_adv,clip_vloss", [(256, True, True), (2048, True, False), (8192, False, True), (3, True, True)])
def test_ppo_loss_random_vs_oracle(native, n, norm_adv, clip_vloss):
    from clip_ppo_b200 import rollout
    gen = torch.Generator().manual_seed(n)
    mk = lambda s=1.0: torch.randn(n, generator=gen) * s
    nlp0, ent0, nv0 = -1.0 + mk(0.3), 1.0 + mk(0.1), mk()
    olp, adv, ret, ov = nlp0 + mk(0.2), mk(), mk(), nv0 + mk(0.15)
    clip_loss = torch.tensor(0.42)

    def run(fn, dev):
        a, b, c = (t.clone().to(dev).requires_grad_(True) for t in (nlp0, ent0, nv0))
        cl = clip_loss.clone().to(dev).requires_grad_(True)
        r = fn(a, b, c, olp.to(dev), adv.to(dev), ret.to(dev), ov.to(dev), clip_loss=cl, clip_lambda=3e-4,
               norm_adv=norm_adv, clip_vloss=clip_vloss)
        r["loss"].backward()
        return r, a.grad.cpu(), b.grad.cpu(), c.grad.cpu(), cl.grad.cpu()

    ref, *rg = run(ol.ppo_loss, "cpu")
    out, *og = run(rollout.ppo_minibatch_loss, "cuda")
    for k in ("loss", "pg_loss", "v_loss", "entropy", "old_approx_kl", "approx_kl", "clipfrac"):
        assert abs(out[k].item() - ref[k].item()) <= 1e-3 * abs(ref[k].item()) + 1e-6, k
    for a, b in zip(og, rg):
        assert torch.allclose(a, b, rtol=1e-3, atol=1e-8)


def test_ppo_loss_with_global_advantage_stats_matches_the_whole_minibatch(native):
    """Data parallelism: two ranks' half-minibatches normalised with the GLOBAL advantage statistics
    (distributed.global_advantage_stats -> ppo_loss(adv_stats=...)) give, averaged, the loss and the gradients of the one-GPU
    run on the whole minibatch (reference clip_ppo_minigrid.py:509 normalises over the minibatch it sees)."""
    from clip_ppo_b200 import losses as L
    from clip_ppo_b200.distributed import global_advantage_stats
    g = torch.Generator(device="cuda").manual_seed(4)
    n = 512
    mk = lambda: torch.randn(n, device="cuda", generator=g)
    nlp, ent, nv, adv, ret, ov = -1 + 0.1 * mk(), mk().abs(), mk(), 2 * mk() + 0.5, mk(), mk()
    olp = nlp + 0.05 * mk()

    def run(sl, stats):
        a, b, c = (t[sl].clone().requires_grad_(True) for t in (nlp, ent, nv))
        out = L.ppo_loss(a, b, c, olp[sl], adv[sl], ret[sl], ov[sl], adv_stats=stats)
        out["loss"].backward()
        return out["loss"].detach(), (a.grad, b.grad, c.grad)

    whole, gw = run(slice(0, n), None)
    stats = global_advantage_stats(adv)                    # single process: the statistics of all rows
    assert torch.allclose(stats[0], adv.mean(), atol=1e-6) and torch.allclose(stats[1], adv.std(), rtol=1e-5)
    (l0, g0), (l1, g1) = run(slice(0, n // 2), stats), run(slice(n // 2, n), stats)
    assert abs(float((l0 + l1) / 2 - whole)) <= 1e-5 * abs(float(whole)) + 1e-7
    for i in range(3):                                     # d(mean over n)/dx = half of d(mean over n/2)/dx
        assert torch.allclose(torch.cat([g0[i], g1[i]]) / 2, gw[i], rtol=1e-4, atol=1e-8)
    # and with each half's OWN statistics the result differs (what adv_stats is for)
    (m0, _), (m1, _) = run(slice(0, n // 2), None), run(slice(n // 2, n), None)
    assert abs(float((m0 + m1) / 2 - whole)) > 1e-6
