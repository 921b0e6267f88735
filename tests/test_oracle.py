"""The CPU oracle against the golden vectors produced by the reference's own code
(oracle/make_goldens.py) and against an independent HF CLIP tower."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import disturb as od
from oracle import losses as ol
from oracle import vit as ov

from conftest import GOLDEN

DISTURB_FILES = sorted(glob.glob(os.path.join(GOLDEN, "disturb_*.npz")))


def load_case(path):
    g = np.load(path)
    u8 = torch.from_numpy(g["u8"])
    x = (u8.float() / 255.0).permute(0, 3, 1, 2)
    if str(g["layout"]) != "nhwc_view":
        x = x.contiguous()
    return g, x


def test_golden_files_present():
    assert len(DISTURB_FILES) == 11
    assert os.path.exists(os.path.join(GOLDEN, "vit_b32_seed0.npz"))
    assert os.path.exists(os.path.join(GOLDEN, "losses.npz"))


@pytest.mark.parametrize("path", DISTURB_FILES, ids=[os.path.basename(p)[:-4] for p in DISTURB_FILES])
def test_disturb_oracle_bit_exact(path):
    g, x = load_case(path)
    cfg = od.SEVERITY_TABLE[str(g["severity"])]
    noise = torch.from_numpy(g["noise"])
    k1d = od.gaussian_kernel1d(int(g["k"]), float(g["sigma_b"]))
    assert torch.equal(k1d, torch.from_numpy(g["k1d"]))
    assert int(g["k"]) == od.blur_kernel_size(cfg["blur_sigma"])
    assert (int(g["ph"]), int(g["pw"])) == od.cutout_patch(x.shape[-2], x.shape[-1], cfg["cutout"])
    out = od.disturb(x, noise, cfg["noise_sigma"], float(g["c"]), k1d, int(g["sh"]), int(g["sw"]), int(g["ph"]), int(g["pw"]))
    assert torch.equal(out, torch.from_numpy(g["out"]))
    # stage pins (stored as fp16)
    s1 = od.add_noise(x, noise, cfg["noise_sigma"])
    s2 = od.contrast(s1, float(g["c"]))
    s3 = od.blur(s2, k1d)
    for mine, key in ((s1, "after_noise"), (s2, "after_contrast"), (s3, "after_blur")):
        assert (mine - torch.from_numpy(g[key]).float()).abs().max() < 1e-3


@pytest.mark.parametrize("path", DISTURB_FILES[:3], ids=["seed0", "seed1", "seed2"])
def test_disturb_rng_replay(path):
    """Seeding like the reference constructor and replaying its draws reproduces the golden."""
    g, x = load_case(path)
    torch.manual_seed(int(g["seed"]))
    out = od.disturb_seeded(x, str(g["severity"]))
    assert torch.equal(out, torch.from_numpy(g["out"]))


def test_patch_sizes_match_survey():
    want84 = {"MILD": (26, 27), "MODERATE": (34, 35), "HARD": (35, 36), "SEVERE": (42, 42)}
    want224 = {"MILD": (70, 71), "MODERATE": (92, 92), "HARD": (95, 95), "SEVERE": (112, 112)}
    for sev, cfg in od.SEVERITY_TABLE.items():
        assert od.cutout_patch(84, 84, cfg["cutout"]) == want84[sev]
        assert od.cutout_patch(224, 224, cfg["cutout"]) == want224[sev]
    assert [od.blur_kernel_size(c["blur_sigma"]) for c in od.SEVERITY_TABLE.values()] == [3, 5, 5, 7]


def test_vit_oracle_against_reference_goldens():
    g = np.load(os.path.join(GOLDEN, "vit_b32_seed0.npz"))
    sd = ov.random_state_dict(ov.VIT_B32, seed=int(g["weights_seed"]))
    for tag in ("84", "224"):
        img = torch.from_numpy(g[f"img{tag}"]).float()
        emb = ov.image_embeddings(sd, img)
        ref = torch.from_numpy(g[f"emb{tag}"])
        assert (emb - ref).abs().max() < 5e-5
        assert torch.sum(emb * ref, dim=-1).min() > 0.99999
    # Atari double /255 path
    rgb = torch.from_numpy(g["gray_atari"]).float().repeat(1, 3, 1, 1) / 255.0
    emb = ov.image_embeddings(sd, rgb)
    assert torch.sum(emb * torch.from_numpy(g["emb_atari"]), dim=-1).min() > 0.99999
    pre = ov.preprocess(torch.from_numpy(g["img84"][:1]).float(), True)
    assert torch.equal(pre, torch.from_numpy(g["pre84_first"]))


def test_vit_oracle_matches_hf_small():
    """Independent implementation check on a small config (fast): restated tower == HF CLIP."""
    transformers = pytest.importorskip("transformers")
    cfg = ov.VitConfig(width=128, layers=2, heads=2, patch=32, image=224, out_dim=64)
    sd = ov.random_state_dict(cfg, seed=5)
    from oracle.make_goldens import _HFTower
    tower = _HFTower(sd)
    x = torch.randn(3, 3, 224, 224, generator=torch.Generator().manual_seed(0))
    a = ov.vision_tower(sd, x)
    b = tower.encode_image(x)
    assert (a - b).abs().max() < 2e-5


def test_text_oracle_against_reference_goldens():
    """tests/golden/text_b32_seed0.npz: the reference's own text branch (shared/clip_ppo_utils.py:132-139) over the
    HF text tower with the seeded weights."""
    from oracle import text as ot
    g = np.load(os.path.join(GOLDEN, "text_b32_seed0.npz"))
    sd = ot.random_state_dict(ot.TEXT_B32, seed=int(g["weights_seed"]))
    tokens = torch.from_numpy(g["tokens"]).long()
    assert tokens.shape == (6, 77) and int(tokens[0].argmax()) == 2 and int(tokens[1].argmax()) == 76
    emb = ot.text_embeddings(sd, tokens)
    ref = torch.from_numpy(g["emb"])
    assert (emb - ref).abs().max() < 5e-5
    assert torch.sum(emb * ref, dim=-1).min() > 0.99999


def test_text_oracle_matches_hf_small():
    """Independent implementation check on a small config (fast): restated text tower == HF CLIPTextModelWithProjection."""
    pytest.importorskip("transformers")
    from oracle import text as ot
    from oracle.make_goldens import _HFTextTower
    cfg = ot.TextConfig(width=128, layers=2, heads=2, context=77, vocab=1000, out_dim=64)
    sd = ot.random_state_dict(cfg, seed=3)
    tokens = ot.random_tokens(5, cfg, seed=1)
    a = ot.text_tower(sd, tokens)
    b = _HFTextTower(sd).encode_text(tokens)
    assert (a - b).abs().max() < 2e-5
    # causal: changing a token AFTER the EOT position must not change the embedding
    t2 = tokens.clone()
    eot = int(tokens[0].argmax())
    if eot + 1 < 77:
        t2[0, eot + 1:] = 5
        assert torch.equal(ot.text_tower(sd, t2)[0], a[0])


def test_loss_oracle_against_reference_goldens():
    g = np.load(os.path.join(GOLDEN, "losses.npz"))
    z = torch.from_numpy(g["cos_z"]).requires_grad_(True)
    c = torch.from_numpy(g["cos_c"]).requires_grad_(True)
    loss = ol.cosine_embedding_loss(z, c)
    loss.backward()
    assert abs(loss.item() - float(g["cos_loss"])) < 1e-6
    assert torch.allclose(c.grad, torch.from_numpy(g["cos_gc"]), atol=1e-7)
    assert torch.allclose(z.grad, torch.from_numpy(g["cos_gz"]), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose([ol.clip_lambda_with_warmup(1e-5, i, 16) for i in range(16)], g["warmup_16"], rtol=0, atol=0)
    np.testing.assert_allclose([ol.clip_lambda_with_warmup(3e-4, i, 100) for i in range(100)], g["warmup_100"], rtol=0, atol=0)
    with pytest.raises(ValueError):
        ol.cosine_embedding_loss(torch.zeros(2, 512), torch.zeros(2, 768))

    t = lambda k: torch.from_numpy(g[k])
    adv, ret = ol.gae(t("gae_rewards"), t("gae_values"), t("gae_dones"), t("gae_next_value"), t("gae_next_done"))
    assert torch.equal(adv, t("gae_advantages")) and torch.equal(ret, t("gae_returns"))

    nlp = t("ppo_newlogprob").requires_grad_(True)
    ent = t("ppo_entropy").requires_grad_(True)
    nv = t("ppo_newvalue").requires_grad_(True)
    r = ol.ppo_loss(nlp, ent, nv, t("ppo_b_logprobs"), t("ppo_b_advantages"), t("ppo_b_returns"), t("ppo_b_values"),
                    clip_loss=torch.tensor(float(g["ppo_clip_loss"])), clip_lambda=float(g["ppo_clip_lambda"]))
    r["loss"].backward()
    for k, gk in (("loss", "ppo_loss"), ("pg_loss", "ppo_pg_loss"), ("v_loss", "ppo_v_loss"), ("entropy", "ppo_entropy_loss"),
                  ("old_approx_kl", "ppo_old_approx_kl"), ("approx_kl", "ppo_approx_kl"), ("clipfrac", "ppo_clipfrac")):
        assert abs(r[k].item() - float(g[gk])) < 1e-6, k
    assert torch.allclose(nlp.grad, t("ppo_g_newlogprob"), atol=1e-8)
    assert torch.allclose(ent.grad, t("ppo_g_entropy"), atol=1e-8)
    assert torch.allclose(nv.grad, t("ppo_g_newvalue"), atol=1e-8)


def test_philox_oracle_known_answers():
    """oracle/philox.py against Random123's published known-answer vectors for philox4x32-10, and the moments of the
    Box-Muller draws built on it (the generator behind apply_disturbances(noise_seed=...))."""
    import numpy as np
    from oracle import philox as P
    for ctr, key, out in P.KAT:
        r = P.philox4x32_10(*[np.uint32(v) for v in ctr], *key)
        assert tuple(int(v) for v in r) == out
    n = P.normal_noise(seed=123456789012345, offset=7, shape=(16, 3, 224, 224))
    assert n.dtype == np.float32 and np.isfinite(n).all()
    n = n.astype(np.float64)                                   # 2.4 M draws: the tolerances are > 4 standard errors
    assert abs(n.mean()) < 3e-3 and abs(n.std() - 1.0) < 3e-3
    assert abs((n ** 3).mean()) < 1e-2 and abs((n ** 4).mean() - 3.0) < 3e-2
    # a shard draws the whole batch's noise
    assert np.array_equal(P.normal_noise(5, 2, (3, 3, 84, 84), first_image=4), P.normal_noise(5, 2, (7, 3, 84, 84))[4:])

