"""Generate tests/golden/*.npz by running the REFERENCE's own code.  TEST INFRASTRUCTURE ONLY.

Run in the build container only (it needs /root/reference, which does not exist on the GPU
box):   python oracle/make_goldens.py

What executes what:
* disturbances  - ``shared.disturbances_gpu.DisturbanceWrapperGPU(device="cpu")`` imported
                  from /root/reference, unmodified.  The randomness it consumed is recovered by
                  replaying the same global-generator draws and stored beside the
                  output, so the CUDA path can be fed identical noise / contrast / cutout values.
* embeddings    - the reference's ``shared.clip_ppo_utils.generate_clip_embeddings`` /
                  ``compute_cosine_embedding_loss`` / ``get_clip_lambda_with_warmup``, imported
                  behind a stub ``clip`` module (openai/CLIP is not installed) whose
                  ``encode_image`` is transformers' CLIPVisionModelWithProjection loaded with the
                  seeded random weights of ``oracle.vit.random_state_dict`` - i.e. an
                  implementation independent of oracle/vit.py.
* GAE, PPO loss - the reference script's own statements
                  (``clip_ppo_minigrid.py`` lines 439-450, 498-531, 559) are read from
                  /root/reference at run time and exec'd on fixed inputs.  No reference source is
                  copied into this repository.
"""
from __future__ import annotations

import os
import sys
import textwrap
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.environ.get("CLIPPPO_GOLDEN_OUT", os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, ROOT)


def _pin_reference_shared():
    """`shared` must mean the REFERENCE's directory while the goldens are generated, not this repository's drop-in
    package of the same name (a regular package, which would win over the reference's namespace package).  Called
    from the generators only - importing this module (tests/test_oracle.py borrows the HF towers) rebinds nothing."""
    if getattr(sys.modules.get("shared"), "_clipppo_reference_pin", False):
        return
    for name in [m for m in sys.modules if m == "shared" or m.startswith("shared.")]:
        del sys.modules[name]
    ref_shared = types.ModuleType("shared")
    ref_shared.__path__ = [os.path.join(REF, "shared")]
    ref_shared._clipppo_reference_pin = True
    sys.modules["shared"] = ref_shared


from oracle import disturb as od  # noqa: E402
from oracle import vit as ov  # noqa: E402


def _import_reference():
    _pin_reference_shared()
    sys.path.insert(0, REF)
    from shared.disturbances_gpu import DisturbanceWrapperGPU  # type: ignore
    from shared.disturbance_types import DisturbanceSeverity, SEVERITY_CONFIGS  # type: ignore
    return DisturbanceWrapperGPU, DisturbanceSeverity, SEVERITY_CONFIGS


def make_disturb_goldens():
    Wrapper, Sev, CFG = _import_reference()
    # the reference's parameter table itself is a golden (README.md:102-107)
    table = {s.value: CFG[s] for s in CFG}
    for name, row in table.items():
        mine = od.SEVERITY_TABLE[name]
        assert row["gaussian_noise_sigma"] == mine["noise_sigma"] and row["gaussian_blur_sigma"] == mine["blur_sigma"]
        assert tuple(row["contrast_range"]) == tuple(mine["contrast"]) and row["cutout_ratio"] == mine["cutout"]

    cases = []
    for sev in ("MILD", "MODERATE", "HARD", "SEVERE"):
        cases.append((sev, (2, 3, 84, 84), "nchw"))
        cases.append((sev, (2, 1, 84, 84), "nchw"))
    cases.append(("MODERATE", (2, 3, 84, 84), "nhwc_view"))   # clip_ppo_minigrid.py:385
    cases.append(("SEVERE", (1, 3, 224, 224), "nchw"))
    cases.append(("HARD", (3, 3, 40, 56), "nchw"))            # ragged / non-square
    for idx, (sev, shape, layout) in enumerate(cases):
        seed = 1000 + idx
        g = torch.Generator().manual_seed(seed)
        B, C, H, W = shape
        u8 = torch.randint(0, 256, (B, H, W, C), generator=g, dtype=torch.uint8)
        if layout == "nhwc_view":
            x = (u8.float() / 255.0).permute(0, 3, 1, 2)          # non-contiguous view
        else:
            x = (u8.float() / 255.0).permute(0, 3, 1, 2).contiguous()
        w = Wrapper(device="cpu", seed=seed, severity=Sev[sev])
        out = w.apply_disturbances(x)
        # replay the draws (the ctor seeded the global generator)
        torch.manual_seed(seed)
        r = od.draw_call_randomness(x, od.SEVERITY_TABLE[sev])
        k1d = od.gaussian_kernel1d(r["k"], r["sigma_b"])
        mine = od.disturb(x, r["noise"], od.SEVERITY_TABLE[sev]["noise_sigma"], r["c"], k1d,
                          r["sh"], r["sw"], r["ph"], r["pw"])
        err = (mine - out).abs().max().item()
        assert err == 0.0, f"oracle restatement differs from the reference: {sev} {shape} {err}"
        # per-stage reference outputs with the same randomness
        torch.manual_seed(seed)
        w2 = Wrapper(device="cpu", seed=None, severity=Sev[sev])
        s1 = w2.apply_gaussian_noise(x)
        s2 = w2.apply_contrast_jitter(s1)
        s3 = w2.apply_gaussian_blur(s2)
        assert torch.equal(w2.apply_cutout(s3), out)
        name = f"disturb_{idx:02d}_{sev}_{C}x{H}x{W}_{layout}.npz"
        np.savez_compressed(
            os.path.join(OUT, name),
            u8=u8.numpy(), layout=layout, severity=sev,
            noise=np.ascontiguousarray(r["noise"].contiguous().numpy()),   # logical NCHW order
            c=np.float64(r["c"]), sigma_b=np.float64(r["sigma_b"]), k=r["k"], k1d=k1d.numpy(),
            sh=r["sh"], sw=r["sw"], ph=r["ph"], pw=r["pw"], seed=seed,
            after_noise=s1.contiguous().numpy().astype(np.float16),      # stage pins (coarse)
            after_contrast=s2.contiguous().numpy().astype(np.float16),
            after_blur=s3.contiguous().numpy().astype(np.float16),
            out=out.contiguous().numpy())
        print("wrote", name, "restatement max-abs", err)


class _HFTower(torch.nn.Module):
    """encode_image via transformers, weights copied from an openai-layout state dict."""

    def __init__(self, sd):
        super().__init__()
        from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
        cfg = ov.config_from_state_dict(sd)
        hf_cfg = CLIPVisionConfig(hidden_size=cfg.width, intermediate_size=4 * cfg.width,
                                  num_hidden_layers=cfg.layers, num_attention_heads=cfg.heads,
                                  image_size=cfg.image, patch_size=cfg.patch, projection_dim=cfg.out_dim,
                                  hidden_act="quick_gelu", layer_norm_eps=1e-5)
        m = CLIPVisionModelWithProjection(hf_cfg).eval()
        D = cfg.width
        t = {}
        t["vision_model.embeddings.patch_embedding.weight"] = sd["visual.conv1.weight"]
        t["vision_model.embeddings.class_embedding"] = sd["visual.class_embedding"]
        t["vision_model.embeddings.position_embedding.weight"] = sd["visual.positional_embedding"]
        t["vision_model.pre_layrnorm.weight"] = sd["visual.ln_pre.weight"]
        t["vision_model.pre_layrnorm.bias"] = sd["visual.ln_pre.bias"]
        t["vision_model.post_layernorm.weight"] = sd["visual.ln_post.weight"]
        t["vision_model.post_layernorm.bias"] = sd["visual.ln_post.bias"]
        t["visual_projection.weight"] = sd["visual.proj"].t().contiguous()
        for i in range(cfg.layers):
            p = f"visual.transformer.resblocks.{i}."
            h = f"vision_model.encoder.layers.{i}."
            W, b = sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"]
            for j, n in enumerate(("q_proj", "k_proj", "v_proj")):
                t[h + f"self_attn.{n}.weight"] = W[j * D:(j + 1) * D]
                t[h + f"self_attn.{n}.bias"] = b[j * D:(j + 1) * D]
            t[h + "self_attn.out_proj.weight"] = sd[p + "attn.out_proj.weight"]
            t[h + "self_attn.out_proj.bias"] = sd[p + "attn.out_proj.bias"]
            t[h + "layer_norm1.weight"] = sd[p + "ln_1.weight"]
            t[h + "layer_norm1.bias"] = sd[p + "ln_1.bias"]
            t[h + "layer_norm2.weight"] = sd[p + "ln_2.weight"]
            t[h + "layer_norm2.bias"] = sd[p + "ln_2.bias"]
            t[h + "mlp.fc1.weight"] = sd[p + "mlp.c_fc.weight"]
            t[h + "mlp.fc1.bias"] = sd[p + "mlp.c_fc.bias"]
            t[h + "mlp.fc2.weight"] = sd[p + "mlp.c_proj.weight"]
            t[h + "mlp.fc2.bias"] = sd[p + "mlp.c_proj.bias"]
        missing, unexpected = m.load_state_dict(t, strict=False)
        assert not unexpected, unexpected
        assert all("position_ids" in k for k in missing), missing
        self.m = m

    @torch.no_grad()
    def encode_image(self, x):
        return self.m(pixel_values=x).image_embeds


def _install_stub_clip(tower):
    """A stand-in for the absent openai ``clip`` package: just enough surface for the
    reference's shared/clip_ppo_utils.py to import and run its image branch."""
    clip = types.ModuleType("clip")
    model = types.ModuleType("clip.model")

    class VisionTransformer(torch.nn.Module):
        pass

    class CLIP(torch.nn.Module):
        pass

    model.VisionTransformer, model.CLIP = VisionTransformer, CLIP
    clip.model = model
    clip.load = lambda name, device="cpu": (tower, None)
    clip.tokenize = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("text tower out of scope"))
    sys.modules["clip"], sys.modules["clip.model"] = clip, model


def make_vit_goldens():
    sd = ov.random_state_dict(ov.VIT_B32, seed=0)
    tower = _HFTower(sd)
    _install_stub_clip(tower)
    _pin_reference_shared()
    sys.path.insert(0, REF)
    import shared.clip_ppo_utils as ref_utils  # type: ignore

    model = ref_utils.load_clip_model("ViT-B/32", device="cpu")
    g = torch.Generator().manual_seed(7)
    img84 = torch.randint(0, 256, (4, 3, 84, 84), generator=g, dtype=torch.uint8)
    img224 = torch.randint(0, 256, (2, 3, 224, 224), generator=g, dtype=torch.uint8)
    out = {}
    for tag, img in (("84", img84), ("224", img224)):
        emb = ref_utils.generate_clip_embeddings(ref_utils.AblationMode.NONE, model, "image",
                                                 img.shape[0], "cpu", images=img.float())
        mine = ov.image_embeddings(sd, img.float())
        cos = torch.sum(emb * mine, dim=-1).min().item()
        err = (emb - mine).abs().max().item()
        print(f"embeddings {tag}: oracle vs reference+HF tower  max-abs {err:.2e}  min-cos {cos:.8f}")
        assert err < 5e-5 and cos > 0.99999
        out[f"img{tag}"] = img.numpy()
        out[f"emb{tag}"] = emb.numpy()
    # Atari quirk (clip_ppo_atari.py:661 + clip_ppo_utils.py:152): frames pre-divided by 255
    gray = torch.randint(0, 256, (3, 1, 84, 84), generator=g, dtype=torch.uint8)
    rgb = gray.float().repeat(1, 3, 1, 1) / 255.0
    emb = ref_utils.generate_clip_embeddings(ref_utils.AblationMode.NONE, model, "image", 3, "cpu", images=rgb)
    out["gray_atari"] = gray.numpy()
    out["emb_atari"] = emb.numpy()
    # the preprocessing alone (reference lines 146-159 run inside generate_clip_embeddings; its
    # intermediate is recovered by calling the tower hook)
    captured = {}
    orig = tower.encode_image
    tower.encode_image = lambda x: (captured.__setitem__("x", x.clone()), orig(x))[1]
    ref_utils.generate_clip_embeddings(ref_utils.AblationMode.NONE, model, "image", 1, "cpu", images=img84[:1].float())
    tower.encode_image = orig
    out["pre84_first"] = captured["x"].numpy().astype(np.float32)
    assert (ov.preprocess(img84[:1].float(), True) - captured["x"]).abs().max().item() == 0.0
    np.savez_compressed(os.path.join(OUT, "vit_b32_seed0.npz"), weights_seed=0, **out)
    print("wrote vit_b32_seed0.npz")
    return ref_utils


class _HFTextTower(torch.nn.Module):
    """encode_text via transformers (CLIPTextModelWithProjection), weights copied from an openai-layout state dict."""

    def __init__(self, sd):
        super().__init__()
        from transformers import CLIPTextConfig, CLIPTextModelWithProjection
        from oracle import text as ot
        cfg = ot.config_from_state_dict(sd)
        hf_cfg = CLIPTextConfig(vocab_size=cfg.vocab, hidden_size=cfg.width, intermediate_size=4 * cfg.width,
                                projection_dim=cfg.out_dim, num_hidden_layers=cfg.layers, num_attention_heads=cfg.heads,
                                max_position_embeddings=cfg.context, hidden_act="quick_gelu", layer_norm_eps=1e-5,
                                bos_token_id=cfg.vocab - 2, eos_token_id=cfg.vocab - 1, pad_token_id=0)
        m = CLIPTextModelWithProjection(hf_cfg).eval()
        D = cfg.width
        t = {"text_model.embeddings.token_embedding.weight": sd["token_embedding.weight"],
             "text_model.embeddings.position_embedding.weight": sd["positional_embedding"],
             "text_model.final_layer_norm.weight": sd["ln_final.weight"],
             "text_model.final_layer_norm.bias": sd["ln_final.bias"],
             "text_projection.weight": sd["text_projection"].t().contiguous()}
        for i in range(cfg.layers):
            p = f"transformer.resblocks.{i}."
            h = f"text_model.encoder.layers.{i}."
            W, b = sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"]
            for j, n in enumerate(("q_proj", "k_proj", "v_proj")):
                t[h + f"self_attn.{n}.weight"] = W[j * D:(j + 1) * D]
                t[h + f"self_attn.{n}.bias"] = b[j * D:(j + 1) * D]
            t[h + "self_attn.out_proj.weight"] = sd[p + "attn.out_proj.weight"]
            t[h + "self_attn.out_proj.bias"] = sd[p + "attn.out_proj.bias"]
            t[h + "layer_norm1.weight"] = sd[p + "ln_1.weight"]
            t[h + "layer_norm1.bias"] = sd[p + "ln_1.bias"]
            t[h + "layer_norm2.weight"] = sd[p + "ln_2.weight"]
            t[h + "layer_norm2.bias"] = sd[p + "ln_2.bias"]
            t[h + "mlp.fc1.weight"] = sd[p + "mlp.c_fc.weight"]
            t[h + "mlp.fc1.bias"] = sd[p + "mlp.c_fc.bias"]
            t[h + "mlp.fc2.weight"] = sd[p + "mlp.c_proj.weight"]
            t[h + "mlp.fc2.bias"] = sd[p + "mlp.c_proj.bias"]
        missing, unexpected = m.load_state_dict(t, strict=False)
        assert not unexpected, unexpected
        assert all("position_ids" in k for k in missing), missing
        self.m = m

    @torch.no_grad()
    def encode_text(self, tokens):
        return self.m(input_ids=tokens).text_embeds


def make_text_goldens():
    """The reference's own text branch (shared/clip_ppo_utils.py:132-139: tokenize -> encode_text -> float ->
    normalize) run on CPU over a stub ``clip`` whose ``encode_text`` is the HF text tower and whose ``tokenize``
    returns seeded ids (the BPE merges file is not available offline)."""
    from oracle import text as ot
    sd = ot.random_state_dict(ot.TEXT_B32, seed=0)
    tower = _HFTextTower(sd)
    _install_stub_clip(tower)
    clip = sys.modules["clip"]
    tokens = ot.random_tokens(6, ot.TEXT_B32, seed=3)
    tokens[0, 2:] = 0
    tokens[0, 2] = ot.EOT                    # shortest caption: SOT, one word, EOT
    tokens[1, :] = torch.randint(1, 40000, (77,), generator=torch.Generator().manual_seed(9))
    tokens[1, 0], tokens[1, 76] = ot.SOT, ot.EOT            # longest: EOT in the last slot
    clip.tokenize = lambda texts, *a, **k: tokens[:len(texts)].clone()
    # by file path: this repository's own drop-in ``shared`` package would shadow the reference's namespace package
    import importlib.util
    spec = importlib.util.spec_from_file_location("reference_clip_ppo_utils", os.path.join(REF, "shared", "clip_ppo_utils.py"))
    ref_utils = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_utils)
    model = ref_utils.load_clip_model("ViT-B/32", device="cpu")
    emb = ref_utils.generate_clip_embeddings(ref_utils.AblationMode.NONE, model, "text", 6, "cpu",
                                             descriptions=[f"caption {i}" for i in range(6)])
    mine = ot.text_embeddings(sd, tokens)
    err, cos = (emb - mine).abs().max().item(), torch.sum(emb * mine, dim=-1).min().item()
    print(f"text embeddings: oracle vs reference+HF text tower  max-abs {err:.2e}  min-cos {cos:.8f}")
    assert err < 5e-5 and cos > 0.99999
    np.savez_compressed(os.path.join(OUT, "text_b32_seed0.npz"), weights_seed=0, tokens=tokens.numpy().astype(np.int32),
                        emb=emb.numpy())
    print("wrote text_b32_seed0.npz")


def _ref_lines(path, lo, hi):
    """Lines lo..hi (1-based, inclusive) of a reference source file, dedented for exec."""
    with open(os.path.join(REF, path)) as f:
        lines = f.read().splitlines()[lo - 1:hi]
    return textwrap.dedent("\n".join(lines))


def make_loss_goldens(ref_utils):
    g = torch.Generator().manual_seed(11)
    out = {}
    # cosine loss + its autograd gradients from the reference function
    z = torch.relu(torch.randn(24, 512, generator=g)).requires_grad_(True)
    c = torch.randn(24, 512, generator=g).requires_grad_(True)
    z.data[3].zero_()                                       # an all-zero latent row (post-ReLU dead unit)
    loss = ref_utils.compute_cosine_embedding_loss(z, c)
    loss.backward()
    out.update(cos_z=z.detach().numpy(), cos_c=c.detach().numpy(), cos_loss=loss.item(),
               cos_gz=z.grad.numpy(), cos_gc=c.grad.numpy())
    out["warmup_16"] = np.array([ref_utils.get_clip_lambda_with_warmup(1e-5, i, 16) for i in range(16)])
    out["warmup_100"] = np.array([ref_utils.get_clip_lambda_with_warmup(3e-4, i, 100) for i in range(100)])

    script = "minigrid_experiments/clip_ppo/clip_ppo_minigrid.py"
    args = types.SimpleNamespace(num_steps=128, gamma=0.99, gae_lambda=0.95, clip_coef=0.1, norm_adv=True,
                                 clip_vloss=True, ent_coef=0.01, vf_coef=0.5)
    T, E = 128, 8
    ns = dict(torch=torch, args=args, device="cpu",
              rewards=(torch.rand(T, E, generator=g) < 0.05).float(),
              dones=(torch.rand(T, E, generator=g) < 0.03).float(),
              values=torch.randn(T, E, generator=g),
              next_value=torch.randn(1, E, generator=g),
              next_done=(torch.rand(E, generator=g) < 0.5).float())
    exec(_ref_lines(script, 439, 450), ns)                   # GAE statements
    for k in ("rewards", "dones", "values", "next_value", "next_done", "advantages", "returns"):
        out["gae_" + k] = ns[k].numpy()

    mb = 256
    ns = dict(torch=torch, args=args, clipfracs=[], mb_inds=torch.arange(mb),
              newlogprob=(-1.5 + 0.3 * torch.randn(mb, generator=g)).requires_grad_(True),
              entropy=(1.0 + 0.2 * torch.rand(mb, generator=g)).requires_grad_(True),
              newvalue=torch.randn(mb, 1, generator=g).requires_grad_(True),
              b_advantages=torch.randn(mb, generator=g), b_returns=torch.randn(mb, generator=g),
              b_values=torch.randn(mb, generator=g))
    ns["b_logprobs"] = ns["newlogprob"].detach() + 0.15 * torch.randn(mb, generator=g)
    leaves = {k: ns[k] for k in ("newlogprob", "entropy", "newvalue")}   # the exec rebinds newvalue
    exec(_ref_lines(script, 498, 531), ns)                   # ratio ... entropy_loss
    ns["current_clip_lambda"] = 1e-5
    ns["clip_loss"] = torch.tensor(0.731)
    exec(_ref_lines(script, 559, 559), ns)                   # total loss
    ns["loss"].backward()
    for k in ("newlogprob", "entropy", "newvalue"):
        out["ppo_" + k] = leaves[k].detach().numpy()
        out["ppo_g_" + k] = leaves[k].grad.numpy()
    for k in ("b_logprobs", "b_advantages", "b_returns", "b_values"):
        out["ppo_" + k] = ns[k].numpy()
    for k in ("loss", "pg_loss", "v_loss", "entropy_loss", "old_approx_kl", "approx_kl"):
        out["ppo_" + k] = ns[k].item()
    out["ppo_clipfrac"] = ns["clipfracs"][0]
    out["ppo_clip_lambda"], out["ppo_clip_loss"] = 1e-5, 0.731
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **out)
    print("wrote losses.npz")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    if sys.argv[1:] == ["text"]:             # the text-tower fixture alone (added after the others)
        make_text_goldens()
        sys.exit(0)
    make_disturb_goldens()
    ref_utils = make_vit_goldens()
    make_loss_goldens(ref_utils)
    make_text_goldens()
