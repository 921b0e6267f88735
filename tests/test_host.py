"""Host-side logic of the drop-in surface that needs no GPU: parameter table, constructor
contract, CPU-generator draw order, warm-up / gating scalars, loud failure without CUDA."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import disturb as od

from conftest import GOLDEN


def test_severity_table_matches_reference_rows():
    from shared.disturbance_types import DisturbanceSeverity, SEVERITY_CONFIGS
    assert [s.value for s in DisturbanceSeverity] == ["NONE", "MILD", "MODERATE", "HARD", "SEVERE"]
    assert DisturbanceSeverity.NONE not in SEVERITY_CONFIGS
    for name, row in od.SEVERITY_TABLE.items():
        mine = SEVERITY_CONFIGS[DisturbanceSeverity[name]]
        assert mine == dict(gaussian_noise_sigma=row["noise_sigma"], gaussian_blur_sigma=row["blur_sigma"],
                            contrast_range=row["contrast"], cutout_ratio=row["cutout"])


def test_wrapper_constructor_contract():
    from shared.disturbances_gpu import DisturbanceWrapperGPU
    from shared.disturbance_types import DisturbanceSeverity
    w = DisturbanceWrapperGPU(device="cpu", severity=DisturbanceSeverity.HARD)
    assert (w.gaussian_noise_sigma, w.gaussian_blur_sigma, w.contrast_range, w.cutout_ratio) == (0.13, 2.1, (0.69, 1.31), 0.18)
    assert w.device == torch.device("cpu")
    assert w.blur_transform.kernel_size == (5, 5) and w.noise_transform.sigma == 0.13
    with pytest.raises(ValueError):
        DisturbanceWrapperGPU(device="cpu", severity=None, gaussian_noise_sigma=0.1)
    c = DisturbanceWrapperGPU(device="cpu", severity=None, gaussian_noise_sigma=0.1, gaussian_blur_sigma=3.0,
                              contrast_range=(0.5, 1.5), cutout_ratio=0.2)
    assert c._kernel_size == 7
    assert DisturbanceWrapperGPU().gaussian_noise_sigma == 0.08          # default severity is MILD


def test_seed_is_global_like_the_reference():
    from shared.disturbances_gpu import DisturbanceWrapperGPU
    DisturbanceWrapperGPU(device="cpu", seed=123)
    a = torch.rand(3)
    torch.manual_seed(123)
    assert torch.equal(a, torch.rand(3))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "disturb_0[0-7]*.npz"))))
def test_cpu_generator_draw_order_matches_golden(path):
    """After the device randn_like, the wrapper's CPU-generator draws (randperm, uniform, uniform,
    randint, randint) must yield the reference's contrast factor, taps and cutout window."""
    from shared.disturbances_gpu import DisturbanceWrapperGPU
    from shared.disturbance_types import DisturbanceSeverity
    g = np.load(path)
    B, H, W, C = g["u8"].shape
    w = DisturbanceWrapperGPU(device="cpu", seed=int(g["seed"]), severity=DisturbanceSeverity[str(g["severity"])])
    torch.randn(B, C, H, W)                   # stands for randn_like(obs) on a CPU tensor
    c = w._draw_contrast()
    taps = w._draw_blur_taps()
    window = w._draw_cutout(H, W, None)
    assert c == float(g["c"])
    assert np.array_equal(np.array(taps, dtype=np.float32), g["k1d"])
    assert window == (int(g["sh"]), int(g["sw"]), int(g["ph"]), int(g["pw"]))


def test_no_cpu_fallback():
    from shared.disturbances_gpu import DisturbanceWrapperGPU, create_disturbance_wrapper
    w = DisturbanceWrapperGPU(device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        w.apply_disturbances(torch.rand(1, 3, 8, 8))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            create_disturbance_wrapper(use_gpu=True)
    import shared.clip_ppo_utils as U
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        U.compute_cosine_embedding_loss(torch.rand(2, 512), torch.rand(2, 512))


def test_clip_utils_scalars_and_errors():
    import shared.clip_ppo_utils as U
    g = np.load(os.path.join(GOLDEN, "losses.npz"))
    assert [U.get_clip_lambda_with_warmup(1e-5, i, 16) for i in range(16)] == list(g["warmup_16"])
    assert [U.get_clip_lambda_with_warmup(3e-4, i, 100) for i in range(100)] == list(g["warmup_100"])
    assert U.CLIP_LOSS_FREQUENCY == 4
    assert U.should_compute_clip_loss(U.AblationMode.NONE, 1e-5)
    assert not U.should_compute_clip_loss(U.AblationMode.NONE, 0.0)
    assert not U.should_compute_clip_loss(U.AblationMode.FROZEN_CLIP, 1e-5)
    assert U.should_compute_clip_loss(U.AblationMode.RANDOM_ENCODER, 1e-5)
    with pytest.raises(ValueError, match="Dimension mismatch"):
        U.compute_cosine_embedding_loss(torch.zeros(2, 512), torch.zeros(2, 768))
    with pytest.raises(ValueError, match="images required"):
        U.generate_clip_embeddings(U.AblationMode.NONE, None, "image", 2, "cpu")
    with pytest.raises(ValueError, match="descriptions required"):
        U.generate_clip_embeddings(U.AblationMode.NONE, None, "text", 2, "cpu")
    with pytest.raises(ValueError, match="Invalid modality"):
        U.generate_clip_embeddings(U.AblationMode.NONE, None, "audio", 2, "cpu")
    e = U.generate_clip_embeddings(U.AblationMode.RANDOM_ENCODER, None, "image", 5, "cpu")
    assert e.shape == (5, 512) and torch.allclose(e.norm(dim=-1), torch.ones(5), atol=1e-6)
    cfg = U.ClipPPOConfig()
    assert (cfg.clip_lambda, cfg.clip_model, cfg.clip_modality, cfg.disturbance_severity) == (1e-5, "ViT-B/32", "text", "MODERATE")
    assert torch.allclose(U._CLIP_MEAN, torch.tensor([0.48145466, 0.4578275, 0.40821073]))
    import clip
    assert isinstance(clip.model.VisionTransformer, type) and isinstance(clip.model.CLIP, type)


def test_compat_state_dict_matches_oracle_weights():
    """The shim's seeded random weights are the oracle's (same generator stream), so GPU-vs-oracle
    parity tests and the golden embeddings refer to the same tower."""
    from clip_ppo_b200.clip_compat.model import random_visual_state_dict
    from oracle import vit as ov
    a = random_visual_state_dict("ViT-B/32", 0)
    b = ov.random_state_dict(ov.VIT_B32, 0)
    assert a.keys() == b.keys()
    for k in ("visual.conv1.weight", "visual.transformer.resblocks.11.mlp.c_proj.weight", "visual.proj"):
        assert torch.equal(a[k], b[k])


def test_compat_text_state_dict_matches_oracle_weights():
    """clip_compat's seeded text-tower weights are the oracle's (same generator, same order), and they travel through
    the compat CLIP module's state dict under upstream's top-level key names."""
    from oracle import text as ot
    from clip_ppo_b200 import clip_compat
    a = clip_compat.random_text_state_dict("ViT-B/32", 0)
    b = ot.random_state_dict(ot.TEXT_B32, 0)
    assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)
    model, _ = clip_compat.load("ViT-B/32", device="cpu")
    sd = model.state_dict()
    assert torch.equal(sd["text_projection"], b["text_projection"]) and "visual.proj" in sd
    assert torch.equal(sd["transformer.resblocks.11.mlp.c_proj.weight"], b["transformer.resblocks.11.mlp.c_proj.weight"])
    other, _ = clip_compat.load("ViT-B/32", device="cpu", seed=1)
    assert not torch.equal(other.state_dict()["ln_final.bias"], sd["ln_final.bias"])
    other.load_state_dict(sd)
    assert torch.equal(other.state_dict()["ln_final.bias"], sd["ln_final.bias"])
    import os
    if not os.environ.get("CLIPPPO_BPE_PATH"):             # the merge list is data of the openai package: absent -> a clear error
        with pytest.raises(FileNotFoundError):
            clip_compat.tokenize(["a red door"])


def test_single_forward_latents_equal_the_scripts_second_forward():
    """a20: rollout.action_value_and_latents == get_action_and_value + get_latent_representation of the reference Agent
    (clip_ppo_minigrid.py:213-271 restated), with one encoder forward."""
    import torch.nn as nn
    from torch.distributions.categorical import Categorical
    from clip_ppo_b200 import rollout

    class Agent(nn.Module):
        def __init__(self):
            super().__init__()
            self.calls = 0
            self.network = nn.Sequential(nn.Conv2d(3, 32, 8, stride=4), nn.ReLU(), nn.Conv2d(32, 64, 4, stride=2), nn.ReLU(),
                                         nn.Conv2d(64, 64, 3, stride=1), nn.ReLU(), nn.Flatten(), nn.Linear(64 * 7 * 7, 512), nn.ReLU())
            self.actor, self.critic = nn.Linear(512, 7), nn.Linear(512, 1)

        def _pre(self, x):
            return x.permute(0, 3, 1, 2).contiguous() / 255.0

        def _get_features(self, x):
            self.calls += 1
            return self.network(x)

        def get_action_and_value(self, x, action=None):
            hidden = self._get_features(self._pre(x))
            probs = Categorical(logits=self.actor(hidden))
            return action, probs.log_prob(action), probs.entropy(), self.critic(hidden)

        def get_latent_representation(self, x):
            return self._get_features(self._pre(x)).detach()

    torch.manual_seed(0)
    agent = Agent()
    obs = torch.randint(0, 256, (5, 84, 84, 3)).float()
    act = torch.randint(0, 7, (5,))
    _, lp, ent, val = agent.get_action_and_value(obs, act)
    lat = agent.get_latent_representation(obs)
    agent.calls = 0
    a2, lp2, ent2, val2, lat2 = rollout.action_value_and_latents(agent, obs, act)
    assert agent.calls == 1
    assert torch.equal(lp, lp2) and torch.equal(ent, ent2) and torch.equal(val, val2) and torch.equal(lat, lat2)
    assert not lat2.requires_grad and lp2.requires_grad


def test_dropin_import_resolution_with_the_scripts_sys_path_hacks(tmp_path):
    """PYTHONPATH=<this repo> against a reference-shaped tree: the scripts put their root and their `shared/`
    directory at the FRONT of sys.path (clip_ppo_minigrid.py:22-29) and import `clip_ppo_utils` as a top-level
    module.  This repository's modules must win for the three names it replaces, the reference's other `shared.*`
    modules (the cv2 twin, checkpoint utils) must stay importable."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = tmp_path / "ref"
    (ref / "shared").mkdir(parents=True)                       # namespace package: no __init__.py, like the reference
    (ref / "exp" / "clip").mkdir(parents=True)
    (ref / "shared" / "clip_ppo_utils.py").write_text("WHO = 'reference'\n")
    (ref / "shared" / "disturbances_gpu.py").write_text("WHO = 'reference'\n")
    (ref / "shared" / "disturbances.py").write_text("from shared.disturbance_types import DisturbanceSeverity\nWHO = 'cv2 twin'\n")
    (ref / "shared" / "checkpoint_utils.py").write_text("WHO = 'reference'\n")
    (ref / "exp" / "clip" / "script.py").write_text(
        "import os, sys\n"
        "sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))\n"
        "sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', '..', 'shared'))\n"
        "import clip_ppo_utils\n"
        "from shared.disturbances_gpu import DisturbanceWrapperGPU\n"
        "from shared.disturbance_types import DisturbanceSeverity\n"
        "from shared import disturbances\n"
        "import shared.checkpoint_utils as ck\n"
        "import shared.clip_ppo_utils as U2\n"
        "print(clip_ppo_utils.__file__); print(sys.modules['shared.disturbances_gpu'].__file__)\n"
        "print(disturbances.WHO, ck.WHO, clip_ppo_utils is U2, clip_ppo_utils.__name__, hasattr(clip_ppo_utils, 'generate_clip_embeddings'))\n")
    script = str(ref / "exp" / "clip" / "script.py")
    # route 1: the opt-in start-up hook (CLIPPPO_DROPIN=1 with the repository on PYTHONPATH), script untouched
    env = dict(os.environ, PYTHONPATH=root, CLIPPPO_DROPIN="1")
    out = subprocess.run([sys.executable, script], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.strip().splitlines()
    assert lines[0] == os.path.join(root, "shared", "clip_ppo_utils.py")
    assert lines[1] == os.path.join(root, "shared", "disturbances_gpu.py")
    assert lines[2] == "cv2 twin reference True shared.clip_ppo_utils True"
    # route 2 (the documented default): two explicit lines in front of the script, no hook
    env = dict(os.environ, PYTHONPATH=root)
    env.pop("CLIPPPO_DROPIN", None)
    code = f"from clip_ppo_b200 import dropin; dropin.install(); import runpy; runpy.run_path({script!r}, run_name='__main__')"
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip().splitlines()[0] == os.path.join(root, "shared", "clip_ppo_utils.py")
    # without either, the repository on the path changes nothing about top-level `clip_ppo_utils`
    out = subprocess.run([sys.executable, script], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().splitlines()[0] == str(ref / "exp" / "clip" / ".." / ".." / "shared" / "clip_ppo_utils.py")


def test_clip_weight_sources(tmp_path, monkeypatch):
    """clip_compat.load: a checkpoint named by CLIPPPO_CLIP_WEIGHTS is what the model holds (round trip through torch.save);
    with no weights at all it RAISES unless random initialisation was asked for - never a silent untrained tower."""
    import warnings
    from clip_ppo_b200 import clip_compat
    sd = clip_compat.random_visual_state_dict("ViT-B/32", 3)
    sd.update(clip_compat.random_text_state_dict("ViT-B/32", 3))
    path = tmp_path / "clip_b32.pt"
    torch.save(sd, path)
    monkeypatch.setenv("CLIPPPO_CLIP_WEIGHTS", str(path))
    monkeypatch.delenv("CLIPPPO_ALLOW_RANDOM_WEIGHTS", raising=False)
    with warnings.catch_warnings():
        warnings.simplefilter("error")                      # loading real weights must not warn
        model, _ = clip_compat.load("ViT-B/32", device="cpu")
    assert model.weight_source == f"checkpoint {path}"
    got = model.state_dict()
    for k in ("visual.conv1.weight", "visual.transformer.resblocks.5.attn.in_proj_weight", "visual.proj", "text_projection",
              "token_embedding.weight"):
        assert torch.equal(got[k], sd[k]), k
    # a pickled module works too (what `torch.save(model)` leaves behind)
    torch.save(model, tmp_path / "module.pt")
    monkeypatch.setenv("CLIPPPO_CLIP_WEIGHTS", str(tmp_path / "module.pt"))
    again, _ = clip_compat.load("ViT-B/32", device="cpu")
    assert torch.equal(again.state_dict()["visual.proj"], sd["visual.proj"])
    # nothing available
    monkeypatch.delenv("CLIPPPO_CLIP_WEIGHTS")
    with pytest.raises(RuntimeError, match="no CLIP weights"):
        clip_compat.load("ViT-B/32", device="cpu")
    with pytest.warns(UserWarning, match="RANDOM"):
        m, _ = clip_compat.load("ViT-B/32", device="cpu", random_init=True)
    assert m.weight_source == "random(seed=0)"
    monkeypatch.setenv("CLIPPPO_ALLOW_RANDOM_WEIGHTS", "1")
    with pytest.warns(UserWarning):
        clip_compat.load("ViT-B/32", device="cpu", seed=2)
