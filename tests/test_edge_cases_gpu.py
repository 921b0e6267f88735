"""Edge cases of the public calls (through the C ABI): empty batches, single frames, ragged batch sizes around the tile
boundaries of the kernels (the reference has no tests for these; the behaviour pinned here is the
reference's PyTorch behaviour on the same inputs - an empty batch gives an empty result of the right shape, a ragged batch
equals the same rows of a larger one)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _wrapper(sev="MODERATE"):
    from shared.disturbances_gpu import DisturbanceWrapperGPU
    from shared.disturbance_types import DisturbanceSeverity
    return DisturbanceWrapperGPU(device="cuda", severity=DisturbanceSeverity[sev])


@pytest.mark.parametrize("shape", [(0, 3, 84, 84), (0, 1, 84, 84), (0, 3, 224, 224)])
def test_disturbances_of_an_empty_batch(native, shape):
    w = _wrapper()
    out = w.apply_disturbances(torch.empty(shape, device="cuda"))
    assert out.shape == shape and out.dtype == torch.float32 and out.is_cuda
    u8 = torch.empty(shape, dtype=torch.uint8, device="cuda")
    assert w.apply_disturbances(u8).shape == shape


def test_embeddings_of_an_empty_batch(native):
    from clip_ppo_b200.clip_compat.model import random_visual_state_dict
    from clip_ppo_b200.vit import VitEngine
    eng = VitEngine(random_visual_state_dict("ViT-B/32", 0), device="cuda")
    for hw in (84, 224):
        out = eng.encode(torch.empty((0, 3, hw, hw), device="cuda"), pre_scale=1.0 / 255.0, l2norm=True)
        assert out.shape == (0, 512) and out.dtype == torch.float32


@pytest.mark.parametrize("n", [1, 2, 3, 127, 129, 257])
def test_ragged_batches_are_rows_of_a_larger_batch(native, n):
    """Odd image counts around the 128-row / image-pair tile boundaries: every row equals, bit for bit, the same image
    encoded inside a batch of 300 (the tower is batch-invariant by construction: fixed reduction orders, no split-K)."""
    from clip_ppo_b200.clip_compat.model import random_visual_state_dict
    from clip_ppo_b200.vit import VitEngine
    eng = VitEngine(random_visual_state_dict("ViT-B/32", 0), device="cuda")
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.rand(300, 3, 84, 84, device="cuda", generator=g) * 255.0
    ref = eng.encode(x, pre_scale=1.0 / 255.0, l2norm=True)
    out = eng.encode(x[:n], pre_scale=1.0 / 255.0, l2norm=True)
    assert torch.equal(out, ref[:n])


@pytest.mark.parametrize("B", [1, 2, 7, 37])
def test_disturbing_a_ragged_batch_equals_the_rows_of_a_larger_one(native, B):
    """Per-image work only (the contrast mean is per image, c / cutout are shared scalars): the first B frames of a batch
    of 64 give the same result alone with the same supplied randomness - up to the last bit of the per-image gray mean,
    whose summation order follows the number of stripes a frame is cut into (which depends on the batch size below 148
    frames: csrc/disturb.cu widen_for_small_batches)."""
    w = _wrapper("SEVERE")
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.rand(64, 3, 84, 84, device="cuda", generator=g)
    noise = torch.randn(64, 3, 84, 84, device="cuda", generator=g)
    ref = w.apply_disturbances(x, noise=noise, contrast_factor=1.23, cutout_start=(5, 9))
    out = w.apply_disturbances(x[:B], noise=noise[:B], contrast_factor=1.23, cutout_start=(5, 9))
    assert (out - ref[:B]).abs().max().item() <= 1e-6


def test_losses_on_the_smallest_inputs(native):
    import shared.clip_ppo_utils as U
    from clip_ppo_b200 import losses as Ls
    z = torch.rand(1, 512, device="cuda", requires_grad=True)
    c = torch.rand(1, 512, device="cuda")
    loss = U.compute_cosine_embedding_loss(z, c)
    ref = 1.0 - torch.nn.functional.cosine_similarity(z.detach(), c).mean()
    assert abs(loss.item() - ref.item()) <= 1e-6
    loss.backward()
    assert torch.isfinite(z.grad).all()
    with pytest.raises(ValueError):                                   # width mismatch: the reference's error (clip_ppo_utils.py:63)
        U.compute_cosine_embedding_loss(torch.rand(4, 512, device="cuda"), torch.rand(4, 256, device="cuda"))
    # GAE over a single step and a single environment
    r = torch.tensor([[1.0]], device="cuda"); d = torch.zeros(1, 1, device="cuda"); v = torch.tensor([[0.5]], device="cuda")
    adv, ret = Ls.gae(r, v, d, torch.tensor([[0.25]], device="cuda"), torch.zeros(1, device="cuda"), 0.99, 0.95)
    assert abs(adv.item() - (1.0 + 0.99 * 0.25 - 0.5)) <= 1e-7 and abs(ret.item() - (adv.item() + 0.5)) <= 1e-7


@pytest.mark.parametrize("name,n,hw", [("ViT-B/32", 1, 84), ("ViT-B/32", 37, 84), ("ViT-B/32", 300, 224), ("ViT-L/14", 9, 224)])
def test_cls_only_last_block_is_bitwise_the_full_pass(native, monkeypatch, name, n, hw):
    """CLIPPPO_VIT_CLS_LAST_BLOCK=1 (opt-in): the last block's out_proj / c_fc / c_proj on the class-token rows alone.
    VisionTransformer.forward reads x[:, 0, :] only, and a row of the GEMM does not depend on the other rows of its tile, so the
    embeddings are bitwise those of the default pass (every token through every block) - eager and graph replay."""
    from clip_ppo_b200.clip_compat.model import random_visual_state_dict
    from clip_ppo_b200.vit import VitEngine
    eng = VitEngine(random_visual_state_dict(name, 0), device="cuda")
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.rand(n, 3, hw, hw, device="cuda", generator=g) * 255.0
    monkeypatch.delenv("CLIPPPO_VIT_CLS_LAST_BLOCK", raising=False)
    full = eng.encode(x, pre_scale=1.0 / 255.0, l2norm=True)
    full_raw = eng.encode(x, pre_scale=1.0 / 255.0, l2norm=False)
    monkeypatch.setenv("CLIPPPO_VIT_CLS_LAST_BLOCK", "1")
    assert torch.equal(eng.encode(x, pre_scale=1.0 / 255.0, l2norm=True), full)
    assert torch.equal(eng.encode(x, pre_scale=1.0 / 255.0, l2norm=False), full_raw)
    if n <= 512:
        assert torch.equal(eng.encode_graphed(x, pre_scale=1.0 / 255.0, l2norm=True), full)
