"""Parity of the frozen CLIP text tower (through the C ABI) with the reference golden and the oracle.
Tolerance: cosine >= 0.999 against the fp32 tower (BASELINE.json north_star, bf16 vs fp32)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import text as ot

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _stream():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("n,T,H", [(3, 77, 8), (2, 77, 12), (5, 16, 8), (2, 64, 8), (2, 65, 16), (1, 130, 8), (2, 257, 16), (4, 7, 8)])
def test_causal_attention_vs_torch(native, n, T, H):
    """clipppo_attention_causal_bf16 == softmax(QK^T / 8 + triu(-inf, 1)) V on the packed [n*T, 3*H*64] bf16 QKV."""
    from clip_ppo_b200 import _native as Nn
    D = H * 64
    gen = torch.Generator(device="cuda").manual_seed(n * 1000 + T)
    qkv = (torch.randn(n * T, 3 * D, device="cuda", generator=gen) * 1.5).bfloat16()
    out = torch.empty(n * T, D, device="cuda", dtype=torch.bfloat16)
    Nn.check(native.clipppo_attention_causal_bf16(qkv.data_ptr(), n, T, H, 64, out.data_ptr(), _stream()))
    torch.cuda.synchronize()
    q, k, v = qkv.float().reshape(n, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    mask = torch.full((T, T), float("-inf"), device="cuda").triu_(1)
    s = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(64) + mask, dim=-1)
    ref = (s @ v).permute(0, 2, 1, 3).reshape(n * T, D)
    err = (out.float() - ref).abs().max().item()
    assert err <= 3e-2, err
    # row 0 of every sequence sees only key 0: its output is V[0] exactly
    v0 = qkv.reshape(n, T, 3, D)[:, 0, 2]
    assert torch.equal(out.reshape(n, T, D)[:, 0], v0)


def _engine(seed=0):
    from clip_ppo_b200.text import TextEngine
    return TextEngine(ot.random_state_dict(ot.TEXT_B32, seed), device="cuda")


def test_text_embeddings_match_reference_golden(native):
    g = np.load(os.path.join(GOLDEN, "text_b32_seed0.npz"))
    eng = _engine(int(g["weights_seed"]))
    tokens = torch.from_numpy(g["tokens"]).cuda()
    emb = eng.encode(tokens, l2norm=True)
    ref = torch.from_numpy(g["emb"])
    assert emb.shape == ref.shape and emb.dtype == torch.float32
    cos = torch.sum(emb.cpu() * ref, dim=-1)
    assert cos.min().item() >= 0.999, cos
    assert torch.allclose(emb.norm(dim=-1).cpu(), torch.ones(len(ref)), atol=1e-5)


@pytest.mark.parametrize("n", [1, 9, 70])
def test_text_embeddings_vs_oracle(native, n):
    eng = _engine(0)
    sd = ot.random_state_dict(ot.TEXT_B32, 0)
    tokens = ot.random_tokens(n, ot.TEXT_B32, seed=n)
    raw = eng.encode(tokens.cuda(), l2norm=False).cpu()
    idx = torch.linspace(0, n - 1, min(n, 6)).long()
    ref = ot.text_tower(sd, tokens[idx])
    cos = torch.nn.functional.cosine_similarity(raw[idx], ref, dim=-1)
    assert cos.min().item() >= 0.999, cos
    assert ((raw[idx].norm(dim=-1) / ref.norm(dim=-1)) - 1).abs().max().item() <= 2e-2
    # causal mask: tokens after the EOT do not matter; batch-size invariance is bitwise
    t2 = tokens.clone()
    for i in range(n):
        e = int(t2[i].argmax())
        t2[i, e + 1:] = 7
    assert torch.equal(eng.encode(t2.cuda(), l2norm=False).cpu(), raw)
    assert torch.equal(eng.encode(tokens[idx[-1:]].cuda(), l2norm=False).cpu()[0], raw[idx[-1]])
    # int64 ids (older clip.tokenize) are accepted
    assert torch.equal(eng.encode(tokens.cuda().long(), l2norm=False).cpu(), raw)


def test_generate_clip_embeddings_text_modality(native):
    """reference shared/clip_ppo_utils.py:132-139 through the drop-in module: pre-tokenised ids in, unit rows out;
    the reference's error behaviour is kept."""
    import shared.clip_ppo_utils as U
    model = U.load_clip_model("ViT-B/32", "cuda")
    tokens = ot.random_tokens(4, ot.TEXT_B32, seed=2)
    e = U.generate_clip_embeddings(U.AblationMode.NONE, model, "text", 4, "cuda", descriptions=tokens)
    ref = ot.text_embeddings(ot.random_state_dict(ot.TEXT_B32, 0), tokens)
    assert e.shape == (4, 512) and e.dtype == torch.float32
    assert torch.sum(e.cpu() * ref, dim=-1).min().item() >= 0.999
    assert torch.allclose(e.norm(dim=-1).cpu(), torch.ones(4), atol=1e-5)
    raw = model.encode_text(tokens.cuda())
    assert torch.allclose(torch.nn.functional.normalize(raw, dim=-1), e, atol=1e-6)
    with pytest.raises(ValueError):
        U.generate_clip_embeddings(U.AblationMode.NONE, model, "text", 4, "cuda")
    import os
    if not os.environ.get("CLIPPPO_BPE_PATH"):                     # the BPE merge list is data of the openai package
        with pytest.raises(FileNotFoundError):
            U.generate_clip_embeddings(U.AblationMode.NONE, model, "text", 1, "cuda", descriptions=["a red door"])
    eng = model.text_engine()
    with pytest.raises(ValueError):
        eng.encode(tokens[:, :50].cuda())
    with pytest.raises(IndexError):
        eng.encode(torch.full((1, 77), 50000, device="cuda"))
    with pytest.raises(TypeError):
        eng.encode(torch.zeros(1, 77, device="cuda"))


def test_repeated_descriptions_are_encoded_once(native, monkeypatch):
    """String descriptions: distinct strings are tokenised and encoded once, rows gathered - same values as
    encoding every row.  (The tokenizer is stubbed: the BPE merges file is not available offline.)"""
    import shared.clip_ppo_utils as U
    model = U.load_clip_model("ViT-B/32", "cuda")
    vocab = {"a red door": 0, "a key": 1, "an empty room": 2}
    table = ot.random_tokens(3, ot.TEXT_B32, seed=8)
    calls = []

    def fake_tokenize(texts, *a, **k):
        calls.append(list(texts))
        return table[[vocab[t] for t in texts]].clone()

    monkeypatch.setattr(U.clip, "tokenize", fake_tokenize, raising=False)
    descriptions = ["a key", "a red door", "a key", "an empty room", "a key", "a red door"] * 50
    e = U.generate_clip_embeddings(U.AblationMode.NONE, model, "text", len(descriptions), "cuda", descriptions=descriptions)
    assert calls == [["a key", "a red door", "an empty room"]]
    every_row = model.text_engine().encode(table[[vocab[t] for t in descriptions]].cuda(), l2norm=True)
    assert e.shape == (300, 512) and torch.equal(e, every_row)


def test_string_descriptions_end_to_end_with_a_merge_list(native, tmp_path, monkeypatch):
    """The MiniGrid default modality (clip_ppo_minigrid.py:44, :394-403): description STRINGS -> BPE ids -> text tower, with the
    tokenizer given a (synthetic) merge list through CLIPPPO_BPE_PATH.  Embeddings equal the oracle's on the same ids."""
    import shared.clip_ppo_utils as U
    from clip_ppo_b200 import clip_compat
    from clip_ppo_b200.clip_compat import tokenizer as T
    b2u = T.bytes_to_unicode()
    merges = [(b2u[ord("t")], b2u[ord("h")]), (b2u[ord("t")] + b2u[ord("h")], b2u[ord("e")] + "</w>"), (b2u[ord("d")], b2u[ord("o")]),
              (b2u[ord("o")], b2u[ord("r")] + "</w>"), (b2u[ord("a")], b2u[ord("g")]), (b2u[ord("e")], b2u[ord("n")])]
    path = tmp_path / "merges.txt"
    path.write_text("#version: 0.2\n" + "\n".join(" ".join(m) for m in merges) + "\n", encoding="utf-8")
    monkeypatch.setenv("CLIPPPO_BPE_PATH", str(path))
    if U.clip is not clip_compat:
        pytest.skip("the real openai clip package is installed: its own tokenizer is used")
    model = U.load_clip_model("ViT-B/32", "cuda")
    texts = ["the agent is in front of the red door", "the agent carries a key", "the agent is in front of the red door"]
    ids = clip_compat.tokenize(texts)
    assert ids.max().item() < 49408 and torch.equal(ids[0], ids[2])
    e = U.generate_clip_embeddings(U.AblationMode.NONE, model, "text", 3, "cuda", descriptions=texts)
    ref = ot.text_embeddings(ot.random_state_dict(ot.TEXT_B32, 0), ids)
    assert e.shape == (3, 512) and torch.sum(e.cpu() * ref, dim=-1).min().item() >= 0.999
    assert torch.equal(e[0], e[2])
