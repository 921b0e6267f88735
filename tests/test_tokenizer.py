"""clip_compat's BPE tokenizer (the host half of §8f-3) against transformers' independent CLIPTokenizer on a synthetic merge
list - the real list is data of the openai package and exists on no box.  CPU only."""
import json
import os

import pytest
import torch

TEXTS = ["a red door to the left of the agent", "The agent is facing a locked yellow door; it carries a key!",
         "agent navigating grid environment", "ball at (3, 4), wall ahead... score: 120", "  multiple   spaces\tand\nnewlines ",
         "don't they're we'll I'm it's", "café naïve 東京 🙂", ""]


def _synthetic_merges(n=300):
    """A plausible merge list learnt greedily from a tiny corpus (the algorithm under test is the encoder, not the learner)."""
    from clip_ppo_b200.clip_compat.tokenizer import bytes_to_unicode
    b2u = bytes_to_unicode()
    corpus = " ".join(TEXTS * 3 + ["the agent navigates the grid to the goal and opens the door with the key"] * 5).lower().split()
    words = [tuple(b2u[b] for b in w.encode("utf-8"))[:-1] + (b2u[w.encode("utf-8")[-1]] + "</w>",) for w in corpus if w]
    merges = []
    for _ in range(n):
        counts = {}
        for w in words:
            for p in zip(w[:-1], w[1:]):
                counts[p] = counts.get(p, 0) + 1
        if not counts:
            break
        best = max(sorted(counts), key=lambda p: counts[p])
        merges.append(best)
        new = []
        for w in words:
            out, i = [], 0
            while i < len(w):
                if i < len(w) - 1 and (w[i], w[i + 1]) == best:
                    out.append(w[i] + w[i + 1]); i += 2
                else:
                    out.append(w[i]); i += 1
            new.append(tuple(out))
        words = new
    return merges


@pytest.fixture(scope="module")
def merge_file(tmp_path_factory):
    d = tmp_path_factory.mktemp("bpe")
    merges = _synthetic_merges()
    path = os.path.join(d, "merges.txt")
    with open(path, "w", encoding="utf-8") as f:
        f.write("#version: 0.2\n" + "\n".join(" ".join(m) for m in merges) + "\n")
    return path, merges


def test_tokenize_matches_hf_clip_tokenizer(merge_file, monkeypatch):
    path, merges = merge_file
    from clip_ppo_b200 import clip_compat
    from clip_ppo_b200.clip_compat import tokenizer as T
    monkeypatch.setenv("CLIPPPO_BPE_PATH", path)
    tk = T.SimpleTokenizer(path)
    # the same vocabulary for transformers' slow tokenizer (pure Python, an implementation independent of ours)
    from transformers import CLIPTokenizer
    vocab_path = os.path.join(os.path.dirname(path), "vocab.json")
    with open(vocab_path, "w", encoding="utf-8") as f:
        json.dump(tk.encoder, f, ensure_ascii=False)
    hf = CLIPTokenizer(vocab_path, path)
    sot, eot = tk.encoder["<|startoftext|>"], tk.encoder["<|endoftext|>"]
    assert (sot, eot) == (len(tk.encoder) - 2, len(tk.encoder) - 1)
    for t in TEXTS[:5]:                                   # plain ASCII: both normalisers agree (HF has no ftfy here either)
        mine = [sot] + tk.encode(t) + [eot]
        assert mine == hf(t)["input_ids"], t
    ids = clip_compat.tokenize(TEXTS)
    assert ids.shape == (len(TEXTS), 77) and ids.dtype == torch.int32
    for row, t in zip(ids, TEXTS):
        n = len(tk.encode(t)) + 2
        assert row[0] == sot and row[n - 1] == eot and int(row[n:].abs().sum()) == 0
        assert int(row.argmax()) == n - 1                 # EOT is the largest id: what encode_text pools on
    # round trip through the byte-level symbols, unicode included
    assert tk.decode(tk.encode("café naïve 東京 🙂")).strip() == "café naïve 東京 🙂"


def test_tokenize_length_and_errors(merge_file, monkeypatch):
    path, _ = merge_file
    from clip_ppo_b200 import clip_compat
    monkeypatch.setenv("CLIPPPO_BPE_PATH", path)
    long_text = "zq " * 100                               # unseen symbols: ~3 tokens per word
    with pytest.raises(RuntimeError):
        clip_compat.tokenize([long_text])
    ids = clip_compat.tokenize([long_text], truncate=True)
    assert ids.shape == (1, 77) and int(ids[0, -1]) == int(ids.max())
    monkeypatch.setenv("CLIPPPO_BPE_PATH", "/nonexistent/bpe.txt.gz")
    with pytest.raises(FileNotFoundError):
        clip_compat.tokenize(["a door"])
    assert not clip_compat.tokenizer_available()
