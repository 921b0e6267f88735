"""Byte-level BPE tokenizer of openai/CLIP (``clip/simple_tokenizer.py`` + ``clip.tokenize``; reference call site
shared/clip_ppo_utils.py:136), restated from its published algorithm.

The algorithm is here; the DATA is not: the merge list (``bpe_simple_vocab_16e6.txt.gz``, 1.3 MB) ships inside the openai
package and cannot be reproduced offline.  Point ``CLIPPPO_BPE_PATH`` at that file (gzip or plain text: a version header
line, then one merge ``a b`` per line) - or at a Hugging Face ``merges.txt`` of a CLIP checkpoint, which is the same list.
With the real file the ids equal ``clip.tokenize``'s; tests/test_tokenizer.py pins the algorithm against transformers'
independent ``CLIPTokenizer`` on a synthetic merge list.

    vocabulary = 256 byte symbols + the same 256 with '</w>' + one entry per merge (the first 48 894) + <|startoftext|>, <|endoftext|>
    text       -> (ftfy if installed) -> html.unescape twice -> collapse whitespace -> lower()
               -> regex pre-tokens -> bytes -> byte symbols -> BPE merges by rank -> ids
    tokenize   = [SOT] + ids + [EOT], zero-padded to context_length (77); too long -> RuntimeError unless truncate=True
"""
from __future__ import annotations

import gzip
import html
import os
from functools import lru_cache
from typing import Dict, Iterable, List, Tuple, Union

import torch

try:
    import regex as re
    _PAT = re.compile(r"""<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+""", re.IGNORECASE)
except ImportError:                                   # pragma: no cover - `regex` is part of every supported image
    import re
    _PAT = re.compile(r"""<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[^\W\d_]+|\d|[^\s\w]+""", re.IGNORECASE)

N_MERGES = 49152 - 256 - 2        # upstream keeps lines [1 : 48895) of the file


@lru_cache()
def bytes_to_unicode() -> Dict[int, str]:
    """The reversible byte -> printable-unicode map of GPT-2 / CLIP: printable latin-1 bytes map to themselves, the other 68
    to code points from 256 up, so no byte becomes whitespace or a control character inside a BPE symbol."""
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return dict(zip(bs, (chr(c) for c in cs)))


def _pairs(word: Tuple[str, ...]):
    return set(zip(word[:-1], word[1:]))


def _clean(text: str) -> str:
    try:
        import ftfy
        text = ftfy.fix_text(text)
    except ImportError:
        pass
    text = html.unescape(html.unescape(text)).strip()
    return " ".join(text.split()).strip()


class SimpleTokenizer:
    def __init__(self, bpe_path: str):
        opener = gzip.open if bpe_path.endswith(".gz") else open
        with opener(bpe_path, "rt", encoding="utf-8") as f:
            lines = f.read().split("\n")
        merges = [tuple(l.split()) for l in lines[1:N_MERGES + 1] if l.strip()]
        vocab = list(bytes_to_unicode().values())
        vocab = vocab + [v + "</w>" for v in vocab]
        vocab += ["".join(m) for m in merges]
        vocab += ["<|startoftext|>", "<|endoftext|>"]
        self.encoder = {s: i for i, s in enumerate(vocab)}
        self.decoder = {i: s for s, i in self.encoder.items()}
        self.bpe_ranks = {m: i for i, m in enumerate(merges)}
        self.byte_encoder = bytes_to_unicode()
        self.byte_decoder = {v: k for k, v in self.byte_encoder.items()}
        self.cache = {"<|startoftext|>": "<|startoftext|>", "<|endoftext|>": "<|endoftext|>"}

    def bpe(self, token: str) -> str:
        if token in self.cache:
            return self.cache[token]
        word = tuple(token[:-1]) + (token[-1] + "</w>",)
        pairs = _pairs(word)
        if not pairs:
            return token + "</w>"
        while True:
            best = min(pairs, key=lambda p: self.bpe_ranks.get(p, float("inf")))
            if best not in self.bpe_ranks:
                break
            a, b = best
            out, i = [], 0
            while i < len(word):
                if i < len(word) - 1 and word[i] == a and word[i + 1] == b:
                    out.append(a + b)
                    i += 2
                else:
                    out.append(word[i])
                    i += 1
            word = tuple(out)
            if len(word) == 1:
                break
            pairs = _pairs(word)
        res = " ".join(word)
        self.cache[token] = res
        return res

    def encode(self, text: str) -> List[int]:
        ids: List[int] = []
        for tok in _PAT.findall(_clean(text).lower()):
            sym = "".join(self.byte_encoder[b] for b in tok.encode("utf-8"))
            ids.extend(self.encoder[p] for p in self.bpe(sym).split(" "))
        return ids

    def decode(self, ids: Iterable[int]) -> str:
        text = "".join(self.decoder[int(i)] for i in ids)
        return bytearray(self.byte_decoder[c] for c in text).decode("utf-8", errors="replace").replace("</w>", " ")


_tokenizer = None


def default_tokenizer() -> SimpleTokenizer:
    """The tokenizer over ``CLIPPPO_BPE_PATH`` (cached).  Raises with the remedy when the merge list is not available."""
    global _tokenizer
    path = os.environ.get("CLIPPPO_BPE_PATH", "")
    if _tokenizer is None or getattr(_tokenizer, "_path", None) != path:
        if not path or not os.path.exists(path):
            raise FileNotFoundError(
                "clip_compat.tokenize needs CLIP's BPE merge list: set CLIPPPO_BPE_PATH to openai/CLIP's "
                "bpe_simple_vocab_16e6.txt.gz (or a CLIP checkpoint's merges.txt), install the openai `clip` package, or pass "
                "pre-tokenised [N, 77] ids to generate_clip_embeddings / encode_text")
        _tokenizer = SimpleTokenizer(path)
        _tokenizer._path = path
    return _tokenizer


def tokenize(texts: Union[str, List[str]], context_length: int = 77, truncate: bool = False, tokenizer: SimpleTokenizer = None) -> torch.Tensor:
    """``clip.tokenize``: [N, context_length] int32 ids, SOT ... EOT, zero padded."""
    if isinstance(texts, str):
        texts = [texts]
    tk = tokenizer or default_tokenizer()
    sot, eot = tk.encoder["<|startoftext|>"], tk.encoder["<|endoftext|>"]
    out = torch.zeros(len(texts), context_length, dtype=torch.int32)
    for i, t in enumerate(texts):
        ids = [sot] + tk.encode(t) + [eot]
        if len(ids) > context_length:
            if not truncate:
                raise RuntimeError(f"Input {t} is too long for context length {context_length}")
            ids = ids[:context_length]
            ids[-1] = eot
        out[i, :len(ids)] = torch.tensor(ids, dtype=torch.int32)
    return out
