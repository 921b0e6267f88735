"""Not a test (pytest does not collect it): text-tower throughput on one B200 - this repository's TextEngine next to the
PyTorch-eager fp16 version of the same tower (what ``clip_model.encode_text`` dispatches on a GPU: cuBLAS GEMMs,
fp32 LayerNorm, SDPA with the causal mask), on the MiniGrid batch (8192 descriptions per iteration,
clip_ppo_minigrid.py:459-470).  Seeded random weights, seeded token ids shaped like clip.tokenize output.

    python tests/text_eager_baseline.py [texts per call] [calls]

Lives under tests/ because only tests/, smoke() and bench.py's CPU legs may import oracle/.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F

from oracle import text as ot
from clip_ppo_b200.text import TextEngine

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
CALLS = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda", 0)
cfg = ot.TEXT_B32
sd32 = ot.random_state_dict(cfg, 0)
sd = {k: v.to(dev).half() for k, v in sd32.items()}
tokens = ot.random_tokens(N, cfg, seed=1).to(dev)
FLOPS = cfg.context * cfg.layers * (24 * cfg.width ** 2 + 4 * cfg.context * cfg.width) + 2 * cfg.width * cfg.out_dim


def ln(x, w, b):
    return F.layer_norm(x.float(), (x.shape[-1],), w.float(), b.float(), 1e-5).to(x.dtype)


@torch.no_grad()
def eager(tok):
    D, H = cfg.width, cfg.heads
    n, T = tok.shape
    X = sd["token_embedding.weight"][tok] + sd["positional_embedding"]
    for i in range(cfg.layers):
        p = f"transformer.resblocks.{i}."
        qkv = F.linear(ln(X, sd[p + "ln_1.weight"], sd[p + "ln_1.bias"]), sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"])
        q, k, v = (t.reshape(n, T, H, D // H).transpose(1, 2) for t in qkv.split(D, dim=-1))
        o = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(n, T, D)
        X = X + F.linear(o, sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])
        h = F.linear(ln(X, sd[p + "ln_2.weight"], sd[p + "ln_2.bias"]), sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"])
        X = X + F.linear(h * torch.sigmoid(1.702 * h), sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])
    X = ln(X, sd["ln_final.weight"], sd["ln_final.bias"])
    e = X[torch.arange(n, device=tok.device), tok.argmax(dim=-1)] @ sd["text_projection"]
    return F.normalize(e.float(), dim=-1)


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(CALLS):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / CALLS, out


eng = TextEngine(sd32, device=dev)
ms_a, a = timed(lambda: eng.encode(tokens, l2norm=True))
ms_b, b = timed(lambda: torch.cat([eager(tokens[i:i + 2048]) for i in range(0, N, 2048)]))
cos = torch.sum(a * b, dim=-1).min().item()
print(f"text tower, {N} x 77 tokens, width {cfg.width}, {cfg.layers} blocks ({FLOPS / 1e9:.2f} GFLOP per text, dense attention counted in full)")
print(f"  this repo      : {ms_a:8.2f} ms  {N / ms_a * 1e3:10.0f} texts/s  {N * FLOPS / ms_a / 1e9:7.1f} TFLOP/s")
print(f"  PyTorch eager  : {ms_b:8.2f} ms  {N / ms_b * 1e3:10.0f} texts/s  {N * FLOPS / ms_b / 1e9:7.1f} TFLOP/s   (fp16 cuBLAS + SDPA, 2048 texts per pass)")
print(f"  ratio {ms_b / ms_a:.2f}x   min cosine between the two: {cos:.6f}")
