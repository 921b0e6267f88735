"""PyTorch-eager CLIP image tower (not a test, not product code): what `clip.load(name, device="cuda")` gives the
reference - an fp16 `VisionTransformer` dispatched op by op through cuDNN / cuBLAS / SDPA - for baselines on the same B200.

openai/CLIP is not installed anywhere (reference requirements.txt:12), so the module is restated here from the published
architecture ([clip] clip/model.py: LayerNorm computed in fp32 and cast back, QuickGELU, nn.MultiheadAttention's fused
attention, class token + positional embedding, ln_post on the class token, projection).  It deliberately does NOT import
oracle/ or this repository's kernels: it is the arm the kernels are compared against.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def _ln(x, w, b):
    return F.layer_norm(x.float(), (x.shape[-1],), w.float(), b.float(), 1e-5).to(x.dtype)


class EagerVisionTransformer(nn.Module):
    """`model.visual` of the openai package, eager, from a state dict in openai key layout (`visual.*`)."""

    def __init__(self, state_dict, device, dtype=torch.float16):
        super().__init__()
        sd = {k[len("visual."):]: v for k, v in state_dict.items() if k.startswith("visual.")}
        self.width = sd["conv1.weight"].shape[0]
        self.patch = sd["conv1.weight"].shape[-1]
        self.layers = len({k.split(".")[2] for k in sd if k.startswith("transformer.resblocks.")})
        self.heads = self.width // 64
        self.out_dim = sd["proj"].shape[1]
        self.p = nn.ParameterDict({k.replace(".", "/"): nn.Parameter(v.to(device=device, dtype=dtype), requires_grad=False)
                                   for k, v in sd.items()})

    def w(self, key):
        return self.p[key.replace(".", "/")]

    @torch.no_grad()
    def forward(self, x):
        D, H, P = self.width, self.heads, self.patch
        N = x.shape[0]
        tok = F.conv2d(x, self.w("conv1.weight"), stride=P)
        tok = tok.reshape(N, D, -1).permute(0, 2, 1)
        X = torch.cat([self.w("class_embedding").to(x.dtype).expand(N, 1, D), tok], 1) + self.w("positional_embedding")
        X = _ln(X, self.w("ln_pre.weight"), self.w("ln_pre.bias"))
        T = X.shape[1]
        for i in range(self.layers):
            p = f"transformer.resblocks.{i}."
            Y = _ln(X, self.w(p + "ln_1.weight"), self.w(p + "ln_1.bias"))
            qkv = F.linear(Y, self.w(p + "attn.in_proj_weight"), self.w(p + "attn.in_proj_bias"))
            q, k, v = (t.reshape(N, T, H, D // H).transpose(1, 2) for t in qkv.split(D, dim=-1))
            o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(N, T, D)
            X = X + F.linear(o, self.w(p + "attn.out_proj.weight"), self.w(p + "attn.out_proj.bias"))
            Y = _ln(X, self.w(p + "ln_2.weight"), self.w(p + "ln_2.bias"))
            h = F.linear(Y, self.w(p + "mlp.c_fc.weight"), self.w(p + "mlp.c_fc.bias"))
            h = h * torch.sigmoid(1.702 * h)
            X = X + F.linear(h, self.w(p + "mlp.c_proj.weight"), self.w(p + "mlp.c_proj.bias"))
        return _ln(X[:, 0, :], self.w("ln_post.weight"), self.w("ln_post.bias")) @ self.w("proj")


class EagerCLIP(nn.Module):
    """The surface of [clip] CLIP that shared/clip_ppo_utils.py uses: `.visual`, `.dtype`, `.encode_image`."""

    def __init__(self, state_dict, device, dtype=torch.float16):
        super().__init__()
        self.visual = EagerVisionTransformer(state_dict, device, dtype)
        self._dtype = dtype

    @property
    def dtype(self):
        return self._dtype

    @torch.no_grad()
    def encode_image(self, image):
        return self.visual(image.type(self._dtype))
