"""Not a test (pytest does not collect it): the PyTorch-eager version of the same path on the SAME B200.

SURVEY.md (headline facts, section 8d): every GPU op of the reference's hot path is a library call
dispatched by PyTorch eager, so the bar on a B200 box is "PyTorch-eager half-precision CLIP tower +
torchvision-style transforms on the same GPU".  The reference itself cannot travel to the GPU box and
openai/CLIP is not installed, so this runs the oracle port (oracle/: the reference's op sequence restated
in torch ops, pinned against the reference's outputs) on cuda:0 with the tower cast to fp16 like
`clip.load` does on CUDA, LayerNorm in fp32 like [clip] LayerNorm, attention through SDPA like
nn.MultiheadAttention's fast path.  Same workload as bench.py: 224x224x3 frames, MODERATE, supplied noise.

    python tests/eager_gpu_baseline.py [frames per step] [steps]

Lives under tests/ because only tests/, smoke() and bench.py's CPU legs may import oracle/.
"""
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F

from oracle import disturb as od, losses as ol, vit as ov

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda", 0)
sd32 = ov.random_state_dict(ov.VIT_B32, 0)
sd = {k: v.to(dev).half() for k, v in sd32.items()}
cfg = ov.VIT_B32


def ln(x, w, b):                                  # [clip] LayerNorm: fp32 inside, cast back
    return F.layer_norm(x.float(), (x.shape[-1],), w.float(), b.float(), 1e-5).to(x.dtype)


def tower_fp16(x):                                # oracle/vit.py:vision_tower, fp16, SDPA
    D, H, P, G = cfg.width, cfg.heads, cfg.patch, cfg.grid
    N = x.shape[0]
    tok = F.conv2d(x, sd["visual.conv1.weight"], stride=P).reshape(N, D, G * G).permute(0, 2, 1)
    X = torch.cat([sd["visual.class_embedding"].expand(N, 1, D), tok], 1) + sd["visual.positional_embedding"]
    X = ln(X, sd["visual.ln_pre.weight"], sd["visual.ln_pre.bias"])
    T = X.shape[1]
    for i in range(cfg.layers):
        p = f"visual.transformer.resblocks.{i}."
        Y = ln(X, sd[p + "ln_1.weight"], sd[p + "ln_1.bias"])
        qkv = F.linear(Y, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"])
        q, k, v = (t.reshape(N, T, H, D // H).transpose(1, 2) for t in qkv.split(D, dim=-1))
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(N, T, D)
        X = X + F.linear(o, sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])
        Y = ln(X, sd[p + "ln_2.weight"], sd[p + "ln_2.bias"])
        h = F.linear(Y, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"])
        h = h * torch.sigmoid(1.702 * h)
        X = X + F.linear(h, sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])
    return ln(X[:, 0, :], sd["visual.ln_post.weight"], sd["visual.ln_post.bias"]) @ sd["visual.proj"]


MEAN = torch.tensor(ov.CLIP_MEAN, device=dev).view(1, 3, 1, 1)
STD = torch.tensor(ov.CLIP_STD, device=dev).view(1, 3, 1, 1)


def preprocess(images):                           # reference shared/clip_ppo_utils.py:146-159 on the device
    u = F.interpolate(images.float() / 255.0, size=(224, 224), mode="bilinear", align_corners=False, antialias=True)
    return (u - MEAN) / STD


g = torch.Generator(device=dev).manual_seed(0)
x = torch.randint(0, 256, (B, 3, 224, 224), device=dev, generator=g, dtype=torch.uint8).float().div_(255.0)
noise = torch.randn(B, 3, 224, 224, device=dev, generator=g)
z = torch.relu(torch.randn(B, 512, device=dev, generator=g))
sev = od.SEVERITY_TABLE["MODERATE"]
k1d = od.gaussian_kernel1d(od.blur_kernel_size(sev["blur_sigma"]), sev["blur_sigma"]).to(dev)
ph, pw = od.cutout_patch(224, 224, sev["cutout"])


@torch.no_grad()
def step():
    d = od.disturb(x, noise, sev["noise_sigma"], 1.1, k1d, 224 // 5, 224 // 4, ph, pw)      # torchvision's op sequence
    pre = preprocess(d * 255.0)                                                         # /255, resize (identity), normalise
    e = F.normalize(tower_fp16(pre.half()).float(), dim=-1)
    return ol.cosine_embedding_loss(z, e)


for _ in range(2):
    loss = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(STEPS):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / STEPS
# parts
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
with torch.no_grad():
    t_d = timeit(lambda: od.disturb(x, noise, sev["noise_sigma"], 1.1, k1d, 44, 56, ph, pw))
    pre = preprocess(x * 255.0).half()
    t_t = timeit(lambda: tower_fp16(pre))
print(f"PyTorch eager on {torch.cuda.get_device_name(0)} (torch {torch.__version__}), {B} frames/step: "
      f"{ms:.1f} ms/step = {B / ms * 1e3:.0f} frames/s   [disturb {t_d:.1f} ms ({12.0 * x.numel() / t_d / 1e6:.0f} GB/s algorithmic), "
      f"fp16 tower {t_t:.1f} ms ({B * 8.8176e9 / t_t / 1e9:.0f} TFLOP/s)]  loss {float(loss):.5f}")
