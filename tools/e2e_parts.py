import sys, torch
sys.path.insert(0, "/root/repo")
B=4096
u8 = torch.randint(0,256,(B,3,224,224),device="cuda",dtype=torch.uint8)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
xf = torch.div(u8, 255.0)
print("div u8->f32      : %.3f ms" % t(lambda: torch.div(u8, 255.0)))
print("randn_like       : %.3f ms" % t(lambda: torch.randn_like(xf)))
print("empty_like alloc : %.3f ms" % t(lambda: torch.empty_like(xf)))

from clip_ppo_b200 import disturb as D
from shared.disturbance_types import DisturbanceSeverity, SEVERITY_CONFIGS
row = SEVERITY_CONFIGS[DisturbanceSeverity.MODERATE]
k = D.blur_kernel_size(row["gaussian_blur_sigma"]); taps = D.gaussian_taps(k, row["gaussian_blur_sigma"])
ph, pw = D.cutout_patch(224, 224, row["cutout_ratio"])
noise = torch.randn_like(xf)
kw = dict(stages=15, noise=noise, noise_sigma=row["gaussian_noise_sigma"], contrast=1.1, taps=taps, window=(3, 5, ph, pw))
print("disturb fp32 in  : %.3f ms" % t(lambda: D.fused_disturb(xf, **kw)))
print("disturb uint8 in : %.3f ms   (9 B / element instead of 12; replaces div + disturb)" % t(lambda: D.fused_disturb(u8, **kw)))
