"""CTA-size sweep of the fast disturbance kernel (CLIPPPO_DISTURB_NTHREADS x CLIPPPO_DISTURB_NSPLIT), one process per point."""
import os, subprocess, sys
CHILD = r'''
import os, sys
sys.path.insert(0, '.')
import torch
from clip_ppo_b200 import disturb as D
from shared.disturbance_types import DisturbanceSeverity, SEVERITY_CONFIGS
def run(B, C, H, W, sev):
    row = SEVERITY_CONFIGS[DisturbanceSeverity[sev]]
    x = torch.rand(B, C, H, W, device="cuda"); n = torch.randn(B, C, H, W, device="cuda")
    k = D.blur_kernel_size(row["gaussian_blur_sigma"]); taps = D.gaussian_taps(k, row["gaussian_blur_sigma"]); ph, pw = D.cutout_patch(H, W, row["cutout_ratio"])
    fn = lambda: D.fused_disturb(x, stages=15, noise=n, noise_sigma=row["gaussian_noise_sigma"], contrast=1.1, taps=taps, window=(3, 5, ph, pw))
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10; gb = 12.0 * B * C * H * W / 1e9
    return f"{sev[:3]} {C}x{H}: {ms*1e3:7.1f} us {gb/ms*1e3/6537.6*100:5.1f}%"
shapes = eval(os.environ["SHAPES"])
print("   " + " | ".join(run(*c) for c in shapes), flush=True)
'''
S84 = "[(16384,3,84,84,'SEVERE'),(16384,3,84,84,'MODERATE'),(16384,3,84,84,'MILD'),(16384,1,84,84,'HARD'),(16384,1,84,84,'SEVERE')]"
S224 = "[(4096,3,224,224,'SEVERE'),(4096,3,224,224,'MODERATE'),(4096,3,224,224,'MILD')]"
for shapes, nts, nss in ((S84, (0, 256, 288, 320, 352, 384), (0, 3, 4, 5, 6)), (S224, (0, 320, 352, 384), (0, 2, 3))):
    for ns in nss:
        for nt in nts:
            print(f"NTHREADS={nt} NSPLIT={ns}", flush=True)
            env = dict(os.environ, CLIPPPO_DISTURB_NTHREADS=str(nt), CLIPPPO_DISTURB_NSPLIT=str(ns), SHAPES=shapes)
            subprocess.run([sys.executable, "-c", CHILD], env=env, check=False)
