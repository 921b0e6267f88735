"""A/B of disturbance-kernel builds on one box:  python tools/disturb_ab.py lib_a.so lib_b.so ...  (each in its own process,
interleaved twice; CUDA events, 10 launches after 3 warm-ups; % of the measured 6537.6 GB/s copy peak at 12 B / element)."""
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
    print(f"   {sev:8s} {B}x{C}x{H}x{W}: {ms*1e3:8.1f} us {gb/ms*1e3/6537.6*100:5.1f}%", flush=True)
    return fn()
outs = []
for cfg in [(4096,3,224,224,"SEVERE"), (4096,3,224,224,"MODERATE"), (4096,3,224,224,"MILD"), (16384,3,84,84,"SEVERE"), (16384,1,84,84,"HARD")]:
    outs.append(run(*cfg)[:8].double().sum().item())
print("   checksum", " ".join(f"{o:.6f}" for o in outs))
'''
libs = sys.argv[1:]
for rep in range(2):
    for lib in libs:
        print(f"[{rep}] {lib}", flush=True)
        env = dict(os.environ, CLIPPPO_LIB=os.path.abspath(lib))
        subprocess.run([sys.executable, "-c", CHILD], env=env, check=False)
