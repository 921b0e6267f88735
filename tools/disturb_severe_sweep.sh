for cfg in "113 0 16" "113 1 16" "113 2 16" "113 3 16" "75 0 16" "75 1 16" "75 2 16" "56 0 16" "56 1 16" "56 2 16" "227 0 16" "227 2 16" "227 4 16"; do
  set -- $cfg
  echo "SMEM_KB=$1 NSPLIT=$2 MAXCL=$3"
  CLIPPPO_DISTURB_SMEM_KB=$1 CLIPPPO_DISTURB_NSPLIT=$2 CLIPPPO_DISTURB_MAXCL=$3 python - <<'PY'
import os,sys
sys.path.insert(0,'.')
import torch
from clip_ppo_b200 import disturb as D
from shared.disturbance_types import DisturbanceSeverity, SEVERITY_CONFIGS
def run(B,C,H,W,sev):
    row=SEVERITY_CONFIGS[DisturbanceSeverity[sev]]
    x=torch.rand(B,C,H,W,device="cuda"); n=torch.randn(B,C,H,W,device="cuda")
    k=D.blur_kernel_size(row["gaussian_blur_sigma"]); taps=D.gaussian_taps(k,row["gaussian_blur_sigma"]); ph,pw=D.cutout_patch(H,W,row["cutout_ratio"])
    fn=lambda: D.fused_disturb(x,stages=15,noise=n,noise_sigma=row["gaussian_noise_sigma"],contrast=1.1,taps=taps,window=(3,5,ph,pw))
    for _ in range(3): fn()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10; gb=12.0*B*C*H*W/1e9
    print(f"  {sev} {B}x{C}x{H}x{W}: {ms*1e3:8.1f} us {gb/ms*1e3/6537.6*100:5.1f}%")
run(4096,3,224,224,"SEVERE"); run(16384,3,84,84,"SEVERE"); run(4096,3,224,224,"MODERATE")
PY
done
