"""Not a test: operand-layout probe of the tcgen05 attention kernel (csrc/attention_tc.cu).

    python tools/probe_attention_tc.py            # tries the descriptor variants, one subprocess each
    python tools/probe_attention_tc.py --one      # checks the variant selected by the CLIPPPO_ATC_* environment

Prints the max abs error against fp32 torch for a few (n, T, heads) per variant."""
import math
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = ((2, 257, 16), (3, 128, 12), (2, 200, 12), (3, 65, 12), (5, 256, 16), (1, 100, 8))


def one():
    import torch
    from clip_ppo_b200 import _native as N
    L = N.lib()
    st = torch.cuda.current_stream().cuda_stream
    for (n, T, H, skew) in [s + (0,) for s in SHAPES] + [(2, 257, 16, 1), (3, 200, 12, 1), (2, 257, 16, 2)]:
        dh, D = 64, H * 64
        gen = torch.Generator(device="cuda").manual_seed(n * 100 + T)
        qkv = torch.randn(n * T, 3 * D, device="cuda", generator=gen)
        if skew:        # later keys score far higher (or lower): every 64-key half moves the running maximum -> the rescale paths
            t = torch.arange(T, device="cuda").repeat(n)
            f = (1.0 + 2.0 * (t // 64).float()) if skew == 1 else (1.0 + 2.0 * ((T - 1 - t) // 64).float())
            qkv[:, D:2 * D] *= f[:, None]
        qkv = qkv.bfloat16()
        out = torch.full((n * T, D), float("nan"), device="cuda", dtype=torch.bfloat16)
        N.check(L.clipppo_attention_bf16(qkv.data_ptr(), n, T, H, dh, out.data_ptr(), st))
        torch.cuda.synchronize()
        q, k, v = qkv.float().reshape(n, T, 3, H, dh).permute(2, 0, 3, 1, 4)
        s = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh), dim=-1)
        ref = (s @ v).permute(0, 2, 1, 3).reshape(n * T, D)
        err = (out.float() - ref).abs()
        rows = err.reshape(n, T, D).amax(dim=(0, 2))
        bad = int((rows > 3e-2).sum())
        print(f"  n={n} T={T} H={H} skew={skew}: max err {err.max().item():.4f}  nan {int(out.float().isnan().sum())}  "
              f"rows over 3e-2: {bad} (first {int((rows > 3e-2).float().argmax()) if bad else -1})", flush=True)


if __name__ == "__main__":
    if "--one" in sys.argv:
        one()
        sys.exit(0)
    for (lbo, sbo, pc) in ((64, 64, 8), (1, 64, 8), (64, 64, 16), (64, 128, 8), (128, 64, 8)):
        env = dict(os.environ, CLIPPPO_ATC_V_LBO=str(lbo), CLIPPPO_ATC_V_SBO=str(sbo), CLIPPPO_ATC_P_COLS=str(pc))
        print(f"variant V_LBO={lbo} V_SBO={sbo} P_COLS={pc}", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env, timeout=120,
                               capture_output=True, text=True)
            print(r.stdout, end="")
            if r.returncode:
                print("  exit", r.returncode, r.stderr[-400:])
        except subprocess.TimeoutExpired:
            print("  timeout")
