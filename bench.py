#!/usr/bin/env python
"""Headline benchmark of the CLIP-PPO observation path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): CLIP-PPO frames/sec through  disturb -> ViT-B/32 embed -> alignment loss.
Workload: BASELINE.json configs[2] - synthetic 224x224 RGB frames, 4096 per GPU per step (weak
scaling), MODERATE disturbances, seeded random ViT-B/32 weights (no CLIP weights exist offline).

One step, through the drop-in API of shared/:
    d    = DisturbanceWrapperGPU.apply_disturbances(x, noise=...)       one fused launch  (D1)
    emb  = frozen tower on d (resize/normalise/im2col + 12 blocks + proj + L2-norm)       (V0-V7)
    loss = compute_cosine_embedding_loss(z, emb)                                           (L1)
    N>1: NCCL all-reduce of a PPO-agent-sized fp32 gradient bucket (the path's only exchange)

`value` times the step with x / noise / z resident in HBM; `e2e` feeds uint8 frames from pinned
host memory every step (H2D inside the timed region, prefetched one step ahead on a copy stream),
draws the noise on the device like the public API does, and reads the loss back to the host.
`--impl reference` times the CPU oracle port of the same path (the reference is Python and cannot
travel to the GPU box; see oracle/__init__.py) on all host threads on a bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch

METRIC = "CLIP-PPO frames/sec (disturb+ViT-B/32 embed+align loss)"
FLOPS_PER_IMAGE = 8.8176e9          # SURVEY.md §8d, ViT-B/32, all 50 tokens through 12 layers
FLOPS_PER_IMAGE_BY_MODEL = {"ViT-B/32": 8.8176e9, "ViT-L/14": 162.03e9}
AGENT_GRAD_ELEMS = 1_686_180        # MiniGrid NatureCNN agent fp32 grads, 7 actions


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="frames per GPU per step")
    ap.add_argument("--hw", type=int, default=224)
    ap.add_argument("--severity", default="MODERATE")
    ap.add_argument("--model", default="ViT-B/32", choices=["ViT-B/32", "ViT-L/14"],
                    help="image tower (the headline is ViT-B/32; ViT-L/14 is BASELINE configs[4]'s variant)")
    ap.add_argument("--ref-batch", type=int, default=64, help="frames per step of the CPU reference arm")
    ap.add_argument("--cpu-sample", type=int, default=192, help="frames of the cpu_baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU oracle arm (cpu_baseline and --impl reference)
# ------------------------------------------------------------------------------------------------
def cpu_path_step(sd, frames01, noise, z, sev):
    from oracle import disturb as od, losses as ol, vit as ov
    cfg = od.SEVERITY_TABLE[sev]
    H, W = frames01.shape[-2:]
    ph, pw = od.cutout_patch(H, W, cfg["cutout"])
    k1d = od.gaussian_kernel1d(od.blur_kernel_size(cfg["blur_sigma"]), cfg["blur_sigma"])
    d = od.disturb(frames01, noise, cfg["noise_sigma"], 1.1, k1d, H // 5, W // 4, ph, pw)
    emb = ov.image_embeddings(sd, d * 255.0)
    return ol.cosine_embedding_loss(z, emb)


def cpu_inputs(n, hw, seed=0):
    g = torch.Generator().manual_seed(seed)
    frames = torch.randint(0, 256, (n, 3, hw, hw), generator=g).float() / 255.0
    noise = torch.randn(n, 3, hw, hw, generator=g)
    z = torch.relu(torch.randn(n, 512, generator=g))
    return frames, noise, z


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import vit as ov
    torch.set_num_threads(os.cpu_count() or 1)
    sd = ov.random_state_dict(ov.VIT_B32, 0)
    n = args.ref_batch
    frames, noise, z = cpu_inputs(n, args.hw)
    with torch.no_grad():
        for _ in range(args.warmup):
            cpu_path_step(sd, frames, noise, z, args.severity)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            loss = cpu_path_step(sd, frames, noise, z, args.severity)
        dt = time.perf_counter() - t0
    value = n * args.steps / dt
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC.replace("ViT-B/32", args.model), "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.batch),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{n} frames/step x {args.steps} steps of the same workload, fp32 CPU oracle port "
                                   f"(oracle/: torch CPU ops, {cores} threads); loss {float(loss):.6f}"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, per_gpu_batch):
    which = "configs[2]" if args.model == "ViT-B/32" else "configs[4]'s ViT-L/14 variant, one GPU's share"
    return {"workload": f"BASELINE {which}: synthetic {args.hw}x{args.hw} RGB frames, disturb({args.severity}) -> "
                        f"{args.model} embed -> cosine alignment loss",
            "per_gpu_batch": per_gpu_batch, "frame": [3, args.hw, args.hw], "severity": args.severity,
            "weights": f"seeded random {args.model} (openai key layout)",
            "l2_policy": "inputs larger than L2 (x + noise = 4.9 GB per step at 4096 frames); no flush needed",
            "parallelism": f"dp{args.gpus} by observation, frozen tower replicated, grad-bucket all-reduce only"}


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from clip_ppo_b200 import _native as N
    import shared.clip_ppo_utils as U
    from shared.disturbances_gpu import DisturbanceWrapperGPU
    from shared.disturbance_types import DisturbanceSeverity
    from clip_ppo_b200 import disturb as D

    L = N.lib()                                   # raises if the CUDA library is missing: no fallback
    B, hw = args.batch, args.hw
    model = U.load_clip_model(args.model, device=dev)
    engine = U._engine_for(model)
    disturber = DisturbanceWrapperGPU(device=dev, seed=1234 + rank, severity=DisturbanceSeverity[args.severity])
    ph, pw = D.cutout_patch(hw, hw, disturber.cutout_ratio)
    window = (hw // 5, hw // 4)

    g = torch.Generator(device=dev).manual_seed(rank)
    x = torch.randint(0, 256, (B, 3, hw, hw), device=dev, generator=g, dtype=torch.uint8).float().div_(255.0)
    noise = torch.randn(B, 3, hw, hw, device=dev, generator=g)
    z = torch.relu(torch.randn(B, engine.cfg.out_dim, device=dev, generator=g))
    grad_bucket = torch.zeros(AGENT_GRAD_ELEMS, device=dev) if world > 1 else None

    def step_device():
        d = disturber.apply_disturbances(x, noise=noise, contrast_factor=1.1, cutout_start=window)
        emb = engine.encode(d, pre_scale=1.0, l2norm=True)      # == generate_clip_embeddings(images = d*255)
        loss = U.compute_cosine_embedding_loss(z, emb)
        if grad_bucket is not None:
            dist.all_reduce(grad_bucket)
        return loss

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if dist is None:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing ------------------------------------------------------------------
    for _ in range(args.warmup):
        loss = step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # pass A - the headline: K steps, launches counted, NO per-kernel events (bracketing every GEMM with
    # an event pair costs 1.5 % of the step: it defeats the programmatic-dependent-launch overlap)
    L.clipppo_prof_begin(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        loss = step_device()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches, gemm_ms, gemm_flops, gemm_launches = C.c_longlong(), C.c_double(), C.c_double(), C.c_longlong()
    N.check(L.clipppo_prof_end(C.byref(launches), None, None, None))
    # pass B - the roofline of the dominant kernel: the same K steps again, every tcgen05 GEMM launch
    # bracketed by CUDA events on its stream (clipppo_prof_begin(1)); still inside the clock-sampled region
    L.clipppo_prof_begin(1)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(args.steps):
        loss = step_device()
    f1.record()
    barrier()
    ms_total_timed = f0.elapsed_time(f1)
    N.check(L.clipppo_prof_end(None, C.byref(gemm_ms), C.byref(gemm_flops), C.byref(gemm_launches)))
    by_shape = []
    for bi in range(64):
        tag, bms, bfl, bn = C.c_longlong(), C.c_double(), C.c_double(), C.c_longlong()
        if L.clipppo_prof_bucket(bi, C.byref(tag), C.byref(bms), C.byref(bfl), C.byref(bn)) != 0:
            break
        by_shape.append({"epilogue": tag.value >> 40, "N": (tag.value >> 20) & 0xFFFFF, "K": tag.value & 0xFFFFF,
                         "launches": bn.value, "ms": round(bms.value, 3),
                         "tflops": round(bfl.value / (bms.value * 1e-3) / 1e12, 1) if bms.value > 0 else 0.0})
    clocks = sampler.stop() if rank == 0 else None
    loss_value = float(loss.item())
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- end-to-end: uint8 frames from pinned host memory, noise drawn on device, loss read back ----
    e2e = None
    if not args.no_e2e:
        host = [torch.randint(0, 256, (B, 3, hw, hw), dtype=torch.uint8).pin_memory() for _ in range(2)]
        dev_u8 = [torch.empty((B, 3, hw, hw), dtype=torch.uint8, device=dev) for _ in range(2)]
        host_loss = torch.zeros(2, dtype=torch.float32).pin_memory()
        copy_stream = torch.cuda.Stream(device=dev)
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        main_stream = torch.cuda.current_stream(dev)

        def prefetch(i):
            b = i & 1
            copy_stream.wait_event(consumed[b])
            with torch.cuda.stream(copy_stream):
                dev_u8[b].copy_(host[b], non_blocking=True)
                ready[b].record(copy_stream)

        def step_e2e(i):
            b = i & 1
            prefetch(i + 1)                                     # next step's frames ride under this step's compute
            main_stream.wait_event(ready[b])
            # uint8 frames straight into the public call: the kernel reads them as `.float() / 255` (bit-identical to
            # the reference benchmark's conversion on the device); randn on device + CPU-generator draws as in the reference
            d = disturber.apply_disturbances(dev_u8[b])
            consumed[b].record(main_stream)
            emb = engine.encode(d, pre_scale=1.0, l2norm=True)
            ls = U.compute_cosine_embedding_loss(z, emb)
            if grad_bucket is not None:
                dist.all_reduce(grad_bucket)
            host_loss[b:b + 1].copy_(ls.reshape(1), non_blocking=True)
            done[b].record(main_stream)
            if i > 0:
                done[(i - 1) & 1].synchronize()                 # host consumes the previous step's loss

        for b in range(2):
            consumed[b].record(main_stream)
        prefetch(0)
        for i in range(args.warmup):
            step_e2e(i)
        barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        base = args.warmup
        for i in range(args.steps):
            step_e2e(base + i)
        done[(base + args.steps - 1) & 1].synchronize()
        t1.record()
        barrier()
        ms_e2e = max_over_ranks(t0.elapsed_time(t1))
        e2e = {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": "frames/s",
               "h2d_bytes_per_step": B * 3 * hw * hw, "d2h_bytes_per_step": 4,
               "ms_per_step": ms_e2e / args.steps,
               "note": "uint8 frames from pinned host memory, H2D prefetched one step ahead on a copy stream, handed to "
                       "apply_disturbances as uint8 (read as .float()/255 inside the kernel); noise drawn on device "
                       "(randn); loss copied back to pinned memory every step"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    pk = peaks()
    tf_ach = gemm_flops.value / (gemm_ms.value * 1e-3) / 1e12 if gemm_ms.value > 0 else 0.0
    tower_share = gemm_ms.value / ms_total_timed if ms_total_timed > 0 else 0.0
    ncu_traffic = None
    ncu_json = os.path.join(ROOT, "profiles", "ncu_gemm_summary.json")
    if os.path.exists(ncu_json):
        try:
            ncu_traffic = json.load(open(ncu_json)).get("dram_bytes_per_launch")
        except Exception:
            ncu_traffic = None
    line = {
        "metric": METRIC.replace("ViT-B/32", args.model), "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, B),
        "clocks": clocks, "gpu_launches": int(launches.value),
        "roofline": {"kernel": "gemm_bf16_kernel (tcgen05.mma + TMA, all fused epilogues)", "bound": "tensor",
                     "achieved": tf_ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                     "frac": tf_ach / pk["tf_sustained"], "traffic": ncu_traffic,
                     "peak_source": f"{pk['source']} bf16 sustained (kernel timed inside a long step)",
                     "launches_timed": int(gemm_launches.value), "share_of_step": tower_share,
                     "timed_pass": f"second pass of the same {args.steps} steps with an event pair around every GEMM launch "
                                   f"({ms_total_timed / args.steps:.2f} ms/step; the headline pass runs without them)",
                     "by_shape": by_shape,
                     "tower_tflops_incl_all_kernels": world * B * args.steps * FLOPS_PER_IMAGE_BY_MODEL[args.model] / (ms_total * 1e-3) / 1e12 / world},
        "loss": loss_value,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if not args.no_cpu_baseline and world == 1:
        from oracle import vit as ov
        torch.set_num_threads(os.cpu_count() or 1)
        sd = ov.random_state_dict(ov.VIT_B32, 0)
        n = args.cpu_sample
        frames, nz, zz = cpu_inputs(n, hw)
        with torch.no_grad():
            cpu_path_step(sd, frames[:16], nz[:16], zz[:16], args.severity)
            t0 = time.perf_counter()
            cpu_path_step(sd, frames, nz, zz, args.severity)
            dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": n / dt, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{n} frames of the same workload through the fp32 CPU oracle port "
                                          f"(torch CPU ops, all host threads), {dt:.1f} s"}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
