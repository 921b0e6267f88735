#!/usr/bin/env python
"""Headline benchmark of the CLIP-PPO observation path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): CLIP-PPO frames/sec through  disturb -> ViT-B/32 embed -> alignment loss.
Workload: BASELINE.json configs[2] - synthetic 224x224 RGB frames, 4096 per GPU per step (weak scaling), MODERATE
disturbances, seeded random ViT-B/32 weights (no CLIP weights exist offline).

One step, every call the PUBLIC drop-in surface of shared/ (the calls the reference's training scripts make):
    obs  = DisturbanceWrapperGPU.apply_disturbances(x, noise=..., out_scale=255)   one fused launch (D1); out_scale = the call sites'
                                                                                   own `* 255` (clip_ppo_minigrid.py:388) done by the
                                                                                   kernel's store, bit-identical to the separate pass
    emb  = generate_clip_embeddings(NONE, model, "image", B, dev, images=obs)      /255 + resize + normalise + tower + L2 norm (V0-V7)
    loss = compute_cosine_embedding_loss(z, emb)                                   (L1)
    N>1: GradBucket.all_reduce_mean() of the PPO agent's real gradients (the path's only exchange)

`value` times the step with x / noise / z resident in HBM; `e2e` feeds uint8 frames from pinned host memory every step (H2D
inside the timed region, prefetched one step ahead on a copy stream), draws the noise on the device like the public API does,
and reads the loss back to the host.  Extra keys of the JSON line:
    roofline      the tcgen05 GEMM (dominant kernel), CUDA-event timed inside a second pass of the same steps
    secondary     short CUDA-event measurements of the other BASELINE configs: disturbance HBM fractions (MODERATE / SEVERE
                  224x224x3, HARD 84x84x1), the learner side of one MiniGrid iteration (configs[1]) next to the script's own
                  statements over the reference's modules on the same GPU, the Atari 84 -> 224 path (configs[3], one GPU's share),
                  ViT-L/14 (configs[4])
    eager_gpu_baseline   the reference's OWN functions (oracle/_ref, staged by oracle/build_ref.py) on the same GPU with an
                  fp16 eager tower (tools/eager_tower.py) - the library-dispatched path the reference runs on a GPU
    cpu_baseline  the reference's own functions on the host cores (fp32 tower restated in oracle/vit.py), bounded sample
    parity, strong   (N > 1) shard-vs-whole equality of embeddings / all-reduced gradients, and the fixed-total-work timing
`--impl reference` times the reference's own CPU path (oracle/_ref; the oracle port if it was never staged) on all host threads.
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
# no CLIP checkpoint exists offline: the towers measured here are seeded random weights of the named architecture
os.environ.setdefault("CLIPPPO_ALLOW_RANDOM_WEIGHTS", "1")

import torch

METRIC = "CLIP-PPO frames/sec (disturb+ViT-B/32 embed+align loss)"
FLOPS_PER_IMAGE_BY_MODEL = {"ViT-B/32": 8.8176e9, "ViT-L/14": 162.03e9}
CHUNK = 256                          # frames per seeded generation chunk (inputs are a function of the GLOBAL frame index)


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
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary / eager_gpu_baseline legs")
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


def event_ms(fn, iters: int, warm: int = 2) -> float:
    """Mean device time of fn() over `iters` calls after `warm` untimed ones (CUDA events on the current stream)."""
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


# ------------------------------------------------------------------------------------------------
# the reference's own functions (oracle/_ref) - CPU arm, cpu_baseline, eager GPU baseline
# ------------------------------------------------------------------------------------------------
class _OracleTower(torch.nn.Module):
    """What `clip.load(name, "cpu")` would return as far as shared/clip_ppo_utils.py can tell: `encode_image` on a
    normalised fp32 batch.  The arithmetic is the fp32 restatement of [clip] VisionTransformer in oracle/vit.py."""

    def __init__(self, sd):
        super().__init__()
        self.sd = sd
        self.anchor = torch.nn.Parameter(torch.zeros(1))

    @torch.no_grad()
    def encode_image(self, x):
        from oracle import vit as ov
        return ov.vision_tower(self.sd, x.float())


def reference_step_fn(device: str, model_name: str, severity: str, seed: int = 0):
    """(step(frames01, z) -> loss, kind): the reference's stock code path - DisturbanceWrapperGPU.apply_disturbances ->
    `* 255` -> generate_clip_embeddings -> compute_cosine_embedding_loss - out of oracle/_ref when it is staged
    (kind "reference"), else the oracle port (kind "port")."""
    from oracle import build_ref, vit as ov
    cfg = ov.VIT_B32 if model_name == "ViT-B/32" else ov.VIT_L14
    sd = ov.random_state_dict(cfg, 0)
    if build_ref.available():
        if device == "cpu":
            tower = _OracleTower(sd)
        else:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            from eager_tower import EagerCLIP
            tower = EagerCLIP(sd, device)
        ref = build_ref.load(tower)
        U, DG, DT = ref.clip_ppo_utils, ref.disturbances_gpu, ref.disturbance_types
        model = U.load_clip_model(model_name, device)
        w = DG.DisturbanceWrapperGPU(device=device, seed=seed, severity=DT.DisturbanceSeverity[severity])

        def step(frames01, z):
            d = w.apply_disturbances(frames01)
            emb = U.generate_clip_embeddings(U.AblationMode.NONE, model, "image", frames01.shape[0], device, images=d * 255.0)
            return U.compute_cosine_embedding_loss(z, emb)
        return step, "reference"
    if device != "cpu":
        return None, "unavailable"
    from oracle import disturb as od, losses as ol
    c = od.SEVERITY_TABLE[severity]

    def step(frames01, z):
        H, W = frames01.shape[-2:]
        ph, pw = od.cutout_patch(H, W, c["cutout"])
        k1d = od.gaussian_kernel1d(od.blur_kernel_size(c["blur_sigma"]), c["blur_sigma"])
        d = od.disturb(frames01, torch.randn_like(frames01), c["noise_sigma"], 1.1, k1d, H // 5, W // 4, ph, pw)
        return ol.cosine_embedding_loss(z, ov.image_embeddings(sd, d * 255.0))
    return step, "port"


def cpu_inputs(n, hw, out_dim=512, seed=0):
    g = torch.Generator().manual_seed(seed)
    frames = torch.randint(0, 256, (n, 3, hw, hw), generator=g).float() / 255.0
    z = torch.relu(torch.randn(n, out_dim, generator=g))
    return frames, z


def run_reference_arm(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    step, kind = reference_step_fn("cpu", args.model, args.severity)
    n = args.ref_batch
    frames, z = cpu_inputs(n, args.hw, 512 if args.model == "ViT-B/32" else 768)
    with torch.no_grad():
        for _ in range(args.warmup):
            step(frames, z)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            loss = step(frames, z)
        dt = time.perf_counter() - t0
    value = n * args.steps / dt
    cores = torch.get_num_threads()
    what = ("the reference's own shared/disturbances_gpu.py + shared/clip_ppo_utils.py (oracle/_ref, unmodified) over the fp32 "
            "tower of oracle/vit.py" if kind == "reference" else "fp32 CPU oracle port (oracle/: torch CPU ops)")
    line = {
        "impl": "reference", "metric": METRIC.replace("ViT-B/32", args.model), "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, n, note=f"CPU arm: a bounded {n}-frame step of the same workload (the GPU arm runs {args.batch})"),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": kind,
                         "sample": f"{n} frames/step x {args.steps} steps of the same workload, {what}, {cores} threads; loss {float(loss):.6f}"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, per_gpu_batch, note=None):
    which = "configs[2]" if args.model == "ViT-B/32" else "configs[4]'s ViT-L/14 variant, one GPU's share"
    cfg = {"workload": f"BASELINE {which}: synthetic {args.hw}x{args.hw} RGB frames, disturb({args.severity}) -> "
                       f"{args.model} embed -> cosine alignment loss",
           "per_gpu_batch": per_gpu_batch, "frame": [3, args.hw, args.hw], "severity": args.severity,
           "weights": f"seeded random {args.model} (openai key layout)",
           "api": "apply_disturbances(out_scale=255: the `* 255` of the call sites inside the kernel's store, bit-identical) -> "
                  "generate_clip_embeddings(images=) -> compute_cosine_embedding_loss (public calls only)",
           "inputs": f"frame / noise / latent chunks of {CHUNK} seeded by GLOBAL chunk index (the same data whatever the GPU count)",
           "l2_policy": "inputs larger than L2 (x + noise = 4.9 GB per step at 4096 frames); no flush needed",
           "parallelism": f"dp{args.gpus} by observation, frozen tower replicated, PPO-agent gradient all-reduce only"}
    if note:
        cfg["note"] = note
        # the reference arm runs the reference's own call sequence
        cfg["api"] = "apply_disturbances -> `* 255` -> generate_clip_embeddings(images=) -> compute_cosine_embedding_loss (the reference's own functions)"
    return cfg


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def seeded_inputs(dev, first_frame: int, n: int, hw: int, out_dim: int, want_noise: bool = True):
    """Frames [n,3,hw,hw] in [0,1] (uint8 values / 255), supplied noise, post-ReLU latents - generated chunk by chunk from
    generators seeded with the GLOBAL chunk index, so rank r of G holds exactly rows [r n, (r+1) n) of the one global batch."""
    xs, ns, zs = [], [], []
    for c0 in range(first_frame, first_frame + n, CHUNK):
        m = min(CHUNK, first_frame + n - c0)
        g = torch.Generator(device=dev).manual_seed(100_003 + c0 // CHUNK)
        xs.append(torch.randint(0, 256, (m, 3, hw, hw), device=dev, generator=g, dtype=torch.uint8))
        if want_noise:
            ns.append(torch.randn(m, 3, hw, hw, device=dev, generator=g))
        zs.append(torch.relu(torch.randn(m, out_dim, device=dev, generator=g)))
    x8 = torch.cat(xs)
    return x8, (torch.cat(ns) if want_noise else None), torch.cat(zs)


class PolicyEncoder(torch.nn.Module):
    """The trainable side of the path: the reference's NatureCNN encoder + heads (clip_ppo_minigrid.py:229-242), 1 686 180
    fp32 parameters with 7 actions - the gradients the data-parallel all-reduce carries."""

    def __init__(self, n_actions: int = 7):
        super().__init__()
        nn = torch.nn
        from clip_ppo_b200.policy import NatureCNN
        # the encoder runs on this repository's fp32 implicit-GEMM kernels (csrc/policy.cu): the gradients the
        # all-reduce carries are produced by them, deterministically
        self.network = NatureCNN(3)
        self.actor = nn.Linear(512, n_actions)
        self.critic = nn.Linear(512, 1)

    def forward(self, obs84):
        h = self.network(obs84)
        return h, self.actor(h), self.critic(h)


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
    from clip_ppo_b200 import rollout as R
    from clip_ppo_b200.distributed import GradBucket

    L = N.lib()                                   # raises if the CUDA library is missing: no fallback
    B, hw = args.batch, args.hw
    model = U.load_clip_model(args.model, device=dev)
    out_dim = U._engine_for(model).cfg.out_dim
    # identically seeded on every rank: the per-call scalars (contrast factor, cutout window) come from the CPU generator
    disturber = DisturbanceWrapperGPU(device=dev, seed=1234, severity=DisturbanceSeverity[args.severity])
    window = (hw // 5, hw // 4)

    # the disturbance kernel against the HBM roofline is a kernel timed ALONE (vs the burst copy peak): measured first, before
    # the sustained passes push the part into its power cap
    disturb_rows = disturb_hbm_rows(dev, peaks()) if (world == 1 and not args.no_secondary) else None

    x8, noise, z = seeded_inputs(dev, rank * B, B, hw, out_dim)
    x = x8.float().div_(255.0)
    del x8

    def path(xf, nz, zz):
        # out_scale=255: the `* 255` between the two calls (the rollout buffers hold 0..255) done by the kernel's store,
        # bit-identical to `apply_disturbances(...) * 255` (tests/test_disturb_gpu.py::test_out_scale_equals_a_separate_multiply_bitwise)
        d255 = disturber.apply_disturbances(xf, noise=nz, contrast_factor=1.1, cutout_start=window, out_scale=255.0)
        emb = U.generate_clip_embeddings(U.AblationMode.NONE, model, "image", xf.shape[0], dev, images=d255)
        return U.compute_cosine_embedding_loss(zz, emb), emb

    # the exchange of the path: the PPO agent's gradients.  One real backward (alignment loss of the agent's latents against
    # the embeddings of this rank's first frames) fills them; every timed step all-reduces that same real gradient set.
    bucket = saved_grads = agent = None
    if world > 1:
        torch.manual_seed(7)
        agent = PolicyEncoder().to(dev)
        with torch.no_grad():
            _, emb0 = path(x[:CHUNK], noise[:CHUNK], z[:CHUNK])
        obs84 = torch.nn.functional.interpolate(x[:CHUNK], size=(84, 84), mode="bilinear", align_corners=False)
        h, logits, v = agent(obs84)
        (U.compute_cosine_embedding_loss(h, emb0) + 1e-3 * logits.square().mean() + 1e-3 * v.square().mean()).backward()
        bucket = GradBucket(agent.parameters())
        saved_grads = [p.grad.clone() for p in bucket.params]

    def step_device():
        loss, _ = path(x, noise, z)
        if bucket is not None:
            torch._foreach_copy_([p.grad for p in bucket.params], saved_grads)
            bucket.all_reduce_mean()
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

        e2e_noise_seed = [None]                                 # None: torch.randn_like's stream (the reference's); int: in-kernel Philox

        def step_e2e(i):
            b = i & 1
            prefetch(i + 1)                                     # next step's frames ride under this step's compute
            main_stream.wait_event(ready[b])
            # uint8 frames straight into the public call: the kernel reads them as `.float() / 255` (bit-identical to
            # the reference benchmark's conversion on the device); CPU-generator draws as in the reference; the noise is
            # torch.randn on the device (default) or generated inside the kernel (opt-in noise_seed=)
            d = disturber.apply_disturbances(dev_u8[b], out_scale=255.0, noise_seed=e2e_noise_seed[0])
            consumed[b].record(main_stream)
            emb = U.generate_clip_embeddings(U.AblationMode.NONE, model, "image", B, dev, images=d)
            ls = U.compute_cosine_embedding_loss(z, emb)
            if bucket is not None:
                torch._foreach_copy_([p.grad for p in bucket.params], saved_grads)
                bucket.all_reduce_mean()
            host_loss[b:b + 1].copy_(ls.reshape(1), non_blocking=True)
            done[b].record(main_stream)
            if i > 0:
                done[(i - 1) & 1].synchronize()                 # host consumes the previous step's loss

        for b in range(2):
            consumed[b].record(main_stream)
        prefetch(0)
        base = [0]

        def timed_e2e():
            for i in range(args.warmup):
                step_e2e(base[0] + i)
            base[0] += args.warmup
            barrier()
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
            for i in range(args.steps):
                step_e2e(base[0] + i)
            done[(base[0] + args.steps - 1) & 1].synchronize()
            t1.record()
            base[0] += args.steps
            barrier()
            return max_over_ranks(t0.elapsed_time(t1))

        ms_e2e = timed_e2e()                                    # the reference's RNG contract: noise = torch.randn on the device
        e2e_noise_seed[0] = 20260000 + rank
        ms_e2e_philox = timed_e2e()                             # opt-in: noise generated inside the disturbance kernel
        e2e = {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": "frames/s",
               "h2d_bytes_per_step": B * 3 * hw * hw, "d2h_bytes_per_step": 4,
               "ms_per_step": ms_e2e / args.steps,
               "note": "uint8 frames from pinned host memory, H2D prefetched one step ahead on a copy stream, handed to "
                       "apply_disturbances as uint8 (read as .float()/255 inside the kernel); noise drawn on device by "
                       "torch.randn (the reference's stream); `* 255` inside the kernel's store (out_scale); "
                       "generate_clip_embeddings as in the scripts; loss copied back to pinned memory every step",
               "in_kernel_noise": {"value": world * B * args.steps / (ms_e2e_philox * 1e-3), "unit": "frames/s",
                                   "ms_per_step": ms_e2e_philox / args.steps,
                                   "note": "the same step with apply_disturbances(noise_seed=): N(0,1) draws generated inside the "
                                           "kernel (Philox4x32-10 + Box-Muller, opt-in, not torch's stream; oracle/philox.py)"}}
        del host, dev_u8

    # ---- N > 1: shard-vs-whole parity and the strong-scaling figure ---------------------------------
    parity = strong = None
    if world > 1:
        parity, strong = multi_gpu_checks(args, dist, dev, rank, world, path, agent, bucket, U, model, out_dim, max_over_ranks, barrier)

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
                     "tower_tflops_incl_all_kernels": B * args.steps * FLOPS_PER_IMAGE_BY_MODEL[args.model] / (ms_total * 1e-3) / 1e12},
        "loss": loss_value,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if parity is not None:
        line["parity"], line["strong"] = parity, strong
    del x, noise
    torch.cuda.empty_cache()
    if world == 1 and not args.no_secondary:
        line["secondary"] = secondary_measurements(args, dev, pk, disturb_rows)
        line["eager_gpu_baseline"] = eager_gpu_baseline(args, dev)
    if not args.no_cpu_baseline and world == 1:
        torch.set_num_threads(os.cpu_count() or 1)
        step, kind = reference_step_fn("cpu", args.model, args.severity)
        n = args.cpu_sample
        frames, zz = cpu_inputs(n, hw, out_dim)
        with torch.no_grad():
            step(frames[:16], zz[:16])
            t0 = time.perf_counter()
            step(frames, zz)
            dt = time.perf_counter() - t0
        what = ("the reference's own functions (oracle/_ref: shared/disturbances_gpu.py, shared/clip_ppo_utils.py, unmodified) over "
                "the fp32 tower of oracle/vit.py" if kind == "reference" else "the fp32 CPU oracle port")
        line["cpu_baseline"] = {"value": n / dt, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": kind,
                                "sample": f"{n} frames of the same workload through {what} (all host threads), {dt:.1f} s"}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def multi_gpu_checks(args, dist, dev, rank, world, path, agent, bucket, U, model, out_dim, max_over_ranks, barrier):
    """(parity, strong).  parity: every rank runs the path on ITS shard of one seeded global batch (CHUNK frames per rank); rank 0
    also runs the whole batch alone.  Embeddings must agree exactly (the tower is batch-invariant bit for bit); the
    GradBucket-averaged agent gradients must equal the single-rank full-batch gradients up to fp32 summation order.
    strong: the same step with the TOTAL work fixed at --batch frames (--batch / N per GPU)."""
    hw = args.hw
    n = CHUNK
    x8, nz, zz = seeded_inputs(dev, rank * n, n, hw, out_dim)
    xs = x8.float() / 255.0
    with torch.no_grad():
        _, emb = path(xs, nz, zz)
    gathered = [torch.empty_like(emb) for _ in range(world)]
    dist.all_gather(gathered, emb)
    # data-parallel gradient: local mean loss, backward, bucket average
    agent.zero_grad(set_to_none=False)
    obs84 = torch.nn.functional.interpolate(xs, size=(84, 84), mode="bilinear", align_corners=False)
    h, _, _ = agent(obs84)
    U.compute_cosine_embedding_loss(h, emb).backward()
    bucket.all_reduce_mean()
    dp_grads = torch.cat([p.grad.reshape(-1) for p in bucket.params]).clone()
    parity = None
    if rank == 0:
        X8, NZ, ZZ = seeded_inputs(dev, 0, n * world, hw, out_dim)
        XS = X8.float() / 255.0
        with torch.no_grad():
            _, emb_all = path(XS, NZ, ZZ)
        emb_max_abs = float((torch.cat(gathered) - emb_all).abs().max().item())
        agent.zero_grad(set_to_none=False)
        H, _, _ = agent(torch.nn.functional.interpolate(XS, size=(84, 84), mode="bilinear", align_corners=False))
        U.compute_cosine_embedding_loss(H, emb_all).backward()
        full = torch.cat([p.grad.reshape(-1) for p in bucket.params])
        grad_rel = float(((dp_grads - full).abs().max() / full.abs().max()).item())
        # what fp32 summation order alone does to the same full-batch gradient: the rows in reversed order, one rank
        agent.zero_grad(set_to_none=False)
        Hr, _, _ = agent(torch.nn.functional.interpolate(XS.flip(0), size=(84, 84), mode="bilinear", align_corners=False))
        U.compute_cosine_embedding_loss(Hr, emb_all.flip(0)).backward()
        rev = torch.cat([p.grad.reshape(-1) for p in bucket.params])
        reorder_rel = float(((rev - full).abs().max() / full.abs().max()).item())
        parity = {"emb_max_abs": emb_max_abs, "grad_rel": grad_rel, "grad_rel_of_row_order_alone": reorder_rel, "frames": n * world,
                  "what": f"{world} ranks x {n}-frame shards of one seeded global batch vs rank 0 running all {n * world} frames alone; "
                          "grad = NatureCNN agent's alignment-loss gradients after GradBucket.all_reduce_mean vs the full-batch backward"}
        del X8, NZ, ZZ, XS
    del x8, nz, zz, xs
    torch.cuda.empty_cache()
    # strong scaling: --batch frames in total
    per = max(CHUNK, args.batch // world // CHUNK * CHUNK)
    x8, nz, zz = seeded_inputs(dev, rank * per, per, hw, out_dim)
    xs = x8.float() / 255.0

    def step():
        loss, _ = path(xs, nz, zz)
        bucket.all_reduce_mean()
        return loss
    for _ in range(args.warmup):
        step()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        step()
    b.record()
    barrier()
    ms = max_over_ranks(a.elapsed_time(b))
    strong = {"value": world * per * args.steps / (ms * 1e-3), "unit": "frames/s", "global_batch": world * per,
              "per_gpu_batch": per, "ms_per_step": ms / args.steps, "scaling": "strong"}
    return parity, strong


def disturb_hbm_rows(dev, pk):
    """The fused disturbance kernel alone against the HBM roofline: 12 algorithmic bytes per element (x + supplied noise + out)."""
    from shared.disturbances_gpu import DisturbanceWrapperGPU
    from shared.disturbance_types import DisturbanceSeverity
    rows = []
    for sev, shape in (("MODERATE", (4096, 3, 224, 224)), ("SEVERE", (4096, 3, 224, 224)), ("HARD", (16384, 1, 84, 84)),
                       ("SEVERE", (16384, 3, 84, 84))):
        g = torch.Generator(device=dev).manual_seed(5)
        xx = torch.rand(shape, device=dev, generator=g)
        nn_ = torch.randn(shape, device=dev, generator=g)
        w = DisturbanceWrapperGPU(device=dev, seed=3, severity=DisturbanceSeverity[sev])
        win = (shape[2] // 5, shape[3] // 4)
        ms = event_ms(lambda: w.apply_disturbances(xx, noise=nn_, contrast_factor=1.1, cutout_start=win), 10, warm=3)
        gbs = 12.0 * xx.numel() / (ms * 1e-3) / 1e9
        rows.append({"severity": sev, "shape": list(shape), "us": round(ms * 1e3, 1), "gb_per_s": round(gbs, 1),
                     "frac_of_hbm_peak": round(gbs / pk["hbm"], 4)})
        if sev == "MODERATE":       # the opt-in variants of the bench shape: in-kernel Philox noise (8 B / element), + uint8 frames (5 B)
            ms = event_ms(lambda: w.apply_disturbances(xx, noise_seed=11, contrast_factor=1.1, cutout_start=win, out_scale=255.0), 10, warm=3)
            rows.append({"severity": sev, "shape": list(shape), "variant": "in-kernel Philox noise + out_scale, 8 B/element",
                         "us": round(ms * 1e3, 1), "gb_per_s": round(8.0 * xx.numel() / (ms * 1e-3) / 1e9, 1),
                         "frac_of_hbm_peak": round(8.0 * xx.numel() / (ms * 1e-3) / 1e9 / pk["hbm"], 4)})
        del xx, nn_
    torch.cuda.empty_cache()
    return rows


def secondary_measurements(args, dev, pk, disturb_rows):
    """Short CUDA-event measurements of the other BASELINE configs (rank 0, N = 1), each reproducible from tools/."""
    import shared.clip_ppo_utils as U
    from shared.disturbances_gpu import DisturbanceWrapperGPU
    from shared.disturbance_types import DisturbanceSeverity
    from clip_ppo_b200 import rollout as R
    out = {}
    out["disturb_hbm"] = {"bytes_per_element": 12, "peak_gb_per_s": pk["hbm"], "peak_source": pk["source"], "rows": disturb_rows,
                          "when": "first thing in the run (kernel timed alone, 10 launches after 3 warm-ups each)"}
    # (1b) BASELINE configs[1]: the learner side of one MiniGrid CLIP-PPO iteration (E = 64 envs x 128 steps, MODERATE)
    mg = minigrid_iteration(dev)
    mg_ref = minigrid_iteration(dev, impl="reference")
    out["minigrid_iteration"] = {
        "config": "BASELINE configs[1]: the learner side of one MiniGrid CLIP-PPO iteration (everything but the environment): E = 64 envs x "
                  "T = 128 steps of 84x84x3 uint8 frames, MODERATE disturbances at every env step, uint8 rollout store, 8192 CLIP "
                  "embeddings per iteration, 4 epochs x 4 minibatches of 2048 (policy forward + backward on csrc/policy.cu, fused GAE / "
                  "PPO-loss / alignment-loss kernels, Adam); synthetic frames stand in for the renderer",
        **mg,
        "reference_eager_same_gpu": {**mg_ref, "what": "the script's own statements over the reference's modules (oracle/_ref) on this GPU: "
                                     "torchvision disturbances, fp32 store, nn.Sequential encoder with two forwards per CLIP minibatch, "
                                     "the 128-step GAE loop, the PPO-loss expressions with their .item() sync, fp16 eager tower"}}
    # (2) BASELINE configs[3]: Atari stacks (84x84 gray x 4, HARD), one GPU's share of 256 envs x 128 steps over 8 GPUs
    model = U.load_clip_model("ViT-B/32", device=dev)
    # (1c) the headline step with the opt-in exact shortcut CLIPPPO_VIT_CLS_LAST_BLOCK=1 (NOT the headline: the headline pass runs
    # every token through every block, as SURVEY.md section 8d counts its FLOPs)
    if args.model == "ViT-B/32":
        nB = args.batch
        x8, nz, zz = seeded_inputs(dev, 0, nB, 224, 512)
        xs = x8.float() / 255.0
        wm = DisturbanceWrapperGPU(device=dev, seed=3, severity=DisturbanceSeverity.MODERATE)

        def b32():
            d = wm.apply_disturbances(xs, noise=nz, contrast_factor=1.1, cutout_start=(44, 56), out_scale=255.0)
            e = U.generate_clip_embeddings(U.AblationMode.NONE, model, "image", nB, dev, images=d)
            return U.compute_cosine_embedding_loss(zz, e)
        t_full, t_cls = [], []
        for _ in range(3):                                      # interleaved: the part drifts with temperature under its power cap
            os.environ.pop("CLIPPPO_VIT_CLS_LAST_BLOCK", None)
            t_full.append(event_ms(b32, 3, warm=1))
            os.environ["CLIPPPO_VIT_CLS_LAST_BLOCK"] = "1"
            t_cls.append(event_ms(b32, 3, warm=1))
        os.environ.pop("CLIPPPO_VIT_CLS_LAST_BLOCK", None)
        ms_full, ms_cls = statistics.median(t_full), statistics.median(t_cls)
        out["cls_rows_only_in_last_block"] = {
            "what": "opt-in, exact: after the last block's attention only the class-token rows go through out_proj / c_fc / c_proj "
                    "(VisionTransformer.forward reads x[:, 0, :] only; embeddings bitwise equal, tests/test_edge_cases_gpu.py); "
                    "the same public calls, same box, three interleaved rounds of three steps each, medians",
            "frames": nB, "ms_full": round(ms_full, 2), "ms": round(ms_cls, 2),
            "frames_per_s_full": round(nB / (ms_full * 1e-3), 1), "frames_per_s": round(nB / (ms_cls * 1e-3), 1)}
        del x8, nz, zz, xs
    stacks = 4096
    g = torch.Generator(device=dev).manual_seed(6)
    obs = torch.randint(0, 256, (stacks, 4, 84, 84), device=dev, generator=g).float()
    zc = torch.relu(torch.randn(stacks, 512, device=dev, generator=g))
    proj = torch.nn.Linear(2048, 512).to(dev)                       # temporal_projection (clip_ppo_atari.py:184-187)
    w = DisturbanceWrapperGPU(device=dev, seed=3, severity=DisturbanceSeverity.HARD)

    def atari():
        with torch.no_grad():
            dobs = R.disturb_atari_stack(w, obs)                                        # clip_ppo_atari.py:568-584
            rgb = R.convert_atari_frames_for_clip(dobs / 255.0)                          # :249-269, :661 (the double /255 quirk)
            e = R.process_multiframe_clip_embeddings(rgb, model, U.AblationMode.NONE, "image", stacks, dev)
            return U.compute_cosine_embedding_loss(zc, proj(e))                          # :730
    ms = event_ms(atari, 3)
    out["atari_84_to_224"] = {"config": "BASELINE configs[3]: Atari stacks 4 x 84x84 gray, HARD disturbances per stacked frame, "
                                        "84 -> 224 up-sampling inside the fused preprocess, 4 ViT-B/32 embeddings per stack, "
                                        "temporal projection, alignment loss; one GPU's share (32 envs x 128 steps)",
                              "stacks": stacks, "clip_frames": 4 * stacks, "ms": round(ms, 2),
                              "clip_frames_per_s": round(4 * stacks / (ms * 1e-3), 1), "stacks_per_s": round(stacks / (ms * 1e-3), 1)}
    del obs, zc, model
    torch.cuda.empty_cache()
    # (3) BASELINE configs[4]'s tower variant: ViT-L/14 (T = 257: the tcgen05 attention kernel), same step at 1024 frames
    if args.model != "ViT-L/14":
        model = U.load_clip_model("ViT-L/14", device=dev)
        nL = 1024
        x8, nz, zz = seeded_inputs(dev, 0, nL, 224, 768)
        xs = x8.float() / 255.0
        w = DisturbanceWrapperGPU(device=dev, seed=3, severity=DisturbanceSeverity.MODERATE)

        def l14():
            d = w.apply_disturbances(xs, noise=nz, contrast_factor=1.1, cutout_start=(44, 56), out_scale=255.0)
            e = U.generate_clip_embeddings(U.AblationMode.NONE, model, "image", nL, dev, images=d)
            return U.compute_cosine_embedding_loss(zz, e)
        ms = event_ms(l14, 3)
        out["vit_l14"] = {"config": "BASELINE configs[4] tower variant: ViT-L/14 (width 1024, 24 blocks, 257 tokens), 1024 frames of 224x224x3, "
                                    "MODERATE, same public calls", "frames": nL, "ms": round(ms, 2),
                          "frames_per_s": round(nL / (ms * 1e-3), 1),
                          "tower_tflops_incl_all_kernels": round(nL * FLOPS_PER_IMAGE_BY_MODEL["ViT-L/14"] / (ms * 1e-3) / 1e12, 1)}
        os.environ["CLIPPPO_VIT_CLS_LAST_BLOCK"] = "1"
        ms_c = event_ms(l14, 3)
        os.environ.pop("CLIPPPO_VIT_CLS_LAST_BLOCK", None)
        out["vit_l14"]["cls_rows_only_in_last_block"] = {"ms": round(ms_c, 2), "frames_per_s": round(nL / (ms_c * 1e-3), 1)}
        del x8, nz, zz, xs, model
        torch.cuda.empty_cache()
    return out


class MiniGridAgent(torch.nn.Module):
    """The reference's MiniGrid Agent (clip_ppo_minigrid.py:213-271) with its encoder on the native kernels."""

    def __init__(self, n_actions: int = 7):
        super().__init__()
        from clip_ppo_b200.policy import NatureCNN
        self.network = NatureCNN(3)
        self.actor = torch.nn.Linear(512, n_actions)
        self.critic = torch.nn.Linear(512, 1)

    def _pre(self, x):                       # [B,H,W,C] 0..255 -> the NCHW view the encoder reads (the / 255 is folded into its first kernel)
        return x.permute(0, 3, 1, 2)

    def _get_features(self, x):
        return self.network(x, in_scale=1.0 / 255.0)


def minigrid_iteration(dev, E: int = 64, T: int = 128, epochs: int = 4, minibatches: int = 4, impl: str = "b200"):
    """Everything of one CLIP-PPO iteration except the environment itself (out of scope), at BASELINE configs[1]'s shape:
    T env steps of [disturb the E frames -> store -> policy forward], GAE, the iteration's CLIP embeddings (84 -> 224), epochs x
    minibatches updates [policy forward, latents, PPO loss, alignment loss on every CLIP_LOSS_FREQUENCY-th minibatch, backward,
    clip_grad_norm_, Adam] (clip_ppo_minigrid.py:378-410, 437-450, 459-470, 486-564).  Synthetic uint8 frames stand in for the renderer.
      impl "b200"       this repository's drop-in calls: disturb_minigrid_obs (one launch, uint8 out), uint8 store, native NatureCNN,
                        the env step's policy forward as a CUDA-graph replay (rollout.PolicyStepGraph), one encoder forward per
                        minibatch, fused GAE / PPO-loss / cosine-loss kernels, the sm_100a tower;
      impl "reference"  the script's own statements over the reference's own modules (oracle/_ref) on the same GPU: torchvision
                        disturbances, fp32 store, nn.Sequential encoder (two forwards per minibatch), the GAE loop, the PPO-loss
                        expressions with their .item() sync, fp16 eager tower."""
    from torch.distributions.categorical import Categorical
    torch.manual_seed(11)
    g = torch.Generator(device=dev).manual_seed(12)
    frames = torch.randint(0, 256, (T, E, 84, 84, 3), device=dev, generator=g, dtype=torch.uint8)
    rewards = (torch.rand(T, E, device=dev, generator=g) < 0.05).float()
    dones = (torch.rand(T, E, device=dev, generator=g) < 0.02).float()
    N_ = T * E
    mbsz = N_ // minibatches
    if impl == "b200":
        import shared.clip_ppo_utils as U
        from shared.disturbances_gpu import DisturbanceWrapperGPU
        from shared.disturbance_types import DisturbanceSeverity
        from clip_ppo_b200 import rollout as R
        model = U.load_clip_model("ViT-B/32", device=dev)
        agent = MiniGridAgent().to(dev)
        w = DisturbanceWrapperGPU(device=dev, seed=5, severity=DisturbanceSeverity.MODERATE)
        store = R.ObsStoreU8(T, E, (84, 84, 3), device=dev)
    else:
        from oracle import build_ref, losses as ol, vit as ov
        if not build_ref.available():
            return {"unavailable": "oracle/_ref not staged"}
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from eager_tower import EagerCLIP
        ref = build_ref.load(EagerCLIP(ov.random_state_dict(ov.VIT_B32, 0), dev))
        U, DG, DT = ref.clip_ppo_utils, ref.disturbances_gpu, ref.disturbance_types
        model = U.load_clip_model("ViT-B/32", dev)
        nn = torch.nn
        agent = nn.Module()
        agent.network = nn.Sequential(nn.Conv2d(3, 32, 8, stride=4), nn.ReLU(), nn.Conv2d(32, 64, 4, stride=2), nn.ReLU(),
                                      nn.Conv2d(64, 64, 3, stride=1), nn.ReLU(), nn.Flatten(), nn.Linear(64 * 7 * 7, 512), nn.ReLU())
        agent.actor, agent.critic = nn.Linear(512, 7), nn.Linear(512, 1)
        agent = agent.to(dev)
        w = DG.DisturbanceWrapperGPU(device=dev, seed=5, severity=DT.DisturbanceSeverity.MODERATE)
        store = torch.zeros((T, E, 84, 84, 3), device=dev)                                       # clip_ppo_minigrid.py:346
        hidden = lambda x: agent.network(x.permute(0, 3, 1, 2) / 255.0)                          # :258-262
    opt = torch.optim.Adam(agent.parameters(), lr=2.5e-4, eps=1e-5)
    lam = U.get_clip_lambda_with_warmup(1e-5, 8, 16)

    policy_step = R.PolicyStepGraph(agent, frames[0]) if impl == "b200" else None      # one graph launch per env step

    def iteration():
        values = torch.empty(T, E, device=dev)
        logprobs = torch.empty(T, E, device=dev)
        actions = torch.empty(T, E, device=dev, dtype=torch.long)
        for t in range(T):
            if impl == "b200":
                d = R.disturb_minigrid_obs(w, frames[t])                                         # clip_ppo_minigrid.py:381-388, one launch
                store[t] = d
                a, lp, _, v = policy_step(d)                                                     # :395-399, graph replay
            else:
                x = frames[t].float() / 255.0                                                    # :383-388, as written
                x = w.apply_disturbances(x.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
                d = (x * 255).byte().float()
                store[t] = d
                with torch.no_grad():
                    h = hidden(d)
                    probs = Categorical(logits=agent.actor(h))
                    a = probs.sample()
                    lp, v = probs.log_prob(a), agent.critic(h)
            values[t], logprobs[t], actions[t] = v.flatten(), lp, a
        with torch.no_grad():
            if impl == "b200":
                next_value = agent.critic(agent._get_features(agent._pre(frames[0].float()))).reshape(1, -1)
                adv, ret = R.compute_gae(rewards, values, dones, next_value, dones[0], 0.99, 0.95)   # :437-450, one launch
                emb = U.generate_clip_embeddings(U.AblationMode.NONE, model, "image", N_, dev,
                                                 images=store.clip_images(torch.arange(N_, device=dev)))   # :459-470
            else:
                next_value = agent.critic(hidden(frames[0].float())).reshape(1, -1)
                adv, ret = ol.gae(rewards, values, dones, next_value, dones[0], 0.99, 0.95)          # the script's T-step loop
                b_obs = store.reshape((-1, 84, 84, 3))
                emb = torch.cat([U.generate_clip_embeddings(U.AblationMode.NONE, model, "image", 2048, dev,
                                                            images=b_obs[i:i + 2048].permute(0, 3, 1, 2))
                                 for i in range(0, N_, 2048)])       # in minibatch-sized pieces: the stock path materialises fp32 224x224 frames
        b_lp, b_adv, b_ret, b_val, b_act = logprobs.reshape(-1), adv.reshape(-1), ret.reshape(-1), values.reshape(-1), actions.reshape(-1)
        counter = 0
        loss = None
        for _ in range(epochs):
            perm = torch.randperm(N_, device=dev)
            for i in range(minibatches):
                mb = perm[i * mbsz:(i + 1) * mbsz]
                do_clip = counter % U.CLIP_LOSS_FREQUENCY == 0
                if impl == "b200":
                    _, nlp, ent, nv, lat = R.action_value_and_latents(agent, store.policy_input(mb), b_act[mb])     # :498, :534 in one forward
                    cl = U.compute_cosine_embedding_loss(lat, emb[mb]) if do_clip else None
                    out = R.ppo_minibatch_loss(nlp, ent, nv.flatten(), b_lp[mb], b_adv[mb], b_ret[mb], b_val[mb], cl, lam)
                else:
                    obs_mb = store.reshape((-1, 84, 84, 3))[mb]
                    h = hidden(obs_mb)                                                               # :498 get_action_and_value
                    probs = Categorical(logits=agent.actor(h))
                    nlp, ent, nv = probs.log_prob(b_act[mb]), probs.entropy(), agent.critic(h)
                    cl = 0.0
                    if do_clip:
                        lat = hidden(obs_mb).detach()                                                # :534 get_latent_representation: a second forward
                        cl = U.compute_cosine_embedding_loss(lat, emb[mb])
                    out = ol.ppo_loss(nlp, ent, nv, b_lp[mb], b_adv[mb], b_ret[mb], b_val[mb], cl, lam)
                    out["clipfrac"].item()                                                           # :505: the script's host sync
                opt.zero_grad(set_to_none=True)
                out["loss"].backward()
                torch.nn.utils.clip_grad_norm_(agent.parameters(), 0.5)
                opt.step()
                counter += 1
                loss = out["loss"]
        return loss

    iteration()
    ms = event_ms(iteration, 2, warm=0)
    res = {"ms_per_iteration": round(ms, 2), "env_frames_per_s": round(N_ / (ms * 1e-3), 1)}
    del model, agent, frames, store
    torch.cuda.empty_cache()
    return res


def eager_gpu_baseline(args, dev):
    """The reference's own functions on this GPU: shared/disturbances_gpu.py (torchvision transforms) and
    shared/clip_ppo_utils.py (F.interpolate, normalise, encode_image, F.normalize, cosine loss) out of oracle/_ref, with the
    fp16 eager tower clip.load would hand them (tools/eager_tower.py).  Same frames, 1024 per step."""
    try:
        step, kind = reference_step_fn(str(dev), args.model, args.severity)
    except Exception as e:          # never lose the headline over the baseline leg
        return {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    if step is None:
        return {"unavailable": "oracle/_ref is not staged (run python oracle/build_ref.py where /root/reference exists)"}
    n = 1024
    x8, _, zz = seeded_inputs(dev, 0, n, args.hw, 512 if args.model == "ViT-B/32" else 768, want_noise=False)
    xs = x8.float() / 255.0
    with torch.no_grad():
        ms = event_ms(lambda: step(xs, zz), 3)
    return {"value": n / (ms * 1e-3), "unit": "frames/s", "frames_per_step": n, "ms_per_step": round(ms, 2), "kind": kind,
            "what": "reference's own DisturbanceWrapperGPU.apply_disturbances (device randn, torchvision ops) -> *255 -> "
                    "generate_clip_embeddings (fp16 eager tower: cuDNN conv, cuBLAS, SDPA) -> compute_cosine_embedding_loss, device-resident frames"}


if __name__ == "__main__":
    main()
