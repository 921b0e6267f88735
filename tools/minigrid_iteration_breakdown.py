import os, sys, time
sys.path.insert(0, '.')
os.environ.setdefault("CLIPPPO_ALLOW_RANDOM_WEIGHTS", "1")
import torch, bench
import shared.clip_ppo_utils as U
from shared.disturbances_gpu import DisturbanceWrapperGPU
from shared.disturbance_types import DisturbanceSeverity
from clip_ppo_b200 import rollout as R
dev = torch.device("cuda", 0)
T, E = 128, 64
model = U.load_clip_model("ViT-B/32", device=dev)
agent = bench.MiniGridAgent().to(dev)
w = DisturbanceWrapperGPU(device=dev, seed=5, severity=DisturbanceSeverity.MODERATE)
frames = torch.randint(0, 256, (T, E, 84, 84, 3), device=dev, dtype=torch.uint8)
store = R.ObsStoreU8(T, E, (84, 84, 3), device=dev)
def sync_time(fn, n=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
def roll_disturb():
    for t in range(T): store[t] = R.disturb_minigrid_obs(w, frames[t])
def roll_policy():
    with torch.no_grad():
        for t in range(T): R.action_value_and_latents(agent, store.data[t].float(), None)
def emb():
    U.generate_clip_embeddings(U.AblationMode.NONE, model, "image", T * E, dev, images=store.clip_images(torch.arange(T * E, device=dev)))
opt = torch.optim.Adam(agent.parameters(), lr=2.5e-4, eps=1e-5)
acts = torch.randint(0, 7, (T * E,), device=dev)
e = torch.nn.functional.normalize(torch.randn(T * E, 512, device=dev), dim=-1)
lp = -torch.rand(T * E, device=dev); adv = torch.randn(T * E, device=dev); ret = torch.randn(T * E, device=dev); val = torch.randn(T * E, device=dev)
def updates():
    perm = torch.randperm(T * E, device=dev)
    for i in range(4):
        mb = perm[i * 2048:(i + 1) * 2048]
        _, nlp, ent, nv, lat = R.action_value_and_latents(agent, store.policy_input(mb), acts[mb])
        cl = U.compute_cosine_embedding_loss(lat, e[mb]) if i == 0 else None
        out = R.ppo_minibatch_loss(nlp, ent, nv.flatten(), lp[mb], adv[mb], ret[mb], val[mb], cl, 1e-5)
        opt.zero_grad(set_to_none=True); out["loss"].backward(); torch.nn.utils.clip_grad_norm_(agent.parameters(), 0.5); opt.step()
print(f"128 x disturb_minigrid_obs + store: {sync_time(roll_disturb):7.2f} ms")
print(f"128 x policy forward (64 frames):   {sync_time(roll_policy):7.2f} ms")
print(f"8192 CLIP embeddings (84->224):     {sync_time(emb):7.2f} ms")
print(f"one epoch = 4 minibatch updates:    {sync_time(updates):7.2f} ms  (x 4 epochs)")
