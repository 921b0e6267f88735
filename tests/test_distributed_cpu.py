"""The N>1 path on CPU: two gloo ranks, observations sharded, gradients all-reduced once."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from clip_ppo_b200.distributed import GradBucket, global_advantage_stats, shard_range


def test_shard_range_partitions():
    for n in (1, 7, 64, 255, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 4))


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(1)
        x, y = torch.randn(64, 16, generator=g), torch.randn(64, 4, generator=g)
        adv = torch.randn(64, generator=g)
        lo, hi = shard_range(64, rank, world)
        m = _model()
        bucket = GradBucket(m.parameters())
        loss = ((m(x[lo:hi]) - y[lo:hi]) ** 2).mean()
        loss.backward()
        bucket.all_reduce_mean()
        mean, std = global_advantage_stats(adv[lo:hi])
        ret[rank] = ([p.grad.clone() for p in m.parameters()], mean.item(), std.item())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gradient_allreduce_matches_single_process():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        results = {k: v for k, v in ret.items()}
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(64, 16, generator=g), torch.randn(64, 4, generator=g)
    adv = torch.randn(64, generator=g)
    m = _model()
    ((m(x) - y) ** 2).mean().backward()                 # equal shard sizes => mean of shard means == global mean
    ref = [p.grad for p in m.parameters()]
    for rank in range(world):
        grads, mean, std = results[rank]
        for a, b in zip(grads, ref):
            assert torch.allclose(a, b, atol=1e-6)
        assert abs(mean - adv.mean().item()) < 1e-6 and abs(std - adv.std().item()) < 1e-5
    for a, b in zip(results[0][0], results[1][0]):
        assert torch.equal(a, b)                        # ranks stay in lock-step


def test_gradbucket_is_noop_without_process_group():
    m = _model()
    (m(torch.ones(2, 16)).sum()).backward()
    before = [p.grad.clone() for p in m.parameters()]
    GradBucket(m.parameters()).all_reduce_mean()
    assert all(torch.equal(a, p.grad) for a, p in zip(before, m.parameters()))
