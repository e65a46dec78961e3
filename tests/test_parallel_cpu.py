"""world_size-2 gloo tests (CPU) of the N>1 host logic: batch sharding and the logits gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from slime_b200.parallel import LogitsGather, balanced_order, gather_logits, shard_bounds


def test_shard_bounds_cover_batch():
    for n in (0, 1, 7, 8, 33, 64):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_bounds(n, r, world)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))
            sizes = [shard_bounds(n, r, world)[1] - shard_bounds(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_balanced_order_is_a_permutation_and_balances():
    costs = [17, 1, 5, 5, 10, 2, 17, 1, 9, 3]
    order = balanced_order(costs, 2)
    assert sorted(order) == list(range(len(costs)))
    lo, hi = shard_bounds(len(costs), 0, 2)
    a = sum(costs[i] for i in order[lo:hi])
    b = sum(costs[i] for i in order[hi:])
    assert abs(a - b) <= max(costs)


def _worker(rank, world, port, n_samples, vocab, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(n_samples * vocab, dtype=torch.float32).view(n_samples, vocab)
    lo, hi = shard_bounds(n_samples, rank, world)
    got = gather_logits(full[lo:hi].clone(), n_samples)
    ok = torch.equal(got, full)
    if n_samples % world == 0:  # the side-stream gatherer of the benchmark (equal blocks), rotating result buffers
        lg = LogitsGather(n_samples // world, vocab, "cpu")
        for step in range(3):
            slot = lg.submit(full[lo:hi] + step)
            ok = ok and torch.equal(lg.result(slot), full + step)
        lg.drain()
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_samples", [4, 5])
def test_gather_logits_gloo_world2(n_samples):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_samples, 16, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
