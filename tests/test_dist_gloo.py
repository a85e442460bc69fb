"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: scene sharding and the packed EMA all-reduce."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from viewformer_b200.dist import shard_range, allreduce_ema_stats, max_over_ranks


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    counts = torch.randint(0, 5, (32,), generator=g).float()
    esum = torch.randn((8, 32), generator=g)
    # reference pattern: two all-reduces (utils_th.py:50-52)
    c2, e2 = counts.clone(), esum.clone()
    dist.all_reduce(c2)
    dist.all_reduce(e2)
    c1, e1 = allreduce_ema_stats(counts, esum)
    ok = torch.equal(c1, c2) and torch.allclose(e1, e2, atol=0, rtol=0)
    ok = ok and c1.shape == counts.shape and e1.shape == esum.shape
    ok = ok and max_over_ranks(float(rank + 1), "cpu") == float(world)
    lo, hi = shard_range(7, rank, world)
    tot = torch.tensor([hi - lo])
    dist.all_reduce(tot)
    ok = ok and int(tot) == 7
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_packed_ema_allreduce_equals_reference_two_call_pattern():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
