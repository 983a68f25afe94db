"""N > 1 host logic on CPU: world_size-2 gloo processes exercise the sharding / gather / loss-statistics
layer (the kernels themselves need a GPU and are covered by -m gpu)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from parc_b200 import sharding


def test_shard_bounds_partition():
    for n in (0, 1, 7, 4096, 65536, 100003):
        for w in (1, 2, 3, 4, 8):
            cuts = [sharding.shard_bounds(n, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_is_identity():
    t = torch.arange(10.0)
    assert sharding.world() == (0, 1)
    assert torch.equal(sharding.shard(t), t)
    assert torch.equal(sharding.all_gather_shards(t, 10), t)
    s = sharding.reduce_loss_stats({"pen": torch.tensor([1.0, 3.0])})
    assert s["pen"] == {"sum": 4.0, "count": 2, "mean": 2.0, "min": 1.0, "max": 3.0}


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        g = torch.Generator().manual_seed(0)
        full = torch.randn(n, 15, 3, generator=g)                 # replicated "global" tensor
        mine = sharding.shard(full)
        lo, hi = sharding.shard_bounds(n, rank, world_size)
        assert mine.shape[0] == hi - lo and torch.equal(mine, full[lo:hi])
        back = sharding.all_gather_shards(mine * 2.0, n)
        assert torch.equal(back, full * 2.0)
        plan = sharding.AllGatherPlan((mine * 3.0).contiguous(), n)       # fixed-buffer form (config 4's per-step gather)
        assert plan.equal == (n % world_size == 0)
        for _ in range(2):
            assert torch.equal(plan.run(), full * 3.0)
        on0 = sharding.gather_shards_to(mine + 1.0, n, dst=0)
        assert (on0 is None) == (rank != 0)
        if rank == 0:
            assert torch.equal(on0, full + 1.0)
        loss = torch.arange(lo, hi, dtype=torch.float32)           # "per-sample losses" of this shard
        st = sharding.reduce_loss_stats({"pen_loss": loss, "contact_loss": -loss})
        q.put((rank, st["pen_loss"], st["contact_loss"]["min"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [11, 64])
def test_gloo_world2_gather_and_stats(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted([q.get(timeout=5) for _ in range(2)], key=lambda x: x[0])
    exp_sum = float(sum(range(n)))
    for _, pen, cmin in got:                                        # every rank sees the global statistics
        assert pen["sum"] == exp_sum and pen["count"] == n and pen["min"] == 0.0 and pen["max"] == float(n - 1)
        assert cmin == -float(n - 1)
