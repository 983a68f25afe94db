#!/usr/bin/env python
"""Peer-memory gather microbenchmark (csrc/peer_gather.cu) against NCCL's all-gather, at BASELINE config 4's exchange:
every rank owns 65536 / N rows of body_pos [.,15,3] and obs [.,441] and every rank ends up with all of them.
Sweeps the push kernel's grid size; one CUDA-event pair around `reps` back-to-back exchanges, max over ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/bench_peer.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from parc_b200 import sharding
    n, J, P = 65536, 15, 441
    lo, hi = sharding.shard_bounds(n, rank, world)
    g = torch.Generator(device=dev).manual_seed(rank)
    local_t = {"body_pos": torch.randn(hi - lo, J, 3, device=dev, generator=g),
               "obs": torch.randn(hi - lo, P, device=dev, generator=g)}
    nbytes = n * (J * 3 + P) * 4
    stream = torch.cuda.current_stream(dev)
    reps = 30

    def timed(fn):
        for _ in range(5):
            fn()
        torch.cuda.synchronize(dev)
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    rows = []
    ag = [sharding.AllGatherPlan(local_t[k], n) for k in ("body_pos", "obs")]
    ms = timed(lambda: [a.run() for a in ag])
    rows.append({"what": "NCCL all_gather_into_tensor x2", "ms": ms, "algbw_GBps": nbytes / ms / 1e6})
    ref = {k: a.out.clone() for k, a in zip(("body_pos", "obs"), ag)}
    for mc in (True, False):
        for blocks in (16, 32, 64, 128, 256, 512):
            pg = sharding.PeerGather({"body_pos": (J, 3), "obs": (P,)}, n, dev, num_blocks=blocks, use_multicast=mc)
            ms = timed(lambda: pg.push(local_t))
            same = bool(torch.equal(pg.out["body_pos"], ref["body_pos"]) and torch.equal(pg.out["obs"], ref["obs"]))
            rows.append({"what": "parc_peer_push", "multicast": bool(pg.multicast), "blocks": blocks, "ms": ms,
                         "algbw_GBps": nbytes / ms / 1e6, "equals_nccl": same})
            del pg
    if rank == 0:
        print(json.dumps({"n_gpus": world, "bytes_total": nbytes, "rows": rows}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
