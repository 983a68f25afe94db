#!/usr/bin/env python
"""Pinned-host read-back microbenchmark: what bounds bench.py's `e2e` figure.

Every rank copies SIZE bytes device -> pinned host memory back to back for ~0.4 s on its own GPU, all ranks at the same
time; prints per-rank and aggregate GB/s for ordinary pinned memory and for write-combined pinned memory
(cudaHostAllocWriteCombined), at the e2e leg's copy size (10.8 MB = every output of a 4096-env step) and at 2 / 64 MB.
One cudaMemcpyAsync per copy.

    python scripts/bench_d2h.py                                   # N = 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/bench_d2h.py
"""
import ctypes as C
import glob
import json
import os
import sys
import time

import torch


def cudart():
    for pat in (os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*"),
                os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"),
                "/usr/local/cuda/lib64/libcudart.so*"):
        hits = sorted(glob.glob(pat))
        if hits:
            return C.CDLL(hits[0])
    raise RuntimeError("libcudart not found")


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    rt = cudart()
    rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
    rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    rt.cudaFreeHost.argtypes = [C.c_void_p]
    stream = torch.cuda.current_stream(dev)
    out = {"n_gpus": world, "rows": []}
    for mb in (2.0, 10.78, 64.0):
        n = int(mb * 1e6) // 4 * 4
        src = torch.empty(n, dtype=torch.uint8, device=dev)
        for kind, flags in (("pinned", 0), ("pinned write-combined", 4)):
            p = C.c_void_p()
            assert rt.cudaHostAlloc(C.byref(p), n, flags) == 0
            for _ in range(3):
                rt.cudaMemcpyAsync(p, src.data_ptr(), n, 2, stream.cuda_stream)
            torch.cuda.synchronize(dev)
            if dist:
                dist.barrier()
            reps = max(4, int(0.4 * 50e9 / n))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                rt.cudaMemcpyAsync(p, src.data_ptr(), n, 2, stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize(dev)
            gbps = n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
            t = torch.tensor([gbps], dtype=torch.float64, device=dev)
            mn = t.clone()
            if dist:
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                dist.all_reduce(mn, op=dist.ReduceOp.MIN)
            out["rows"].append({"MB": mb, "host_memory": kind, "aggregate_GBps": t.item(), "slowest_rank_GBps": mn.item()})
            rt.cudaFreeHost(p)
    if rank == 0:
        print(json.dumps(out))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
