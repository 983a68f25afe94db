"""Multi-GPU layer: one process per GPU, queries / samples / clips sharded contiguously across ranks.

The path has no exchange step: every (clip id, time) query, every loss sample
and every clip is independent, the frame tables and the global heightfield are replicated.  So the data
path runs with NO collective; `torch.distributed` (NCCL over NVLink on the GPUs, gloo in the CPU tests)
is used only
  * to gather results onto one rank / all ranks when a caller wants them in one place (NCCL, or the library's own
    peer-memory kernels: `PeerGather`), and
  * to reduce loss statistics (sum / min / max / count in fp64), mirroring what the reference's logger
    does with its all-reduce (util/logger.py:164-183, util/mp_util.py:90-105).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of `n` items owned by `rank`; the first n % world ranks get one extra."""
    assert 0 <= rank < world_size
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(t: torch.Tensor, rank: Optional[int] = None, world_size: Optional[int] = None) -> torch.Tensor:
    """This rank's contiguous slice of dim 0 (a view)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    lo, hi = shard_bounds(t.shape[0], rank, world_size)
    return t[lo:hi]


def all_gather_shards(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Inverse of `shard`: every rank ends up with the full [n_total, ...] tensor.  Shards may differ by
    one row, so they are padded to the largest shard for a single all_gather_into_tensor call."""
    rank, w = world()
    if w == 1:
        assert local.shape[0] == n_total
        return local
    per = (n_total + w - 1) // w
    rest = tuple(local.shape[1:])
    padded = local.new_zeros((per,) + rest)
    padded[:local.shape[0]] = local
    out = local.new_empty((w * per,) + rest)
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    pieces = []
    for r in range(w):
        lo, hi = shard_bounds(n_total, r, w)
        pieces.append(out[r * per:r * per + (hi - lo)])
    return torch.cat(pieces, dim=0)


class AllGatherPlan:
    """`all_gather_shards` over FIXED buffers, for callers that gather every step (config 4: each rank's slice of
    `body_pos` / `obs` -> the full batch on every rank).  The global tensor is allocated once; with equal shards the
    collective writes straight into it (one `all_gather_into_tensor`, no staging copy, no concatenation); ragged
    shards go through one padded staging buffer.  `local` must be the same tensor on every call (e.g. a
    `MotionQueryPlan` output)."""

    def __init__(self, local: torch.Tensor, n_total: int, group=None):
        self.rank, self.world = world()
        self.local, self.n_total, self.group = local, int(n_total), group
        rest = tuple(local.shape[1:])
        lo, hi = shard_bounds(self.n_total, self.rank, self.world)
        assert local.shape[0] == hi - lo and local.is_contiguous()
        self.equal = self.n_total % self.world == 0
        if self.world == 1:
            self.out = local
        elif self.equal:
            self.out = local.new_empty((self.n_total,) + rest)
        else:
            per = (self.n_total + self.world - 1) // self.world
            self._padded = local.new_zeros((per,) + rest)
            self._stage = local.new_empty((self.world * per,) + rest)
            self.out = local.new_empty((self.n_total,) + rest)
            self._per = per

    def run(self) -> torch.Tensor:
        """Enqueue the gather behind the work already queued on the current stream; returns the global tensor."""
        if self.world == 1:
            return self.out
        if self.equal:
            dist.all_gather_into_tensor(self.out, self.local, group=self.group)
            return self.out
        self._padded[:self.local.shape[0]].copy_(self.local)
        dist.all_gather_into_tensor(self._stage, self._padded, group=self.group)
        for r in range(self.world):
            lo, hi = shard_bounds(self.n_total, r, self.world)
            self.out[lo:hi].copy_(self._stage[r * self._per:r * self._per + (hi - lo)])
        return self.out


class PeerGather:
    """All-gather of query shards over NVLink / NVSwitch peer memory with the library's own kernels
    (csrc/peer_gather.cu) instead of NCCL: the gathered tensors live in symmetric memory
    (torch.distributed._symmetric_memory allocates and maps it; PyTorch is plumbing here), every rank stores its rows
    into all ranks' copies -- one store to the NVSwitch MULTICAST address where the fabric has one, else one store per
    peer pointer -- and a release / acquire hand-shake on per-block signal slots ends the launch.

        pg = PeerGather({"body_pos": (J, 3), "obs": (P,)}, n_total, device)
        pg.push({"body_pos": local_bp, "obs": local_obs})     # after the query, same stream
        pg.out["body_pos"]                                    # [n_total, J, 3] on every rank
    or, "direct": hand `pg.direct_ptr(name)` to `MotionQueryPlan.redirect_output` so that the query kernel's own
    stores go to the multicast address, and call `pg.barrier()` after the launch.

    Consecutive steps must alternate between two PeerGather objects when a consumer of step s may still be reading
    while step s + 1 is pushed (a rank passes the hand-shake of step s + 1 only after every rank has launched its push
    of step s + 1, which is stream-ordered behind that rank's consumer of step s)."""

    def __init__(self, row_shapes: Dict[str, Tuple[int, ...]], n_total: int, device, group=None, num_blocks: int = 64,
                 use_multicast: Optional[bool] = None, timeout_s: float = 10.0):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        self.rank, self.world = world()
        assert 2 <= self.world <= _lib.PARC_MAX_PEERS, "PeerGather needs an initialised process group of 2..16 ranks"
        self.device = torch.device(device)
        self.n_total = int(n_total)
        self.lo, self.hi = shard_bounds(self.n_total, self.rank, self.world)
        self.num_blocks = int(num_blocks)
        self.names = list(row_shapes)
        assert 1 <= len(self.names) <= _lib.PARC_MAX_PUSH_SEGMENTS
        off, self._off, self._row_bytes = 0, {}, {}
        for k in self.names:
            rb = 4
            for d in row_shapes[k]:
                rb *= int(d)
            self._off[k], self._row_bytes[k] = off, rb
            off = (off + self.n_total * rb + 255) // 256 * 256
        self._slots = self.num_blocks + 1                     # last slot: the stand-alone barrier
        self._sig_off = off
        total = off + 8 * self._slots
        grp = group if group is not None else dist.group.WORLD
        name = grp if isinstance(grp, str) else grp.group_name
        self.buf = symm.empty(total, dtype=torch.uint8, device=self.device)
        self.buf.zero_()
        torch.cuda.synchronize(self.device)
        try:
            self.hdl = symm.rendezvous(self.buf, name)
        except Exception:                                     # older releases want the group enabled first
            symm.enable_symm_mem_for_group(name)
            self.hdl = symm.rendezvous(self.buf, name)
        self.hdl.barrier()
        torch.cuda.synchronize(self.device)
        # A multicast store also loops the rank's own rows back through the switch: world / (world - 1) times the
        # ingress of a peer-pointer push, which in turn issues `world` stores per load.  Measured on B200 (65 536 envs):
        # 2 GPUs 0.197 ms multicast / 0.150 ms peer pointers; 8 GPUs 0.192 / 0.209 ms -- multicast from 8 ranks up.
        if use_multicast is None:
            use_multicast = self.world >= 8
        mc = int(self.hdl.multicast_ptr) if (use_multicast and getattr(self.hdl, "has_multicast_support", False)) else 0
        self.multicast = mc != 0
        self._mc = mc
        self._peers = [int(p) for p in self.hdl.buffer_ptrs]
        assert self._peers[self.rank] == self.buf.data_ptr()
        self.out = {k: self.buf[self._off[k]:self._off[k] + self.n_total * self._row_bytes[k]].view(torch.float32)
                    .view(self.n_total, *row_shapes[k]) for k in self.names}
        self.epoch = torch.zeros(self._slots, dtype=torch.int64, device=self.device)
        sig = _lib.ParcPeerSignals()
        sig.multicast_signal = (mc + self._sig_off) if mc else None
        for r in range(self.world):
            sig.peer_signal[r] = self._peers[r] + self._sig_off
        sig.local_signal = self.buf.data_ptr() + self._sig_off
        sig.epoch = self.epoch.data_ptr()
        sig.world, sig.num_slots, sig.rank = self.world, self._slots, self.rank
        # a rank that never launches its side must not hang this GPU: the hand-shake gives up after `timeout_s`
        self.timeout_flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        sig.timeout_ns, sig.timeout_flag = int(timeout_s * 1e9), self.timeout_flag.data_ptr()
        self._sig = sig
        self._segs, self._segs_key = None, None
        self._lib, self._C = _lib, C

    def direct_ptr(self, name: str) -> int:
        """Multicast address of THIS rank's rows of gathered tensor `name` (for stores issued by the producing kernel)."""
        if not self.multicast:
            raise RuntimeError("direct stores need an NVSwitch multicast mapping; use push()")
        return self._mc + self._off[name] + self.lo * self._row_bytes[name]

    def _stream(self, stream):
        return torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream

    def push(self, local: Dict[str, torch.Tensor], stream: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """Enqueue the push of this rank's shards (behind the work already on the stream); when it has run, `self.out`
        holds every rank's rows on this rank."""
        key = tuple((k, local[k].data_ptr()) for k in self.names)
        if key != self._segs_key:
            segs = (self._lib.ParcPeerSegment * len(self.names))()
            for i, k in enumerate(self.names):
                t = local[k]
                assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.shape[0] == self.hi - self.lo
                assert t.numel() * 4 == (self.hi - self.lo) * self._row_bytes[k], (k, tuple(t.shape))
                rel = self._off[k] + self.lo * self._row_bytes[k]
                segs[i].src = t.data_ptr()
                segs[i].dst_multicast = (self._mc + rel) if self.multicast else None
                for r in range(self.world):
                    segs[i].dst_peer[r] = self._peers[r] + rel
                segs[i].bytes = t.numel() * 4
            self._segs, self._segs_key, self._keep = segs, key, [local[k] for k in self.names]
        rc = self._lib.load().parc_peer_push(self._segs, len(self.names), self._C.byref(self._sig), self.num_blocks,
                                             self._stream(stream))
        self._lib.check(rc, "parc_peer_push")
        return self.out

    def check(self):
        """Raise if a hand-shake since the last check gave up waiting for a peer (synchronises)."""
        if int(self.timeout_flag.item()):
            self.timeout_flag.zero_()
            raise RuntimeError("PeerGather: a rank did not arrive within the time-out; the gathered tensors are incomplete")

    def barrier(self, stream: Optional[int] = None):
        """Hand-shake alone: every rank's earlier stores into the gathered tensors are visible once it has run."""
        rc = self._lib.load().parc_peer_barrier(self._C.byref(self._sig), self.num_blocks, self._stream(stream))
        self._lib.check(rc, "parc_peer_barrier")


def gather_shards_to(local: torch.Tensor, n_total: int, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Like all_gather_shards but only `dst` receives the result (others get None)."""
    rank, w = world()
    if w == 1:
        return local
    per = (n_total + w - 1) // w
    rest = tuple(local.shape[1:])
    padded = local.new_zeros((per,) + rest)
    padded[:local.shape[0]] = local
    bufs = [local.new_empty((per,) + rest) for _ in range(w)] if rank == dst else None
    dist.gather(padded.contiguous(), bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([bufs[r][:shard_bounds(n_total, r, w)[1] - shard_bounds(n_total, r, w)[0]] for r in range(w)], dim=0)


def reduce_loss_stats(values: Dict[str, torch.Tensor], group=None) -> Dict[str, Dict[str, float]]:
    """Per key: global sum / mean / min / max / count over every rank's local [n_local] tensor.
    Three collectives in total regardless of the number of keys (sum+count, min, max), in fp64."""
    keys = sorted(values)
    _, w = world()
    dev = values[keys[0]].device if keys else torch.device("cpu")
    sums = torch.zeros(2 * len(keys), dtype=torch.float64, device=dev)
    mins = torch.full((len(keys),), float("inf"), dtype=torch.float64, device=dev)
    maxs = torch.full((len(keys),), float("-inf"), dtype=torch.float64, device=dev)
    for i, k in enumerate(keys):
        v = values[k].detach().to(torch.float64).reshape(-1)
        sums[2 * i] = v.sum()
        sums[2 * i + 1] = v.numel()
        if v.numel() > 0:
            mins[i], maxs[i] = v.min(), v.max()
    if w > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(mins, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(maxs, op=dist.ReduceOp.MAX, group=group)
    sums, mins, maxs = sums.cpu(), mins.cpu(), maxs.cpu()
    out = {}
    for i, k in enumerate(keys):
        cnt = sums[2 * i + 1].item()
        out[k] = {"sum": sums[2 * i].item(), "count": int(cnt), "mean": sums[2 * i].item() / max(cnt, 1.0),
                  "min": mins[i].item(), "max": maxs[i].item()}
    return out


class ShardedMotionQuery:
    """Tracker-shaped sharding (configs 2 and 4): the global batch of environments is split
    contiguously across ranks; each rank queries only its slice on its own GPU."""

    def __init__(self, motion_lib, hf_desc=None, obs_tmpl=None):
        self.mlib = motion_lib
        self.hf_desc = hf_desc
        self.obs_tmpl = obs_tmpl
        self._out = {}

    def query_local(self, global_ids: torch.Tensor, global_times: torch.Tensor) -> dict:
        """Inputs are the GLOBAL [N] tensors (replicated on every rank, already on this rank's device);
        returns this rank's outputs only.  No communication."""
        return self.mlib.calc_motion_frame_fk_obs(shard(global_ids).contiguous(), shard(global_times).contiguous(),
                                                  hf_desc=self.hf_desc, obs_tmpl=self.obs_tmpl, out=self._out)

    def query_gathered(self, global_ids, global_times, keys=("body_pos", "obs")) -> dict:
        """query_local + one all-gather per requested output."""
        local = self.query_local(global_ids, global_times)
        n = int(global_ids.shape[0])
        return {k: all_gather_shards(local[k], n) for k in keys if k in local}
