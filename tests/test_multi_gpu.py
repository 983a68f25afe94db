"""N > 1 on real GPUs: one process per GPU over NCCL (skipped with fewer than 2 GPUs; the gloo world-size-2 tests of
test_sharding.py cover the same host logic on CPU).  Config 4's shape scaled down: a global batch of environments is
split contiguously across the ranks, each rank queries only its slice through the C ABI, and the gathered result
must equal the single-GPU query bit for bit; loss statistics are reduced with three collectives."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ASSET

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world_size, device_id=dev)
    try:
        from parc_b200 import ops, sharding
        from parc_b200.anim.kin_char_model import KinCharModel
        from parc_b200.anim.motion_lib import LoopMode, MotionLib
        from parc_b200.tools.procgen.mdm_path import body_points_desc
        from parc_b200.util import geom_util, synth
        km = KinCharModel(dev)
        km.load_char_file(ASSET)
        rng = np.random.default_rng(11)
        hf = synth.rolling_terrain(rng, 128, 128, num_boxes=60)
        frames, contacts = synth.synth_clips(km, 64, seed=12, hf=hf)
        lib = MotionLib(torch.from_numpy(frames).to(dev), km, dev, init_type="motion_frames", loop_mode=LoopMode.WRAP,
                        fps=30, contact_info=True, contacts=torch.from_numpy(contacts).to(dev))
        gen = torch.Generator().manual_seed(5)
        ids = torch.randint(0, 64, (n,), generator=gen).to(dev)          # the GLOBAL batch, replicated
        times = (torch.rand(n, generator=gen) * 20.0 - 3.0).to(dev)
        hfd = ops.HeightfieldDesc(hf=torch.from_numpy(hf).to(dev), min_x=0.0, min_y=0.0, dx=0.4, dy=0.4)
        tmpl = geom_util.get_xy_points_cone(torch.zeros(2, device=dev), 0.05, 2, 60, 3, 3, 0.26179938779)
        sq = sharding.ShardedMotionQuery(lib, hf_desc=hfd, obs_tmpl=tmpl)
        local = sq.query_local(ids, times)
        lo, hi = sharding.shard_bounds(n, rank, world_size)
        assert local["body_pos"].shape[0] == hi - lo
        gathered = sq.query_gathered(ids, times, keys=("root_pos", "body_pos", "body_rot", "obs", "contacts"))
        full = lib.calc_motion_frame_fk_obs(ids, times, hf_desc=hfd, obs_tmpl=tmpl)
        same = all(torch.equal(gathered[k], full[k]) for k in gathered)
        on0 = sharding.gather_shards_to(local["obs"], n, dst=0)
        ok0 = (on0 is None) if rank != 0 else torch.equal(on0, full["obs"])
        # the library's own peer-memory gather (NVSwitch multicast where available, peer pointers otherwise) against the
        # single-GPU result; ragged shards (n odd) exercise the 4-byte path, n_even the 16-byte one
        peer = {}
        J, P = km.get_num_joints(), int(tmpl.shape[0])
        lb = {k: local[k].clone() for k in ("body_pos", "obs")}
        for mode, mc in (("auto", True), ("p2p", False)):
            pg = sharding.PeerGather({"body_pos": (J, 3), "obs": (P,)}, n, dev, use_multicast=mc)
            for rep in range(3):                              # replays reuse the per-slot epochs
                pg.push(lb)
            torch.cuda.synchronize(dev)
            pg.check()                                        # no hand-shake gave up waiting
            peer[mode] = bool(torch.equal(pg.out["body_pos"], full["body_pos"]) and torch.equal(pg.out["obs"], full["obs"]))
            peer["multicast"] = bool(pg.multicast) if mode == "auto" else peer["multicast"]
            dist.barrier()
        n_even = (n // (16 * world_size)) * 16 * world_size
        lo2, hi2 = sharding.shard_bounds(n_even, rank, world_size)
        pg = sharding.PeerGather({"body_pos": (J, 3), "obs": (P,)}, n_even, dev, use_multicast=True)
        pg.push({"body_pos": full["body_pos"][lo2:hi2].contiguous(), "obs": full["obs"][lo2:hi2].contiguous()})
        torch.cuda.synchronize(dev)
        peer["vec16"] = bool(torch.equal(pg.out["body_pos"], full["body_pos"][:n_even]) and
                             torch.equal(pg.out["obs"], full["obs"][:n_even]))
        dist.barrier()
        if pg.multicast:
            # direct form: the query kernel's own body_pos / obs stores go to the multicast address of this rank's rows
            pg.buf.zero_()
            torch.cuda.synchronize(dev)
            dist.barrier()
            plan = lib.make_query_plan(ids[lo2:hi2].contiguous(), times[lo2:hi2].contiguous(), hf_desc=hfd, obs_tmpl=tmpl)
            plan.redirect_output("body_pos", pg.direct_ptr("body_pos"))
            plan.redirect_output("obs", pg.direct_ptr("obs"))
            for rep in range(2):
                plan.launch()
                pg.barrier()
            torch.cuda.synchronize(dev)
            peer["direct"] = bool(torch.equal(pg.out["body_pos"], full["body_pos"][:n_even]) and
                                  torch.equal(pg.out["obs"], full["obs"][:n_even]))
            dist.barrier()
        else:
            peer["direct"] = None
        # loss statistics of sharded samples (config 3's shape, small)
        B, F = 32, 24
        base = [synth.box_terrain(np.random.default_rng(100 + i)) for i in range(B)]
        s = synth.synth_motion_samples(km, B, F, base[0], (0.0, 0.0), (0.4, 0.4), seed=21)
        sl = slice(*sharding.shard_bounds(B, rank, world_size))
        hfs = torch.tensor(np.stack(base)).to(dev)
        tb = ops.make_terrain_batch(hfs[sl].contiguous(), torch.zeros(sl.stop - sl.start, 2, device=dev), (0.4, 0.4), base_z=-10.0)
        pts = body_points_desc(km, geom_util.get_char_point_samples(km))
        tot, pen, con = ops.body_loss(km.c_model(), pts, tb, torch.tensor(s["root_pos"][sl]).to(dev),
                                      ops.exp_map_to_quat(torch.tensor(s["root_exp"][sl]).to(dev)),
                                      km.dof_to_rot(torch.tensor(s["joint_dof"][sl]).to(dev)),
                                      torch.tensor(s["contacts"][sl]).to(dev), 0.1, 0.1)
        st = sharding.reduce_loss_stats({"pen": pen, "con": con})
        q.put((rank, bool(same), bool(ok0), st["pen"], st["con"]["count"], pen.double().sum().item(), peer))
    finally:
        dist.destroy_process_group()


def test_nccl_world2_sharded_query_and_loss_stats():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n = 8191                                                     # odd: the shards differ by one row
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    got = sorted([q.get(timeout=10) for _ in range(2)], key=lambda x: x[0])
    assert all(g[1] and g[2] for g in got)
    for g in got:                                                # peer-memory gather == single-GPU query, every mode
        assert g[6]["auto"] and g[6]["p2p"] and g[6]["vec16"] and g[6]["direct"] in (True, None), g[6]
    assert got[0][3] == got[1][3] and got[0][4] == got[1][4] == 32 * 1          # one loss value per sample
    assert abs(got[0][3]["sum"] - (got[0][5] + got[1][5])) <= 1e-9 * max(1.0, abs(got[0][3]["sum"]))
    assert got[0][3]["min"] >= 0.0 and got[0][3]["max"] >= got[0][3]["mean"] >= got[0][3]["min"]
