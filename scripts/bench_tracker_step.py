#!/usr/bin/env python
"""the kinematic side of one tracker control step (parc_b200/envs/ig_parkour/step_assembly.py)
on the bench's library -- reference frame + 6 future targets + FK, simulated character's observation and ray
heightmap, target observation, DeepMimic reward terms, episode flags -- eager (one launch per operator) and as a
captured CUDA graph, next to the same sequence composed from the CPU oracle on all host cores.

    python scripts/bench_tracker_step.py [--envs 4096] [--clips 2048] [--steps 50]
L2 is flushed between timed steps (256 MiB write); per-step CUDA events.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

HF_DIM, HF_DX = 1536, 0.4


def bytes_per_env_step(J, D, K, S, P):
    """Algorithmic bytes (each logical input read once, each output written once, fp32/int64)."""
    q = (S + 1) * (36 + 2 * (3 + 4 + 4 * (J - 1) + J) * 4 + (6 + D) * 4          # clip meta + two key frames + velocities
                   + (3 + 4 + 3 + 3 + 4 * (J - 1) + D + J) * 4 + J * 7 * 4) + 12 + 8   # frame out + FK out; id, t, xy offset
    d2r = D * 4 + (J - 1) * 16
    ray = P * 8 + 20
    char_w = 12 + 6 * (J - 1) + D + 3 * K
    tar_w = 9 + 6 * (J - 1) + 3 * K
    char = (13 + 4 * (J - 1) + D + 3 * K) * 4 + char_w * 4
    tar = S * (7 + 4 * (J - 1) + 3 * K) * 4 + 28 + S * tar_w * 4
    W = char_w + S * tar_w + S * J + J + P
    cat = 2 * (S * J + J) * 4            # the two contact-flag copies; every other block is written in place
    rew = 2 * (13 + 4 * (J - 1) + D + 3 * K) * 4 + 20
    done = 4 + 2 * J * 12 + 32 + J * 12 + 12 + J * 4 + 4
    return dict(query=q, dof_to_rot=d2r, ray_obs=ray, char_obs=char, tar_obs=tar, contact_copies=cat, reward=rew, done=done,
                total=q + d2r + ray + char + tar + cat + rew + done, obs_width=W)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--clips", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--cpu-reps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU oracle leg (profiling runs)")
    ap.add_argument("--fuse-tar-obs", action="store_true", help="target observation written by the query kernel (3 launches)")
    ap.add_argument("--single-sim-launch", action="store_true",
                    help="the simulated character's share as ONE launch behind the query (round 1's 4-launch step)")
    ap.add_argument("--query-variant", type=int, default=0, help="instantiation of the query kernel (tuning)")
    args = ap.parse_args()
    import __graft_entry__ as entry
    entry.ensure_built()
    from parc_b200 import _lib
    from parc_b200.anim.kin_char_model import KinCharModel
    from parc_b200.anim.motion_lib import LoopMode, MotionLib
    from parc_b200.envs.ig_parkour.step_assembly import TrackerStep
    from parc_b200.util import geom_util, synth
    from parc_b200.util.terrain_util import SubTerrain

    dev = torch.device("cuda", 0)
    km = KinCharModel(dev)
    km.load_char_file(os.path.join(ROOT, "parc_b200", "assets", "humanoid.xml"))
    rng = np.random.default_rng(1234)
    hf_np = synth.rolling_terrain(rng, HF_DIM, HF_DIM, num_boxes=6000)
    frames, contacts = synth.synth_clips(km, args.clips, seed=1235, hf=hf_np, min_xy=(0.0, 0.0), dxdy=(HF_DX, HF_DX))
    mlib = MotionLib(torch.from_numpy(frames).to(dev), km, dev, init_type="motion_frames", loop_mode=LoopMode.CLAMP, fps=30,
                     contact_info=True, contacts=torch.from_numpy(contacts).to(dev))
    terrain = SubTerrain("global", x_dim=HF_DIM, y_dim=HF_DIM, dx=HF_DX, dy=HF_DX, min_x=0.0, min_y=0.0, device=dev)
    terrain.hf = torch.from_numpy(hf_np).to(dev)
    tmpl = geom_util.get_xy_points_cone(center=torch.zeros(2, device=dev), dx=0.05, num_neg=2, num_pos=60, num_rays_neg=3,
                                        num_rays_pos=3, angle_between_rays=0.26179938779)
    n, J, D = args.envs, km.get_num_joints(), km.get_dof_size()
    key_ids = [km.get_body_id(b) for b in ("right_hand", "left_hand", "right_foot", "left_foot")]
    feet = key_ids[2:]
    jw = torch.tensor([1.0, 0.6, 0.6, 0.4, 0.0, 0.6, 0.4, 0.0, 1.0, 0.6, 0.4, 1.0, 0.6, 0.4])
    ptd = torch.tensor([0.7, 1.0, 0.7, 0.7, 0.7, 0.7, 0.7, 0.7, 1.0, 1.2, 10.0, 1.0, 1.2, 10.0])
    steps = [1, 2, 3, 10, 20, 30]
    ts = TrackerStep(mlib, terrain, n, 1.0 / 30.0, steps, key_ids, tmpl, joint_err_w=jw, pose_termination_dist=ptd,
                     contact_body_ids=feet, fuse_tar_obs=args.fuse_tar_obs, query_variant=args.query_variant,
                     split_sim=not args.single_sim_launch)
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, args.clips, (n,), generator=g)
    times = torch.rand(n, generator=g) * (264.0 / 30.0)
    ts.motion_ids.copy_(ids); ts.motion_times.copy_(times)
    ts.motion_xy_offset.copy_(torch.randn(n, 2, generator=g) * 0.2)
    # simulated character = reference pose + noise
    fr = mlib.calc_motion_frame(ts.motion_ids, ts.motion_times)
    root_pos = fr[0] + 0.05 * torch.randn(n, 3, generator=g).to(dev)
    root_rot = fr[1].clone()
    root_vel, root_ang_vel = fr[2].clone(), fr[3].clone()
    dof_pos = km.rot_to_dof(fr[4]) + 0.05 * torch.randn(n, D, generator=g).to(dev)
    dof_vel = fr[5].clone()
    body_pos = km.forward_kinematics(root_pos, root_rot, km.dof_to_rot(dof_pos))[0]
    forces = (torch.randn(n, J, 3, generator=g) * (torch.rand(n, J, 1, generator=g) < 0.2)).to(dev)
    time_buf = (torch.rand(n, generator=g) * 11.0).to(dev)
    env_off = torch.zeros(n, 3, device=dev)
    char_contacts = (torch.rand(n, J, generator=g) < 0.3).float().to(dev)
    state = (root_pos, root_rot, root_vel, root_ang_vel, dof_pos, dof_vel, body_pos, forces, time_buf, env_off, char_contacts)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(args.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / args.steps

    l0 = _lib.LAUNCHES[0]
    res = ts.step(*state)
    launches = _lib.LAUNCHES[0] - l0
    ms_eager = timed(lambda: ts.step(*state))
    ts.capture(*state)
    ms_graph = timed(ts.replay)
    torch.cuda.synchronize()
    K, S, P = len(key_ids), len(steps), int(tmpl.shape[0])
    b = bytes_per_env_step(J, D, K, S, P)
    assert res["obs"].shape[1] == b["obs_width"]

    # ---- CPU: the same sequence composed from the oracle on all host cores (the leg lives in bench.py, the one
    #      place that may execute oracle/) ----
    cpu = None
    if not args.no_cpu:
        import bench
        best, cores, (o_obs, o_rew, o_done) = bench.cpu_leg_tracker_step(
            mlib, hf_np, HF_DX, state, ids, times, ts.motion_xy_offset.cpu(), ts.time_offsets.cpu(), key_ids, feet, jw,
            ts.dof_err_w.cpu(), ptd, tmpl.cpu(), args.cpu_reps)
        cpu = dict(cores=cores, ms_per_step=best * 1e3, env_steps_per_s=n / best,
                   sample=f"the full {n}-env step, best of {args.cpu_reps}, oracle composition (torch CPU fp32)",
                   obs_max_abs_diff=float((res["obs"].cpu() - o_obs).abs().max()),
                   reward_max_abs_diff=float((res["reward_terms"].cpu() - o_rew).abs().max()),
                   done_mismatches=int((res["done"].cpu() != o_done).sum()))
    peak = 6555.5
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    out = {
        "workload": f"tracker control step, kinematic side: {n} envs, {args.clips} clips x 265 frames, 1 + {S} frame queries + "
                    f"FK, char/target observation (W = {b['obs_width']}), {P}-pt ray heightmap, reward terms, done flags",
        "parc_launches_per_step": launches, "target_observation": "fused into the query kernel" if args.fuse_tar_obs else "own launch",
        "query_variant": args.query_variant,
        "sim_step": "one launch behind the query" if args.single_sim_launch else "split: observation half beside the query, reward / done behind it",
        "eager": {"ms_per_step": ms_eager, "env_steps_per_s": n / (ms_eager * 1e-3)},
        "cuda_graph": {"ms_per_step": ms_graph, "env_steps_per_s": n / (ms_graph * 1e-3),
                       "achieved_GBps": n * b["total"] / (ms_graph * 1e-3) / 1e9,
                       "roofline_frac": n * b["total"] / (ms_graph * 1e-3) / 1e9 / peak, "hbm_peak_GBps": peak},
        "algorithmic_bytes_per_env_step": b,
        "l2": "flushed between steps (256 MiB write)",
        "cpu_oracle": cpu,
    }
    if cpu:
        out["speedup_graph_vs_cpu"] = cpu["ms_per_step"] / ms_graph
    print(json.dumps(out))


if __name__ == "__main__":
    main()
