#!/usr/bin/env python
"""Config 3 of BASELINE.json: kin-gen penetration/contact loss forward+backward for 1024 synthetic MDM
samples on procedural box/stair terrains (16x16 @0.4 m, one terrain per sample).  Reports samples/s,
point-cell SDF evaluations/s and the CPU oracle on a stated sub-sample.  Not the headline metric
(that is bench.py); numbers go to profiles/ and DESIGN.md.

    python scripts/bench_loss.py [--batch 1024] [--frames 200] [--steps 5] [--cpu-samples 1]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--frames", type=int, default=200)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--cpu-samples", type=int, default=1)
    ap.add_argument("--cpu-frames", type=int, default=20)
    args = ap.parse_args()

    from parc_b200 import ops
    from parc_b200.anim.kin_char_model import KinCharModel
    from parc_b200.tools.procgen.mdm_path import body_points_desc
    from parc_b200.util import geom_util, synth

    dev = torch.device("cuda", 0)
    km = KinCharModel(dev)
    km.load_char_file(os.path.join(ROOT, "parc_b200", "assets", "humanoid.xml"))
    rng = np.random.default_rng(3)
    B, F = args.batch, args.frames
    # 32 distinct terrains cycled over the batch (per-sample terrain pointers all differ in the kernel)
    base = [synth.box_terrain(rng) if i % 2 == 0 else synth.stairs_terrain(rng) for i in range(32)]
    hfs = np.stack([base[i % 32] for i in range(B)])
    smp = synth.synth_motion_samples(km, B, F, base[0], (0.0, 0.0), (0.4, 0.4), seed=11)
    leaves = [torch.tensor(smp[k]).to(dev).requires_grad_(True) for k in ("root_pos", "root_exp", "joint_dof")]
    contacts = torch.tensor(smp["contacts"]).to(dev)
    pts = body_points_desc(km, geom_util.get_char_point_samples(km))
    tb = ops.make_terrain_batch(torch.tensor(hfs).to(dev), torch.zeros(B, 2, device=dev), (0.4, 0.4), base_z=-10.0)
    model = km.c_model()

    def step():
        for t in leaves:
            t.grad = None
        rq = ops.exp_map_to_quat(leaves[1])
        jr = km.dof_to_rot(leaves[2])
        total, pen, con = ops.body_loss(model, pts, tb, leaves[0], rq, jr, contacts, 0.1, 0.1)
        total.sum().backward()
        return total

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    S = int(pts.points.shape[0])
    evals = B * F * S * 256 * 2            # point x cell x {air, solid}
    # forward only
    with torch.no_grad():
        rq = ops.exp_map_to_quat(leaves[1])
        jr = km.dof_to_rot(leaves[2])
        e0.record()
        for _ in range(args.steps):
            ops.body_loss(model, pts, tb, leaves[0], rq, jr, contacts, 0.1, 0.1)
        e1.record()
        torch.cuda.synchronize()
    ms_fwd = e0.elapsed_time(e1) / args.steps

    # CPU leg (oracle port) on a sub-sample: lives in bench.py, the one place that may execute oracle/
    import bench
    cpu = bench.cpu_leg_loss(smp, hfs, args.cpu_samples, args.cpu_frames, F)
    cpu_samples_per_s = cpu["samples_per_s"]
    print(json.dumps({
        "workload": f"cfg3: pen/contact loss fwd+bwd, B={B} samples x F={F} frames, 304 body points, 16x16 terrain/sample",
        "gpu_ms_fwd_bwd": ms, "gpu_ms_fwd": ms_fwd, "samples_per_s_fwd_bwd": B / (ms * 1e-3),
        "point_cell_evals_per_s": evals / (ms * 1e-3),
        "cpu_oracle": cpu,
        "speedup_vs_cpu": (B / (ms * 1e-3)) / cpu_samples_per_s}))


if __name__ == "__main__":
    main()
