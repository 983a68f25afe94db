#!/usr/bin/env python
"""The kinematic motion optimisation loop (parc_2_kin_gen.py:445 -> motion_contact_optimization): a 254-frame clip
on a 50x50 terrain, default weights of PARC/kin_gen_default.yaml:26-37, CUDA-graph replay vs eager, next to ONE
iteration of the oracle's restatement of the reference loss (forward + backward + Adam) on the host cores.

    python scripts/bench_motion_opt.py [--iters 300]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=300)
    args = ap.parse_args()
    from parc_b200.anim.kin_char_model import KinCharModel
    from parc_b200.tools.motion_opt.motion_optimization import motion_contact_optimization
    from parc_b200.util import geom_util
    from parc_b200.util.terrain_util import SubTerrain

    dev = torch.device("cuda", 0)
    km = KinCharModel(dev)
    km.load_char_file(os.path.join(ROOT, "parc_b200", "assets", "humanoid.xml"))
    civ = np.load(os.path.join(ROOT, "tests", "golden", "clip_civilization.npz"))
    frames = torch.tensor(civ["frames"]).to(dev)
    frames[:, 2] -= 0.05
    contacts = torch.tensor(civ["contacts"]).to(dev)
    t = SubTerrain("civ", x_dim=50, y_dim=50, dx=0.4, dy=0.4, min_x=0.0, min_y=0.0, device=dev)
    t.hf = torch.tensor(civ["hf"]).to(dev)
    pts = geom_util.get_char_point_samples(km)
    W = dict(w_root_pos=1.0, w_root_rot=10.0, w_joint_rot=1.0, w_smoothness=10.0, w_penetration=1000.0, w_contact=1000.0,
             w_sliding=10.0, w_body_constraints=1000.0, w_jerk=1000.0)

    def run(graph, iters):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = motion_contact_optimization(frames.clone(), contacts, pts, t, km, num_iters=iters, step_size=0.001,
                                          body_constraints=None, max_jerk=1000.0, exp_name="bench", use_wandb=False,
                                          log_file=None, use_cuda_graph=graph, quiet=True, **W)
        torch.cuda.synchronize()
        return time.perf_counter() - t0, out

    run(True, 30)
    s_graph, out_g = run(True, args.iters)
    s_eager, out_e = run(False, args.iters)

    import bench          # the CPU leg (oracle port) lives in bench.py, the one place that may execute oracle/
    cpu_iter_s, cores = bench.cpu_leg_motion_opt(frames.cpu(), contacts.cpu(), torch.tensor(civ["hf"]), W)
    print(json.dumps({
        "workload": f"motion_contact_optimization: 254-frame clip, 50x50 terrain, {args.iters} Adam iterations, all loss terms",
        "gpu_ms_per_iter_cuda_graph": s_graph / args.iters * 1e3, "gpu_ms_per_iter_eager": s_eager / args.iters * 1e3,
        "graph_vs_eager_max_abs_diff": (out_g - out_e).abs().max().item(),
        "cpu_oracle_s_per_iter": cpu_iter_s, "cores": cores,
        "speedup_vs_cpu": cpu_iter_s / (s_graph / args.iters),
        "projected_3000_iters": {"gpu_s": s_graph / args.iters * 3000, "cpu_h": cpu_iter_s * 3000 / 3600}}))


if __name__ == "__main__":
    main()
