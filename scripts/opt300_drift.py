#!/usr/bin/env python
"""Distance of the GPU optimiser's trajectory from the reference's, at the checkpoints of
tests/golden/motion_opt300_golden.npz (16 frames, 16 x 16 terrain, all nine terms, 300 Adam iterations)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from parc_b200.anim.kin_char_model import KinCharModel
    from parc_b200.tools.motion_opt.motion_optimization import motion_contact_optimization, motion_terrain_contact_loss, source_constants
    from parc_b200.util import geom_util
    from parc_b200.util.terrain_util import SubTerrain
    g = np.load(os.path.join(ROOT, "tests", "golden", "motion_opt300_golden.npz"))
    dev = torch.device("cuda", 0)
    km = KinCharModel(dev)
    km.load_char_file(os.path.join(ROOT, "parc_b200", "assets", "humanoid.xml"))
    W = {str(k): float(v) for k, v in zip(g["weight_names"], g["weights"])}
    t = SubTerrain("crop", x_dim=16, y_dim=16, dx=0.4, dy=0.4, min_x=float(g["min_point"][0]), min_y=float(g["min_point"][1]), device=dev)
    t.hf = torch.tensor(g["hf"]).to(dev)
    pts = geom_util.get_char_point_samples(km)
    src, cts = torch.tensor(g["src_frames"]).to(dev), torch.tensor(g["contacts"]).to(dev)
    rq, jr, bv, brv = source_constants(src, km)

    def objective(fr):
        with torch.no_grad():
            return float(motion_terrain_contact_loss(fr[:, 0:3], fr[:, 3:6], fr[:, 6:], src[:, 0:3], rq, jr, bv, brv, cts, t, pts, km,
                                                     body_constraints=None, max_jerk=1000.0, **W)[0])
    rows = []
    for it, ref in zip(g["checkpoints"].tolist(), g["frames"]):
        out = motion_contact_optimization(src.clone(), cts, pts, t, km, num_iters=it, step_size=0.001, body_constraints=None,
                                          max_jerk=1000.0, use_cuda_graph=True, quiet=True, **W)
        ref = torch.tensor(ref).to(dev)
        d = (out - ref).abs()
        upd = (ref - src).abs()
        rows.append({"iterations": it, "max_abs_diff": float(d.max()), "mean_abs_diff": float(d.mean()), "rms_diff": float(d.pow(2).mean().sqrt()),
                     "ref_update_max": float(upd.max()), "ref_update_mean": float(upd.mean()),
                     "objective_ours": objective(out), "objective_ref": objective(ref), "objective_src": objective(src)})
        print(rows[-1], file=sys.stderr)
    print(json.dumps({"rows": rows}))


if __name__ == "__main__":
    main()
