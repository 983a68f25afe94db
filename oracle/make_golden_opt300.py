"""300 Adam iterations of the reference's motion optimisation on a 16 x 16 terrain: checkpoints for the GPU test.

TEST INFRASTRUCTURE; runs only in the authoring container (needs /root/reference):
    python -m oracle.make_golden_opt300
The loop of tools/motion_opt/motion_optimization.py:404-500 built from the REAL reference's
motion_terrain_contact_loss (all nine terms, kin_gen_default.yaml weights) + torch.optim.Adam, 16 frames of the
in-repo clip on a 16 x 16 crop of its terrain.  The frames after 1 / 4 / 25 / 100 / 300 iterations and the objective at
each checkpoint go to tests/golden/motion_opt300_golden.npz.  The oracle restatement is run beside it and its distance
from the reference at every checkpoint is recorded (`oracle_vs_ref_max_abs`): Adam on this non-smooth objective
(arg-min cells, clamps, first-index ties) amplifies last-bit differences, so two fp32 implementations that agree to
1e-7 per iteration drift apart over hundreds of iterations -- that drift between two CPU implementations is the
yardstick for the GPU path's own drift.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

ref_shim.activate()
import anim.kin_char_model as ref_kcm  # noqa: E402
import util.geom_util as ref_geom  # noqa: E402
import util.terrain_util as ref_terrain  # noqa: E402
import util.torch_util as ref_tu  # noqa: E402
import tools.motion_opt.motion_optimization as ref_mopt  # noqa: E402

from oracle import parc_oracle as O  # noqa: E402

CHECKPOINTS = (1, 4, 25, 100, 300)
W = dict(w_root_pos=1.0, w_root_rot=10.0, w_joint_rot=1.0, w_smoothness=10.0, w_penetration=1000.0, w_contact=1000.0,
         w_sliding=10.0, w_body_constraints=1000.0, w_jerk=1000.0)


def main():
    torch.manual_seed(0)
    km = ref_kcm.KinCharModel("cpu")
    km.load_char_file(os.path.join(ref_shim.REFERENCE_ROOT, "data/assets/humanoid.xml"))
    model = O.CharModel.from_npz(os.path.join(GOLD, "humanoid_model.npz"))
    body_points = ref_geom.get_char_point_samples(km)
    civ = np.load(os.path.join(GOLD, "clip_civilization.npz"))
    F0, F = 40, 16
    fr = torch.tensor(civ["frames"][F0:F0 + F]).clone()
    cts = torch.tensor(civ["contacts"][F0:F0 + F]).clone()
    hf_full = torch.tensor(civ["hf"])
    # crop the terrain to 16 x 16 cells around the clip segment
    gx = int(((fr[:, 0].mean() - float(civ["min_point"][0])) / 0.4).round()) - 8
    gy = int(((fr[:, 1].mean() - float(civ["min_point"][1])) / 0.4).round()) - 8
    gx, gy = max(0, min(gx, 50 - 16)), max(0, min(gy, 50 - 16))
    terr = ref_terrain.SubTerrain("crop", x_dim=16, y_dim=16, dx=0.4, dy=0.4, min_x=float(civ["min_point"][0]) + 0.4 * gx,
                                  min_y=float(civ["min_point"][1]) + 0.4 * gy, device="cpu")
    terr.hf = hf_full[gx:gx + 16, gy:gy + 16].clone()
    fr[:, 2] -= 0.04                                        # push the character into the ground a little

    def source_constants(tu, dof_to_rot, fk):
        rq = tu(fr[:, 3:6]); jr = dof_to_rot(fr[:, 6:34])
        bp, br = fk(fr[:, 0:3], rq, jr)
        return rq, jr, bp[1:] - bp[:-1], br

    s_rq = ref_tu.exp_map_to_quat(fr[:, 3:6]); s_jr = km.dof_to_rot(fr[:, 6:34])
    s_bp, s_br = km.forward_kinematics(fr[:, 0:3], s_rq, s_jr)
    s_bv = s_bp[1:] - s_bp[:-1]; s_brv = ref_tu.quat_diff_angle(s_br[1:], s_br[:-1])

    leaves = [fr[:, 0:3].clone().requires_grad_(True), fr[:, 3:6].clone().requires_grad_(True),
              fr[:, 6:34].clone().requires_grad_(True)]
    opt = torch.optim.Adam(leaves, lr=0.001)
    ref_ck, ref_loss = {}, {}
    for it in range(1, max(CHECKPOINTS) + 1):
        opt.zero_grad()
        l_, _d = ref_mopt.motion_terrain_contact_loss(leaves[0], leaves[1], leaves[2], fr[:, 0:3], s_rq, s_jr, s_bv, s_brv,
                                                     cts, terr, body_points, km, body_constraints=None, max_jerk=1000.0, **W)
        l_.backward()
        opt.step()
        if it in CHECKPOINTS:
            ref_ck[it] = torch.cat([t.detach() for t in leaves], dim=-1).clone()
            ref_loss[it] = float(l_.item())            # objective BEFORE this iteration's update
            print(f"reference iteration {it}: loss {ref_loss[it]:.6f}", flush=True)
    drift = {}
    for it in CHECKPOINTS:
        o = O.motion_contact_optimization(model, fr, cts, terr.hf, terr.min_point, terr.dxdy, it, 0.001, W, 1000.0)
        drift[it] = float((o - ref_ck[it]).abs().max())
        print(f"oracle vs reference after {it}: max abs {drift[it]:.3e} (update size {float((ref_ck[it] - fr).abs().max()):.3e})", flush=True)
    np.savez_compressed(
        os.path.join(GOLD, "motion_opt300_golden.npz"), src_frames=fr.numpy(), contacts=cts.numpy(), hf=terr.hf.numpy(),
        min_point=terr.min_point.numpy(), dxdy=terr.dxdy.numpy(), checkpoints=np.array(CHECKPOINTS),
        frames=np.stack([ref_ck[i].numpy() for i in CHECKPOINTS]), loss=np.array([ref_loss[i] for i in CHECKPOINTS]),
        oracle_vs_ref_max_abs=np.array([drift[i] for i in CHECKPOINTS]),
        weights=np.array([W[k] for k in sorted(W)]), weight_names=np.array(sorted(W)))
    print("wrote motion_opt300_golden.npz")


if __name__ == "__main__":
    main()
