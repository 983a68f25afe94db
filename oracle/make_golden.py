"""Generate tests/golden/*.npz from the REAL reference and pin the oracle against it.

TEST INFRASTRUCTURE; runs only in the authoring container (needs /root/reference):
    python -m oracle.make_golden
Every fixture below is an output of the imported reference's own functions on CPU (torch fp32).  While
generating, the oracle restatement (oracle/parc_oracle.py) is run on the same inputs and must match the
reference BIT-EXACTLY (torch.equal); any mismatch aborts.  The summary is written to
tests/golden/PIN_REPORT.txt.
"""
from __future__ import annotations

import os
import pickle
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

ref_shim.activate()
import anim.kin_char_model as ref_kcm  # noqa: E402
import anim.motion_lib as ref_mlib  # noqa: E402
import util.geom_util as ref_geom  # noqa: E402
import util.terrain_util as ref_terrain  # noqa: E402
import util.torch_util as ref_tu  # noqa: E402
import util.motion_util as ref_mu  # noqa: E402
import tools.procgen.mdm_path as ref_mdm_path  # noqa: E402
import tools.motion_opt.motion_optimization as ref_mopt  # noqa: E402
import zmotion_editing_tools.motion_edit_lib as ref_mel  # noqa: E402
import envs.ig_char_env as ref_char_env  # noqa: E402
import envs.ig_parkour.mgdm_dm_util as ref_dm_util  # noqa: E402

from oracle import parc_oracle as O  # noqa: E402

REPORT = []


def pin(name, ref, mine):
    ref = torch.as_tensor(ref)
    mine = torch.as_tensor(mine)
    same = ref.shape == mine.shape and torch.equal(ref, mine)
    if not same:
        both_nan = torch.isnan(ref) & torch.isnan(mine) if ref.is_floating_point() else torch.zeros_like(ref, dtype=torch.bool)
        same = ref.shape == mine.shape and bool(((ref == mine) | both_nan).all())
    REPORT.append(f"{'OK ' if same else 'FAIL'} {name} shape={tuple(ref.shape)}")
    if not same:
        diff = (ref.double() - mine.double()).abs().max().item() if ref.shape == mine.shape else float('nan')
        print("\n".join(REPORT))
        raise SystemExit(f"oracle != reference for {name}: max abs diff {diff}")


def pin_close(name, ref, mine, rtol):
    """max|ref - mine| <= rtol * max|ref| -- for sums of many autograd branches, whose accumulation ORDER (and so
    the last bits) depends on graph construction order rather than on the algorithm."""
    ref = torch.as_tensor(ref).double()
    mine = torch.as_tensor(mine).double()
    scale = max(ref.abs().max().item(), 1e-30)
    err = (ref - mine).abs().max().item() if ref.shape == mine.shape else float("inf")
    ok = err <= rtol * scale
    REPORT.append(f"{'OK~' if ok else 'FAIL'} {name} shape={tuple(ref.shape)} max|err|/max|ref|={err / scale:.2e} (bar {rtol:g})")
    if not ok:
        print("\n".join(REPORT))
        raise SystemExit(f"oracle !~ reference for {name}")


def npf(t):
    return t.detach().cpu().numpy()


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    os.makedirs(GOLD, exist_ok=True)
    ref_root = ref_shim.REFERENCE_ROOT

    # ------------------------------------------------------------------ 1. character model
    km = ref_kcm.KinCharModel("cpu")
    km.load_char_file(os.path.join(ref_root, "data/assets/humanoid.xml"))
    body_points = ref_geom.get_char_point_samples(km)
    min_points = ref_geom.get_minimal_char_point_samples(km)
    J = km.get_num_joints()
    jt = [km._joints[j].joint_type.value for j in range(J)]
    axes = np.zeros((J, 3), np.float32)
    for j in range(J):
        if km._joints[j].axis is not None:
            axes[j] = npf(km._joints[j].axis)
    geom_rows = []
    for b in range(J):
        for g in km.get_geoms(b):
            d = npf(g._dims).reshape(-1)
            d3 = np.zeros(3, np.float32)
            d3[:d.shape[0]] = d
            geom_rows.append([b, g._shape_type.value, *npf(g._offset).tolist(), *d3.tolist(),
                              -1.0 if g._radius is None else float(g._radius)])
    np.savez_compressed(
        os.path.join(GOLD, "humanoid_model.npz"),
        body_names=np.array(km._body_names), parents=npf(km._parent_indices),
        local_translation=npf(km._local_translation), local_rotation=npf(km._local_rotation),
        joint_type=np.array(jt, np.int64), joint_axis=axes,
        dof_idx=np.array([km._joints[j].dof_idx for j in range(J)], np.int64),
        dof_dim=np.array([km._joints[j].get_dof_dim() for j in range(J)], np.int64),
        lower_dof_limits=npf(km._lower_dof_limits), upper_dof_limits=npf(km._upper_dof_limits),
        body_points=npf(torch.cat(body_points, 0)), body_point_counts=np.array([p.shape[0] for p in body_points]),
        min_body_points=npf(torch.cat(min_points, 0)),
        min_body_point_counts=np.array([p.shape[0] for p in min_points]),
        geoms=np.array(geom_rows, np.float32))
    model = O.CharModel.from_npz(os.path.join(GOLD, "humanoid_model.npz"))

    # ------------------------------------------------------------------ 2. the two in-repo clips
    clips_raw = {}
    for name in ("civilization", "TEASER_TERRAIN"):
        with open(os.path.join(ref_root, "data/terrains", name + ".pkl"), "rb") as f:
            d = pickle.load(f)
        t = d["terrain"]
        t.update_old()
        t.to_torch("cpu")
        fps = d.get("fps", 30)
        fps = fps.item() if isinstance(fps, np.ndarray) else fps
        clips_raw[name] = dict(frames=np.asarray(d["frames"], np.float32), contacts=np.asarray(d["contacts"], np.float32),
                               hf=npf(t.hf), min_point=npf(t.min_point), dxdy=npf(t.dxdy), fps=float(fps),
                               loop_mode=d.get("loop_mode", "CLAMP"))
        np.savez_compressed(os.path.join(GOLD, f"clip_{name.lower()}.npz"), **{
            k: (np.array(v) if not isinstance(v, np.ndarray) else v) for k, v in clips_raw[name].items()})

    # a library of 3 clips: civilization (CLAMP), teaser played as WRAP, civilization's first 40 frames as WRAP@60fps
    tmp = tempfile.mkdtemp()
    lib_spec = [("civilization", clips_raw["civilization"]["frames"], clips_raw["civilization"]["contacts"], 30.0, "CLAMP", 1.0),
                ("teaser_wrap", clips_raw["TEASER_TERRAIN"]["frames"], clips_raw["TEASER_TERRAIN"]["contacts"], 30.0, "WRAP", 2.0),
                ("civ_short", clips_raw["civilization"]["frames"][:40], clips_raw["civilization"]["contacts"][:40], 60.0, "WRAP", 0.5)]
    yaml_lines = ["motions:"]
    for nm, fr, ct, fps, loop, w in lib_spec:
        path = os.path.join(tmp, nm + ".pkl")
        with open(path, "wb") as f:
            pickle.dump({"frames": fr, "contacts": ct, "fps": fps, "loop_mode": loop}, f)
        yaml_lines += [f"- file: {path}", f"  weight: {w}"]
    ypath = os.path.join(tmp, "lib.yaml")
    with open(ypath, "w") as f:
        f.write("\n".join(yaml_lines) + "\n")
    mlib = ref_mlib.MotionLib(ypath, km, "cpu", init_type="motion_file", contact_info=True)

    tb = O.build_tables(model, [O.Clip(fr, ct, fps, O.WRAP if loop == "WRAP" else O.CLAMP, w)
                                for _, fr, ct, fps, loop, w in lib_spec])
    for k_ref, k_o in [("_frame_root_pos", "root_pos"), ("_frame_root_rot", "root_rot"), ("_frame_joint_rot", "joint_rot"),
                       ("_frame_root_vel", "root_vel"), ("_frame_root_ang_vel", "root_ang_vel"),
                       ("_frame_dof_vel", "dof_vel"), ("_frame_contacts", "contacts"), ("_motion_frames", "frames"),
                       ("_motion_num_frames", "num_frames"), ("_motion_start_idx", "start_idx"),
                       ("_motion_lengths", "lengths"), ("_motion_loop_modes", "loop_modes"),
                       ("_motion_root_pos_delta", "root_pos_delta"), ("_motion_weights", "weights")]:
        pin("tables." + k_o, getattr(mlib, k_ref), getattr(tb, k_o))
    np.savez_compressed(os.path.join(GOLD, "tables_golden.npz"),
                        lib_fps=np.array([s[3] for s in lib_spec]), lib_loop=np.array([s[4] for s in lib_spec]),
                        lib_weight=np.array([s[5] for s in lib_spec]), lib_short_frames=np.array(40),
                        root_rot=npf(mlib._frame_root_rot), joint_rot=npf(mlib._frame_joint_rot),
                        root_vel=npf(mlib._frame_root_vel), root_ang_vel=npf(mlib._frame_root_ang_vel),
                        dof_vel=npf(mlib._frame_dof_vel), lengths=npf(mlib._motion_lengths),
                        start_idx=npf(mlib._motion_start_idx), root_pos_delta=npf(mlib._motion_root_pos_delta),
                        weights=npf(mlib._motion_weights))

    # the motion_frames loader (with its fps-as-dt quirk, anim/motion_lib.py:178)
    mf = torch.tensor(np.stack([clips_raw["civilization"]["frames"][:30], clips_raw["civilization"]["frames"][100:130]]))
    mc = torch.tensor(np.stack([clips_raw["civilization"]["contacts"][:30], clips_raw["civilization"]["contacts"][100:130]]))
    mlib2 = ref_mlib.MotionLib(mf, km, "cpu", init_type="motion_frames", loop_mode=ref_mlib.LoopMode.CLAMP, fps=30,
                               contact_info=True, contacts=mc)
    np.savez_compressed(os.path.join(GOLD, "tables_motion_frames_golden.npz"), frames=npf(mf), contacts=npf(mc),
                        root_rot=npf(mlib2._frame_root_rot), joint_rot=npf(mlib2._frame_joint_rot),
                        root_vel=npf(mlib2._frame_root_vel), root_ang_vel=npf(mlib2._frame_root_ang_vel),
                        dof_vel=npf(mlib2._frame_dof_vel), root_pos_delta=npf(mlib2._motion_root_pos_delta),
                        lengths=npf(mlib2._motion_lengths))

    # ------------------------------------------------------------------ 3. frame queries + FK
    g = torch.Generator().manual_seed(7)
    lens = mlib._motion_lengths
    edge_ids = torch.tensor([0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 2, 2, 2, 2], dtype=torch.long)
    edge_t = torch.tensor([0.0, 8.4333, 8.5, 100.0, -1.0, 4.2,
                           0.0, lens[1].item(), lens[1].item() * 2.5, -0.7, 1.0 / 30.0,
                           0.0, lens[2].item() * 7.25, -lens[2].item() * 1.5, 0.3], dtype=torch.float32)
    rnd_ids = torch.randint(0, 3, (241,), generator=g)
    rnd_t = (torch.rand(241, generator=g) * 1.6 - 0.3) * lens[rnd_ids]
    # exact key-frame times (blend == 0 and the i/fps rounding cases)
    key_ids = torch.zeros(44, dtype=torch.long)
    key_t = torch.arange(44, dtype=torch.float32) * 6.0 / 30.0
    ids = torch.cat([edge_ids, rnd_ids, key_ids])
    times = torch.cat([edge_t, rnd_t, key_t])
    i0, i1, bl = mlib._calc_frame_blend(ids, times)
    fr = mlib.calc_motion_frame(ids, times)
    body_pos, body_rot = km.forward_kinematics(fr[0], fr[1], fr[4])
    o_i0, o_i1, o_bl = O.frame_blend(tb, ids, times)
    o_fr = O.calc_motion_frame(tb, ids, times)
    o_bp, o_br = O.forward_kinematics(model, o_fr[0], o_fr[1], o_fr[4])
    pin("query.idx0", i0, o_i0); pin("query.idx1", i1, o_i1); pin("query.blend", bl, o_bl)
    for nm, a, b in zip(("root_pos", "root_rot", "root_vel", "root_ang_vel", "joint_rot", "dof_vel", "contacts"), fr, o_fr):
        pin("query." + nm, a, b)
    pin("fk.body_pos", body_pos, o_bp); pin("fk.body_rot", body_rot, o_br)
    pin("query.phase", mlib.calc_motion_phase(ids, times), O.motion_phase(tb, ids, times))
    fidx = torch.randint(0, 40, (64,), generator=g)
    fids = torch.randint(0, 3, (64,), generator=g)
    gf = mlib.get_motion_frame(fids, fidx)
    for nm, a, b in zip(("root_pos", "root_rot", "root_vel", "root_ang_vel", "joint_rot", "dof_vel", "contacts"), gf,
                        O.get_motion_frame(tb, fids, fidx)):
        pin("get_frame." + nm, a, b)
    np.savez_compressed(os.path.join(GOLD, "query_golden.npz"), ids=npf(ids), times=npf(times), idx0=npf(i0), idx1=npf(i1),
                        blend=npf(bl), root_pos=npf(fr[0]), root_rot=npf(fr[1]), root_vel=npf(fr[2]),
                        root_ang_vel=npf(fr[3]), joint_rot=npf(fr[4]), dof_vel=npf(fr[5]), contacts=npf(fr[6]),
                        body_pos=npf(body_pos), body_rot=npf(body_rot), get_ids=npf(fids), get_fidx=npf(fidx),
                        get_root_pos=npf(gf[0]), get_joint_rot=npf(gf[4]), get_dof_vel=npf(gf[5]),
                        get_contacts=npf(gf[6]))

    # dof_to_rot / rot_to_dof / exp_map_to_quat on the clip's own frames
    frames_t = torch.tensor(clips_raw["civilization"]["frames"])
    jr = km.dof_to_rot(frames_t[:, 6:])
    pin("dof_to_rot", jr, O.dof_to_rot(model, frames_t[:, 6:]))
    pin("rot_to_dof", km.rot_to_dof(jr), O.rot_to_dof(model, jr))
    rq = ref_tu.exp_map_to_quat(frames_t[:, 3:6])
    pin("exp_map_to_quat", rq, O.exp_map_to_quat(frames_t[:, 3:6]))
    np.savez_compressed(os.path.join(GOLD, "dof_golden.npz"), joint_rot=npf(jr), root_quat=npf(rq),
                        dof_back=npf(km.rot_to_dof(jr)))

    # ------------------------------------------------------------------ 4. heightfield observations
    civ = clips_raw["civilization"]
    terr = ref_terrain.SubTerrain("t", x_dim=50, y_dim=50, dx=0.4, dy=0.4, min_x=0.0, min_y=0.0, device="cpu")
    terr.hf = torch.tensor(civ["hf"]); terr.min_point = torch.tensor(civ["min_point"]); terr.dxdy = torch.tensor(civ["dxdy"])
    ot = O.Terrain(hf=terr.hf, min_point=terr.min_point, dxdy=terr.dxdy)
    tmpl = ref_geom.get_xy_points_cone(center=torch.zeros(2), dx=0.05, num_neg=2, num_pos=60, num_rays_neg=3,
                                       num_rays_pos=3, angle_between_rays=0.26179938779)
    pin("cone_template", tmpl, O.cone_template(0.05, 2, 60, 3, 3, 0.26179938779))
    sel = torch.arange(0, 254, 4)
    root_pos_sel = fr_root = torch.tensor(civ["frames"][sel.numpy(), 0:3])
    root_q_sel = ref_tu.exp_map_to_quat(torch.tensor(civ["frames"][sel.numpy(), 3:6]))
    heading = ref_tu.calc_heading(root_q_sel)
    pin("calc_heading", heading, O.calc_heading(root_q_sel))

    # restated env method (envs/ig_parkour/mgdm_dm_util.py:158-179) composed from the REAL reference functions
    def ref_ray_obs(root_xyz, hd):
        n = root_xyz.shape[0]
        rp = tmpl.unsqueeze(0).expand(n, -1, -1)
        h = hd.unsqueeze(-1).expand(-1, rp.shape[1])
        xy = ref_tu.rotate_2d_vec(rp, h) + root_xyz[..., 0:2].unsqueeze(1)
        xy = xy.view(-1, 2)
        z = ref_terrain.get_local_hf_from_terrain(xy, terr).view(n, -1)
        return torch.clamp(z - root_xyz[..., 2].unsqueeze(-1), min=-3.0, max=3.0), xy

    obs, obs_xy = ref_ray_obs(root_pos_sel, heading)
    pin("ray_obs", obs, O.ray_obs(ot, root_pos_sel, heading, tmpl))
    gi = terr.get_grid_index(obs_xy)
    pin("grid_index", gi, O.grid_index(ot, obs_xy))
    probe_xy = torch.tensor([[0.2, 0.6], [-5.0, 3.0], [100.0, 0.19999], [0.6, 1.0], [19.6, 19.8], [1.0, 1.4]])
    pin("grid_index.probe", terr.get_grid_index(probe_xy), O.grid_index(ot, probe_xy))
    gobs = ref_terrain.sample_hf_z_on_terrain(terr, root_pos_sel[:16, 0:2], heading[:16], 0.2, 0.2, 15, 15, 15, 15)
    gt = O.grid_template(0.2, 0.2, 15, 15, 15, 15)
    pin("grid_obs", gobs, O.grid_obs(ot, root_pos_sel[:16, 0:2], heading[:16], gt))
    np.savez_compressed(os.path.join(GOLD, "obs_golden.npz"), tmpl=npf(tmpl), root_pos=npf(root_pos_sel),
                        root_quat=npf(root_q_sel), heading=npf(heading), ray_obs=npf(obs), ray_xy=npf(obs_xy),
                        ray_grid_index=npf(gi), ray_grid_coord=npf(O.grid_coord(ot, obs_xy)),
                        probe_xy=npf(probe_xy), probe_index=npf(terr.get_grid_index(probe_xy)),
                        grid_obs=npf(gobs), grid_tmpl=npf(gt))

    # ------------------------------------------------------------------ 5. SDF + losses
    # (a) the survey's 4x4 probe
    hf4 = torch.zeros(1, 4, 4); hf4[0, 2, 2] = 1.0
    pts4 = torch.tensor([[[0.8, 0.8, 0.5], [0.8, 0.8, 1.5], [0.0, 0.0, -0.3], [5.0, 5.0, 0.5], [0.4, 0.6, 0.2]]])
    mc4 = torch.zeros(1, 2); dxdy = torch.tensor([0.4, 0.4])
    sd4_inv = ref_terrain.points_hf_sdf(pts4, hf4, mc4, dxdy, base_z=-10.0, inverted=True)
    sd4_sol = ref_terrain.points_hf_sdf(pts4, hf4, mc4, dxdy, base_z=-10.0, inverted=False)
    pin("sdf.probe.inv", sd4_inv, O.points_hf_sdf(pts4, hf4, mc4, dxdy, -10.0, True))
    pin("sdf.probe.sol", sd4_sol, O.points_hf_sdf(pts4, hf4, mc4, dxdy, -10.0, False))
    # (b) body points of 6 clip frames against the clip's 50x50 terrain
    fsel = [0, 50, 100, 150, 200, 250]
    rp6 = torch.tensor(civ["frames"][fsel, 0:3]); rq6 = ref_tu.exp_map_to_quat(torch.tensor(civ["frames"][fsel, 3:6]))
    jr6 = km.dof_to_rot(torch.tensor(civ["frames"][fsel, 6:]))
    bp6, br6 = km.forward_kinematics(rp6, rq6, jr6)
    wp = torch.cat([ref_tu.quat_rotate(br6[:, b].unsqueeze(1), body_points[b].unsqueeze(0)) + bp6[:, b].unsqueeze(1)
                    for b in range(J)], dim=1).reshape(1, -1, 3)
    hfb = terr.hf.unsqueeze(0); mcb = terr.min_point.unsqueeze(0)
    sdf_inv = ref_terrain.points_hf_sdf(wp, hfb, mcb, terr.dxdy, base_z=-10.0, inverted=True)
    sdf_sol = ref_terrain.points_hf_sdf(wp, hfb, mcb, terr.dxdy, base_z=-10.0, inverted=False)
    pin("sdf.clip.inv", sdf_inv, O.points_hf_sdf(wp, hfb, mcb, terr.dxdy, -10.0, True))
    pin("sdf.clip.sol", sdf_sol, O.points_hf_sdf(wp, hfb, mcb, terr.dxdy, -10.0, False))
    np.savez_compressed(os.path.join(GOLD, "sdf_golden.npz"), probe_points=npf(pts4), probe_hf=npf(hf4),
                        probe_inv=npf(sd4_inv), probe_sol=npf(sd4_sol), clip_points=npf(wp), clip_inv=npf(sdf_inv),
                        clip_sol=npf(sdf_sol))

    # (c) compute_motion_loss, B=2, F=10 (frames 60..69 and 200..209, second one sunk 0.15 m so it penetrates)
    def mf_batch(starts, F, dz):
        rp = torch.stack([torch.tensor(civ["frames"][s:s + F, 0:3]) for s in starts])
        for i, d in enumerate(dz):
            rp[i, :, 2] += d
        rq = torch.stack([ref_tu.exp_map_to_quat(torch.tensor(civ["frames"][s:s + F, 3:6])) for s in starts])
        jr = torch.stack([km.dof_to_rot(torch.tensor(civ["frames"][s:s + F, 6:])) for s in starts])
        ct = torch.stack([torch.tensor(civ["contacts"][s:s + F]) for s in starts])
        return rp, rq, jr, ct

    rp, rq, jr, ct = mf_batch([60, 200], 10, [0.0, -0.15])
    ct = ct.clone(); ct[0, 3, 11] = -0.02          # a small negative contact, as MDM output can have
    mframes = ref_mu.MotionFrames(root_pos=rp, root_rot=rq, joint_rot=jr, contacts=ct)
    ml = ref_mdm_path.compute_motion_loss(mframes, None, terr, km, body_points, w_contact=0.1, w_pen=0.1, w_path=0.0,
                                          verbose=False)
    oml = O.compute_motion_loss(model, rp, rq, jr, ct, terr.hf, terr.min_point, terr.dxdy, 0.1, 0.1)
    for k in ("total_loss", "contact_loss", "pen_loss"):
        pin("compute_motion_loss." + k, ml[k], oml[k])

    # (d) motion_terrain_contact_loss, F=8 frames 200..207 sunk by 0.12 m: value + gradients of the leaves
    F = 8
    s = 200
    tgt_rp = torch.tensor(civ["frames"][s:s + F, 0:3]).clone(); tgt_rp[:, 2] -= 0.12
    tgt_re = torch.tensor(civ["frames"][s:s + F, 3:6]).clone()
    tgt_jd = torch.tensor(civ["frames"][s:s + F, 6:]).clone()
    cts = torch.tensor(civ["contacts"][s:s + F]).clone()
    src_rp = tgt_rp.clone() + 0.01
    src_rq = ref_tu.exp_map_to_quat(tgt_re * 1.05)
    src_jr = km.dof_to_rot(tgt_jd * 0.97)
    sbp, sbr = km.forward_kinematics(src_rp, src_rq, src_jr)
    src_bv = sbp[1:] - sbp[:-1]
    src_brv = ref_tu.quat_diff_angle(sbr[1:], sbr[:-1])

    def run_ref(w_pen, w_con, others):
        a, b, c = (t.clone().requires_grad_(True) for t in (tgt_rp, tgt_re, tgt_jd))
        loss, ld = ref_mopt.motion_terrain_contact_loss(
            a, b, c, src_rp, src_rq, src_jr, src_bv, src_brv, cts, terr, body_points, km,
            w_root_pos=others, w_root_rot=others, w_joint_rot=others, w_smoothness=others, w_penetration=w_pen,
            w_contact=w_con, w_sliding=others, w_body_constraints=0.0, w_jerk=others, body_constraints=None,
            max_jerk=1000.0)
        loss.backward()
        return loss.detach(), ld, a.grad, b.grad, c.grad

    loss_pc, ld_pc, g_rp, g_re, g_jd = run_ref(1000.0, 1000.0, 0.0)
    a, b, c = (t.clone().requires_grad_(True) for t in (tgt_rp, tgt_re, tgt_jd))
    o_loss, o_pen, o_con = O.motion_opt_pen_contact(model, a, b, c, cts, terr.hf, terr.min_point, terr.dxdy, 1000.0, 1000.0)
    o_loss.backward()
    pin("motion_opt.loss(pen+contact)", loss_pc, o_loss.detach())
    pin("motion_opt.pen", torch.tensor(ld_pc[ref_mopt.LossType.PENETRATION_LOSS]), o_pen.detach())
    pin("motion_opt.grad_root_pos", g_rp, a.grad); pin("motion_opt.grad_root_rot", g_re, b.grad)
    pin("motion_opt.grad_joint_dof", g_jd, c.grad)
    loss_all, ld_all, ga_rp, ga_re, ga_jd = run_ref(1000.0, 1000.0, 1.0)
    np.savez_compressed(
        os.path.join(GOLD, "loss_golden.npz"),
        ml_root_pos=npf(rp), ml_root_rot=npf(rq), ml_joint_rot=npf(jr), ml_contacts=npf(ct),
        ml_total=npf(ml["total_loss"]), ml_contact=npf(ml["contact_loss"]), ml_pen=npf(ml["pen_loss"]),
        mo_root_pos=npf(tgt_rp), mo_root_exp=npf(tgt_re), mo_joint_dof=npf(tgt_jd), mo_contacts=npf(cts),
        mo_src_root_pos=npf(src_rp), mo_src_root_quat=npf(src_rq), mo_src_joint_rot=npf(src_jr),
        mo_src_body_vels=npf(src_bv), mo_src_body_rot_vels=npf(src_brv),
        mo_loss_pc=npf(loss_pc), mo_pen=np.array(ld_pc[ref_mopt.LossType.PENETRATION_LOSS]),
        mo_contact=np.array(float(ld_pc[ref_mopt.LossType.CONTACT_LOSS])),
        mo_grad_root_pos=npf(g_rp), mo_grad_root_exp=npf(g_re), mo_grad_joint_dof=npf(g_jd),
        mo_loss_all=npf(loss_all), mo_all_grad_root_pos=npf(ga_rp), mo_all_grad_root_exp=npf(ga_re),
        mo_all_grad_joint_dof=npf(ga_jd),
        mo_all_terms=np.array([float(ld_all[k]) for k in ref_mopt.LossType if k in ld_all], np.float64),
        mo_all_term_names=np.array([k.name for k in ref_mopt.LossType if k in ld_all]))

    # ------------------------------------------------------------------ 6. contact labelling + hf masks (8(f)-1)
    lframes = torch.tensor(civ["frames"][::4].copy())             # 64 frames
    lframes[:, 2] -= 0.03                                         # sink a little so the feet penetrate somewhere
    upd, fc = ref_mel.compute_hf_foot_contacts_and_correct_pen(lframes, terr, km)
    hc = ref_mel.compute_motion_terrain_hand_contacts(lframes, terr, km)
    terr2 = ref_terrain.SubTerrain("t2", x_dim=50, y_dim=50, dx=0.4, dy=0.4, min_x=0.0, min_y=0.0, device="cpu")
    terr2.hf = torch.tensor(civ["hf"]) + 0.75                     # raised terrain: hands do touch
    terr2.min_point = terr.min_point.clone(); terr2.dxdy = terr.dxdy.clone()
    hc2 = ref_mel.compute_motion_terrain_hand_contacts(lframes, terr2, km)
    inds, minh = ref_terrain.compute_hf_mask_inds(lframes[:24], terr, km, body_points)
    mask = ref_terrain.compute_hf_mask_from_inds(terr, inds)
    t3 = terr.torch_copy()
    ref_terrain.compute_hf_extra_vals(lframes[:24], t3, km, body_points)
    lf, rf = km.get_body_id("left_foot"), km.get_body_id("right_foot")
    lh, rh = km.get_body_id("left_hand"), km.get_body_id("right_hand")
    feet = [(b, km._geoms[b][0]._dims.tolist(), km._geoms[b][0]._offset.tolist()) for b in (lf, rf)]
    hands = [(b, km._geoms[b][0]._dims.item()) for b in (lh, rh)]
    o_upd, o_fc, _ = O.foot_contacts_and_pen(model, lframes, ot, feet)
    pin("label.foot_contacts", fc, o_fc); pin("label.updated_frames", upd, o_upd)
    pin("label.hand_contacts", hc, O.hand_contacts(model, lframes, ot, hands))
    pin("label.hand_contacts_raised", hc2,
        O.hand_contacts(model, lframes, O.Terrain(hf=terr2.hf, min_point=terr2.min_point, dxdy=terr2.dxdy), hands))
    o_inds, o_minh = O.hf_mask_inds(model, lframes[:24], ot)
    pin("label.min_body_heights", minh, o_minh)
    pin("label.hf_mask_inds(concat)", torch.cat(inds), torch.cat(o_inds))
    pin("label.hf_mask_counts", torch.tensor([i.shape[0] for i in inds]), torch.tensor([i.shape[0] for i in o_inds]))
    np.savez_compressed(
        os.path.join(GOLD, "label_golden.npz"), frames=npf(lframes), foot_contacts=npf(fc), updated_z=npf(upd[:, 2]),
        hand_contacts=npf(hc), hand_contacts_raised=npf(hc2), raised_by=np.float32(0.75),
        mask_inds=npf(torch.cat(inds)), mask_counts=np.array([i.shape[0] for i in inds]), min_body_heights=npf(minh),
        hf_mask=npf(mask), extra_hf_mask=npf(t3.hf_mask), extra_hf_maxmin=npf(t3.hf_maxmin),
        feet_body=np.array([lf, rf]), feet_half=np.array([f[1] for f in feet], np.float32),
        feet_offset=np.array([f[2] for f in feet], np.float32), hands_body=np.array([lh, rh]),
        hands_radius=np.array([h[1] for h in hands], np.float32))

    # ------------------------------------------------------------------ 7. full objective incl. body constraints + Adam loop (8(f)-2)
    W = dict(w_root_pos=1.0, w_root_rot=10.0, w_joint_rot=1.0, w_smoothness=10.0, w_penetration=1000.0,
             w_contact=1000.0, w_sliding=10.0, w_body_constraints=1000.0, w_jerk=1000.0)      # PARC/kin_gen_default.yaml:28-37
    bcs = [[] for _ in range(J)]

    def mk_bc(s_, e_, pt):
        c = ref_mopt.BodyConstraint()
        c.start_frame_idx, c.end_frame_idx, c.constraint_point = s_, e_, torch.tensor(pt, dtype=torch.float32)
        return c

    with torch.no_grad():
        tq = ref_tu.exp_map_to_quat(tgt_re)
        tj = km.dof_to_rot(tgt_jd)
        tbp, _ = km.forward_kinematics(tgt_rp, tq, tj)
    bcs[lf] = [mk_bc(1, 4, (tbp[2, lf] + torch.tensor([0.02, -0.01, -0.06])).tolist())]
    bcs[rh] = [mk_bc(2, 6, (tbp[4, rh] + torch.tensor([0.05, 0.03, -0.04])).tolist())]

    def run_ref_bc():
        a, b, c = (t.clone().requires_grad_(True) for t in (tgt_rp, tgt_re, tgt_jd))
        loss, ld = ref_mopt.motion_terrain_contact_loss(
            a, b, c, src_rp, src_rq, src_jr, src_bv, src_brv, cts, terr, body_points, km, body_constraints=bcs,
            max_jerk=1000.0, **W)
        loss.backward()
        return loss.detach(), ld, a.grad, b.grad, c.grad

    loss_bc, ld_bc, gb_rp, gb_re, gb_jd = run_ref_bc()
    geom0 = [(km._geoms[b][0]._shape_type.value, npf(km._geoms[b][0]._offset).tolist(),
              npf(km._geoms[b][0]._dims).reshape(-1).tolist() if km._geoms[b][0]._dims.dim() > 0 else float(km._geoms[b][0]._dims))
             for b in range(J)]
    o_bcs = [[(c.start_frame_idx, c.end_frame_idx, c.constraint_point.tolist()) for c in bcs[b]] for b in range(J)]
    a, b, c = (t.clone().requires_grad_(True) for t in (tgt_rp, tgt_re, tgt_jd))
    o_loss, o_terms = O.motion_terrain_contact_loss_full(model, a, b, c, src_rp, src_rq, src_jr, src_bv, src_brv, cts, terr.hf,
                                                         terr.min_point, terr.dxdy, W, 1000.0, o_bcs, geom0)
    o_loss.backward()
    pin("motion_opt_full.loss", loss_bc, o_loss.detach())
    pin("motion_opt_full.body_constraint_term", torch.tensor(float(ld_bc[ref_mopt.LossType.BODY_CONSTRAINT_LOSS])),
        torch.tensor(float(o_terms["body_constraint"])))
    pin_close("motion_opt_full.grad_root_pos", gb_rp, a.grad, 1e-6); pin_close("motion_opt_full.grad_root_rot", gb_re, b.grad, 1e-6)
    pin_close("motion_opt_full.grad_joint_dof", gb_jd, c.grad, 1e-6)

    # the Adam loop (motion_optimization.py:404-500) for 4 iterations, built from the reference's own loss function
    src_fr = torch.cat([tgt_rp, tgt_re, tgt_jd], dim=-1)
    s_rq = ref_tu.exp_map_to_quat(src_fr[:, 3:6]); s_jr = km.dof_to_rot(src_fr[:, 6:34])
    s_bp, s_br = km.forward_kinematics(src_fr[:, 0:3], s_rq, s_jr)
    s_bv = s_bp[1:] - s_bp[:-1]; s_brv = ref_tu.quat_diff_angle(s_br[1:], s_br[:-1])
    leaves = [src_fr[:, 0:3].clone().requires_grad_(True), src_fr[:, 3:6].clone().requires_grad_(True),
              src_fr[:, 6:34].clone().requires_grad_(True)]
    opt = torch.optim.Adam(leaves, lr=0.001)
    for _ in range(4):
        opt.zero_grad()
        l_, _d = ref_mopt.motion_terrain_contact_loss(leaves[0], leaves[1], leaves[2], src_fr[:, 0:3], s_rq, s_jr, s_bv, s_brv,
                                                     cts, terr, body_points, km, body_constraints=bcs, max_jerk=1000.0, **W)
        l_.backward()
        opt.step()
    ref_opt = torch.cat([t.detach() for t in leaves], dim=-1)
    o_opt = O.motion_contact_optimization(model, src_fr, cts, terr.hf, terr.min_point, terr.dxdy, 4, 0.001, W, 1000.0, o_bcs, geom0)
    pin_close("motion_opt_full.adam4_frames", ref_opt - src_fr, o_opt - src_fr, 1e-4)    # the 4-step UPDATE itself
    np.savez_compressed(
        os.path.join(GOLD, "motion_opt_golden.npz"), src_frames=npf(src_fr), contacts=npf(cts),
        src_root_pos=npf(src_rp), src_root_quat=npf(src_rq), src_joint_rot=npf(src_jr), src_body_vels=npf(src_bv),
        src_body_rot_vels=npf(src_brv), weights=np.array([W[k] for k in sorted(W)]), weight_names=np.array(sorted(W)),
        bc_body=np.array([lf, rh]), bc_start=np.array([1, 2]), bc_end=np.array([4, 6]),
        bc_point=np.stack([npf(bcs[lf][0].constraint_point), npf(bcs[rh][0].constraint_point)]),
        loss=npf(loss_bc), body_constraint_term=np.array(float(ld_bc[ref_mopt.LossType.BODY_CONSTRAINT_LOSS])),
        grad_root_pos=npf(gb_rp), grad_root_exp=npf(gb_re), grad_joint_dof=npf(gb_jd), adam4_frames=npf(ref_opt))

    # ------------------------------------------------------------------ 9. tracker step assembly
    g = torch.Generator().manual_seed(99)
    NE, dt_ctrl = 96, 1.0 / 30.0
    steps = torch.tensor([1, 2, 3, 10, 20, 30], dtype=torch.float32)         # dm_env_default.yaml:166-172
    e_ids = torch.randint(0, 3, (NE,), generator=g)
    e_t = torch.rand(NE, generator=g) * lens[e_ids] * 1.1
    e_t[:4] = 0.0                                                            # first-step envs never fail
    cur = mlib.calc_motion_frame(e_ids, e_t)
    ref_bp, ref_br = km.forward_kinematics(cur[0], cur[1], cur[4])
    # simulated character = reference pose + noise of mixed size so that every done / reward branch is exercised
    amp = torch.where(torch.arange(NE) % 3 == 0, 0.6, 0.05).unsqueeze(-1)
    sim_root_pos = cur[0] + amp * 0.5 * torch.randn(NE, 3, generator=g)
    sim_root_rot = ref_tu.quat_mul(cur[1], ref_tu.exp_map_to_quat(amp * 1.5 * torch.randn(NE, 3, generator=g)))
    sim_root_vel = cur[2] + 0.3 * torch.randn(NE, 3, generator=g)
    sim_root_ang_vel = cur[3] + 0.3 * torch.randn(NE, 3, generator=g)
    sim_dof = km.rot_to_dof(cur[4]) + amp * 0.4 * torch.randn(NE, km.get_dof_size(), generator=g)
    sim_joint_rot = km.dof_to_rot(sim_dof)
    sim_dof_vel = cur[5] + 0.5 * torch.randn(NE, km.get_dof_size(), generator=g)
    sim_bp, _ = km.forward_kinematics(sim_root_pos, sim_root_rot, sim_joint_rot)
    key_ids_t = torch.tensor([km.get_body_id(nm) for nm in ("right_hand", "left_hand", "right_foot", "left_foot")])
    tar_rp, tar_rr, tar_jr, tar_ct = ref_dm_util.fetch_tar_obs_data(e_ids, e_t, mlib, dt_ctrl, steps)
    tar_bp, _ = km.forward_kinematics(tar_rp.reshape(-1, 3), tar_rr.reshape(-1, 4), tar_jr.reshape(-1, J - 1, 4))
    tar_key = tar_bp.reshape(NE, steps.shape[0], J, 3)[:, :, key_ids_t]
    none = torch.zeros([0])
    ts = {}
    for gl in (False, True):
        for rh in (False, True):
            r = ref_char_env.compute_char_obs(sim_root_pos, sim_root_rot, sim_root_vel, sim_root_ang_vel, sim_joint_rot,
                                              sim_dof_vel, sim_bp[:, key_ids_t], gl, rh)
            o = O.compute_char_obs(sim_root_pos, sim_root_rot, sim_root_vel, sim_root_ang_vel, sim_joint_rot,
                                   sim_dof_vel, sim_bp[:, key_ids_t], gl, rh)
            pin(f"step.char_obs(global={gl},root_h={rh})", r, o)
            ts[f"char_obs_g{int(gl)}_h{int(rh)}"] = npf(r)
            r = ref_dm_util.compute_tar_obs(sim_root_pos, sim_root_rot, tar_rp.clone(), tar_rr.clone(), tar_jr, tar_key.clone(), gl, rh)
            o = O.compute_tar_obs(sim_root_pos, sim_root_rot, tar_rp, tar_rr, tar_jr, tar_key, gl, rh)
            pin(f"step.tar_obs(global={gl},tar_h={rh})", r, o)
            ts[f"tar_obs_g{int(gl)}_h{int(rh)}"] = npf(r)
    r = ref_char_env.compute_char_obs(sim_root_pos, sim_root_rot, sim_root_vel, sim_root_ang_vel, sim_joint_rot, sim_dof_vel, none, False, False)
    pin("step.char_obs(no key bodies)", r, O.compute_char_obs(sim_root_pos, sim_root_rot, sim_root_vel, sim_root_ang_vel,
                                                               sim_joint_rot, sim_dof_vel, none, False, False))
    ts["char_obs_nokey"] = npf(r)
    r = ref_dm_util.compute_tar_obs(sim_root_pos, sim_root_rot, tar_rp.clone(), tar_rr.clone(), tar_jr, none, False, False)
    pin("step.tar_obs(no key bodies)", r, O.compute_tar_obs(sim_root_pos, sim_root_rot, tar_rp, tar_rr, tar_jr, none, False, False))
    ts["tar_obs_nokey"] = npf(r)
    jw = torch.tensor([1.0, 0.6, 0.6, 0.4, 0.0, 0.6, 0.4, 0.0, 1.0, 0.6, 0.4, 1.0, 0.6, 0.4])     # dm_env_default.yaml:99-113
    dw = torch.zeros(km.get_dof_size())
    for j in range(1, J):
        if km.get_joint_dof_dim(j) > 0:
            dw[km.get_joint_dof_idx(j):km.get_joint_dof_idx(j) + km.get_joint_dof_dim(j)] = jw[j - 1]
    for tr in (True, False):
        for th in (True, False):
            args = (sim_root_pos, sim_root_rot, sim_root_vel, sim_root_ang_vel, sim_joint_rot, sim_dof_vel, sim_bp[:, key_ids_t],
                    cur[0], cur[1], cur[2], cur[3], cur[4], cur[5], ref_bp[:, key_ids_t], jw, dw, th, tr)
            r = ref_dm_util.compute_deepmimic_reward(*args)
            pin(f"step.reward(track_root={tr},track_h={th})", r, O.compute_deepmimic_reward(*args))
            ts[f"reward_r{int(tr)}_h{int(th)}"] = npf(r)
    try:        # without key bodies the reference's torch.stack sees a [0] tensor next to [N] ones and raises
        ref_dm_util.compute_deepmimic_reward(sim_root_pos, sim_root_rot, sim_root_vel, sim_root_ang_vel, sim_joint_rot, sim_dof_vel,
                                             none, cur[0], cur[1], cur[2], cur[3], cur[4], cur[5], none, jw, dw, True, True)
        raise SystemExit("reference reward unexpectedly accepts an empty key-body set")
    except RuntimeError:
        REPORT.append("OK  step.reward(no key bodies) raises in the reference")
    # done flags, with the termination-height lookup of RefCharEnv.update_done (mgdm_dm_util.py:205-230)
    terr_s = ref_terrain.SubTerrain("step", 40, 32, 0.4, 0.4, -3.0, -2.0, device="cpu")
    # heights straddle the lowest bodies so that the fall test (height AND contact force) fires for some envs
    terr_s.hf[...] = torch.rand(40, 32, generator=g) * 0.8 + (sim_bp[..., 2].min(dim=-1)[0].median() - 0.4)
    env_off = torch.randn(NE, 3, generator=g) * 0.5
    o_terr = O.Terrain(hf=terr_s.hf.clone(), min_point=terr_s.min_point.clone(), dxdy=terr_s.dxdy.clone())
    ptd = torch.tensor([0.7, 1.0, 0.7, 0.7, 0.7, 0.7, 0.7, 0.7, 1.0, 1.2, 10.0, 1.0, 1.2, 10.0])   # dm_env_default.yaml:129-143
    time_buf = torch.rand(NE, generator=g) * 12.0
    time_buf[:4] = 0.0
    forces = torch.randn(NE, J, 3, generator=g) * (torch.rand(NE, J, 1, generator=g) < 0.2)
    feet = torch.tensor([km.get_body_id("right_foot"), km.get_body_id("left_foot")])
    done_in = torch.zeros(NE, dtype=torch.int)
    for tag, cids, pt, tr in (("default", torch.zeros([0], dtype=torch.long), True, True), ("feet", feet, True, True),
                              ("feet_nopose", feet, False, True), ("noroot", feet, True, False)):
        gpos = sim_bp[..., 0:2] + env_off[:, 0:2].unsqueeze(1)
        gi = terr_s.get_grid_index(gpos)
        th_ref = terr_s.hf[gi[..., 0], gi[..., 1]] + 0.15
        th_o = O.termination_heights(o_terr, sim_bp, env_off, 0.15)
        pin(f"step.termination_heights[{tag}]", th_ref, th_o)
        r = ref_dm_util.compute_done(done_in, time_buf, 10.0, sim_root_rot, sim_bp, sim_root_pos, cur[1], ref_bp, forces, cids,
                                     th_ref, pt, ptd, False, True, tr, 0.6, 1.309)
        o = O.compute_done(done_in, time_buf, 10.0, sim_root_rot, sim_bp, cur[1], ref_bp, forces, cids, th_o, pt, ptd, True,
                           tr, 0.6, 1.309)
        pin(f"step.done[{tag}]", r, o)
        ts[f"done_{tag}"] = npf(r)
        ts["term_heights"] = npf(th_ref)
    r = ref_dm_util.compute_done(done_in, time_buf, 10.0, sim_root_rot, sim_bp, sim_root_pos, cur[1], ref_bp, forces, feet,
                                 th_ref, True, ptd, False, False, True, 0.6, 1.309)
    pin("step.done[no early termination]", r, O.compute_done(done_in, time_buf, 10.0, sim_root_rot, sim_bp, cur[1], ref_bp, forces,
                                                             feet, th_o, True, ptd, False, True, 0.6, 1.309))
    ts["done_noearly"] = npf(r)
    np.savez_compressed(
        os.path.join(GOLD, "tracker_step_golden.npz"), ids=npf(e_ids), times=npf(e_t), steps=npf(steps), dt=np.float32(dt_ctrl),
        key_ids=npf(key_ids_t), root_pos=npf(sim_root_pos), root_rot=npf(sim_root_rot), root_vel=npf(sim_root_vel),
        root_ang_vel=npf(sim_root_ang_vel), joint_rot=npf(sim_joint_rot), dof_vel=npf(sim_dof_vel), body_pos=npf(sim_bp),
        ref_root_pos=npf(cur[0]), ref_root_rot=npf(cur[1]), ref_root_vel=npf(cur[2]), ref_root_ang_vel=npf(cur[3]),
        ref_joint_rot=npf(cur[4]), ref_dof_vel=npf(cur[5]), ref_body_pos=npf(ref_bp),
        tar_root_pos=npf(tar_rp), tar_root_rot=npf(tar_rr), tar_joint_rot=npf(tar_jr), tar_key_pos=npf(tar_key),
        tar_contacts=npf(tar_ct), joint_err_w=npf(jw), dof_err_w=npf(dw), hf=npf(terr_s.hf), hf_min=npf(terr_s.min_point),
        hf_dxdy=npf(terr_s.dxdy), env_offsets=npf(env_off), pose_termination_dist=npf(ptd), time_buf=npf(time_buf),
        contact_forces=npf(forces), feet=npf(feet), **ts)

    with open(os.path.join(GOLD, "PIN_REPORT.txt"), "w") as f:
        f.write("oracle/parc_oracle.py vs the imported reference (torch %s, CPU, fp32) -- torch.equal on every line\n"
                % torch.__version__)
        f.write("\n".join(REPORT) + "\n")
    print("\n".join(REPORT))
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
