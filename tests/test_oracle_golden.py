"""The oracle (oracle/parc_oracle.py) against the committed golden vectors.

The vectors are outputs of the REAL reference run in the authoring container by oracle/make_golden.py
(which also asserted torch.equal there).  Here they are re-checked wherever the suite runs: integer /
index results and IEEE-only arithmetic bit-exactly, transcendental-bearing values to 1e-6 (the host CPU's
vector math library may differ from the authoring container's).
"""
import numpy as np
import torch

from conftest import assert_close, golden, lib_clips_from_golden
from oracle import parc_oracle as O


def T(x):
    return torch.as_tensor(np.asarray(x))


def test_pin_report_is_all_ok():
    import os
    from conftest import GOLDEN
    lines = open(os.path.join(GOLDEN, "PIN_REPORT.txt")).read().splitlines()[1:]
    assert len(lines) >= 60 and all(l.startswith("OK") for l in lines)     # 'OK ' = torch.equal, 'OK~' = tolerance pin


def test_tables(oracle_model):
    g = golden("tables_golden.npz")
    tb = O.build_tables(oracle_model, lib_clips_from_golden())
    assert torch.equal(tb.start_idx, T(g["start_idx"])) and torch.equal(tb.lengths, T(g["lengths"]))
    assert torch.equal(tb.weights, T(g["weights"])) and torch.equal(tb.root_pos_delta, T(g["root_pos_delta"]))
    for k in ("root_rot", "joint_rot", "root_vel", "root_ang_vel", "dof_vel"):
        assert_close(getattr(tb, k), g[k], rtol=1e-6, atol=1e-6, what=k)


def test_query_indices_bit_exact_and_values(oracle_model):
    g = golden("query_golden.npz")
    tg = golden("tables_golden.npz")
    tb = O.build_tables(oracle_model, lib_clips_from_golden())
    # decouple from the host's libm: query on the golden tables themselves
    tb.root_rot, tb.joint_rot = T(tg["root_rot"]), T(tg["joint_rot"])
    tb.root_vel, tb.root_ang_vel, tb.dof_vel = T(tg["root_vel"]), T(tg["root_ang_vel"]), T(tg["dof_vel"])
    ids, times = T(g["ids"]), T(g["times"])
    i0, i1, bl = O.frame_blend(tb, ids, times)
    assert torch.equal(i0, T(g["idx0"])) and torch.equal(i1, T(g["idx1"])) and torch.equal(bl, T(g["blend"]))
    fr = O.calc_motion_frame(tb, ids, times)
    for k, t in zip(("root_pos", "root_rot", "root_vel", "root_ang_vel", "joint_rot", "dof_vel", "contacts"), fr):
        assert_close(t, g[k], rtol=1e-6, atol=1e-6, what=k)
    assert torch.equal(fr[0], T(g["root_pos"])) and torch.equal(fr[6], T(g["contacts"]))
    bp, br = O.forward_kinematics(oracle_model, T(g["root_pos"]), T(g["root_rot"]), T(g["joint_rot"]))
    assert_close(bp, g["body_pos"], rtol=1e-6, atol=1e-6, what="body_pos")
    assert_close(br, g["body_rot"], rtol=1e-6, atol=1e-6, what="body_rot")
    gf = O.get_motion_frame(tb, T(g["get_ids"]), T(g["get_fidx"]))
    assert torch.equal(gf[4], T(g["get_joint_rot"])) and torch.equal(gf[6], T(g["get_contacts"]))


def test_known_answers_from_survey(oracle_model):
    """SURVEY.md appendix A1 probe on the 254-frame clip."""
    tb = O.build_tables(oracle_model, lib_clips_from_golden()[:1])
    t = torch.tensor([0, 8.4333, 8.5, 100, -1, 4.2], dtype=torch.float32)
    i0, i1, bl = O.frame_blend(tb, torch.zeros(6, dtype=torch.long), t)
    assert i0.tolist() == [0, 252, 253, 253, 0, 125] and i1.tolist() == [1, 253, 253, 253, 1, 126]
    assert_close(bl, torch.tensor([0, 0.99900818, 0, 0, 0, 0.99999237]), rtol=0, atol=1e-7, what="blend")


def test_dof_conversions(oracle_model):
    civ, g = golden("clip_civilization.npz"), golden("dof_golden.npz")
    assert_close(O.dof_to_rot(oracle_model, T(civ["frames"][:, 6:])), g["joint_rot"], rtol=1e-6, atol=1e-6, what="dof_to_rot")
    assert_close(O.rot_to_dof(oracle_model, T(g["joint_rot"])), g["dof_back"], rtol=1e-6, atol=2e-6, what="rot_to_dof")
    assert_close(O.exp_map_to_quat(T(civ["frames"][:, 3:6])), g["root_quat"], rtol=1e-6, atol=1e-6, what="exp_map")


def test_heightfield(oracle_model):
    civ, g = golden("clip_civilization.npz"), golden("obs_golden.npz")
    t = O.Terrain(hf=T(civ["hf"]), min_point=T(civ["min_point"]), dxdy=T(civ["dxdy"]))
    assert torch.equal(O.grid_index(t, T(g["ray_xy"])), T(g["ray_grid_index"]))
    assert torch.equal(O.grid_index(t, T(g["probe_xy"])), T(g["probe_index"]))
    assert O.grid_index(t, torch.tensor([[0.2, 0.6]])).tolist() == [[0, 2]]     # half-to-even
    assert_close(O.cone_template(0.05, 2, 60, 3, 3, 0.26179938779), g["tmpl"], rtol=1e-6, atol=1e-7, what="template")
    obs = O.ray_obs(t, T(g["root_pos"]), T(g["heading"]), T(g["tmpl"]))
    assert (obs != T(g["ray_obs"])).float().mean() < 1e-3
    gobs = O.grid_obs(t, T(g["root_pos"][:16, 0:2]), T(g["heading"][:16]), T(g["grid_tmpl"]))
    assert (gobs != T(g["grid_obs"])).float().mean() < 2e-3
    assert_close(O.calc_heading(T(g["root_quat"])), g["heading"], rtol=1e-6, atol=1e-6, what="heading")


def test_sdf_and_losses(oracle_model):
    civ, g, L = golden("clip_civilization.npz"), golden("sdf_golden.npz"), golden("loss_golden.npz")
    dxdy = torch.tensor([0.4, 0.4])
    # SURVEY A8 probe values
    inv = O.points_hf_sdf(T(g["probe_points"]), T(g["probe_hf"]), torch.zeros(1, 2), dxdy, -10.0, True)
    sol = O.points_hf_sdf(T(g["probe_points"]), T(g["probe_hf"]), torch.zeros(1, 2), dxdy, -10.0, False)
    assert torch.equal(inv, T(g["probe_inv"])) and torch.equal(sol, T(g["probe_sol"]))
    assert_close(inv[0, :3], torch.tensor([-0.2, 0.2, -0.3]), rtol=1e-5, what="probe inverted")
    assert_close(sol[0, :2], torch.tensor([-0.2, 0.5]), rtol=1e-5, what="probe solid")
    hf, mp = T(civ["hf"]), T(civ["min_point"])
    assert torch.equal(O.points_hf_sdf(T(g["clip_points"]), hf[None], mp[None], dxdy, -10.0, True), T(g["clip_inv"]))
    assert torch.equal(O.points_hf_sdf(T(g["clip_points"]), hf[None], mp[None], dxdy, -10.0, False), T(g["clip_sol"]))
    out = O.compute_motion_loss(oracle_model, T(L["ml_root_pos"]), T(L["ml_root_rot"]), T(L["ml_joint_rot"]),
                                T(L["ml_contacts"]), hf, mp, dxdy, 0.1, 0.1)
    assert_close(out["total_loss"], L["ml_total"], rtol=1e-6, what="total")
    assert_close(out["pen_loss"], L["ml_pen"], rtol=1e-6, what="pen")
    a, b, c = (T(L[k]).clone().requires_grad_(True) for k in ("mo_root_pos", "mo_root_exp", "mo_joint_dof"))
    loss, pen, con = O.motion_opt_pen_contact(oracle_model, a, b, c, T(L["mo_contacts"]), hf, mp, dxdy, 1000.0, 1000.0)
    loss.backward()
    assert_close(loss, L["mo_loss_pc"], rtol=1e-6, what="motion-opt loss")
    assert_close(a.grad, L["mo_grad_root_pos"], rtol=1e-5, atol=1e-3, what="grad root_pos")
    assert_close(b.grad, L["mo_grad_root_exp"], rtol=1e-5, atol=1e-3, what="grad root_exp")
    assert_close(c.grad, L["mo_grad_joint_dof"], rtol=1e-5, atol=1e-3, what="grad joint_dof")


def test_contact_labelling_and_masks(oracle_model):
    civ, g = golden("clip_civilization.npz"), golden("label_golden.npz")
    t = O.Terrain(hf=T(civ["hf"]), min_point=T(civ["min_point"]), dxdy=T(civ["dxdy"]))
    frames = T(g["frames"])
    feet = [(int(b), h.tolist(), o.tolist()) for b, h, o in zip(g["feet_body"], g["feet_half"], g["feet_offset"])]
    hands = [(int(b), float(r)) for b, r in zip(g["hands_body"], g["hands_radius"])]
    upd, fc, _ = O.foot_contacts_and_pen(oracle_model, frames, t, feet)
    assert (fc != T(g["foot_contacts"])).sum() <= 1               # thresholded: tolerate one borderline frame
    assert_close(upd[:, 2], g["updated_z"], rtol=1e-6, atol=1e-6, what="updated root z")
    t2 = O.Terrain(hf=t.hf + float(g["raised_by"]), min_point=t.min_point, dxdy=t.dxdy)
    assert (O.hand_contacts(oracle_model, frames, t2, hands) != T(g["hand_contacts_raised"])).sum() <= 1
    assert (O.hand_contacts(oracle_model, frames, t, hands) != T(g["hand_contacts"])).sum() <= 1
    inds, minh = O.hf_mask_inds(oracle_model, frames[:24], t)
    assert [i.shape[0] for i in inds] == g["mask_counts"].tolist()
    assert torch.equal(torch.cat(inds), T(g["mask_inds"]))
    assert_close(minh, g["min_body_heights"], rtol=1e-6, atol=1e-6, what="min body heights")


def _step_states(g):
    sim = tuple(T(g[k]) for k in ("root_pos", "root_rot", "root_vel", "root_ang_vel", "joint_rot", "dof_vel"))
    ref = tuple(T(g["ref_" + k]) for k in ("root_pos", "root_rot", "root_vel", "root_ang_vel", "joint_rot", "dof_vel"))
    key_ids = T(g["key_ids"]).long()
    return sim, ref, key_ids


def test_tracker_step_obs_reward_done():
    """§8(f)-3: compute_char_obs / compute_tar_obs / compute_deepmimic_reward / compute_done restatements against
    the reference's outputs (flags bit-exact; values to 1e-6 across libm builds)."""
    g = golden("tracker_step_golden.npz")
    sim, ref, key_ids = _step_states(g)
    key = T(g["body_pos"])[:, key_ids]
    none = torch.zeros([0])
    tar = tuple(T(g["tar_" + k]) for k in ("root_pos", "root_rot", "joint_rot", "key_pos"))
    for gl in (0, 1):
        for h in (0, 1):
            assert_close(O.compute_char_obs(*sim, key, bool(gl), bool(h)), g[f"char_obs_g{gl}_h{h}"], what="char_obs")
            assert_close(O.compute_tar_obs(sim[0], sim[1], *tar, bool(gl), bool(h)), g[f"tar_obs_g{gl}_h{h}"],
                         what="tar_obs")
    assert_close(O.compute_char_obs(*sim, none, False, False), g["char_obs_nokey"])
    assert_close(O.compute_tar_obs(sim[0], sim[1], tar[0], tar[1], tar[2], none, False, False), g["tar_obs_nokey"])
    jw, dw = T(g["joint_err_w"]), T(g["dof_err_w"])
    ref_key = T(g["ref_body_pos"])[:, key_ids]
    for tr in (1, 0):
        for th in (1, 0):
            r = O.compute_deepmimic_reward(*sim, key, *ref, ref_key, jw, dw, bool(th), bool(tr))
            assert_close(r, g[f"reward_r{tr}_h{th}"], what="reward")
    terr = O.Terrain(hf=T(g["hf"]), min_point=T(g["hf_min"]), dxdy=T(g["hf_dxdy"]))
    th = O.termination_heights(terr, T(g["body_pos"]), T(g["env_offsets"]), 0.15)
    assert torch.equal(th, T(g["term_heights"]))
    feet = T(g["feet"]).long()
    done_in = torch.zeros(sim[0].shape[0], dtype=torch.int)
    common = (done_in, T(g["time_buf"]), 10.0, sim[1], T(g["body_pos"]), ref[1], T(g["ref_body_pos"]),
              T(g["contact_forces"]))
    ptd = T(g["pose_termination_dist"])
    cases = {"default": (torch.zeros([0], dtype=torch.long), True, True, True), "feet": (feet, True, True, True),
             "feet_nopose": (feet, False, True, True), "noroot": (feet, True, True, False),
             "noearly": (feet, True, False, True)}
    for tag, (cids, pose, early, track) in cases.items():
        d = O.compute_done(*common, cids, th, pose, ptd, early, track, 0.6, 1.309)
        assert d.dtype == torch.int32 and torch.equal(d, T(g["done_" + tag])), tag
    assert set(np.unique(g["done_feet"])) == {0, 1, 3}                    # every flag value is exercised


def _a16_points(g):
    cnt = g["minimal_point_count"].tolist()
    pts, s0 = [], 0
    for n in cnt:
        pts.append(T(g["minimal_points"][s0:s0 + n]))
        s0 += n
    return pts


def test_a16_pin_report_is_all_ok():
    import os
    from conftest import GOLDEN
    lines = open(os.path.join(GOLDEN, "PIN_REPORT_a16.txt")).read().splitlines()
    assert len(lines) == 12 and all(l.startswith("OK") for l in lines)


def test_mdm_hf_collision_loss_and_gradients(oracle_model):
    """util/terrain_util.py:1895-1949 and the MDM's 0.5 * sum(clamp(sdf, max=0)^2) through
    points_hf_sdf (diffusion/mdm.py:729-737, :1484-1496), values and autograd gradients of the reference."""
    g = golden("a16_golden.npz")
    hf, mc, dxdy = T(g["hf"]), T(g["min_center"]), T(g["dxdy"])
    pts = _a16_points(g)
    assert sum(p.shape[0] for p in pts) == 42
    for interior, tag in ((True, "int"), (False, "ext")):
        mf = T(g["frames"]).clone().requires_grad_(True)
        loss, wp, sdf = O.motion_frames_hf_sdf_loss(oracle_model, mf, pts, hf, mc, dxdy, interior_distance=interior)
        loss.sum().backward()
        assert_close(loss, g[f"loss_{tag}"], rtol=1e-5, atol=1e-7, what=f"loss {tag}")
        assert_close(sdf, g[f"sdf_{tag}"], rtol=1e-5, atol=2e-6, what=f"sdf {tag}")
        gr = T(g[f"grad_frames_{tag}"])
        assert (mf.grad - gr).abs().max() <= 2e-5 * gr.abs().max()
    for tag in ("guidance", "train"):
        p = T(g["world_points"]).clone().requires_grad_(True)
        sdf = O.points_hf_sdf(p, hf, mc, dxdy, base_z=float(g[f"base_z_{tag}"]), inverted=True)
        O.hf_collision_loss(sdf).sum().backward()
        assert torch.equal(sdf, T(g[f"psdf_{tag}"])) and torch.equal(p.grad, T(g[f"pgrad_{tag}"]))


def _sampler_case(g):
    terr, mms, inds = [], [], []
    for c in range(3):
        terr.append(O.Terrain(hf=T(g[f"hf{c}"]), min_point=T(g[f"min_point{c}"]), dxdy=T(g[f"dxdy{c}"])))
        mms.append(T(g[f"maxmin{c}"]))
        flat, cnt = T(g[f"mask_inds{c}"]), g[f"mask_count{c}"].tolist()
        per, s0 = [], 0
        for n in cnt:
            per.append(flat[s0:s0 + n])
            s0 += n
        inds.append(per)
    return terr, mms, inds


def test_mdm_sampler_terrain_gather():
    """SURVEY 8(f)-4 clause: diffusion/mdm_heightfield_contact_motion_sampler.py:414-474 on three clips with their own
    terrains (50x50, 102x102, 16x16) and body-cover masks, both relative-z styles."""
    g = golden("sampler_golden.npz")
    terr, mms, inds = _sampler_case(g)
    n = int(g["num_neg"])
    for tag, rel_root in (("relative_to_root_floor", False), ("relative_to_root", True)):
        hfs, ch, mm = O.clip_hfs_from_data(terr, mms, inds, T(g["ids"]), T(g["root_pos"]), T(g["root_rot"]), T(g["canon_z"]),
                                           T(g["mti"]), T(g["grid"]), n, n, float(g["max_h"]), rel_root)
        # host libm may round the heading's sin / cos differently from the authoring container: compare off-border
        coord = torch.as_tensor(g["grid_coord"]).double()
        border = ((coord - torch.floor(coord) - 0.5).abs() <= 8 * 1.2e-7 * coord.abs().clamp(min=1.0)).any(dim=-1)
        assert torch.equal(hfs[~border], T(g[f"hfs_{tag}"])[~border])
        assert torch.equal(mm[~border], T(g[f"mm_{tag}"])[~border])
        assert (ch != T(g[f"center_h_{tag}"])).sum() <= 1
    lines = open(__import__("os").path.join(__import__("conftest").GOLDEN, "PIN_REPORT_sampler.txt")).read().splitlines()
    assert len(lines) == 6 and all(l.startswith("OK") for l in lines)
