"""Parity tests proper: the CUDA path (through the C ABI) against the oracle and the golden vectors.

Tolerances (north_star): frame indices / grid indices bit-exact; fp32 values within 1e-5 relative
(assert_close: atol 1e-6 + rtol 1e-5); gradients 1e-5 of the tensor's max magnitude (norm-wise).
Heightmap observations are exact except where the sample lands within a few ulps of a cell border,
where sin/cos of the GPU and of the host may legitimately round the coordinate to the other cell.
"""
import os

import numpy as np
import pytest
import torch

from conftest import (assert_close, assert_close_normwise, golden, lib_clips_from_golden, write_clip_library)

pytestmark = pytest.mark.gpu

FRAME_KEYS = ("root_pos", "root_rot", "root_vel", "root_ang_vel", "joint_rot", "dof_vel", "contacts")


@pytest.fixture(scope="module")
def O():
    from oracle import parc_oracle
    return parc_oracle


@pytest.fixture(scope="module")
def golden_lib(gpu_model, tmp_path_factory):
    from parc_b200.anim.motion_lib import MotionLib
    y = write_clip_library(tmp_path_factory.mktemp("lib"), lib_clips_from_golden())
    return MotionLib(y, gpu_model, "cuda:0", init_type="motion_file", contact_info=True)


@pytest.fixture(scope="module")
def oracle_tables(O, oracle_model):
    return O.build_tables(oracle_model, lib_clips_from_golden())


def dev(x, dtype=None):
    t = torch.as_tensor(np.asarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.to("cuda:0")


# ----------------------------------------------------------------------------------------- tables
def test_packed_tables_match_golden(golden_lib):
    g = golden("tables_golden.npz")
    assert torch.equal(golden_lib._frame_root_rot.cpu(), torch.tensor(g["root_rot"]))
    assert torch.equal(golden_lib._frame_joint_rot.cpu(), torch.tensor(g["joint_rot"]))
    assert torch.equal(golden_lib._frame_dof_vel.cpu(), torch.tensor(g["dof_vel"]))
    assert torch.equal(golden_lib._motion_lengths.cpu(), torch.tensor(g["lengths"]))
    lay = golden_lib._packed.layout
    rows = golden_lib._packed.rows.cpu()
    assert rows.shape == (352, lay.row_floats) and lay.row_floats % 8 == 0
    assert torch.equal(rows[:, 4:8], torch.tensor(g["root_rot"]))
    assert torch.equal(rows[:, 8:64].reshape(-1, 14, 4), torch.tensor(g["joint_rot"]))
    v0 = lay.vel_slot * 4
    assert torch.equal(rows[:, v0:v0 + 3], torch.tensor(g["root_vel"]))
    assert torch.equal(rows[:, v0 + 4:v0 + 7], torch.tensor(g["root_ang_vel"]))
    assert torch.equal(rows[:, v0 + 8:v0 + 36], torch.tensor(g["dof_vel"]))


# ----------------------------------------------------------------------------------------- query
def test_frame_index_bit_exact_vs_golden(golden_lib):
    g = golden("query_golden.npz")
    i0, i1, bl = golden_lib._calc_frame_blend(dev(g["ids"]), dev(g["times"]))
    assert torch.equal(i0.cpu(), torch.tensor(g["idx0"]))
    assert torch.equal(i1.cpu(), torch.tensor(g["idx1"]))
    assert torch.equal(bl.cpu(), torch.tensor(g["blend"]))          # blend is exact too (IEEE ops only)


def test_calc_motion_frame_vs_golden(golden_lib):
    g = golden("query_golden.npz")
    out = golden_lib.calc_motion_frame(dev(g["ids"]), dev(g["times"]))
    assert len(out) == 7
    for k, t in zip(FRAME_KEYS, out):
        assert_close(t, g[k], what=f"calc_motion_frame.{k}")
    # gathers and lerps involve IEEE ops only -> exact
    assert torch.equal(out[2].cpu(), torch.tensor(g["root_vel"]))
    assert torch.equal(out[5].cpu(), torch.tensor(g["dof_vel"]))
    assert torch.equal(out[0].cpu(), torch.tensor(g["root_pos"]))
    assert torch.equal(out[6].cpu(), torch.tensor(g["contacts"]))


def test_get_motion_frame_vs_golden(golden_lib):
    g = golden("query_golden.npz")
    out = golden_lib.get_motion_frame(dev(g["get_ids"]), dev(g["get_fidx"]))
    assert torch.equal(out[0].cpu(), torch.tensor(g["get_root_pos"]))
    assert torch.equal(out[4].cpu(), torch.tensor(g["get_joint_rot"]))
    assert torch.equal(out[5].cpu(), torch.tensor(g["get_dof_vel"]))
    assert torch.equal(out[6].cpu(), torch.tensor(g["get_contacts"]))


def test_fk_vs_golden(golden_lib, gpu_model):
    g = golden("query_golden.npz")
    bp, br = gpu_model.forward_kinematics(dev(g["root_pos"]), dev(g["root_rot"]), dev(g["joint_rot"]))
    assert_close(bp, g["body_pos"], what="fk.body_pos")
    assert_close(br, g["body_rot"], what="fk.body_rot")
    # arbitrary leading dims
    bp2, br2 = gpu_model.forward_kinematics(dev(g["root_pos"]).view(3, 100, 3), dev(g["root_rot"]).view(3, 100, 4),
                                            dev(g["joint_rot"]).view(3, 100, 14, 4))
    assert bp2.shape == (3, 100, 15, 3) and torch.equal(bp2.view(300, 15, 3), bp)
    assert torch.equal(br2.view(300, 15, 4), br)


def test_fused_query_fk_matches_separate(golden_lib, gpu_model):
    g = golden("query_golden.npz")
    r = golden_lib.calc_motion_frame_fk_obs(dev(g["ids"]), dev(g["times"]))
    assert_close(r["body_pos"], g["body_pos"], what="fused.body_pos")
    assert_close(r["body_rot"], g["body_rot"], what="fused.body_rot")
    bp, br = gpu_model.forward_kinematics(r["root_pos"], r["root_rot"], r["joint_rot"])
    assert torch.equal(bp, r["body_pos"]) and torch.equal(br, r["body_rot"])


def test_query_vs_oracle_random_and_edges(golden_lib, O, oracle_tables, oracle_model):
    gen = torch.Generator().manual_seed(2024)
    n = 20000
    ids = torch.randint(0, 3, (n,), generator=gen)
    lens = oracle_tables.lengths[ids]
    times = (torch.rand(n, generator=gen) * 3.0 - 1.0) * lens
    # exact multiples of the frame period and of the clip length (blend == 0 / wrap boundaries)
    k = torch.arange(n) % 400
    times[::7] = (k[::7].float() / 30.0)
    times[::11] = lens[::11] * (k[::11] % 5).float()
    i0, i1, bl = O.frame_blend(oracle_tables, ids, times)
    ref = O.calc_motion_frame(oracle_tables, ids, times)
    rbp, rbr = O.forward_kinematics(oracle_model, ref[0], ref[1], ref[4])
    r = golden_lib.calc_motion_frame_fk_obs(ids.cuda(), times.cuda())
    g0, g1, gb = golden_lib._calc_frame_blend(ids.cuda(), times.cuda())
    assert torch.equal(g0.cpu(), i0) and torch.equal(g1.cpu(), i1) and torch.equal(gb.cpu(), bl)
    for k_, t in zip(FRAME_KEYS, ref):
        assert_close(r[k_], t, what=f"query.{k_}")
    assert_close(r["body_pos"], rbp, what="query.body_pos")
    assert_close(r["body_rot"], rbr, what="query.body_rot")


def test_tracker_step_form_matches_separate_queries(golden_lib, O, oracle_tables):
    """parc_motion_query_steps: N envs x 7 time offsets (current + tar_obs_steps look-aheads) in one launch ==
    the reference's tiled ids / `motion_times + timestep * tar_obs_steps` (mgdm_dm_util.py:279-302)."""
    g = golden("obs_golden.npz")
    t = _civ_terrain()
    gen = torch.Generator().manual_seed(31)
    n = 1001                                                        # odd: exercises the half-empty last warp
    ids = torch.randint(0, 3, (n,), generator=gen)
    times = torch.rand(n, generator=gen) * oracle_tables.lengths[ids]
    timestep = 1.0 / 30.0
    steps = torch.tensor([0, 1, 2, 3, 10, 20, 30], dtype=torch.float32)
    offsets = timestep * steps                                       # fp32, as the reference computes it
    plan = golden_lib.make_query_plan(ids.cuda(), times.cuda(), hf_desc=t.hf_desc(), obs_tmpl=dev(g["tmpl"]),
                                      time_offsets=offsets.cuda())
    out = plan.launch()
    ids_t = torch.broadcast_to(ids.unsqueeze(-1), (n, 7)).flatten()
    times_t = (times.unsqueeze(-1) + offsets).flatten()
    sep = golden_lib.calc_motion_frame_fk_obs(ids_t.cuda(), times_t.cuda(), hf_desc=t.hf_desc(), obs_tmpl=dev(g["tmpl"]))
    for k in FRAME_KEYS + ("body_pos", "body_rot"):
        assert torch.equal(out[k], sep[k]), k
    assert out["obs"].shape == (n, 441)
    assert torch.equal(out["obs"], sep["obs"].view(n, 7, 441)[:, 0])
    ref = O.calc_motion_frame(oracle_tables, ids_t, times_t)
    assert torch.equal(out["root_pos"].cpu(), ref[0]) and torch.equal(out["contacts"].cpu(), ref[6])
    assert_close(out["joint_rot"], ref[4], what="steps joint_rot")


def test_slerp_branches_bit_exact_decisions(golden_lib, O, oracle_tables):
    """FIXED joints (identical key frames: |cos| >= 1 -> q0) and slow joints (sin < 1e-3 -> midpoint)
    take the same branch as the reference: these outputs involve IEEE ops only, so they are exact."""
    gen = torch.Generator().manual_seed(5)
    ids = torch.randint(0, 3, (4096,), generator=gen)
    times = torch.rand(4096, generator=gen) * oracle_tables.lengths[ids]
    ref = O.calc_motion_frame(oracle_tables, ids, times)
    out = golden_lib.calc_motion_frame(ids.cuda(), times.cuda())
    i0, i1, bl = O.frame_blend(oracle_tables, ids, times)
    q0, q1 = oracle_tables.joint_rot[i0], oracle_tables.joint_rot[i1]
    c = torch.sum(q0 * q1, dim=-1).abs()
    s = torch.sqrt(1.0 - c * c)
    non_slerp = (c >= 1) | (s < 0.001)
    assert non_slerp.any()
    assert torch.equal(out[4].cpu()[non_slerp], ref[4][non_slerp])


def test_empty_and_single_query(golden_lib):
    e = golden_lib.calc_motion_frame(torch.zeros(0, dtype=torch.long, device="cuda"), torch.zeros(0, device="cuda"))
    assert e[0].shape == (0, 3) and e[4].shape == (0, 14, 4)
    one = golden_lib.calc_motion_frame(torch.tensor([1], device="cuda"), torch.tensor([0.5], device="cuda"))
    assert one[4].shape == (1, 14, 4) and torch.isfinite(one[4]).all()


def test_cpu_tensor_raises(golden_lib):
    from parc_b200._lib import ParcLibraryError
    with pytest.raises(ParcLibraryError):
        golden_lib.calc_motion_frame(torch.tensor([0]), torch.tensor([0.1]))


def test_motion_frames_loader_quirk(gpu_model):
    """init_type="motion_frames" reproduces the reference's fps-as-dt dof_vel (anim/motion_lib.py:178).  Host frames
    are built on the CPU (bit-identical tables); CUDA frames are built on the GPU (ulp-level differences, as the
    reference itself shows between devices) -- frame indices are exact either way."""
    from parc_b200.anim.motion_lib import LoopMode, MotionLib
    g = golden("tables_motion_frames_golden.npz")
    host = MotionLib(torch.tensor(g["frames"]), gpu_model, "cuda:0", init_type="motion_frames", loop_mode=LoopMode.CLAMP,
                     fps=30, contact_info=True, contacts=torch.tensor(g["contacts"]))
    assert torch.equal(host._frame_dof_vel.cpu(), torch.tensor(g["dof_vel"]))
    assert torch.equal(host._frame_root_ang_vel.cpu(), torch.tensor(g["root_ang_vel"]))
    assert torch.equal(host._frame_joint_rot.cpu(), torch.tensor(g["joint_rot"]))
    lib = MotionLib(dev(g["frames"]), gpu_model, "cuda:0", init_type="motion_frames", loop_mode=LoopMode.CLAMP, fps=30,
                    contact_info=True, contacts=dev(g["contacts"]))
    # finite differences multiply the ulp-level quaternion differences by fps (30): angular velocities get a
    # correspondingly wider absolute bar
    for k, a, atol in (("root_rot", lib._frame_root_rot, 2e-6), ("joint_rot", lib._frame_joint_rot, 2e-6),
                       ("root_vel", lib._frame_root_vel, 1e-6), ("root_ang_vel", lib._frame_root_ang_vel, 1e-4),
                       ("dof_vel", lib._frame_dof_vel, 2e-6)):
        assert a.is_cuda
        assert_close(a, g[k], atol=atol, what=f"device-built {k}")
    assert torch.equal(lib._motion_lengths.cpu(), torch.tensor(g["lengths"]))
    assert_close(lib._motion_root_pos_delta, g["root_pos_delta"], what="delta")
    ids = torch.tensor([0, 1, 1, 0], device="cuda")
    tms = torch.tensor([0.5, 0.25, 0.9, 0.0333], device="cuda")
    a, b = lib._calc_frame_blend(ids, tms), host._calc_frame_blend(ids, tms)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    out, ref = lib.calc_motion_frame(ids, tms), host.calc_motion_frame(ids, tms)
    assert out[5].shape == (4, 28)
    for x, y in zip(out, ref):
        assert_close(x, y, atol=1e-4, what="query on device-built tables")


# ----------------------------------------------------------------------------------------- dof <-> rot
def test_dof_to_rot_and_back(gpu_model):
    civ, g = golden("clip_civilization.npz"), golden("dof_golden.npz")
    dof = dev(civ["frames"][:, 6:])
    jr = gpu_model.dof_to_rot(dof)
    assert_close(jr, g["joint_rot"], what="dof_to_rot")
    assert_close(gpu_model.rot_to_dof(dev(g["joint_rot"])), g["dof_back"], atol=2e-6, what="rot_to_dof (kernel)")
    jr_g = dev(g["joint_rot"]).requires_grad_(True)
    d_t = gpu_model.rot_to_dof(jr_g)                                        # autograd path (torch ops)
    assert d_t.requires_grad
    assert_close(d_t, g["dof_back"], atol=2e-6, what="rot_to_dof (torch path)")
    assert gpu_model.rot_to_dof(dev(g["joint_rot"]).view(2, 127, 14, 4)).shape == (2, 127, 28)
    ident = torch.zeros(3, 14, 4, device="cuda"); ident[..., 3] = 1.0          # identity -> zero DoFs, no NaN
    assert (gpu_model.rot_to_dof(ident) == 0).all()
    from parc_b200 import ops
    assert_close(ops.exp_map_to_quat(dev(civ["frames"][:, 3:6])), g["root_quat"], what="exp_map_to_quat")
    z = gpu_model.dof_to_rot(torch.zeros(2, 28, device="cuda"))       # exact zeros: identity, no NaN
    assert torch.equal(z[..., 3], torch.ones(2, 14, device="cuda")) and (z[..., :3] == 0).all()


def test_fk_and_dof_gradients_vs_oracle(gpu_model, O, oracle_model):
    civ = golden("clip_civilization.npz")
    fr = torch.tensor(civ["frames"][40:72])
    gen = torch.Generator().manual_seed(3)
    wp = torch.randn(32, 15, 3, generator=gen)
    wr = torch.randn(32, 15, 4, generator=gen)

    def run(fk, d2r, e2q, to):
        a, b, c = (fr[:, 0:3].clone().to(to).requires_grad_(True), fr[:, 3:6].clone().to(to).requires_grad_(True),
                   fr[:, 6:].clone().to(to).requires_grad_(True))
        bp, br = fk(a, e2q(b), d2r(c))
        loss = (bp * wp.to(to)).sum() + (br * wr.to(to)).sum()
        loss.backward()
        return loss.detach(), a.grad, b.grad, c.grad

    from parc_b200 import ops
    ref = run(lambda p, q, j: O.forward_kinematics(oracle_model, p, q, j), lambda d: O.dof_to_rot(oracle_model, d),
              O.exp_map_to_quat, "cpu")
    got = run(gpu_model.forward_kinematics, gpu_model.dof_to_rot, ops.exp_map_to_quat, "cuda:0")
    assert_close(got[0], ref[0], rtol=1e-5, what="fk loss")
    for nm, a, b in zip(("root_pos", "root_exp", "joint_dof"), got[1:], ref[1:]):
        assert_close_normwise(a, b, what=f"fk grad {nm}")


# ----------------------------------------------------------------------------------------- heightfield
def _civ_terrain():
    from parc_b200.util.terrain_util import SubTerrain
    civ = golden("clip_civilization.npz")
    t = SubTerrain("civ", x_dim=50, y_dim=50, dx=0.4, dy=0.4, min_x=0.0, min_y=0.0, device="cuda:0")
    t.hf = dev(civ["hf"])
    t.min_point = dev(civ["min_point"])
    t.dxdy = dev(civ["dxdy"])
    return t


def test_grid_index_bit_exact():
    g = golden("obs_golden.npz")
    t = _civ_terrain()
    assert torch.equal(t.get_grid_index(dev(g["ray_xy"])).cpu(), torch.tensor(g["ray_grid_index"]))
    assert torch.equal(t.get_grid_index(dev(g["probe_xy"])).cpu(), torch.tensor(g["probe_index"]))
    from parc_b200.util.terrain_util import get_local_hf_from_terrain
    z = get_local_hf_from_terrain(dev(g["ray_xy"]), t).cpu()
    hf = torch.tensor(golden("clip_civilization.npz")["hf"])
    gi = torch.tensor(g["ray_grid_index"])
    assert torch.equal(z, hf[gi[:, 0], gi[:, 1]])
    assert torch.equal(t.get_hf_val_from_points(dev(g["ray_xy"]).view(64, 441, 2)).cpu(), z.view(64, 441))


@pytest.mark.parametrize("mn,d,dim", [(0.0, 0.4, 1536), (0.0, 0.4, 50), (-3.2, 0.4, 16), (0.0, 0.05, 4000),
                                       (1.7, 0.2, 31), (0.0, 1.0, 256), (-100.5, 0.3, 977), (12.0, 0.7, 1)])
def test_fast_grid_index_equals_ieee_division_for_every_float(mn, d, dim):
    """The observation loops divide by the (loop-invariant) cell size with a hoisted reciprocal + 3 FMAs;
    the resulting cell index must equal the reference form (true IEEE division) for ALL 2^32 inputs."""
    from parc_b200 import ops
    assert ops.selftest_grid_index(mn, d, dim) == 0


def _border_mask(coord, ulps=8):
    """True where the pre-rounding grid coordinate is within a few fp32 ulps of a half-integer."""
    c = torch.as_tensor(coord).double()
    frac = (c - torch.floor(c) - 0.5).abs()
    tol = ulps * 1.2e-7 * c.abs().clamp(min=1.0)
    return (frac <= tol).any(dim=-1)


def test_ray_obs_vs_golden():
    from parc_b200.envs.ig_parkour.mgdm_dm_util import refresh_ray_obs_hfs
    g = golden("obs_golden.npz")
    t = _civ_terrain()
    obs = refresh_ray_obs_hfs(dev(g["root_pos"]), dev(g["heading"]), dev(g["tmpl"]), t, -3.0, 3.0).cpu()
    exp = torch.tensor(g["ray_obs"])
    mism = (obs != exp)
    border = _border_mask(g["ray_grid_coord"]).view(64, 441)
    assert not (mism & ~border).any(), f"{int((mism & ~border).sum())} mismatches away from cell borders"
    assert mism.float().mean() < 1e-3


def test_grid_obs_vs_golden(O):
    from parc_b200.util.terrain_util import sample_hf_z_on_terrain
    g = golden("obs_golden.npz")
    t = _civ_terrain()
    z = sample_hf_z_on_terrain(t, dev(g["root_pos"][:16, 0:2]), dev(g["heading"][:16]), 0.2, 0.2, 15, 15, 15, 15).cpu()
    exp = torch.tensor(g["grid_obs"])
    assert z.shape == (16, 31, 31)
    # every differing sample must sit within 8 ulp of a cell border (GPU vs host sin / cos of the heading)
    ot = O.Terrain(hf=t.hf.cpu(), min_point=t.min_point.cpu(), dxdy=t.dxdy.cpu())
    tmpl = O.grid_template(0.2, 0.2, 15, 15, 15, 15)
    pts = O.rotate_2d(tmpl, torch.tensor(g["heading"][:16]).view(16, 1, 1)) + torch.tensor(g["root_pos"][:16, 0:2]).view(16, 1, 1, 2)
    assert torch.equal(O.hf_sample(ot, pts), exp), "oracle restatement of the golden grid observation"
    border = _border_mask(O.grid_coord(ot, pts)).view(16, 31, 31)
    mism = z != exp
    assert not (mism & ~border).any(), f"{int((mism & ~border).sum())} mismatches away from cell borders"
    assert mism.float().mean() < 2e-3


@pytest.mark.parametrize("fast_heading", [False, True])
def test_fused_obs_vs_oracle(golden_lib, O, oracle_tables, fast_heading):
    """The fused observation against the oracle, two ways.  (1) Against the oracle evaluated on the kernel's OWN
    root position / rotation outputs (which are checked to 1e-5 elsewhere): the only differences left are the last
    bits of atan2 / sin / cos, so every mismatching sample must lie within 8 ulp of a cell border -- the bar of the
    stand-alone observation kernel.  (2) Against the all-oracle chain (CPU slerp -> heading -> samples), where the
    heading additionally inherits the few-ulp difference of the slerped rotation: 64-ulp mask."""
    g = golden("obs_golden.npz")
    t = _civ_terrain()
    gen = torch.Generator().manual_seed(11)
    n = 2048
    ids = torch.zeros(n, dtype=torch.long)
    times = torch.rand(n, generator=gen) * oracle_tables.lengths[0]
    r = golden_lib.calc_motion_frame_fk_obs(ids.cuda(), times.cuda(), hf_desc=t.hf_desc(), obs_tmpl=dev(g["tmpl"]),
                                            fast_heading=fast_heading)
    ot = O.Terrain(hf=t.hf.cpu(), min_point=t.min_point.cpu(), dxdy=t.dxdy.cpu())
    tmpl = torch.tensor(g["tmpl"])
    obs = r["obs"].cpu()
    # (1) oracle on the kernel's own frame outputs
    rp, rr = r["root_pos"].cpu(), r["root_rot"].cpu()
    heading = O.calc_heading(rr)
    exp = O.ray_obs(ot, rp, heading, tmpl)
    coord = O.grid_coord(ot, O.ray_obs_points(rp, heading, tmpl))
    mism = obs != exp
    border = _border_mask(coord, ulps=8 if not fast_heading else 16).view(n, 441)
    assert not (mism & ~border).any(), f"{int((mism & ~border).sum())} mismatches away from cell borders (own frame)"
    assert mism.float().mean() < 5e-4
    # (2) all-oracle chain
    ref = O.calc_motion_frame(oracle_tables, ids, times)
    heading = O.calc_heading(ref[1])
    exp = O.ray_obs(ot, ref[0], heading, tmpl)
    coord = O.grid_coord(ot, O.ray_obs_points(ref[0], heading, tmpl))
    mism = obs != exp
    border = _border_mask(coord, ulps=64).view(n, 441)
    assert not (mism & ~border).any(), f"{int((mism & ~border).sum())} mismatches away from cell borders"
    assert mism.float().mean() < 1e-3
    assert_close(r["body_pos"], O.forward_kinematics(O.CharModel.from_npz(__import__("os").path.join(
        __import__("conftest").GOLDEN, "humanoid_model.npz")), ref[0], ref[1], ref[4])[0], what="fused body_pos")


def test_out_of_range_ids_are_flagged_not_read(golden_lib):
    """The reference raises IndexError (CPU) / a device assert (CUDA) on a clip id or frame outside the tables.
    Here the kernel answers the entry with clip 0 / the nearest valid frame, never reads out of bounds, and sets an
    error bit the mirror turns into IndexError."""
    M = golden_lib.num_motions()
    ids = torch.tensor([0, 1, M, -1, 2], device="cuda:0")
    times = torch.zeros(5, device="cuda:0")
    r = golden_lib.calc_motion_frame(ids, times)           # lazily checked: no exception yet
    with pytest.raises(IndexError, match="motion id out of range"):
        golden_lib.check_query_errors()
    golden_lib.check_query_errors()                        # the word is cleared by the raise
    ok = golden_lib.calc_motion_frame(torch.tensor([0, 1, 0, 0, 2], device="cuda:0"), times)
    for a, b in zip(r, ok):
        assert torch.equal(a, b)                           # bad ids answered as clip 0
    nf = int(golden_lib._motion_num_frames[1].item())
    fr = golden_lib.get_motion_frame(torch.tensor([1, 1, 1], device="cuda:0"),
                                     torch.tensor([nf, -3, nf - 1], device="cuda:0"))
    with pytest.raises(IndexError, match="frame index out of range"):
        golden_lib.check_query_errors()
    ok = golden_lib.get_motion_frame(torch.tensor([1, 1, 1], device="cuda:0"),
                                     torch.tensor([nf - 1, 0, nf - 1], device="cuda:0"))
    for a, b in zip(fr, ok):
        assert torch.equal(a, b)
    golden_lib.check_query_errors()
    golden_lib.validate_ids = True
    try:
        with pytest.raises(IndexError):
            golden_lib.calc_motion_frame(ids, times)
        golden_lib.calc_motion_frame(ids.clamp(0, M - 1), times)
    finally:
        golden_lib.validate_ids = False


def test_empty_observation_template_is_a_no_op(golden_lib):
    """P == 0: the sweep must not read an unstaged template; the frame / FK outputs are unaffected."""
    t = _civ_terrain()
    ids = torch.zeros(33, dtype=torch.long, device="cuda:0")
    times = torch.linspace(0, 3, 33, device="cuda:0")
    a = golden_lib.calc_motion_frame_fk_obs(ids, times, hf_desc=t.hf_desc(), obs_tmpl=torch.zeros(0, 2, device="cuda:0"))
    b = golden_lib.calc_motion_frame_fk_obs(ids, times)
    assert a["obs"].shape == (33, 0)
    for k in ("root_pos", "joint_rot", "body_pos", "body_rot"):
        assert torch.equal(a[k], b[k])


def test_query_variants_pdl_and_output_selection_are_bit_identical(golden_lib):
    """Every instantiation (sweep depth / characters per warp), the programmatic-dependent-launch forms and a plan
    restricted to some outputs must produce the same bits as the default launch."""
    g = golden("obs_golden.npz")
    t = _civ_terrain()
    gen = torch.Generator().manual_seed(5)
    n = 3001
    ids = torch.randint(0, golden_lib.num_motions(), (n,), generator=gen).cuda()
    times = (torch.rand(n, generator=gen) * 9.0 - 0.5).cuda()
    tmpl = dev(g["tmpl"])
    base = golden_lib.make_query_plan(ids, times, hf_desc=t.hf_desc(), obs_tmpl=tmpl).launch()
    torch.cuda.synchronize()
    base = {k: v.clone() for k, v in base.items()}
    for variant in (1, 2, 3, 4, 5, 6):
        got = golden_lib.make_query_plan(ids, times, hf_desc=t.hf_desc(), obs_tmpl=tmpl, variant=variant).launch()
        for k, v in base.items():
            assert torch.equal(got[k], v), f"variant {variant}: {k}"
    for early, variant in ((False, 0), (True, 0), (True, 5), (True, 6), (True, 2)):
        pl = golden_lib.make_query_plan(ids, times, hf_desc=t.hf_desc(), obs_tmpl=tmpl, pdl=True, pdl_early_inputs=early,
                                        variant=variant)
        for _ in range(3):                         # back to back: each launch overlaps the previous one's tail
            got = pl.launch()
        for k, v in base.items():
            assert torch.equal(got[k], v), f"pdl early={early} variant={variant}: {k}"
    from parc_b200 import ops
    plans = [golden_lib.make_query_plan(ids, times, hf_desc=t.hf_desc(), obs_tmpl=tmpl, pdl=True, pdl_early_inputs=True,
                                        out={}) for _ in range(4)]
    graph = ops.capture_launches(plans)
    for pl in plans:
        for v in pl.out.values():
            v.zero_()
    graph.replay()
    for pl in plans:
        for k, v in base.items():
            assert torch.equal(pl.out[k], v), f"graph of pdl launches: {k}"
    sel = golden_lib.make_query_plan(ids, times, hf_desc=t.hf_desc(), obs_tmpl=tmpl, outputs=("body_pos", "obs"), out={})
    got = sel.launch()
    assert set(got) == {"body_pos", "obs"}
    assert torch.equal(got["body_pos"], base["body_pos"]) and torch.equal(got["obs"], base["obs"])


def test_mdm_sampler_terrain_gather_vs_golden():
    """diffusion/mdm_heightfield_contact_motion_sampler.py:414-474 (get_hfs_from_data, augmentation off) in one
    launch over packed per-clip terrains: heights, centre heights and the augmenter's (max, min) bands under the
    frame-window body masks, both relative-z styles, against the reference's outputs.  Samples within 8 ulp of a cell
    border may land in the neighbouring cell (GPU vs host sin / cos of the heading); everything else is exact."""
    from parc_b200.diffusion.mdm_heightfield_contact_motion_sampler import (ClipHeightfieldSampler, ClipTerrainPack,
                                                                             RelativeZStyle, mask_inds_to_bits)
    from parc_b200.util.terrain_util import SubTerrain
    g = golden("sampler_golden.npz")
    terrains, bits = [], []
    for c in range(3):
        hf = g[f"hf{c}"]
        t = SubTerrain(f"c{c}", x_dim=hf.shape[0], y_dim=hf.shape[1], dx=float(g[f"dxdy{c}"][0]), dy=float(g[f"dxdy{c}"][1]),
                       min_x=float(g[f"min_point{c}"][0]), min_y=float(g[f"min_point{c}"][1]), device="cuda:0")
        t.hf, t.hf_maxmin = dev(hf), dev(g[f"maxmin{c}"])
        terrains.append(t)
        flat, cnt = torch.tensor(g[f"mask_inds{c}"]), g[f"mask_count{c}"].tolist()
        per, s0 = [], 0
        for n in cnt:
            per.append(flat[s0:s0 + n])
            s0 += n
        W = (hf.shape[0] * hf.shape[1] + 31) // 32
        bits.append(torch.from_numpy(mask_inds_to_bits(per, hf.shape[1], W).view(np.int32)))
    pack = ClipTerrainPack(terrains, bits, "cuda:0")
    coord = torch.as_tensor(g["grid_coord"]).double()
    border = ((coord - torch.floor(coord) - 0.5).abs() <= 8 * 1.2e-7 * coord.abs().clamp(min=1.0)).any(dim=-1)
    n = int(g["num_neg"])
    for style, tag in ((RelativeZStyle.RELATIVE_TO_ROOT_FLOOR, "relative_to_root_floor"),
                       (RelativeZStyle.RELATIVE_TO_ROOT, "relative_to_root")):
        smp = ClipHeightfieldSampler(pack, float(g["dx"]), n, n, n, n, float(g["max_h"]), style)
        assert torch.equal(smp._generic_heightmap.cpu(), torch.tensor(g["grid"]))
        hfs, ch, mm = smp.get_hfs_from_data(dev(g["ids"]), dev(g["root_pos"]), dev(g["root_rot"]), dev(g["canon_z"]),
                                            dev(g["mti"]))
        e_hfs, e_ch, e_mm = torch.tensor(g[f"hfs_{tag}"]), torch.tensor(g[f"center_h_{tag}"]), torch.tensor(g[f"mm_{tag}"])
        centre_ok = ch.cpu() == e_ch
        assert (~centre_ok).sum() <= 1
        if style == RelativeZStyle.RELATIVE_TO_ROOT_FLOOR:          # a flipped centre cell shifts its whole sample
            border = border | (~centre_ok).view(-1, 1, 1)
        bad_h = (hfs.cpu() != e_hfs) & ~border
        bad_m = (mm.cpu() != e_mm).any(dim=-1) & ~border
        assert not bad_h.any() and not bad_m.any(), f"{int(bad_h.sum())} heights / {int(bad_m.sum())} bands differ off-border"
        assert (hfs.cpu() != e_hfs).float().mean() < 2e-3
        assert (e_mm[..., 0] + (e_ch if style == RelativeZStyle.RELATIVE_TO_ROOT_FLOOR else torch.tensor(g["canon_z"])).view(-1, 1, 1)
                < 2.0 * float(g["max_h"]) - 1e-3).float().mean() > 0.02           # the masks really select bands
    two = smp.get_hfs_from_data(dev(g["ids"]), dev(g["root_pos"]), dev(g["root_rot"]), dev(g["canon_z"]), dev(g["mti"]),
                                want_maxmin=False)
    assert len(two) == 2 and torch.equal(two[0], hfs)
    with pytest.raises(IndexError):
        smp.get_hfs_from_data(torch.tensor([3], device="cuda:0"), dev(g["root_pos"][:1]), dev(g["root_rot"][:1]),
                              dev(g["canon_z"][:1]), dev(g["mti"][:1]))


# ----------------------------------------------------------------------------------------- SDF + losses
def test_points_hf_sdf_vs_golden():
    from parc_b200.util.terrain_util import points_hf_sdf
    g = golden("sdf_golden.npz")
    dxdy = torch.tensor([0.4, 0.4], device="cuda")
    for inv, key in ((True, "probe_inv"), (False, "probe_sol")):
        sd = points_hf_sdf(dev(g["probe_points"]), dev(g["probe_hf"]), torch.zeros(1, 2, device="cuda"), dxdy,
                           base_z=-10.0, inverted=inv)
        assert_close(sd, g[key], what=key)
    t = _civ_terrain()
    for inv, key in ((True, "clip_inv"), (False, "clip_sol")):
        sd = points_hf_sdf(dev(g["clip_points"]), t.hf.unsqueeze(0), t.min_point.unsqueeze(0), t.dxdy, base_z=-10.0,
                           inverted=inv)
        assert_close(sd, g[key], what=key)


def _a16_points(g):
    pts, s0 = [], 0
    for n in g["minimal_point_count"].tolist():
        pts.append(dev(g["minimal_points"][s0:s0 + n]))
        s0 += n
    return pts


def test_points_hf_sdf_gradient_vs_golden():
    """Row a16: the MDM back-propagates 0.5 * sum(clamp(sdf, max=0)^2) through points_hf_sdf (diffusion/mdm.py:729-737,
    :1484-1496); golden = the reference's own autograd gradients on its 31 x 31 @ 0.2 m local grid."""
    from parc_b200.util.terrain_util import points_hf_sdf
    g = golden("a16_golden.npz")
    hf, mc, dxdy = dev(g["hf"]), dev(g["min_center"]), dev(g["dxdy"])
    for tag in ("guidance", "train"):
        p = dev(g["world_points"]).clone().requires_grad_(True)
        sdf = points_hf_sdf(p, hf, mc, dxdy, base_z=float(g[f"base_z_{tag}"]))
        loss = 0.5 * torch.sum(torch.square(torch.clamp(sdf, max=0.0)))
        loss.backward()
        assert_close(sdf, g[f"psdf_{tag}"], rtol=1e-5, atol=2e-6, what=f"sdf {tag}")
        assert_close_normwise(p.grad, g[f"pgrad_{tag}"], 1e-5, what=f"d loss / d points {tag}")
        assert (torch.tensor(g[f"pgrad_{tag}"]).abs().sum(-1) > 0).sum() > 20       # the fixture really penetrates
    p = dev(g["world_points"]).clone().requires_grad_(True)
    sdf = points_hf_sdf(p, hf, mc, dxdy, base_z=-10.0, inverted=False)
    (sdf * dev(g["upstream_solid"])).sum().backward()
    assert_close(sdf, g["psdf_solid"], rtol=1e-5, atol=2e-6, what="solid sdf")
    assert_close_normwise(p.grad, g["pgrad_solid"], 1e-5, what="solid d/d points")
    # no graph when nothing requires grad; want_arg still served
    from parc_b200 import ops
    tb = ops.make_terrain_batch(hf, mc, (0.2, 0.2), base_z=-10.0)
    v, a = ops.points_hf_sdf(dev(g["world_points"]), tb, True, want_arg=True)
    assert not v.requires_grad and a.dtype == torch.int32 and int(a.max()) < 31 * 31


def test_motion_frames_hf_sdf_loss_vs_golden(gpu_model):
    """util/terrain_util.py:1895-1949 with get_minimal_char_point_samples: value, per-point sdf, world points and
    the gradient with respect to the motion frames, against the reference's own outputs."""
    from parc_b200.util.terrain_util import motion_frames_hf_sdf_loss
    g = golden("a16_golden.npz")
    hf, mc, dxdy = dev(g["hf"]), dev(g["min_center"]), dev(g["dxdy"])
    pts = _a16_points(g)
    for interior, tag in ((True, "int"), (False, "ext")):
        mf = dev(g["frames"]).clone().requires_grad_(True)
        loss, wp, sdf = motion_frames_hf_sdf_loss(mf, pts, hf, mc, dxdy, gpu_model, ret_vis_info=True,
                                                  interior_distance=interior)
        loss.sum().backward()
        assert_close(loss, g[f"loss_{tag}"], rtol=1e-5, atol=1e-7, what=f"loss {tag}")
        assert_close(sdf, g[f"sdf_{tag}"], rtol=1e-5, atol=2e-6, what=f"sdf {tag}")
        if interior:
            assert_close(wp, g["world_points"], rtol=1e-5, atol=1e-6, what="world points")
        assert_close_normwise(mf.grad, g[f"grad_frames_{tag}"], 1e-5, what=f"d loss / d motion_frames {tag}")
    only = motion_frames_hf_sdf_loss(dev(g["frames"]), pts, hf, mc, dxdy, gpu_model)
    assert_close(only, g["loss_int"], rtol=1e-5, atol=1e-7, what="loss (no vis info)")


def test_body_points_world_layout_and_gradient_vs_oracle(gpu_model, O, oracle_model):
    from parc_b200 import ops
    from parc_b200.tools.procgen.mdm_path import body_points_desc
    from parc_b200.util import geom_util
    gen = torch.Generator().manual_seed(4)
    B, F = 3, 5
    bp = torch.randn(B, F, 15, 3, generator=gen)
    br = torch.nn.functional.normalize(torch.randn(B, F, 15, 4, generator=gen), dim=-1)
    up = torch.randn(B, F * 304, 3, generator=gen)
    a, b = bp.clone().requires_grad_(True), br.clone().requires_grad_(True)
    want = O.world_body_points(a, b, oracle_model.body_points)
    (want * up).sum().backward()
    pts = body_points_desc(gpu_model, geom_util.get_char_point_samples(gpu_model))
    c, d = bp.cuda().requires_grad_(True), br.cuda().requires_grad_(True)
    got = ops.body_points_world(c, d, pts)
    (got * up.cuda()).sum().backward()
    assert_close(got, want, rtol=1e-5, atol=1e-6, what="world body points")
    assert_close_normwise(c.grad, a.grad, 1e-5, what="d/d body_pos")
    assert_close_normwise(d.grad, b.grad, 1e-5, what="d/d body_rot")


def test_sdf_on_terrains_beyond_shared_memory_and_huge_batches(O):
    """No size limits: a 300 x 260 tile (312 KB > one SM's shared memory) is scanned from global memory
    with the same exact result as the oracle's brute-force min over all cells -- values, arg-min and gradient; a batch
    beyond the grid's 65 535-sample y limit is chunked."""
    from parc_b200 import ops
    gen = torch.Generator().manual_seed(8)
    X, Y = 300, 260
    hf = torch.zeros(1, X, Y)
    for _ in range(400):
        x0, y0 = int(torch.randint(0, X - 12, (1,), generator=gen)), int(torch.randint(0, Y - 12, (1,), generator=gen))
        hf[0, x0:x0 + int(torch.randint(2, 12, (1,), generator=gen)), y0:y0 + int(torch.randint(2, 12, (1,), generator=gen))] = \
            float(torch.rand(1, generator=gen) * 2.0 - 0.7)
    mc = torch.tensor([[-3.0, 1.5]])
    dxdy = torch.tensor([0.4, 0.4])
    n = 600
    p = torch.rand(1, n, 3, generator=gen) * torch.tensor([X * 0.4 + 4.0, Y * 0.4 + 4.0, 2.4]) + torch.tensor([-5.0, -0.5, -0.9])
    tb = ops.make_terrain_batch(hf.cuda(), mc.cuda(), (0.4, 0.4), base_z=-10.0)
    for inverted in (True, False):
        want = O.points_hf_sdf(p, hf, mc, dxdy, base_z=-10.0, inverted=inverted, chunk=64)
        got, arg = ops.points_hf_sdf(p.cuda(), tb, inverted, want_arg=True)
        assert_close(got, want, rtol=1e-5, atol=2e-6, what=f"large-terrain sdf inverted={inverted}")
        pg = p.clone().requires_grad_(True)
        O.points_hf_sdf(pg, hf, mc, dxdy, base_z=-10.0, inverted=inverted, chunk=64).sum().backward()
        pc = p.cuda().requires_grad_(True)
        ops.points_hf_sdf(pc, tb, inverted).sum().backward()
        assert_close_normwise(pc.grad, pg.grad, 1e-5, what=f"large-terrain gradient inverted={inverted}")
    # batch > 65 535: chunked launches
    Bn = 70001
    small = torch.rand(Bn, 4, 4, generator=gen).cuda()
    pts = (torch.rand(Bn, 3, 3, generator=gen) * 1.6 - 0.2).cuda()
    tbs = ops.make_terrain_batch(small, torch.zeros(Bn, 2).cuda(), (0.4, 0.4), base_z=-10.0)
    full = ops.points_hf_sdf(pts, tbs, True)
    for lo in (0, 65535, 70000):
        one = ops.points_hf_sdf(pts[lo:lo + 1].contiguous(),
                                ops.make_terrain_batch(small[lo:lo + 1].contiguous(), torch.zeros(1, 2).cuda(), (0.4, 0.4),
                                                       base_z=-10.0), True)
        assert torch.equal(full[lo:lo + 1], one)


def test_pruned_sdf_scan_equals_brute_force_on_adversarial_terrains(O):
    """The scan visits blocks / cells only where a bound allows (parc_sdf.cuh::scan_cells): per-mode reach, per-block
    height ranges, bit-exact monotone vertical bounds.  Against the oracle's brute-force min over ALL cells on terrains
    built to stress that: dimensions that are not multiples of the 4-cell block, tall one-cell spikes and pits
    (block max / min far from the neighbours), plateaus, points on cell
    borders and corners, points inside the ground, far outside the tile, and a base plane close to the surface (where
    the rounding of (h +- base) / 2 is smallest) as well as the usual -10 m.  Values to 1e-5; the arg-min cell must be the
    oracle's wherever the runner-up is further than 1e-5 away, and must attain the minimum everywhere."""
    from parc_b200 import ops
    gen = torch.Generator().manual_seed(23)
    for (X, Y), base_z in (((16, 16), -10.0), ((31, 31), -10.0), ((37, 50), -10.0), ((13, 9), -3.0), ((50, 50), -1000.0)):
        hf = torch.zeros(1, X, Y)
        for _ in range(12):                               # plateaus
            x0, y0 = int(torch.randint(0, X - 3, (1,), generator=gen)), int(torch.randint(0, Y - 3, (1,), generator=gen))
            hf[0, x0:x0 + int(torch.randint(2, 7, (1,), generator=gen)), y0:y0 + int(torch.randint(2, 7, (1,), generator=gen))] = \
                float(torch.randint(-4, 8, (1,), generator=gen)) * 0.25
        for _ in range(10):                               # one-cell spikes and pits
            hf[0, int(torch.randint(0, X, (1,), generator=gen)), int(torch.randint(0, Y, (1,), generator=gen))] = \
                float(torch.rand(1, generator=gen) * 4.0 - 2.0)
        mc = torch.tensor([[0.7, -1.1]])
        dxdy = torch.tensor([0.4, 0.4])
        n = 1500
        ext = torch.tensor([X * 0.4, Y * 0.4, 3.5])
        p = torch.rand(1, n, 3, generator=gen) * (ext + torch.tensor([1.6, 1.6, 0.0])) + torch.tensor([-0.1, -1.9, -1.5])
        # a third of the points snapped onto cell borders / corners in xy (exact ties between neighbouring columns)
        k = n // 3
        p[0, :k, 0] = torch.round((p[0, :k, 0] - 0.7) / 0.2) * 0.2 + 0.7
        p[0, :k // 2, 1] = torch.round((p[0, :k // 2, 1] + 1.1) / 0.2) * 0.2 - 1.1
        p[0, -8:, :2] += torch.tensor([40.0, -35.0])      # far outside the tile
        tb = ops.make_terrain_batch(hf.cuda(), mc.cuda(), (0.4, 0.4), base_z=base_z)
        centres, half = O.hf_cell_boxes(hf, mc, dxdy, base_z, False)
        centres_i, half_i = O.hf_cell_boxes(hf, mc, dxdy, base_z, True)
        for inverted, (cc, hh) in ((False, (centres, half)), (True, (centres_i, half_i))):
            rel = p.unsqueeze(2) - cc.unsqueeze(1)
            sd_all = O.sd_box(rel, hh.unsqueeze(1).expand_as(rel))[0]             # [n, M]
            want, want_arg = torch.min(sd_all, dim=-1)
            got, arg = ops.points_hf_sdf(p.cuda(), tb, inverted, want_arg=True)
            got, arg = got[0].cpu(), arg[0].cpu().long()
            sign = -1.0 if inverted else 1.0
            tol = 1e-5 * want.abs() + 2e-6
            assert ((got * sign - want).abs() <= tol).all(), f"{X}x{Y} base {base_z} inverted={inverted}: value"
            # the reported cell attains the minimum ...
            at_arg = sd_all.gather(1, arg.view(-1, 1))[:, 0]
            assert ((at_arg - want).abs() <= tol).all(), f"{X}x{Y} base {base_z} inverted={inverted}: arg does not attain the min"
            # ... and is the oracle's own first-index choice wherever the minimum is clearly separated
            second = sd_all.scatter(1, want_arg.view(-1, 1), float("inf")).min(dim=-1)[0]
            clear = (second - want) > 1e-5
            assert clear.float().mean() > 0.3
            assert torch.equal(arg[clear], want_arg[clear]), f"{X}x{Y} base {base_z} inverted={inverted}: arg-min cell"
            # (exact ties -- plateaus, border points -- are covered by the golden gradient tests: which of two equal
            # cells wins decides the sign of a gradient component there)


def test_compute_motion_loss_vs_golden(gpu_model):
    from parc_b200.tools.procgen.mdm_path import compute_motion_loss
    from parc_b200.util import geom_util
    from parc_b200.util.motion_util import MotionFrames
    g = golden("loss_golden.npz")
    mf = MotionFrames(root_pos=dev(g["ml_root_pos"]), root_rot=dev(g["ml_root_rot"]), joint_rot=dev(g["ml_joint_rot"]),
                      contacts=dev(g["ml_contacts"]))
    pts = geom_util.get_char_point_samples(gpu_model)
    out = compute_motion_loss(mf, None, _civ_terrain(), gpu_model, pts, w_contact=0.1, w_pen=0.1, w_path=0.0)
    assert_close(out["pen_loss"], g["ml_pen"], what="pen_loss")
    assert_close(out["contact_loss"], g["ml_contact"], what="contact_loss")
    assert_close(out["total_loss"], g["ml_total"], what="total_loss")


def test_motion_opt_loss_and_gradients_vs_golden(gpu_model):
    from parc_b200.tools.motion_opt.motion_optimization import LossType, motion_terrain_contact_loss
    from parc_b200.util import geom_util
    g = golden("loss_golden.npz")
    pts = geom_util.get_char_point_samples(gpu_model)
    t = _civ_terrain()

    def run(others):
        a, b, c = (dev(g[k]).clone().requires_grad_(True) for k in ("mo_root_pos", "mo_root_exp", "mo_joint_dof"))
        loss, ld = motion_terrain_contact_loss(
            a, b, c, dev(g["mo_src_root_pos"]), dev(g["mo_src_root_quat"]), dev(g["mo_src_joint_rot"]),
            dev(g["mo_src_body_vels"]), dev(g["mo_src_body_rot_vels"]), dev(g["mo_contacts"]), t, pts, gpu_model,
            w_root_pos=others, w_root_rot=others, w_joint_rot=others, w_smoothness=others, w_penetration=1000.0,
            w_contact=1000.0, w_sliding=others, w_body_constraints=0.0, w_jerk=others, body_constraints=None,
            max_jerk=1000.0)
        loss.backward()
        return loss.detach(), ld, a.grad, b.grad, c.grad

    loss, ld, ga, gb, gc = run(0.0)
    assert_close(loss, g["mo_loss_pc"], what="loss (pen+contact)")
    assert_close(torch.tensor(ld[LossType.PENETRATION_LOSS]), g["mo_pen"], what="pen term")
    assert_close(torch.tensor(float(ld[LossType.CONTACT_LOSS])), g["mo_contact"], what="contact term")
    assert_close_normwise(ga, g["mo_grad_root_pos"], what="grad root_pos")
    assert_close_normwise(gb, g["mo_grad_root_exp"], what="grad root_rot")
    assert_close_normwise(gc, g["mo_grad_joint_dof"], what="grad joint_dof")
    # all terms on (tracking, smoothness, sliding, jerk): same bar
    loss, ld, ga, gb, gc = run(1.0)
    assert_close(loss, g["mo_loss_all"], what="loss (all terms)")
    assert_close_normwise(ga, g["mo_all_grad_root_pos"], what="grad root_pos (all)")
    assert_close_normwise(gb, g["mo_all_grad_root_exp"], what="grad root_rot (all)")
    assert_close_normwise(gc, g["mo_all_grad_joint_dof"], what="grad joint_dof (all)")


def test_body_loss_vs_oracle_synthetic(gpu_model, O, oracle_model):
    """Config-3-shaped inputs at a size the oracle finishes in seconds: B=3 samples x F=6 frames on
    three different 16x16 box / stair terrains (per-sample terrain), fwd + bwd."""
    from parc_b200 import ops
    from parc_b200.tools.procgen.mdm_path import body_points_desc
    from parc_b200.util import geom_util, synth
    rng = np.random.default_rng(42)
    B, F = 3, 6
    hfs = np.stack([synth.box_terrain(rng), synth.stairs_terrain(rng), synth.box_terrain(rng, h_range=(-0.5, 0.6))])
    smp = [synth.synth_motion_samples(gpu_model, 1, F, hfs[i], (0.0, 0.0), (0.4, 0.4), seed=100 + i) for i in range(B)]
    cat = lambda k: torch.tensor(np.concatenate([s[k] for s in smp], axis=0))
    root_pos, root_exp, joint_dof, contacts = cat("root_pos"), cat("root_exp"), cat("joint_dof"), cat("contacts")
    dxdy = torch.tensor([0.4, 0.4])
    mins = torch.zeros(B, 2)

    # oracle: per-sample terrain -> loop over samples
    leaves = [t.clone().requires_grad_(True) for t in (root_pos, root_exp, joint_dof)]
    tot = 0.0
    pens, cons = [], []
    for i in range(B):
        rq = O.exp_map_to_quat(leaves[1][i])
        jr = O.dof_to_rot(oracle_model, leaves[2][i])
        bp, br = O.forward_kinematics(oracle_model, leaves[0][i], rq, jr)
        pen, con = O.pen_contact_terms(bp.unsqueeze(0), br.unsqueeze(0), contacts[i].unsqueeze(0),
                                       oracle_model.body_points, torch.tensor(hfs[i]), mins[i], dxdy, -10.0)
        pens.append(pen[0]); cons.append(con[0])
        tot = tot + 0.7 * pen[0] + 1.3 * con[0]
    tot.backward()

    g_leaves = [t.clone().cuda().requires_grad_(True) for t in (root_pos, root_exp, joint_dof)]
    pts = body_points_desc(gpu_model, geom_util.get_char_point_samples(gpu_model))
    tb = ops.make_terrain_batch(torch.tensor(hfs).cuda(), mins.cuda(), (0.4, 0.4), base_z=-10.0)
    rq = ops.exp_map_to_quat(g_leaves[1])
    jr = gpu_model.dof_to_rot(g_leaves[2])
    total, pen, con = ops.body_loss(gpu_model.c_model(), pts, tb, g_leaves[0], rq, jr, contacts.cuda(), 0.7, 1.3)
    total.sum().backward()
    assert_close(pen, torch.stack(pens).detach(), what="pen")
    assert_close(con, torch.stack(cons).detach(), what="contact")
    assert float(torch.stack(pens).sum()) > 0 and float(torch.stack(cons).abs().sum()) > 0
    for nm, a, b in zip(("root_pos", "root_exp", "joint_dof"), g_leaves, leaves):
        assert_close_normwise(a.grad, b.grad, what=f"body_loss grad {nm}")


@pytest.mark.parametrize("kind", ["box", "stairs"])
def test_cfg3_shape_loss_and_gradients_vs_oracle(gpu_model, O, oracle_model, kind):
    """BASELINE configs[2] at its own shape, subsampled in the batch only: B = 8 MDM-style samples x F = 150 frames,
    each on its own 16 x 16 procedural terrain (boxes / stairs as kin_gen_default.yaml:98-121), contacts in
    [-0.05, 1], ranking weights 0.1 / 0.1 -- forward values and the gradients of all three leaves against autograd
    through the oracle (the a15 formulation batched over samples)."""
    from parc_b200 import ops
    from parc_b200.tools.procgen.mdm_path import body_points_desc
    from parc_b200.util import geom_util, synth
    rng = np.random.default_rng(31 if kind == "box" else 32)
    B, F = 8, 150
    hfs = np.stack([synth.box_terrain(rng) if kind == "box" else synth.stairs_terrain(rng) for _ in range(B)])
    smp = synth.synth_motion_samples(gpu_model, B, F, hfs[0], (0.0, 0.0), (0.4, 0.4), seed=77)
    pts = body_points_desc(gpu_model, geom_util.get_char_point_samples(gpu_model))
    tb = ops.make_terrain_batch(torch.tensor(hfs).cuda(), torch.zeros(B, 2).cuda(), (0.4, 0.4), base_z=-10.0)
    leaves = [torch.tensor(smp[k]).cuda().requires_grad_(True) for k in ("root_pos", "root_exp", "joint_dof")]
    total, pen, con = ops.body_loss(gpu_model.c_model(), pts, tb, leaves[0], ops.exp_map_to_quat(leaves[1]),
                                    gpu_model.dof_to_rot(leaves[2]), torch.tensor(smp["contacts"]).cuda(), 0.1, 0.1)
    total.sum().backward()
    pens, cons = [], []
    for i in range(B):
        cpu = [torch.tensor(smp[k][i]).requires_grad_(True) for k in ("root_pos", "root_exp", "joint_dof")]
        loss, p_i, c_i = O.motion_opt_pen_contact(oracle_model, cpu[0], cpu[1], cpu[2], torch.tensor(smp["contacts"][i]),
                                                  torch.tensor(hfs[i]), torch.zeros(2), torch.tensor([0.4, 0.4]), 0.1, 0.1)
        loss.backward()
        pens.append(p_i.detach())
        cons.append(c_i.detach())
        for a, c, nm in zip(leaves, cpu, ("root_pos", "root exp-map", "joint dofs")):
            assert_close_normwise(a.grad[i], c.grad, 1e-5, what=f"{kind} sample {i} d/d {nm}")
    assert_close(pen, torch.stack(pens), rtol=1e-5, atol=1e-6, what=f"{kind} pen")
    assert_close(con, torch.stack(cons), rtol=1e-5, atol=2e-6, what=f"{kind} contact")
    assert float(torch.stack(pens).sum()) > 0 and float(torch.stack(cons).abs().sum()) > 0


def test_body_loss_linearity_full_size(gpu_model):
    """Size-independent property at config-3 scale (B=64 here x F=150): the loss is a sum over frames, so
    evaluating two halves of the frames separately and adding must reproduce the whole; and gradients of
    frame f do not depend on other frames."""
    from parc_b200 import ops
    from parc_b200.tools.procgen.mdm_path import body_points_desc
    from parc_b200.util import geom_util, synth
    rng = np.random.default_rng(1)
    B, F = 64, 150
    hf = synth.box_terrain(rng)
    s = synth.synth_motion_samples(gpu_model, B, F, hf, (0.0, 0.0), (0.4, 0.4), seed=5)
    rp = torch.tensor(s["root_pos"]).cuda()
    rq = ops.exp_map_to_quat(torch.tensor(s["root_exp"]).cuda())
    jr = gpu_model.dof_to_rot(torch.tensor(s["joint_dof"]).cuda())
    ct = torch.tensor(s["contacts"]).cuda()
    pts = body_points_desc(gpu_model, geom_util.get_char_point_samples(gpu_model))
    tb = ops.make_terrain_batch(torch.tensor(hf).cuda(), torch.zeros(2).cuda(), (0.4, 0.4), base_z=-10.0)
    m = gpu_model.c_model()
    _, pen, con = ops.body_loss(m, pts, tb, rp, rq, jr, ct, 1.0, 1.0)
    h = F // 2
    _, pen_a, con_a = ops.body_loss(m, pts, tb, rp[:, :h].contiguous(), rq[:, :h].contiguous(), jr[:, :h].contiguous(),
                                    ct[:, :h].contiguous(), 1.0, 1.0)
    _, pen_b, con_b = ops.body_loss(m, pts, tb, rp[:, h:].contiguous(), rq[:, h:].contiguous(), jr[:, h:].contiguous(),
                                    ct[:, h:].contiguous(), 1.0, 1.0)
    assert_close(pen_a + pen_b, pen, rtol=1e-5, what="pen additivity")
    assert_close(con_a + con_b, con, rtol=1e-5, atol=1e-5, what="contact additivity")
    assert (pen >= 0).all() and torch.isfinite(pen).all() and torch.isfinite(con).all()


# ----------------------------------------------------------------------------------------- dataset sweep (8(f)-1)
def test_frames_fk_matches_three_step_path(gpu_model):
    from parc_b200 import ops
    civ, q = golden("clip_civilization.npz"), golden("dof_golden.npz")
    fr = dev(civ["frames"])
    bp, br, rr, jr = ops.frames_fk(gpu_model.c_model(), fr, want_rot=True)
    assert_close(rr, q["root_quat"], what="frames_fk root_rot")
    assert_close(jr, q["joint_rot"], what="frames_fk joint_rot")
    bp2, br2 = gpu_model.forward_kinematics(fr[:, 0:3], ops.exp_map_to_quat(fr[:, 3:6]), gpu_model.dof_to_rot(fr[:, 6:]))
    assert torch.equal(bp, bp2) and torch.equal(br, br2)
    bp3, _ = gpu_model.frames_forward_kinematics(fr.view(2, 127, 34))
    assert bp3.shape == (2, 127, 15, 3) and torch.equal(bp3.view(254, 15, 3), bp)


# Thresholded outputs (contact labels, cell membership) can legitimately differ from the CPU reference only where the
# decision is numerically borderline.  These helpers compute, with the oracle, how far every decision is from its
# threshold, so the tests can PROVE each tolerated mismatch is such a case.
LABEL_TOL = 1e-5        # metres: GPU vs host FK / SDF differ by a few 1e-7


def _foot_decision(O, om, frames, terr, feet, eps=0.04):
    """-> {body: (worst margin [F] = min over the 8 box corners of z - (h + eps); on_border [F])}: the label is
    `worst margin < 0`; a corner whose xy sits within a few ulp of a cell border may read a neighbouring height."""
    bp, br = O.frames_fk(om, frames)
    out = {}
    for b, half, off in feet:
        pts = O.box_corners(bp[:, b], br[:, b], torch.tensor(half, dtype=torch.float32), torch.tensor(off, dtype=torch.float32))
        h = O.hf_sample(terr, pts[..., 0:2])
        coord = O.grid_coord(terr, pts[..., 0:2])
        out[b] = ((pts[..., 2] - (h + eps)).min(dim=-1)[0], _border_mask(coord.reshape(-1, 2), ulps=16).view(-1, 8).any(dim=-1))
    return out


def _hand_decision(O, om, frames, terr, hands, eps=0.04):
    bp, _ = O.frames_fk(om, frames)
    base_z = torch.min(terr.hf).item() - 10.0
    out = {}
    for b, radius in hands:
        sd = O.points_hf_sdf(bp[:, b].unsqueeze(0), terr.hf.unsqueeze(0), terr.min_point.unsqueeze(0), terr.dxdy, base_z, False)
        out[b] = sd[0] - radius - eps              # label = margin < 0
    return out


def _assert_label_mismatches_are_borderline(got, exp, foot_dec, hand_dec, what):
    """Every (frame, body) where the labels differ must have its decision within LABEL_TOL of the threshold (or a foot
    corner on a cell border); returns the number of such justified mismatches."""
    diff = (got != exp).nonzero().tolist()
    for f, b in diff:
        if b in foot_dec:
            margin, border = foot_dec[b]
            assert abs(float(margin[f])) <= LABEL_TOL or bool(border[f]), \
                f"{what}: foot label of frame {f} body {b} differs although its margin is {float(margin[f]):.3e}"
        elif b in hand_dec:
            assert abs(float(hand_dec[b][f])) <= LABEL_TOL, \
                f"{what}: hand label of frame {f} body {b} differs although its margin is {float(hand_dec[b][f]):.3e}"
        else:
            raise AssertionError(f"{what}: label of a body that is neither foot nor hand differs (frame {f} body {b})")
    return len(diff)


def _uncertain_cells(O, om, frames, terr):
    """bool [F, X, Y]: cells a surface point within 16 ulp of a cell border could be attributed to (the 3 x 3 block
    around the cell the oracle put it in)."""
    bp, br = O.frames_fk(om, frames)
    X, Y = terr.hf.shape
    out = torch.zeros(frames.shape[0], X, Y, dtype=torch.bool)
    for f in range(frames.shape[0]):
        pts = torch.cat([O.quat_rotate(br[f, b].unsqueeze(0), om.body_points[b]) + bp[f, b] for b in range(om.num_bodies)], dim=0)
        near = _border_mask(O.grid_coord(terr, pts[:, 0:2]), ulps=16)
        g = O.grid_index(terr, pts[near][:, 0:2])
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                out[f, (g[:, 0] + dx).clamp(0, X - 1), (g[:, 1] + dy).clamp(0, Y - 1)] = True
    return out


def _golden_feet_hands(g):
    feet = [(int(b), h.tolist(), o.tolist()) for b, h, o in zip(g["feet_body"], g["feet_half"], g["feet_offset"])]
    hands = [(int(b), float(r)) for b, r in zip(g["hands_body"], g["hands_radius"])]
    return feet, hands


def test_contact_labelling_vs_golden(gpu_model, O, oracle_model):
    from parc_b200.zmotion_editing_tools.motion_edit_lib import (compute_hf_foot_contacts_and_correct_pen,
                                                                 compute_motion_terrain_hand_contacts)
    g = golden("label_golden.npz")
    t = _civ_terrain()
    ot = O.Terrain(hf=t.hf.cpu(), min_point=t.min_point.cpu(), dxdy=t.dxdy.cpu())
    feet, hands = _golden_feet_hands(g)
    frames = dev(g["frames"])
    upd, fc = compute_hf_foot_contacts_and_correct_pen(frames, t, gpu_model)
    assert fc.shape == (64, 15)
    # labels are exact except where the thresholded quantity sits on its threshold -- proven per mismatch
    foot_dec = _foot_decision(O, oracle_model, torch.tensor(g["frames"]), ot, feet)
    n = _assert_label_mismatches_are_borderline(fc.cpu(), torch.tensor(g["foot_contacts"]), foot_dec, {}, "feet")
    assert n <= 2
    assert fc[:, [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 13]].abs().sum() == 0
    assert_close(upd[:, 2], g["updated_z"], atol=2e-6, what="penetration-corrected root z")
    assert torch.equal(upd[:, :2], frames[:, :2]) and torch.equal(upd[:, 3:], frames[:, 3:])
    hc = compute_motion_terrain_hand_contacts(frames, t, gpu_model)
    hand_dec = _hand_decision(O, oracle_model, torch.tensor(g["frames"]), ot, hands)
    assert _assert_label_mismatches_are_borderline(hc.cpu(), torch.tensor(g["hand_contacts"]), {}, hand_dec, "hands") <= 2
    t.hf = t.hf + float(g["raised_by"])
    ot2 = O.Terrain(hf=t.hf.cpu(), min_point=t.min_point.cpu(), dxdy=t.dxdy.cpu())
    hc2 = compute_motion_terrain_hand_contacts(frames, t, gpu_model)
    hand_dec2 = _hand_decision(O, oracle_model, torch.tensor(g["frames"]), ot2, hands)
    assert _assert_label_mismatches_are_borderline(hc2.cpu(), torch.tensor(g["hand_contacts_raised"]), {}, hand_dec2,
                                                   "hands, raised terrain") <= 2
    assert hc2.sum() > 10


def test_hf_mask_inds_vs_golden(gpu_model, O, oracle_model):
    from parc_b200.util import geom_util
    from parc_b200.util.terrain_util import compute_hf_extra_vals, compute_hf_mask_from_inds, compute_hf_mask_inds
    g = golden("label_golden.npz")
    t = _civ_terrain()
    ot = O.Terrain(hf=t.hf.cpu(), min_point=t.min_point.cpu(), dxdy=t.dxdy.cpu())
    frames = dev(g["frames"][:24])
    pts = geom_util.get_char_point_samples(gpu_model)
    inds, minh = compute_hf_mask_inds(frames, t, gpu_model, pts)
    assert len(inds) == 24 and all(i.dtype == torch.int64 and i.shape[1] == 2 for i in inds)
    # cell membership is exact except for surface points within a few ulp of a cell border: every differing cell of
    # every frame must be one such a point could be attributed to
    unc = _uncertain_cells(O, oracle_model, torch.tensor(g["frames"][:24]), ot)
    exp_counts = g["mask_counts"].tolist()
    exp_flat, s0 = torch.tensor(g["mask_inds"]), 0
    for f, n in enumerate(exp_counts):
        exp_m = torch.zeros(50, 50, dtype=torch.bool)
        e = exp_flat[s0:s0 + n]
        exp_m[e[:, 0], e[:, 1]] = True
        got_m = torch.zeros(50, 50, dtype=torch.bool)
        got_m[inds[f][:, 0].cpu(), inds[f][:, 1].cpu()] = True
        assert not ((got_m != exp_m) & ~unc[f]).any(), f"frame {f}: a cell differs that no border point explains"
        s0 += n
    assert sum(abs(a.shape[0] - b) for a, b in zip(inds, exp_counts)) <= 3
    mask = compute_hf_mask_from_inds(t, inds).cpu()
    anyf = unc.any(dim=0)
    assert not ((mask != torch.tensor(g["hf_mask"])) & ~anyf).any() and (mask != torch.tensor(g["hf_mask"])).sum() <= 2
    exp_minh = torch.tensor(g["min_body_heights"])
    same = (minh.cpu() - exp_minh).abs() <= 1e-5 * exp_minh.abs().clamp(min=1.0)
    assert not (~same & ~anyf).any() and (~same).sum() <= 2
    compute_hf_extra_vals(frames, t, gpu_model, pts)
    assert not ((t.hf_mask.cpu() != torch.tensor(g["extra_hf_mask"])) & ~anyf).any()
    mm = ((t.hf_maxmin.cpu() - torch.tensor(g["extra_hf_maxmin"])).abs() > 1e-4).any(dim=-1)
    assert not (mm & ~anyf).any() and mm.sum() <= 4


@pytest.mark.parametrize("B,F", [(5, 40), (32, 265)])
def test_label_clips_batched_vs_oracle(gpu_model, O, oracle_model, B, F):
    """Config-5-shaped: a batch of clips, one 16x16 terrain per clip, feet + hands + body hf + masks in one launch;
    (32, 265) is a subsample of BASELINE configs[4] at its own clip length.  Labels / cell memberships are exact
    except where the oracle's decision margin shows the case is borderline."""
    from parc_b200 import ops
    from parc_b200.util import geom_util, synth
    from parc_b200.zmotion_editing_tools.motion_edit_lib import label_clips
    g = golden("label_golden.npz")
    rng = np.random.default_rng(8)
    hfs = np.stack([synth.box_terrain(rng, h_range=(-0.4, 0.7)) if i % 2 else synth.stairs_terrain(rng) for i in range(B)])
    fr = np.concatenate([synth.synth_clips(gpu_model, 1, seed=50 + i, num_frames=F, hf=hfs[i])[0] for i in range(B)])
    fr[..., 2] -= 0.02
    tb = ops.make_terrain_batch(torch.tensor(hfs).cuda(), torch.zeros(B, 2).cuda(), (0.4, 0.4),
                                base_z=(torch.tensor(hfs).amin(dim=(1, 2)) - 10.0).cuda())
    pts = geom_util.get_char_point_samples(gpu_model)
    out = label_clips(torch.tensor(fr).cuda(), tb, gpu_model, body_points=pts, want_masks=True, want_fk=True)
    feet, hands = _golden_feet_hands(g)
    bad = 0
    mask_clips = range(B) if B <= 8 else range(0, B, 8)            # the oracle's mask restatement loops over frames
    for i in range(B):
        t = O.Terrain(hf=torch.tensor(hfs[i]), min_point=torch.zeros(2), dxdy=torch.tensor([0.4, 0.4]))
        f_i = torch.tensor(fr[i])
        _, fc, corr = O.foot_contacts_and_pen(oracle_model, f_i, t, feet)
        hc = O.hand_contacts(oracle_model, f_i, t, hands)
        got = out["contacts"][i].cpu()
        if (got != fc + hc).any():
            bad += _assert_label_mismatches_are_borderline(got, fc + hc, _foot_decision(O, oracle_model, f_i, t, feet),
                                                           _hand_decision(O, oracle_model, f_i, t, hands), f"clip {i}")
        # the correction is a min over corners of z - h: a corner on a cell border may read the neighbouring height
        dec = _foot_decision(O, oracle_model, f_i, t, feet)
        on_border = torch.stack([d[1] for d in dec.values()]).any(dim=0)
        pc = out["pen_correction"][i].cpu()
        assert ((pc - corr).abs() <= 2e-6)[~on_border].all(), f"pen_correction clip {i}"
        bp, _ = O.frames_fk(oracle_model, f_i)
        assert_close(out["body_pos"][i], bp, what="label body_pos")
        bhf = O.hf_sample(t, bp[..., 0:2])
        origin_border = _border_mask(O.grid_coord(t, bp[..., 0:2]).reshape(-1, 2), ulps=16).view(F, -1)
        assert not ((out["body_hf"][i].cpu() != bhf) & ~origin_border).any()
        if i in mask_clips:
            inds, minh = O.hf_mask_inds(oracle_model, f_i, t)
            unc = _uncertain_cells(O, oracle_model, f_i, t)
            masks = ops.unpack_frame_masks(out["frame_mask_bits"][i], 16, 16).cpu()
            exp_m = torch.zeros(F, 16, 16, dtype=torch.bool)
            for f, ind in enumerate(inds):
                exp_m[f, ind[:, 0], ind[:, 1]] = True
            assert not ((masks != exp_m) & ~unc).any(), f"clip {i}: a mask cell differs that no border point explains"
            same = (out["min_body_heights"][i].cpu() - minh).abs() <= 1e-5 * minh.abs().clamp(min=1.0)
            assert not (~same & ~unc.any(dim=0)).any()
    assert bad <= max(3, B * F // 1000), f"{bad} borderline contact labels"
    assert out["contacts"].sum() > 0


# ----------------------------------------------------------------------------------------- generic character (J > 15)
def test_generic_21_body_character_vs_oracle(O):
    """Characters with more than 15 bodies take the one-character-per-warp (G = 32) instantiation of the fused
    kernel; random tree, hinge/spherical/fixed mix, non-identity local rotations.  Query + FK + obs, the
    stand-alone FK forward/backward and the DoF conversions, all against the oracle."""
    from conftest import make_random_tree_model
    from parc_b200 import ops
    om, plain = make_random_tree_model(21)
    m = ops.make_char_model(**plain)
    J, D = 21, om.dof_size
    rng = np.random.default_rng(12)
    clips = []
    for c in range(3):
        F = [40, 17, 64][c]
        fr = np.zeros((F, 6 + D), np.float32)
        t = np.arange(F)[:, None] / 30.0
        fr[:, 0:3] = np.cumsum(rng.normal(scale=0.02, size=(F, 3)), axis=0) + [3.0, 3.0, 1.0]
        fr[:, 3:6] = 0.3 * np.sin(t * rng.uniform(0.5, 2, 3) + rng.uniform(0, 6, 3)) + 0.05
        fr[:, 6:] = 0.8 * np.sin(t * rng.uniform(0.5, 3, D) + rng.uniform(0, 6, D)) + 0.01
        ct = (rng.uniform(size=(F, J)) < 0.4).astype(np.float32)
        clips.append(O.Clip(fr, ct, 30.0, O.WRAP if c == 1 else O.CLAMP, 1.0))
    tb = O.build_tables(om, clips)
    # pack the oracle's tables with the library and query through the C ABI
    dv = lambda x: x.cuda()
    rows, lay = ops.pack_frames(m, dv(tb.root_pos), dv(tb.root_rot), dv(tb.joint_rot), dv(tb.contacts), dv(tb.root_vel),
                                dv(tb.root_ang_vel), dv(tb.dof_vel))
    meta = ops.build_clip_meta(tb.num_frames, tb.loop_modes, tb.start_idx, tb.lengths, tb.root_pos_delta, "cuda:0")
    packed = ops.PackedTables(rows=rows, clips=meta, total_frames=int(rows.shape[0]), num_clips=3, layout=lay)
    gen = torch.Generator().manual_seed(9)
    n = 777
    ids = torch.randint(0, 3, (n,), generator=gen)
    times = (torch.rand(n, generator=gen) * 2.5 - 0.5) * tb.lengths[ids]
    hf = torch.tensor(np.random.default_rng(3).uniform(-0.3, 0.5, size=(24, 20)).astype(np.float32))
    hfd = ops.HeightfieldDesc(hf=hf.cuda(), min_x=-1.0, min_y=0.5, dx=0.4, dy=0.3)
    tmpl = O.cone_template(0.05, 2, 60, 3, 3, 0.26179938779)
    r = ops.motion_query(packed, m, ids.cuda(), motion_times=times.cuda(), want_index=True, want_fk=True, hf=hfd,
                         obs_tmpl=tmpl.cuda())
    i0, i1, bl = O.frame_blend(tb, ids, times)
    assert torch.equal(r["frame_idx0"].cpu(), i0) and torch.equal(r["frame_idx1"].cpu(), i1)
    assert torch.equal(r["blend"].cpu(), bl)
    ref = O.calc_motion_frame(tb, ids, times)
    for k, t in zip(FRAME_KEYS, ref):
        assert_close(r[k], t, what=f"generic.{k}")
    assert torch.equal(r["dof_vel"].cpu(), ref[5]) and torch.equal(r["contacts"].cpu(), ref[6])
    bp, br = O.forward_kinematics(om, ref[0], ref[1], ref[4])
    assert_close(r["body_pos"], bp, what="generic.body_pos")
    assert_close(r["body_rot"], br, what="generic.body_rot")
    ot = O.Terrain(hf=hf, min_point=torch.tensor([-1.0, 0.5]), dxdy=torch.tensor([0.4, 0.3]))
    heading = O.calc_heading(ref[1])
    exp_obs = O.ray_obs(ot, ref[0], heading, tmpl)
    mism = r["obs"].cpu() != exp_obs
    border = _border_mask(O.grid_coord(ot, O.ray_obs_points(ref[0], heading, tmpl)), ulps=64).view(n, 441)
    assert not (mism & ~border).any() and mism.float().mean() < 2e-3
    # integer-frame lookup
    fi = torch.randint(0, 17, (n,), generator=gen)
    g = ops.motion_query(packed, m, ids.cuda(), frame_idxs=fi.cuda())
    gref = O.get_motion_frame(tb, ids, fi)
    assert torch.equal(g["joint_rot"].cpu(), gref[4]) and torch.equal(g["root_pos"].cpu(), gref[0])
    # stand-alone FK forward + VJP and DoF conversions on the raw frames
    fr = torch.tensor(clips[2].frames)
    wp = torch.randn(64, J, 3, generator=gen)
    wr = torch.randn(64, J, 4, generator=gen)

    def run(fk, d2r, e2q, to):
        a, b, c = (fr[:, 0:3].clone().to(to).requires_grad_(True), fr[:, 3:6].clone().to(to).requires_grad_(True),
                   fr[:, 6:].clone().to(to).requires_grad_(True))
        p_, r_ = fk(a, e2q(b), d2r(c))
        ((p_ * wp.to(to)).sum() + (r_ * wr.to(to)).sum()).backward()
        return p_.detach(), r_.detach(), a.grad, b.grad, c.grad

    cpu = run(lambda p_, q_, j_: O.forward_kinematics(om, p_, q_, j_), lambda d: O.dof_to_rot(om, d), O.exp_map_to_quat, "cpu")
    gpu = run(lambda p_, q_, j_: ops.forward_kinematics(m, p_, q_, j_), lambda d: ops.dof_to_rot(m, d), ops.exp_map_to_quat,
              "cuda:0")
    assert_close(gpu[0], cpu[0], what="generic fk pos")
    assert_close(gpu[1], cpu[1], what="generic fk rot")
    for nm, a, b in zip(("root_pos", "root_exp", "joint_dof"), gpu[2:], cpu[2:]):
        assert_close_normwise(a, b, what=f"generic grad {nm}")
    bp2, br2 = ops.frames_fk(m, fr.cuda())
    assert_close(bp2, cpu[0], what="generic frames_fk pos")


# ----------------------------------------------------------------------------------------- full objective + Adam loop (8(f)-2)
def _motion_opt_case(gpu_model):
    from parc_b200.tools.motion_opt.motion_optimization import BodyConstraint
    g = golden("motion_opt_golden.npz")
    W = {str(k): float(v) for k, v in zip(g["weight_names"], g["weights"])}
    bcs = [[] for _ in range(15)]
    for b, s_, e_, pt in zip(g["bc_body"], g["bc_start"], g["bc_end"], g["bc_point"]):
        c = BodyConstraint()
        c.start_frame_idx, c.end_frame_idx, c.constraint_point = int(s_), int(e_), dev(pt)
        bcs[int(b)].append(c)
    return g, W, bcs


def test_full_objective_with_body_constraints_vs_golden(gpu_model):
    from parc_b200.tools.motion_opt.motion_optimization import LossType, motion_terrain_contact_loss
    from parc_b200.util import geom_util
    g, W, bcs = _motion_opt_case(gpu_model)
    pts = geom_util.get_char_point_samples(gpu_model)
    fr = dev(g["src_frames"])
    a, b, c = (fr[:, 0:3].clone().requires_grad_(True), fr[:, 3:6].clone().requires_grad_(True),
               fr[:, 6:].clone().requires_grad_(True))
    loss, ld = motion_terrain_contact_loss(a, b, c, dev(g["src_root_pos"]), dev(g["src_root_quat"]), dev(g["src_joint_rot"]),
                                           dev(g["src_body_vels"]), dev(g["src_body_rot_vels"]), dev(g["contacts"]),
                                           _civ_terrain(), pts, gpu_model, body_constraints=bcs, max_jerk=1000.0, **W)
    loss.backward()
    assert_close(loss, g["loss"], what="full objective")
    assert_close(torch.tensor(ld[LossType.BODY_CONSTRAINT_LOSS]), g["body_constraint_term"], what="body constraint term")
    assert ld[LossType.BODY_CONSTRAINT_LOSS] > 0
    assert_close_normwise(a.grad, g["grad_root_pos"], what="full grad root_pos")
    assert_close_normwise(b.grad, g["grad_root_exp"], what="full grad root_rot")
    assert_close_normwise(c.grad, g["grad_joint_dof"], what="full grad joint_dof")


@pytest.mark.parametrize("use_graph", [True, False])
def test_motion_contact_optimization_vs_golden(gpu_model, use_graph):
    """4 Adam iterations of the reference's loop (tools/motion_opt/motion_optimization.py:404-500); the CUDA-graph
    replay path and the eager path must both land on the reference's frames."""
    from parc_b200.tools.motion_opt.motion_optimization import motion_contact_optimization
    from parc_b200.util import geom_util
    g, W, bcs = _motion_opt_case(gpu_model)
    pts = geom_util.get_char_point_samples(gpu_model)
    src = dev(g["src_frames"])
    out = motion_contact_optimization(src.clone(), dev(g["contacts"]), pts, _civ_terrain(), gpu_model, num_iters=4,
                                      step_size=0.001, body_constraints=bcs, max_jerk=1000.0, exp_name="t",
                                      use_wandb=False, log_file=None, use_cuda_graph=use_graph, quiet=True, **W)
    assert out.shape == (8, 34)
    upd, exp_upd = (out - src).cpu(), torch.tensor(g["adam4_frames"] - g["src_frames"])
    assert exp_upd.abs().max() > 1e-3                              # the loop actually moved the pose
    # Adam normalises every gradient to ~+-lr per step.  Leaves whose gradient is decisive moved ~4 * lr and must
    # match tightly; leaves whose true gradient cancels to rounding level (|g| ~ eps = 1e-8) get g / (|g| + eps),
    # which turns last-bit differences into a fraction of one step -- in the reference just as much.
    decisive = exp_upd.abs() >= 2e-3
    assert decisive.float().mean() > 0.5
    assert ((upd - exp_upd).abs()[decisive] <= 0.02 * exp_upd.abs()[decisive]).all(), "decisive leaves differ"
    assert (upd - exp_upd).abs().max() <= 0.3e-3, "a noise-level leaf moved by more than 0.3 of one step"
    assert torch.equal(src, dev(g["src_frames"]))                  # the input is not modified

    # and the objective at our optimum equals the objective at the reference's optimum
    from parc_b200.tools.motion_opt.motion_optimization import motion_terrain_contact_loss

    def objective(fr):
        with torch.no_grad():
            return motion_terrain_contact_loss(fr[:, 0:3], fr[:, 3:6], fr[:, 6:], dev(g["src_root_pos"]), dev(g["src_root_quat"]),
                                               dev(g["src_joint_rot"]), dev(g["src_body_vels"]), dev(g["src_body_rot_vels"]),
                                               dev(g["contacts"]), _civ_terrain(), pts, gpu_model, body_constraints=bcs,
                                               max_jerk=1000.0, **W)[0]
    l_ours, l_ref, l_src = objective(out), objective(dev(g["adam4_frames"])), objective(src)
    assert l_ref < l_src and l_ours < l_src
    assert_close(l_ours, l_ref, rtol=2e-3, what="objective after 4 iterations")


def test_fused_objective_vs_oracle_autograd_with_every_term_active(gpu_model, O, oracle_model):
    """The fused objective kernel (csrc/motion_opt.cu) against autograd through the oracle on a 20-frame segment with
    every stencil exercised: a small max_jerk (the jerk clamp is active on most bodies), sliding on contact bodies,
    a sphere-body and a box-body constraint (whose frames pay no sliding), negative contact labels."""
    from parc_b200.tools.motion_opt.motion_optimization import BodyConstraint, LossType, motion_terrain_contact_loss
    from parc_b200.util import geom_util
    civ = golden("clip_civilization.npz")
    F = 20
    gen = torch.Generator().manual_seed(21)
    src = torch.tensor(civ["frames"][60:60 + F]).clone()
    cts = torch.tensor(civ["contacts"][60:60 + F]).clone()
    cts[3:6, 11] = -0.2                                    # MDM-style negative labels (clamped in the sliding weight only)
    tgt = src + 0.02 * torch.randn(F, 34, generator=gen)   # perturbed leaves: velocities, jerk and tracking all non-zero
    tgt[:, 2] -= 0.03
    W = dict(w_root_pos=1.0, w_root_rot=10.0, w_joint_rot=1.0, w_smoothness=10.0, w_penetration=1000.0, w_contact=1000.0,
             w_sliding=10.0, w_body_constraints=1000.0, w_jerk=1000.0)
    max_jerk = 30.0
    rq, jr = O.exp_map_to_quat(src[:, 3:6]), O.dof_to_rot(oracle_model, src[:, 6:])
    bp, br = O.forward_kinematics(oracle_model, src[:, 0:3], rq, jr)
    bv, brv = bp[1:] - bp[:-1], O.quat_diff_angle(br[1:], br[:-1])
    lf, rh = 14, 5                                         # left_foot (box geom), right_hand (sphere geom)
    pt_lf = (bp[8, lf] + torch.tensor([0.03, -0.02, -0.05])).tolist()
    pt_rh = (bp[12, rh] + torch.tensor([0.04, 0.02, -0.03])).tolist()
    o_bcs = [[] for _ in range(15)]
    o_bcs[lf], o_bcs[rh] = [(6, 11, pt_lf)], [(10, 25, pt_rh)]            # the second one runs past the last frame
    g = golden("humanoid_model.npz")
    geom0 = []
    for b in range(15):
        gm = gpu_model.get_geoms(b)[0]
        dims = gm._dims.detach().cpu().reshape(-1).tolist()
        geom0.append((1 if gm._shape_type.name == "SPHERE" else (0 if gm._shape_type.name == "BOX" else 2),
                      gm._offset.detach().cpu().tolist(), dims if len(dims) > 1 else dims[0]))
    hf, mp, dxdy = torch.tensor(civ["hf"]), torch.tensor(civ["min_point"]), torch.tensor(civ["dxdy"])
    a, b, c = (tgt[:, 0:3].clone().requires_grad_(True), tgt[:, 3:6].clone().requires_grad_(True),
               tgt[:, 6:].clone().requires_grad_(True))
    want, wt = O.motion_terrain_contact_loss_full(oracle_model, a, b, c, src[:, 0:3], rq, jr, bv, brv, cts, hf, mp, dxdy, W,
                                                  max_jerk, o_bcs, geom0)
    want.backward()
    bcs = [[] for _ in range(15)]
    for body, (s_, e_, pt) in ((lf, o_bcs[lf][0]), (rh, o_bcs[rh][0])):
        k = BodyConstraint()
        k.start_frame_idx, k.end_frame_idx, k.constraint_point = s_, e_, dev(np.array(pt, dtype=np.float32))
        bcs[body].append(k)
    pts = geom_util.get_char_point_samples(gpu_model)
    x, y, z = (dev(tgt[:, 0:3].numpy()).requires_grad_(True), dev(tgt[:, 3:6].numpy()).requires_grad_(True),
               dev(tgt[:, 6:].numpy()).requires_grad_(True))
    got, ld = motion_terrain_contact_loss(x, y, z, dev(src[:, 0:3].numpy()), rq.cuda(), jr.cuda(), bv.cuda(), brv.cuda(),
                                          cts.cuda(), _civ_terrain(), pts, gpu_model, body_constraints=bcs, max_jerk=max_jerk, **W)
    got.backward()
    assert_close(got, want.detach(), rtol=2e-5, what="objective, every term active")
    for key, name in ((LossType.JERK_LOSS, "jerk"), (LossType.SLIDING_LOSS, "sliding"), (LossType.SMOOTHNESS_LOSS, "smoothness"),
                      (LossType.BODY_CONSTRAINT_LOSS, "body_constraint"), (LossType.ROOT_ROT_LOSS, "root_rot"),
                      (LossType.JOINT_ROT_LOSS, "joint_rot")):
        assert float(wt[name]) > 0, name
        assert_close(torch.tensor(float(ld[key])), torch.tensor(float(wt[name])), rtol=2e-5, atol=1e-6, what=name)
    assert_close_normwise(x.grad, a.grad, 1e-5, what="d/d root_pos")
    assert_close_normwise(y.grad, b.grad, 1e-5, what="d/d root exp-map")
    assert_close_normwise(z.grad, c.grad, 1e-5, what="d/d joint dofs")


def test_motion_optimisation_graph_replay_equals_eager_launches_bit_for_bit(gpu_model):
    """The CUDA-graph path replays exactly the four launches the eager path enqueues; every kernel is deterministic
    (no atomics in the gradient), so 60 iterations land on identical bits (the ATen version differed by 0.1
    after 300 iterations between its capturable-Adam graph and eager forms)."""
    from parc_b200.tools.motion_opt.motion_optimization import motion_contact_optimization
    from parc_b200.util import geom_util
    g, W, bcs = _motion_opt_case(gpu_model)
    pts = geom_util.get_char_point_samples(gpu_model)
    src = dev(g["src_frames"])
    outs = [motion_contact_optimization(src.clone(), dev(g["contacts"]), pts, _civ_terrain(), gpu_model, num_iters=60,
                                        step_size=0.001, body_constraints=bcs, max_jerk=1000.0, use_cuda_graph=ug,
                                        quiet=True, **W) for ug in (True, False, True)]
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert (outs[0] - src).abs().max() > 5e-3


def test_300_adam_iterations_track_the_reference_trajectory(gpu_model):
    """tests/golden/motion_opt300_golden.npz: the reference's own loop (its motion_terrain_contact_loss + torch Adam) for
    300 iterations on a 16 x 16 terrain, checkpointed (oracle/make_golden_opt300.py; the oracle port reproduces it bit
    for bit there).  Adam on this non-smooth objective (arg-min cells, clamps, g / (|g| + eps) normalisation) amplifies
    last-bit differences between fp32 implementations: iteration 1 agrees to rounding, after that the two trajectories
    separate slowly while descending the same objective.  Bars: the first update is bit-level identical; the objective
    at every checkpoint agrees within 0.1 % / 1 % / 1 % / 5 %; the mean distance stays a fraction of the mean
    distance travelled.  (Measured on B200: profiles/r2_opt300_drift.json.)"""
    from parc_b200.tools.motion_opt.motion_optimization import (motion_contact_optimization, motion_terrain_contact_loss,
                                                                 source_constants)
    from parc_b200.util import geom_util
    from parc_b200.util.terrain_util import SubTerrain
    g = golden("motion_opt300_golden.npz")
    W = {str(k): float(v) for k, v in zip(g["weight_names"], g["weights"])}
    t = SubTerrain("crop", x_dim=16, y_dim=16, dx=0.4, dy=0.4, min_x=float(g["min_point"][0]),
                   min_y=float(g["min_point"][1]), device="cuda:0")
    t.hf = dev(g["hf"])
    pts = geom_util.get_char_point_samples(gpu_model)
    src, cts = dev(g["src_frames"]), dev(g["contacts"])
    rq, jr, bv, brv = source_constants(src, gpu_model)

    def objective(fr):
        with torch.no_grad():
            return float(motion_terrain_contact_loss(fr[:, 0:3], fr[:, 3:6], fr[:, 6:], src[:, 0:3], rq, jr, bv, brv, cts, t,
                                                     pts, gpu_model, body_constraints=None, max_jerk=1000.0, **W)[0])
    obj_tol = {1: 1e-5, 4: 1e-3, 25: 1e-2, 100: 1e-2, 300: 5e-2}
    dist_tol = {1: 1e-3, 4: 0.02, 25: 0.10, 100: 0.25, 300: 0.35}
    l_src = objective(src)
    for it, ref in zip(g["checkpoints"].tolist(), g["frames"]):
        out = motion_contact_optimization(src.clone(), cts, pts, t, gpu_model, num_iters=it, step_size=0.001,
                                          body_constraints=None, max_jerk=1000.0, quiet=True, **W)
        ref = dev(ref)
        travelled = (ref - src).abs().mean().item()
        assert (out - ref).abs().mean().item() <= dist_tol[it] * travelled, f"{it} iterations: trajectory distance"
        l_ours, l_ref = objective(out), objective(ref)
        assert l_ours < l_src and abs(l_ours - l_ref) <= obj_tol[it] * l_ref, f"{it} iterations: objective {l_ours} vs {l_ref}"
        if it == 1:
            assert (out - ref).abs().max().item() <= 1e-7
    assert l_ours < 0.02 * l_src                      # 300 iterations removed > 98 % of the objective, as in the reference


# ----------------------------------------------------------------------------------------- edge cases with hand-built tables
def _pack_oracle_tables(gpu_model, tb):
    from parc_b200 import ops
    m = gpu_model.c_model()
    rows, lay = ops.pack_frames(m, tb.root_pos.cuda(), tb.root_rot.cuda(), tb.joint_rot.cuda(), tb.contacts.cuda(),
                                tb.root_vel.cuda(), tb.root_ang_vel.cuda(), tb.dof_vel.cuda())
    meta = ops.build_clip_meta(tb.num_frames, tb.loop_modes, tb.start_idx, tb.lengths, tb.root_pos_delta, "cuda:0")
    return m, ops.PackedTables(rows=rows, clips=meta, total_frames=int(rows.shape[0]), num_clips=int(tb.num_frames.shape[0]),
                               layout=lay)


def test_slerp_edge_cases_through_the_abi(gpu_model, O, oracle_model):
    """Key-frame pairs crafted to hit every slerp branch of util/torch_util.py:443-468: identical quaternions
    (|cos| >= 1 -> q0), antipodal ones (cos < 0 -> sign flip, then q0), half-angles around the sin < 1e-3 midpoint
    threshold, large rotations; plus 2-frame CLAMP and WRAP clips, negative / huge / exact-boundary times.  The tables are edited AFTER loading, so they need not be reachable from DoF frames."""
    from parc_b200 import ops
    civ = golden("clip_civilization.npz")
    clips = [O.Clip(civ["frames"][:6], civ["contacts"][:6], 30.0, O.CLAMP, 1.0),
             O.Clip(civ["frames"][10:12], civ["contacts"][10:12], 30.0, O.CLAMP, 1.0),        # 2 frames: the minimum the
                                                                                              # reference can load
             O.Clip(civ["frames"][20:22], civ["contacts"][20:22], 30.0, O.WRAP, 1.0)]
    tb = O.build_tables(oracle_model, clips)
    gen = torch.Generator().manual_seed(77)
    q = torch.nn.functional.normalize(torch.randn(14, 4, generator=gen), dim=-1)
    q = torch.where(q[:, 3:] < 0, -q, q)

    def rotated(qb, half_angle):        # qb (x) small rotation about a random axis, half-angle given
        ax = torch.nn.functional.normalize(torch.randn(qb.shape[0], 3, generator=gen), dim=-1)
        d = torch.cat([ax * torch.sin(half_angle).unsqueeze(-1), torch.cos(half_angle).unsqueeze(-1)], dim=-1)
        return O.quat_mul(qb, d)

    tb.joint_rot[0] = q
    tb.joint_rot[1] = q.clone()                                                   # identical -> q0
    tb.joint_rot[2] = -q                                                          # antipodal -> flip -> q0
    ha = torch.tensor([2e-4, 5e-4, 9e-4, 9.9e-4, 1.0e-3, 1.01e-3, 1.1e-3, 2e-3, 1e-2, 0.3, 1.0, 1.5, 1.57, 0.7])
    tb.joint_rot[3] = rotated(tb.joint_rot[2], ha)                                # around the midpoint threshold
    tb.joint_rot[4] = -rotated(tb.joint_rot[3], ha)                               # negative dot AND a real angle
    tb.root_rot[1] = tb.root_rot[0].clone()
    tb.root_rot[2] = -tb.root_rot[1]
    m, packed = _pack_oracle_tables(gpu_model, tb)
    L = tb.lengths
    ids = torch.tensor([0] * 40 + [1] * 6 + [2] * 10)
    times = torch.cat([torch.linspace(0.0, L[0].item(), 33), torch.tensor([-5.0, -0.0, 1e9, L[0].item() * 0.999999, 0.01, 0.05, 0.1]),
                       torch.tensor([0.0, 1.0, -1.0, 1e-30, 3.4e38, 0.5]),
                       torch.tensor([0.0, L[2].item(), L[2].item() * 3, -L[2].item() * 2.5, 0.02, 0.033, 0.0333, 10.0, -10.0, 1e6])])
    r = ops.motion_query(packed, m, ids.cuda(), motion_times=times.cuda(), want_index=True, want_fk=True)
    i0, i1, bl = O.frame_blend(tb, ids, times)
    ref = O.calc_motion_frame(tb, ids, times)
    finite = ~torch.isnan(times / L[ids])
    assert torch.equal(r["frame_idx0"].cpu()[finite], i0[finite]) and torch.equal(r["frame_idx1"].cpu()[finite], i1[finite])
    assert torch.equal(r["blend"].cpu()[finite], bl[finite])
    ok = finite & ~torch.isnan(ref[4]).any(dim=-1).any(dim=-1)
    assert ok.sum() >= 50
    for k, t in zip(FRAME_KEYS, ref):
        assert_close(r[k][ok.cuda()], t[ok], what=f"edge.{k}")
    # branch-exact outputs: identical / antipodal key frames and sub-threshold angles are IEEE-only
    q0, q1 = tb.joint_rot[i0], tb.joint_rot[i1]
    c = torch.sum(q0 * q1, dim=-1).abs()
    non_slerp = ((c >= 1) | (torch.sqrt(1.0 - c * c) < 0.001)) & ok.unsqueeze(-1)
    assert non_slerp.sum() > 100
    assert torch.equal(r["joint_rot"].cpu()[non_slerp], ref[4][non_slerp])
    bp, br = O.forward_kinematics(oracle_model, ref[0], ref[1], ref[4])
    assert_close(r["body_pos"][ok.cuda()], bp[ok], what="edge.body_pos")


def test_heightfield_cell_borders_and_out_of_range(gpu_model):
    """Points exactly on cell borders (round half to even), outside the grid, at +-inf and NaN."""
    from parc_b200.util.terrain_util import SubTerrain
    t = SubTerrain("b", x_dim=5, y_dim=4, dx=0.4, dy=0.25, min_x=-0.8, min_y=1.0, device="cuda:0")
    t.hf = torch.arange(20, dtype=torch.float32, device="cuda").view(5, 4)
    xs = torch.tensor([-0.8 + 0.4 * k + 0.2 for k in range(-2, 6)] + [float("inf"), -float("inf"), float("nan"), 1e30, -1e30])
    ys = torch.tensor([1.0 + 0.25 * k + 0.125 for k in range(-2, 5)] + [float("inf"), -float("inf"), float("nan"), 0.0, 1e9, 1.3])
    pts = torch.stack([xs, ys], dim=-1)
    idx = t.get_grid_index(pts.cuda()).cpu()
    exp = torch.round((pts - torch.tensor([-0.8, 1.0])) / torch.tensor([0.4, 0.25])).to(torch.int64)
    exp = torch.clamp(exp, torch.zeros(2, dtype=torch.int64), torch.tensor([4, 3]))
    assert torch.equal(idx, exp)
    z = t.get_hf_val_from_points(pts.cuda()).cpu()
    assert torch.equal(z, t.hf.cpu()[exp[:, 0], exp[:, 1]])


# ----------------------------------------------------------------------------------------- BASELINE full sizes: properties
@pytest.fixture(scope="module")
def big_lib(gpu_model):
    """The bench's library: 2048 synthetic 265-frame clips (260 MB packed), built on the GPU."""
    from parc_b200.anim.motion_lib import LoopMode, MotionLib
    from parc_b200.util import synth
    rng = np.random.default_rng(1234)
    hf = synth.rolling_terrain(rng, 512, 512, num_boxes=800)
    frames, contacts = synth.synth_clips(gpu_model, 2048, seed=1235, hf=hf)
    lib = MotionLib(torch.from_numpy(frames).cuda(), gpu_model, "cuda:0", init_type="motion_frames",
                    loop_mode=LoopMode.WRAP, fps=30, contact_info=True, contacts=torch.from_numpy(contacts).cuda())
    return lib, torch.from_numpy(hf).cuda()


def test_full_size_query_is_batch_order_and_split_invariant(big_lib, gpu_model):
    """Config 4 size (65 536 envs): every query is independent, so permuting the batch permutes the outputs
    bit-exactly, querying two halves equals querying the whole, and the fused FK equals the stand-alone FK."""
    from parc_b200 import ops
    from parc_b200.util import geom_util
    lib, hf = big_lib
    n = 65536
    gen = torch.Generator().manual_seed(1)
    ids = torch.randint(0, 2048, (n,), generator=gen).cuda()
    times = ((torch.rand(n, generator=gen) * 3.0 - 1.0) * (264.0 / 30.0)).cuda()
    hfd = ops.HeightfieldDesc(hf=hf, min_x=0.0, min_y=0.0, dx=0.4, dy=0.4)
    tmpl = geom_util.get_xy_points_cone(torch.zeros(2, device="cuda"), 0.05, 2, 60, 3, 3, 0.26179938779)
    keys = FRAME_KEYS + ("body_pos", "body_rot", "obs")
    full = {k: v.clone() for k, v in lib.calc_motion_frame_fk_obs(ids, times, hf_desc=hfd, obs_tmpl=tmpl).items()}
    perm = torch.randperm(n, generator=gen).cuda()
    shuf = lib.calc_motion_frame_fk_obs(ids[perm], times[perm], hf_desc=hfd, obs_tmpl=tmpl)
    for k in keys:
        assert torch.equal(shuf[k], full[k][perm]), f"permutation changed {k}"
    lo = lib.calc_motion_frame_fk_obs(ids[:30001].contiguous(), times[:30001].contiguous(), hf_desc=hfd, obs_tmpl=tmpl)
    for k in keys:
        assert torch.equal(lo[k], full[k][:30001]), f"split changed {k}"
    bp, br = gpu_model.forward_kinematics(full["root_pos"], full["root_rot"], full["joint_rot"])
    assert torch.equal(bp, full["body_pos"]) and torch.equal(br, full["body_rot"])
    assert torch.isfinite(full["body_pos"]).all() and (full["obs"].abs() <= 3.0).all()
    # unit quaternions stay (nearly) unit through slerp + FK for these smooth clips
    assert (full["body_rot"].norm(dim=-1) - 1).abs().max() < 1e-3
    i0, i1, bl = lib._calc_frame_blend(ids, times)
    assert ((i1 - i0 == 1) | (i1 == i0)).all() and (bl >= 0).all() and (bl < 1).all()
    assert (i0 // 265 == ids).all() and (i1 // 265 == ids).all()          # indices never leave their clip


def test_full_size_loss_frames_are_independent(gpu_model):
    """Config 3 size (1024 samples x 200 frames, one terrain per sample): per-frame terms and gradients do not
    depend on which other frames / samples share the launch."""
    from parc_b200 import ops
    from parc_b200.tools.procgen.mdm_path import body_points_desc
    from parc_b200.util import geom_util, synth
    rng = np.random.default_rng(3)
    B, F = 1024, 200
    base = [synth.box_terrain(rng) if i % 2 == 0 else synth.stairs_terrain(rng) for i in range(16)]
    hfs = torch.tensor(np.stack([base[i % 16] for i in range(B)])).cuda()
    s = synth.synth_motion_samples(gpu_model, B, F, base[0], (0.0, 0.0), (0.4, 0.4), seed=11)
    rp = torch.tensor(s["root_pos"]).cuda()
    rq = ops.exp_map_to_quat(torch.tensor(s["root_exp"]).cuda())
    jr = gpu_model.dof_to_rot(torch.tensor(s["joint_dof"]).cuda())
    ct = torch.tensor(s["contacts"]).cuda()
    pts = body_points_desc(gpu_model, geom_util.get_char_point_samples(gpu_model))
    m = gpu_model.c_model()
    tb = ops.make_terrain_batch(hfs, torch.zeros(B, 2).cuda(), (0.4, 0.4), base_z=-10.0)
    full = ops._body_loss_launch(m, pts, tb, rp, rq, jr, ct, 0.1, 0.1, True)
    sub = slice(100, 164)
    fr = slice(37, 90)
    tb2 = ops.make_terrain_batch(hfs[sub].contiguous(), torch.zeros(64, 2).cuda(), (0.4, 0.4), base_z=-10.0)
    part = ops._body_loss_launch(m, pts, tb2, rp[sub, fr].contiguous(), rq[sub, fr].contiguous(), jr[sub, fr].contiguous(),
                                 ct[sub, fr].contiguous(), 0.1, 0.1, True)
    for a, b in zip(full, part):
        assert torch.equal(a[sub, fr], b)
    assert torch.isfinite(full[0]).all() and (full[0] >= 0).all() and full[0].sum() > 0
    # a launch of few frames runs the CTA-per-frame form of the kernel (the motion optimiser's regime): same bits
    sub, fr = slice(700, 704), slice(0, 200)
    tb3 = ops.make_terrain_batch(hfs[sub].contiguous(), torch.zeros(4, 2).cuda(), (0.4, 0.4), base_z=-10.0)
    few = ops._body_loss_launch(m, pts, tb3, rp[sub, fr].contiguous(), rq[sub, fr].contiguous(), jr[sub, fr].contiguous(),
                                ct[sub, fr].contiguous(), 0.1, 0.1, True)
    for a, b in zip(full, few):
        assert torch.equal(a[sub, fr], b)


def test_full_size_frames_fk_split_and_cross_kernel_consistency(gpu_model):
    """Config 5 per-GPU shard (12 500 clips x 265 frames): the one-launch raw-frame FK equals the three-step path
    (exp_map_to_quat, dof_to_rot, forward_kinematics) bit-exactly, and is split invariant."""
    from parc_b200 import ops
    from parc_b200.util import synth
    fr_np, _ = synth.synth_clips(gpu_model, 250, seed=77)
    fr = torch.tensor(fr_np).cuda().repeat(50, 1, 1).reshape(-1, 34)        # 3 312 500 frames
    assert fr.shape[0] == 12500 * 265
    bp, br = ops.frames_fk(gpu_model.c_model(), fr)
    sel = torch.randint(0, fr.shape[0], (20000,), generator=torch.Generator().manual_seed(2)).cuda()
    f2 = fr[sel]
    bp2, br2 = gpu_model.forward_kinematics(f2[:, 0:3], ops.exp_map_to_quat(f2[:, 3:6]), gpu_model.dof_to_rot(f2[:, 6:]))
    assert torch.equal(bp[sel], bp2) and torch.equal(br[sel], br2)
    bp3, _ = ops.frames_fk(gpu_model.c_model(), fr[1000001:2000000])
    assert torch.equal(bp3, bp[1000001:2000000])
    assert torch.isfinite(bp).all()


# ----------------------------------------------------------------------------------------- f3: tracker step assembly
def _cu(x):
    return torch.as_tensor(np.asarray(x)).cuda()


def _step_golden():
    g = golden("tracker_step_golden.npz")
    names = ("root_pos", "root_rot", "root_vel", "root_ang_vel", "joint_rot", "dof_vel")
    sim = tuple(_cu(g[k]) for k in names)
    ref = tuple(_cu(g["ref_" + k]) for k in names)
    return g, sim, ref, _cu(g["key_ids"]).long()


def test_step_char_and_tar_obs_match_reference():
    from parc_b200.envs import ig_char_env
    from parc_b200.envs.ig_parkour import mgdm_dm_util as dm
    g, sim, ref, key_ids = _step_golden()
    key = _cu(g["body_pos"])[:, key_ids]
    none = torch.zeros([0], device="cuda")
    tar = tuple(_cu(g["tar_" + k]) for k in ("root_pos", "root_rot", "joint_rot", "key_pos"))
    for gl in (0, 1):
        for h in (0, 1):
            o = ig_char_env.compute_char_obs(*sim, key, bool(gl), bool(h))
            assert_close(o, g[f"char_obs_g{gl}_h{h}"], what=f"char_obs g{gl} h{h}")
            o = dm.compute_tar_obs(sim[0], sim[1], *tar, bool(gl), bool(h))
            assert_close(o, g[f"tar_obs_g{gl}_h{h}"], what=f"tar_obs g{gl} h{h}")
    assert_close(ig_char_env.compute_char_obs(*sim, none, False, False), g["char_obs_nokey"])
    assert_close(dm.compute_tar_obs(sim[0], sim[1], tar[0], tar[1], tar[2], none, False, False), g["tar_obs_nokey"])
    d = dm.compute_deepmimic_obs(*sim, key, False, False, True, *tar)
    assert list(d) == ["char_obs", "tar_obs"]
    assert_close(d["tar_obs"], g["tar_obs_g0_h0"])
    # empty batch (an empty key tensor means "no key bodies", as in the reference)
    e = ig_char_env.compute_char_obs(*(t[:0] for t in sim), key[:0], False, False)
    assert e.shape == (0, g["char_obs_nokey"].shape[1])


def test_step_reward_matches_reference():
    from parc_b200.envs.ig_parkour import mgdm_dm_util as dm
    g, sim, ref, key_ids = _step_golden()
    key, ref_key = _cu(g["body_pos"])[:, key_ids], _cu(g["ref_body_pos"])[:, key_ids]
    jw, dw = _cu(g["joint_err_w"]), _cu(g["dof_err_w"])
    for tr in (1, 0):
        for th in (1, 0):
            r = dm.compute_deepmimic_reward(*sim, key, *ref, ref_key, jw, dw, bool(th), bool(tr))
            assert_close(r, g[f"reward_r{tr}_h{th}"], what=f"reward track_root={tr} track_h={th}")
    with pytest.raises(ValueError):
        none = torch.zeros([0], device="cuda")
        dm.compute_deepmimic_reward(*sim, none, *ref, none, jw, dw, True, True)
    # identical states: every error is exactly zero -> all five rewards are exactly 1
    r = dm.compute_deepmimic_reward(*ref, ref_key, *ref, ref_key, jw, dw, True, True)
    assert torch.equal(r, torch.ones_like(r))


def test_step_done_flags_bit_exact():
    from parc_b200.envs.ig_parkour import mgdm_dm_util as dm
    from parc_b200.util.terrain_util import SubTerrain
    g, sim, ref, key_ids = _step_golden()
    n = sim[0].shape[0]
    feet = _cu(g["feet"]).long()
    bp, rbp = _cu(g["body_pos"]), _cu(g["ref_body_pos"])
    tm, cf, ptd = _cu(g["time_buf"]), _cu(g["contact_forces"]), _cu(g["pose_termination_dist"])
    th = _cu(g["term_heights"])
    done_in = torch.zeros(n, dtype=torch.int, device="cuda")
    cases = {"default": (torch.zeros([0], dtype=torch.long, device="cuda"), True, True, True),
             "feet": (feet, True, True, True), "feet_nopose": (feet, False, True, True),
             "noroot": (feet, True, True, False), "noearly": (feet, True, False, True)}
    terr = SubTerrain("step", g["hf"].shape[0], g["hf"].shape[1], float(g["hf_dxdy"][0]), float(g["hf_dxdy"][1]),
                      float(g["hf_min"][0]), float(g["hf_min"][1]), device="cuda")
    terr.hf[...] = _cu(g["hf"])
    for tag, (cids, pose, early, track) in cases.items():
        d = dm.compute_done(done_in, tm, 10.0, sim[1], bp, sim[0], ref[1], rbp, cf, cids, th, pose, ptd, False, early,
                            track, 0.6, 1.309)
        assert d.dtype == torch.int32 and torch.equal(d.cpu(), torch.as_tensor(g["done_" + tag])), tag
        # the fused form: heights sampled in the same launch
        buf = torch.full((n,), 7, dtype=torch.int, device="cuda")
        dm.update_done(buf, tm, terr, _cu(g["env_offsets"]), 0.15, 10.0, cids, pose, ptd, False, early, track, 0.6, 1.309,
                       sim[1], bp, ref[1], rbp, cf)
        assert torch.equal(buf.cpu(), torch.as_tensor(g["done_" + tag])), tag + " (fused heights)"
    from parc_b200 import ops
    _, hts = ops.done_flags(tm, 10.0, sim[1], bp, ref[1], rbp, cf, [], True, ptd, True, True, 0.6, 1.309,
                            hf=terr.hf_desc(), env_offsets=_cu(g["env_offsets"]), termination_height=0.15,
                            want_heights=True)
    assert torch.equal(hts.cpu(), torch.as_tensor(g["term_heights"]))
    d64 = dm.compute_done(done_in.long(), tm, 10.0, sim[1], bp, sim[0], ref[1], rbp, cf, feet, th, True, ptd, False, True,
                          True, 0.6, 1.309)
    assert d64.dtype == torch.int64


def test_step_assembly_matches_oracle_on_a_generic_tree():
    """A 21-body random tree, 1000 envs, 3 key bodies, thresholds chosen from the data's own quantiles so that every
    branch splits the batch: values to 1e-5, flags bit-exact."""
    from parc_b200 import ops
    from conftest import make_random_tree_model
    from oracle import parc_oracle as O
    model_o, _ = make_random_tree_model(21, seed=9)
    J, D = 21, model_o.dof_size
    gen = torch.Generator().manual_seed(5)
    n, S = 1000, 4

    def rq(*shape):
        q = torch.randn(*shape, 4, generator=gen)
        return q / q.norm(dim=-1, keepdim=True)

    def state(noise_of=None, amp=0.3):
        if noise_of is None:
            return [torch.randn(n, 3, generator=gen), rq(n), torch.randn(n, 3, generator=gen),
                    torch.randn(n, 3, generator=gen), rq(n, J - 1), torch.randn(n, D, generator=gen)]
        out = []
        for t in noise_of:
            x = t + amp * torch.rand(n, *([1] * (t.dim() - 1)), generator=gen) * torch.randn(t.shape, generator=gen)
            out.append(x / x.norm(dim=-1, keepdim=True) if t.shape[-1] == 4 else x)
        return out

    ref = state()
    sim = state(ref)
    rbp, _ = O.forward_kinematics(model_o, ref[0], ref[1], ref[4])
    sbp, _ = O.forward_kinematics(model_o, sim[0], sim[1], sim[4])
    key_ids = torch.tensor([4, 11, 20])
    jw, dw = torch.rand(J - 1, generator=gen), torch.rand(D, generator=gen)
    tar = [torch.randn(n, S, 3, generator=gen), rq(n, S), rq(n, S, J - 1), torch.randn(n, S, 3, 3, generator=gen)]
    c = lambda ts: [t.cuda() for t in ts]
    for gl in (False, True):
        # randn positions reach |4|: components that cancel to ~0 carry the rounding of their O(4) terms -> atol 4e-6
        assert_close(ops.char_obs(*c(sim), sbp[:, key_ids].cuda(), gl, True),
                     O.compute_char_obs(*sim, sbp[:, key_ids], gl, True), atol=4e-6, what="char_obs")
        assert_close(ops.tar_obs(sim[0].cuda(), sim[1].cuda(), *c(tar), gl, not gl),
                     O.compute_tar_obs(sim[0], sim[1], *tar, gl, not gl), atol=4e-6, what="tar_obs")
    for tr in (True, False):
        r = ops.deepmimic_reward(tuple(c(sim)) + (sbp[:, key_ids].cuda(),), tuple(c(ref)) + (rbp[:, key_ids].cuda(),),
                                 jw.cuda(), dw.cuda(), tr, tr)
        assert_close(r, O.compute_deepmimic_reward(*sim, sbp[:, key_ids], *ref, rbp[:, key_ids], jw, dw, tr, tr),
                     what="reward")
    # done: thresholds at data quantiles
    d = (rbp[:, 1:] - rbp[:, :1]) - (sbp[:, 1:] - sbp[:, :1])
    ptd = (d * d).sum(-1).sqrt().quantile(0.97, dim=0)
    root_d = float((sbp[:, 0] - rbp[:, 0]).norm(dim=-1).quantile(0.8))
    ang = float(O.quat_diff_angle(sim[1], ref[1]).abs().quantile(0.8))
    hf = torch.rand(24, 20, generator=gen) * 2 - 1
    terr_o = O.Terrain(hf=hf, min_point=torch.tensor([-2.0, -2.0]), dxdy=torch.tensor([0.25, 0.25]))
    off = torch.randn(n, 3, generator=gen)
    th = O.termination_heights(terr_o, sbp, off, 0.2)
    tm = torch.rand(n, generator=gen) * 6
    tm[:10] = 0
    cf = torch.randn(n, J, 3, generator=gen) * (torch.rand(n, J, 1, generator=gen) < 0.05)
    allowed = torch.tensor([0, 7, 8, 19])
    hfd = ops.HeightfieldDesc(hf=hf.cuda(), min_x=-2.0, min_y=-2.0, dx=0.25, dy=0.25)
    seen = set()
    for cids in (allowed, torch.zeros([0], dtype=torch.long)):
        for pose in (True, False):
            for track in (True, False):
                want = O.compute_done(torch.zeros(n, dtype=torch.int), tm, 5.0, sim[1], sbp, ref[1], rbp, cf, cids, th,
                                      pose, ptd, True, track, root_d, ang)
                got = ops.done_flags(tm.cuda(), 5.0, sim[1].cuda(), sbp.cuda(), ref[1].cuda(), rbp.cuda(), cf.cuda(),
                                     cids.tolist(), pose, ptd.cuda(), True, track, root_d, ang, hf=hfd,
                                     env_offsets=off.cuda(), termination_height=0.2)
                mism = (got.cpu() != want)
                # a root angle within 1e-6 of its threshold may legitimately land on either side (libm atan2)
                close = (O.quat_diff_angle(sim[1], ref[1]).abs() - ang).abs() < 1e-6
                assert not (mism & ~close).any(), (int(mism.sum()), cids.tolist(), pose, track)
                seen |= set(want.tolist())
    assert seen == {0, 1, 3}


@pytest.mark.parametrize("variant", [3, 5])
@pytest.mark.parametrize("global_obs", [False, True])
def test_query_fused_target_observation_equals_the_standalone_operator(golden_lib, variant, global_obs):
    """ParcTarObsSpec: the step-form query writes compute_tar_obs of steps 1..S from its registers.  Against parc_tar_obs
    run on the query's own stored targets (same device functions, same inputs): every instantiation that carries it
    (half / whole sweep in flight), local and global frames, a strided output block, envs that do not fill the
    last warp."""
    from parc_b200 import ops
    gen = torch.Generator().manual_seed(31 + variant)
    n, S = 333, 4
    ids = torch.randint(0, golden_lib.num_motions(), (n,), generator=gen).cuda()
    times = (torch.rand(n, generator=gen) * 9.0 - 0.5).cuda()
    offs = torch.tensor([0.0, 1 / 30, 3 / 30, 0.7]).cuda()
    sim_pos = (torch.randn(n, 3, generator=gen) * 0.5).cuda()
    q = torch.randn(n, 4, generator=gen)
    sim_rot = (q / q.norm(dim=-1, keepdim=True)).cuda()
    keys = torch.tensor([4, 7, 10, 13], dtype=torch.int32).cuda()
    J = 15
    W = 9 + 6 * (J - 1) + 3 * 4
    wide = torch.full((n, 7 + (S - 1) * W + 5), -7.0, device="cuda")
    blk = wide[:, 7:7 + (S - 1) * W]
    plan = golden_lib.make_query_plan(ids, times, time_offsets=offs, variant=variant,
                                      tar_obs=dict(sim_root_pos=sim_pos, sim_root_rot=sim_rot, key_body_ids=keys, out=blk,
                                                   global_obs=global_obs, global_tar_root_h=False))
    out = plan.launch()
    v = {k: t.view(n, S, *t.shape[1:]) for k, t in out.items() if k != "obs"}
    want = ops.tar_obs(sim_pos, sim_rot, v["root_pos"][:, 1:], v["root_rot"][:, 1:], v["joint_rot"][:, 1:],
                       v["body_pos"][:, 1:], global_obs, False, key_body_ids=keys)
    got = blk.reshape(n, S - 1, W)
    assert_close(got, want, rtol=1e-6, atol=1e-6, what=f"fused target observation variant {variant}")
    assert (got - want).abs().max().item() <= 2e-6
    assert (wide[:, :7] == -7.0).all() and (wide[:, 7 + (S - 1) * W:] == -7.0).all()      # stays inside its block
    # the query's ordinary outputs are unchanged by the extra work
    plain = golden_lib.make_query_plan(ids, times, time_offsets=offs, variant=variant).launch()
    for k in ("root_pos", "root_rot", "joint_rot", "body_pos", "body_rot", "dof_vel"):
        assert torch.equal(out[k], plain[k]), k


@pytest.mark.parametrize("fused,fuse_tar,split_sim", [(False, False, False), (True, False, False), (True, False, True),
                                                      (True, True, False), (True, True, True)])
def test_tracker_step_sequence_matches_oracle_composition(golden_lib, gpu_model, O, oracle_tables, oracle_model, fused,
                                                          fuse_tar, split_sim):
    """The whole kinematic side of a control step (envs/ig_parkour/step_assembly.py): reference frame + 6 targets
    with the per-env terrain placement, simulated character's observation, ray heightmap, reward terms and done
    flags -- against the same sequence composed from the oracle; then the CUDA-graph replay against the eager run."""
    from parc_b200.envs.ig_parkour.step_assembly import TrackerStep
    from parc_b200.util import geom_util
    from parc_b200.util.terrain_util import SubTerrain
    g, sim, ref_g, key_ids = _step_golden()
    n = sim[0].shape[0]
    gen = torch.Generator().manual_seed(77)
    terr = SubTerrain("step", g["hf"].shape[0], g["hf"].shape[1], float(g["hf_dxdy"][0]), float(g["hf_dxdy"][1]),
                      float(g["hf_min"][0]), float(g["hf_min"][1]), device="cuda")
    terr.hf[...] = _cu(g["hf"])
    o_terr = O.Terrain(hf=torch.as_tensor(g["hf"]), min_point=torch.as_tensor(g["hf_min"]),
                       dxdy=torch.as_tensor(g["hf_dxdy"]))
    tmpl = geom_util.get_xy_points_cone(torch.zeros(2), 0.05, 2, 60, 3, 3, 0.26179938779)
    jw, ptd = torch.as_tensor(g["joint_err_w"]), torch.as_tensor(g["pose_termination_dist"])
    feet = [int(i) for i in g["feet"]]
    ts = TrackerStep(golden_lib, terr, n, float(g["dt"]), g["steps"].tolist(), key_ids.tolist(), tmpl, joint_err_w=jw,
                     pose_termination_dist=ptd, contact_body_ids=feet, fused=fused, fuse_tar_obs=fuse_tar,
                     split_sim=split_sim)
    ids, times = torch.as_tensor(g["ids"]).long(), torch.as_tensor(g["times"])
    xy_off = torch.randn(n, 2, generator=gen) * 0.3
    env_off = torch.as_tensor(g["env_offsets"])
    ts.motion_ids.copy_(ids); ts.motion_times.copy_(times); ts.motion_xy_offset.copy_(xy_off)
    sim_c = [t.cpu() for t in sim]
    dof_pos = O.rot_to_dof(oracle_model, sim_c[4])
    body_pos = O.forward_kinematics(oracle_model, sim_c[0], sim_c[1], O.dof_to_rot(oracle_model, dof_pos))[0]
    forces, tm = torch.as_tensor(g["contact_forces"]), torch.as_tensor(g["time_buf"])
    char_contacts = (torch.rand(n, 15, generator=gen) < 0.3).float()
    state = [t.cuda() for t in (sim_c[0], sim_c[1], sim_c[2], sim_c[3], dof_pos, sim_c[5], body_pos, forces, tm, env_off,
                                char_contacts)]
    res = ts.step(*state)

    # ---- the same sequence from the oracle ----
    S = len(g["steps"])
    offs = torch.cat([torch.zeros(1), torch.as_tensor(g["dt"]) * torch.as_tensor(g["steps"])])
    ids_t = ids.unsqueeze(-1).expand(n, S + 1).flatten()
    times_t = (times.unsqueeze(-1) + offs).flatten()
    fr = list(O.calc_motion_frame(oracle_tables, ids_t, times_t))
    fr[0] = fr[0].clone()
    fr[0][:, 0:2] += xy_off.repeat_interleave(S + 1, dim=0)
    bp_all = O.forward_kinematics(oracle_model, fr[0], fr[1], fr[4])[0]
    v = lambda t: t.view(n, S + 1, *t.shape[1:])
    rp, rr, rv, rw, jr, dv, ct, bp = (v(t) for t in (fr[0], fr[1], fr[2], fr[3], fr[4], fr[5], fr[6], bp_all))
    assert torch.equal(res["ref_root_pos"].cpu(), rp[:, 0])                       # placement is one exact fp32 add
    sim_jr = O.dof_to_rot(oracle_model, dof_pos)
    kid = key_ids.cpu()
    o_char = O.compute_char_obs(sim_c[0], sim_c[1], sim_c[2], sim_c[3], sim_jr, sim_c[5], body_pos[:, kid], False, False)
    o_tar = O.compute_tar_obs(sim_c[0], sim_c[1], rp[:, 1:], rr[:, 1:], jr[:, 1:], bp[:, 1:][:, :, kid], False, False)
    o_ray = O.ray_obs(o_terr, sim_c[0] + env_off, O.calc_heading(sim_c[1]), tmpl)
    o_obs = torch.cat([o_char, o_tar.reshape(n, -1), ct[:, 1:].reshape(n, -1), char_contacts, o_ray], dim=-1)
    assert res["obs"].shape == o_obs.shape
    # world coordinates reach ~20 m (ulp 1.9e-6); differences of them that land near 0 keep that granularity
    assert_close(res["obs"], o_obs, atol=6e-6, what="policy observation")
    dw = ts.dof_err_w.cpu()
    assert torch.equal(dw, torch.as_tensor(g["dof_err_w"]))
    o_rew = O.compute_deepmimic_reward(sim_c[0], sim_c[1], sim_c[2], sim_c[3], sim_jr, sim_c[5], body_pos[:, kid],
                                       rp[:, 0], rr[:, 0], rv[:, 0], rw[:, 0], jr[:, 0], dv[:, 0], bp[:, 0][:, kid], jw, dw,
                                       True, True)
    assert_close(res["reward_terms"], o_rew, what="reward terms")
    th = O.termination_heights(o_terr, body_pos, env_off, 0.15)
    o_done = O.compute_done(torch.zeros(n, dtype=torch.int), tm, 10.0, sim_c[1], body_pos, rr[:, 0], bp[:, 0], forces,
                            torch.tensor(feet), th, True, ptd, True, True, 0.6, 1.309)
    assert torch.equal(res["done"].cpu(), o_done) and set(o_done.tolist()) == {0, 1, 3}

    # ---- CUDA graph: one launch per step, same numbers; inputs are re-read from the same buffers ----
    eager = {k: t.clone() for k, t in res.items() if k in ("obs", "reward_terms", "done")}
    ts.capture(*state)
    ts.motion_times.add_(0.25)
    state[0].add_(0.01)
    out2 = ts.replay()
    assert not torch.equal(out2["obs"], eager["obs"])
    ts.motion_times.sub_(0.25)
    state[0].sub_(0.01)
    out3 = ts.replay()
    # (x + 0.01) - 0.01 is not always x in fp32: compare against a fresh eager run on the current buffers
    again = ts.step(*state)
    for k in ("obs", "reward_terms", "done"):
        assert torch.equal(out3[k], again[k]), k


# ----------------------------------------------------------------------------------------- f4: GPU loader + packed file
def test_gpu_loader_matches_host_built_tables(golden_lib, gpu_model, tmp_path):
    """MotionLib(..., build_on_device=True): every clip's conversions and finite differences in one launch
    (csrc/table_build.cu) against the host-built tables (which are bit-identical to the reference's).  Mixed
    lengths (254 / 58 / 40 frames), mixed fps (30 / 60), CLAMP and WRAP."""
    from parc_b200.anim.motion_lib import MotionLib
    y = write_clip_library(tmp_path, lib_clips_from_golden())
    lib = MotionLib(y, gpu_model, "cuda:0", init_type="motion_file", contact_info=True, build_on_device=True)
    h = golden_lib
    for k in ("_motion_num_frames", "_motion_start_idx", "_motion_lengths", "_motion_loop_modes", "_motion_weights",
              "_motion_fps", "_motion_dt", "_frame_root_pos", "_frame_contacts", "_motion_frames", "_frame_root_vel"):
        assert torch.equal(getattr(lib, k), getattr(h, k)), k          # IEEE-only arithmetic: exact
    assert torch.equal(lib._motion_root_pos_delta, h._motion_root_pos_delta)
    for k, atol in (("_frame_root_rot", 2e-6), ("_frame_joint_rot", 2e-6), ("_frame_root_ang_vel", 2e-4),
                    ("_frame_dof_vel", 2e-4)):
        # velocities = ulp-level quaternion differences x fps (up to 60)
        assert_close(getattr(lib, k), getattr(h, k), atol=atol, what=k)
    assert (lib._frame_joint_rot[..., 3] >= 0).all()                    # quat_pos applied
    # last frame of every clip repeats the previous velocity
    last = (lib._motion_start_idx + lib._motion_num_frames - 1)
    assert torch.equal(lib._frame_dof_vel[last], lib._frame_dof_vel[last - 1])
    assert torch.equal(lib._frame_root_ang_vel[last], lib._frame_root_ang_vel[last - 1])
    # the rows the kernel wrote are what packing the unpacked views would give
    from parc_b200 import ops
    rows2, _ = ops.pack_frames(gpu_model.c_model(), lib._frame_root_pos, lib._frame_root_rot, lib._frame_joint_rot,
                               lib._frame_contacts, lib._frame_root_vel, lib._frame_root_ang_vel, lib._frame_dof_vel)
    assert torch.equal(rows2, lib._packed.rows)
    gen = torch.Generator().manual_seed(8)
    ids = torch.randint(0, 3, (2000,), generator=gen).cuda()
    times = (torch.rand(2000, generator=gen) * 9.0 - 0.5).cuda()
    a, b = lib._calc_frame_blend(ids, times), h._calc_frame_blend(ids, times)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    # Query values: the reference's slerp is DISCONTINUOUS at sin(half angle) = 1e-3 (midpoint vs true slerp,
    # util/torch_util.py:460-466), a jump of up to |t - 0.5| * |q1 - q0| ~ 5e-4; key frames that differ in the last
    # bit can sit on either side.  So: everything within that jump, and all but a sliver within the normal bar.
    for name, x, y in zip(FRAME_KEYS, lib.calc_motion_frame(ids, times), h.calc_motion_frame(ids, times)):
        err = (x.double() - y.double()).abs()
        assert float(err.max()) < 1e-3, name
        assert float((err > 2e-4).double().mean()) < 1e-3, name


def test_packed_file_round_trip(golden_lib, gpu_model, tmp_path):
    """save_packed -> init_type="packed_file": identical rows, clip records, terrains and query results; a file
    packed for another character is refused."""
    from parc_b200.anim.motion_lib import MotionLib
    from parc_b200.util.terrain_util import SubTerrain
    lib = golden_lib
    t = SubTerrain("t0", 8, 6, 0.4, 0.4, -1.0, 2.0, device="cuda:0")
    t.hf[...] = torch.rand(8, 6, device="cuda")
    old_terrains, old_masks, old_extras = lib._terrains, lib._hf_mask_inds, lib._motion_extras
    lib._terrains = [t, None, None]
    # per-frame body-cover cell lists (ragged) and a JSON-able extra travel too
    masks0 = [torch.tensor([[0, 1], [3, 2]], device="cuda"), torch.zeros(0, 2, dtype=torch.long, device="cuda"),
              torch.tensor([[7, 5]], device="cuda")]
    lib._hf_mask_inds = [masks0, None, None]
    lib._motion_extras = [{"tag": "a", "k": [1, 2]}, None, None]
    path = str(tmp_path / "lib.parcpack")
    try:
        lib.save_packed(path)
    finally:
        lib._terrains, lib._hf_mask_inds, lib._motion_extras = old_terrains, old_masks, old_extras
    assert os.path.getsize(path) > lib._packed.rows.numel() * 4
    back = MotionLib(path, gpu_model, "cuda:0", init_type="packed_file", contact_info=True)
    assert torch.equal(back._packed.rows, lib._packed.rows) and torch.equal(back._packed.clips, lib._packed.clips)
    for k in ("_motion_num_frames", "_motion_start_idx", "_motion_lengths", "_motion_loop_modes", "_motion_weights",
              "_motion_fps", "_motion_dt", "_motion_root_pos_delta", "_frame_root_pos", "_frame_root_rot",
              "_frame_joint_rot", "_frame_root_vel", "_frame_root_ang_vel", "_frame_dof_vel", "_frame_contacts",
              "_motion_frames"):
        assert torch.equal(getattr(back, k), getattr(lib, k)), k
    assert back._motion_names == lib._motion_names and back.num_motions() == 3
    assert back._terrains[1] is None and torch.equal(back._terrains[0].hf, t.hf)
    assert back._hf_mask_inds[1] is None and len(back._hf_mask_inds[0]) == 3
    assert all(torch.equal(x, y) for x, y in zip(back._hf_mask_inds[0], masks0))
    assert back._motion_extras == [{"tag": "a", "k": [1, 2]}, None, None]
    assert torch.equal(back._terrains[0].min_point, t.min_point) and torch.equal(back._terrains[0].dxdy, t.dxdy)
    gen = torch.Generator().manual_seed(9)
    ids = torch.randint(0, 3, (777,), generator=gen).cuda()
    times = (torch.rand(777, generator=gen) * 9.0).cuda()
    for x, y in zip(back.calc_motion_frame(ids, times), lib.calc_motion_frame(ids, times)):
        assert torch.equal(x, y)
    assert back.sample_motions(5).shape == (5,)
    # another character model -> refused
    other = gpu_model.get_copy("cuda:0")
    other._local_translation[3, 0] += 0.01                       # a limb of different length
    with pytest.raises(ValueError):
        MotionLib(path, other, "cuda:0", init_type="packed_file", contact_info=True)
    # truncated file -> refused
    data = open(path, "rb").read()
    bad = str(tmp_path / "short.parcpack")
    open(bad, "wb").write(data[:len(data) // 2])
    with pytest.raises(ValueError):
        MotionLib(bad, gpu_model, "cuda:0", init_type="packed_file", contact_info=True)


def test_frame_index_path_is_bit_exact_over_random_clip_shapes(gpu_model, O, oracle_model, tmp_path):
    """Indices, blend, lerped root position / contacts and the un-blended velocities are IEEE-only arithmetic and
    must be bit-identical to the reference's for ANY clip shape: 160 clips with 2..400 frames at 24 / 29.97 / 30 /
    60 / 120 fps, CLAMP and WRAP, queried at adversarial times (exact key-frame times, clip ends and their float
    neighbours, negative and multi-cycle times, huge times)."""
    from parc_b200.anim.motion_lib import MotionLib
    rng = np.random.default_rng(2024)
    D = oracle_model.dof_size
    clips = []
    for c in range(160):
        n = int(rng.choice([2, 3, 5, 17, 64, 100, 265, 400])) if c < 40 else int(rng.integers(2, 401))
        fps = float(rng.choice([24.0, 29.97, 30.0, 60.0, 120.0]))
        fr = np.zeros((n, 6 + D), np.float32)
        t = np.arange(n)[:, None] / fps
        fr[:, 0:3] = np.cumsum(rng.normal(scale=0.03, size=(n, 3)), axis=0) + rng.uniform(-50, 50, 3)
        fr[:, 3:6] = 0.4 * np.sin(t * rng.uniform(0.5, 2, 3) + rng.uniform(0, 6, 3)) + 0.05
        fr[:, 6:] = 0.7 * np.sin(t * rng.uniform(0.5, 3, D) + rng.uniform(0, 6, D)) + 0.01
        ct = (rng.uniform(size=(n, 15)) < 0.4).astype(np.float32)
        clips.append(O.Clip(fr, ct, fps, O.WRAP if c % 3 == 1 else O.CLAMP, float(rng.uniform(0.1, 2.0))))
    tb = O.build_tables(oracle_model, clips)
    lib = MotionLib(write_clip_library(tmp_path, clips), gpu_model, "cuda:0", init_type="motion_file", contact_info=True)
    assert torch.equal(lib._motion_lengths.cpu(), tb.lengths) and torch.equal(lib._motion_start_idx.cpu(), tb.start_idx)
    M = len(clips)
    ids, times = [], []
    f32 = lambda x: np.float32(x)
    for c in range(M):
        n, fps, L = int(tb.num_frames[c]), float(tb.fps[c]), float(tb.lengths[c])
        ks = rng.integers(0, n, 6)
        cand = [0.0, L, np.nextafter(f32(L), f32(0)), np.nextafter(f32(L), f32(1e9)), -L, -0.5 * L, 2.0 * L, 7.25 * L,
                -3.5 * L, 1e6, -1e6, 1e-30, L * 0.5]
        cand += [k / fps for k in ks] + [float(np.nextafter(f32(k / fps), f32(1e9))) for k in ks]
        cand += [float(f32(k) * f32(1.0 / fps)) for k in ks] + list(rng.uniform(-2 * L, 3 * L, 20))
        ids += [c] * len(cand)
        times += [float(x) for x in cand]
    ids_t, times_t = torch.tensor(ids, dtype=torch.long), torch.tensor(times, dtype=torch.float32)
    i0, i1, bl = O.frame_blend(tb, ids_t, times_t)
    ref = O.calc_motion_frame(tb, ids_t, times_t)
    g0, g1, gb = lib._calc_frame_blend(ids_t.cuda(), times_t.cuda())
    assert torch.equal(g0.cpu(), i0) and torch.equal(g1.cpu(), i1) and torch.equal(gb.cpu(), bl)
    out = lib.calc_motion_frame(ids_t.cuda(), times_t.cuda())
    for k in (0, 2, 3, 5, 6):                          # root_pos, root_vel, root_ang_vel, dof_vel, contacts
        assert torch.equal(out[k].cpu(), ref[k]), FRAME_KEYS[k]
    assert torch.equal(lib.calc_motion_phase(ids_t.cuda(), times_t.cuda()).cpu(), O.motion_phase(tb, ids_t, times_t))
    assert (bl == 0).any() and (i0 == i1).any() and len(times) > 8000
    # rotations: slerp value path, 1e-5 (away from the reference's own branch discontinuity, see the loader test)
    err = (out[4].cpu().double() - ref[4].double()).abs()
    assert float((err > 2e-5).double().mean()) < 2e-3 and float(err.max()) < 1e-3


def test_strided_outputs_stay_inside_their_blocks(gpu_model):
    """Canary test for the row-strided outputs (char_obs / tar_obs / hf_obs write a column block of a wider buffer)
    and for the loader's row buffer: sentinels around the blocks must survive, the blocks must equal the dense
    results bit for bit."""
    import ctypes as C
    from parc_b200 import _lib, ops
    g, sim, ref, key_ids = _step_golden()
    n = sim[0].shape[0]
    key = _cu(g["body_pos"])[:, key_ids]
    tar = tuple(_cu(g["tar_" + k]) for k in ("root_pos", "root_rot", "joint_rot", "key_pos"))
    S = tar[0].shape[1]
    SENT = 12345.0
    dense_c = ops.char_obs(*sim, key, False, True)
    dense_t = ops.tar_obs(sim[0], sim[1], *tar, False, False)
    hf = torch.rand(30, 20, device="cuda")
    hfd = ops.HeightfieldDesc(hf=hf, min_x=-3.0, min_y=-2.0, dx=0.4, dy=0.4)
    tmpl = torch.randn(77, 2, device="cuda")
    dense_h = ops.hf_obs(hfd, tmpl, sim[0], None, relative=True, root_rot=sim[1])
    wc, wt, wh = dense_c.shape[1], S * dense_t.shape[2], 77
    buf = torch.full((n, 5 + wc + 3 + wt + 2 + wh + 4), SENT, device="cuda")
    c0, t0, h0 = 5, 5 + wc + 3, 5 + wc + 3 + wt + 2
    ops.char_obs(*sim, key, False, True, out=buf[:, c0:c0 + wc])
    ops.tar_obs(sim[0], sim[1], *tar, False, False, out=buf[:, t0:t0 + wt])
    ops.hf_obs(hfd, tmpl, sim[0], None, relative=True, root_rot=sim[1], out=buf[:, h0:h0 + wh])
    assert torch.equal(buf[:, c0:c0 + wc], dense_c)
    assert torch.equal(buf[:, t0:t0 + wt], dense_t.view(n, -1))
    assert torch.equal(buf[:, h0:h0 + wh], dense_h)
    gaps = torch.cat([buf[:, :c0], buf[:, c0 + wc:t0], buf[:, t0 + wt:h0], buf[:, h0 + wh:]], dim=1)
    assert (gaps == SENT).all()
    # zero-copy input views: step 0 / steps 1.. of an [n, S+1, ...] buffer give the same results as dense copies
    big = {k: torch.cat([torch.full_like(t[:, :1], SENT), t], dim=1) for k, t in
           zip(("root_pos", "root_rot", "joint_rot"), tar[:3])}
    body = torch.full((n, S + 1, 15, 3), SENT, device="cuda")
    body[:, 1:, key_ids] = tar[3]
    v = ops.tar_obs(sim[0], sim[1], big["root_pos"][:, 1:], big["root_rot"][:, 1:], big["joint_rot"][:, 1:], body[:, 1:],
                    False, False, key_body_ids=key_ids)
    assert torch.equal(v, dense_t)
    # loader: rows_out followed by a canary
    m = gpu_model.c_model()
    lay = ops.row_layout(m)
    civ = golden("clip_civilization.npz")
    fr = dev(civ["frames"])
    total = fr.shape[0]
    raw = torch.full((total * lay.row_floats + 64,), SENT, device="cuda")
    nf = torch.tensor([total], dtype=torch.long, device="cuda")
    clip = torch.zeros(total, dtype=torch.int32, device="cuda")
    start = torch.zeros(1, dtype=torch.long, device="cuda")
    fps = torch.tensor([30.0], device="cuda")
    dt = torch.tensor([1.0 / 30.0], device="cuda")
    rc = _lib.load().parc_build_tables(fr.data_ptr(), total, fr.shape[1], None, clip.data_ptr(), start.data_ptr(),
                                       nf.data_ptr(), fps.data_ptr(), dt.data_ptr(), 1, C.byref(m), raw.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    assert (raw[total * lay.row_floats:] == SENT).all() and not (raw[:total * lay.row_floats] == SENT).any()
    rows, _, _ = ops.build_tables(m, fr, None, nf, fps, dt)
    assert torch.equal(raw[:total * lay.row_floats].view(total, -1), rows)


@pytest.mark.parametrize("loop", ["CLAMP", "WRAP"])
def test_config0_single_265_frame_clip(gpu_model, O, oracle_model, tmp_path, loop):
    """BASELINE.json configs[0]: frame query + FK on ONE synthetic 265-frame 34-DoF humanoid clip (SURVEY §8(d)
    cfg 1): every integer frame through get_motion_frame and 4240 random times through calc_motion_frame, FK on
    all of them, against the oracle."""
    from parc_b200.anim.motion_lib import MotionLib
    from parc_b200.util import synth
    frames, contacts = synth.synth_clips(gpu_model, 1, seed=1234)
    assert frames.shape == (1, 265, 34)
    mode = O.WRAP if loop == "WRAP" else O.CLAMP
    clips = [O.Clip(frames[0], contacts[0], 30.0, mode, 1.0)]
    tb = O.build_tables(oracle_model, clips)
    lib = MotionLib(write_clip_library(tmp_path, clips), gpu_model, "cuda:0", init_type="motion_file", contact_info=True)
    zeros = torch.zeros(265, dtype=torch.long)
    fidx = torch.arange(265)
    got = lib.get_motion_frame(zeros.cuda(), fidx.cuda())
    want = O.get_motion_frame(tb, zeros, fidx)
    for k, (a, b) in enumerate(zip(got, want)):
        assert torch.equal(a.cpu(), b), ("get_motion_frame", FRAME_KEYS[k])        # pure gathers
    gen = torch.Generator().manual_seed(1234)
    ids = torch.zeros(4240, dtype=torch.long)
    times = (torch.rand(4240, generator=gen) * 1.4 - 0.2) * tb.lengths[0]
    out = lib.calc_motion_frame_fk_obs(ids.cuda(), times.cuda())
    i0, i1, bl = O.frame_blend(tb, ids, times)
    g0, g1, gb = lib._calc_frame_blend(ids.cuda(), times.cuda())
    assert torch.equal(g0.cpu(), i0) and torch.equal(g1.cpu(), i1) and torch.equal(gb.cpu(), bl)
    ref = O.calc_motion_frame(tb, ids, times)
    for k in (0, 2, 3, 5, 6):
        assert torch.equal(out[FRAME_KEYS[k]].cpu(), ref[k]), FRAME_KEYS[k]
    assert_close(out["root_rot"], ref[1], what="root_rot")
    assert_close(out["joint_rot"], ref[4], what="joint_rot")
    bp, br = O.forward_kinematics(oracle_model, ref[0], ref[1], ref[4])
    assert_close(out["body_pos"], bp, atol=2e-6, what="body_pos")
    assert_close(out["body_rot"], br, what="body_rot")
    # FK of the integer frames through the stand-alone operator
    bp2, br2 = gpu_model.forward_kinematics(got[0], got[1], got[4])
    wp, wr = O.forward_kinematics(oracle_model, want[0], want[1], want[4])
    assert_close(bp2, wp, atol=2e-6, what="fk body_pos")
    assert_close(br2, wr, what="fk body_rot")


def test_small_api_helpers(golden_lib, gpu_model, O, oracle_model, oracle_tables):
    """motion_util.motion_frames_from_mlib_format / cat_motion_frames and MotionLib._calc_loop_offset."""
    from parc_b200.util import motion_util
    civ = golden("clip_civilization.npz")
    fr = dev(civ["frames"]).view(2, 127, 34)
    mf = motion_util.motion_frames_from_mlib_format(fr, gpu_model, contacts=dev(civ["contacts"]).view(2, 127, 15))
    f_cpu = torch.as_tensor(civ["frames"])
    rq, jr = O.exp_map_to_quat(f_cpu[:, 3:6]), O.dof_to_rot(oracle_model, f_cpu[:, 6:])
    bp, br = O.forward_kinematics(oracle_model, f_cpu[:, 0:3], rq, jr)
    assert mf.root_pos.shape == (2, 127, 3) and mf.joint_rot.shape == (2, 127, 14, 4)
    assert_close(mf.root_rot.reshape(-1, 4), rq, what="root_rot")
    assert_close(mf.joint_rot.reshape(-1, 14, 4), jr, what="joint_rot")
    assert_close(mf.body_pos.reshape(-1, 15, 3), bp, atol=2e-6, what="body_pos")
    assert_close(mf.body_rot.reshape(-1, 15, 4), br, what="body_rot")
    both = motion_util.cat_motion_frames([mf.get_slice(slice(0, 100)), mf.get_slice(slice(100, 127))])
    for k in ("root_pos", "root_rot", "joint_rot", "body_pos", "body_rot", "contacts"):
        assert torch.equal(getattr(both, k), getattr(mf, k)), k
    gen = torch.Generator().manual_seed(4)
    ids = torch.randint(0, 3, (500,), generator=gen)
    times = (torch.rand(500, generator=gen) * 6 - 2) * oracle_tables.lengths[ids]
    assert torch.equal(golden_lib._calc_loop_offset(ids.cuda(), times.cuda()).cpu(), O.loop_offset(oracle_tables, ids, times))


@pytest.mark.parametrize("global_obs,root_h,track_root,track_h,feet_only,pose", [
    (False, False, True, True, True, True), (True, True, False, False, False, True),
    (False, True, False, True, True, False), (True, False, True, False, False, False)])
def test_sim_step_equals_the_standalone_operators(gpu_model, global_obs, root_h, track_root, track_h, feet_only, pose):
    """parc_sim_step (one launch) against dof_to_rot + char_obs + deepmimic_reward + done_flags + the two contact
    copies, over the flag combinations; reference frame read in place as step 0 of an [n, 3, ...] buffer."""
    from parc_b200 import ops
    g, sim, ref, key_ids = _step_golden()
    n, J, D = sim[0].shape[0], 15, 28
    kid = key_ids.to(torch.int32)
    dof_pos = gpu_model.rot_to_dof(sim[4])
    body_pos = gpu_model.forward_kinematics(sim[0], sim[1], gpu_model.dof_to_rot(dof_pos))[0].contiguous()
    SENT = 777.0
    names = ("root_pos", "root_rot", "root_vel", "root_ang_vel", "joint_rot", "dof_vel")
    big = {k: torch.full((n, 3) + tuple(t.shape[1:]), SENT, device="cuda") for k, t in zip(names, ref)}
    for k, t in zip(names, ref):
        big[k][:, 0] = t
    big["body_pos"] = torch.full((n, 3, J, 3), SENT, device="cuda")
    big["body_pos"][:, 0] = _cu(g["ref_body_pos"])
    big["contacts"] = torch.rand(n, 3, J, device="cuda")
    refv = {k: v[:, 0] for k, v in big.items()}
    forces, tm, off = _cu(g["contact_forces"]), _cu(g["time_buf"]), _cu(g["env_offsets"])
    char_contacts = (torch.rand(n, J, device="cuda") < 0.5).float()
    hfd = ops.HeightfieldDesc(hf=_cu(g["hf"]), min_x=float(g["hf_min"][0]), min_y=float(g["hf_min"][1]),
                              dx=float(g["hf_dxdy"][0]), dy=float(g["hf_dxdy"][1]))
    jw, dw, ptd = _cu(g["joint_err_w"]), _cu(g["dof_err_w"]), _cu(g["pose_termination_dist"])
    allowed = [int(i) for i in g["feet"]] if feet_only else []
    cfg = dict(global_obs=global_obs, root_height_obs=root_h, track_root=track_root, track_root_h=track_h,
               pose_termination=pose, enable_early_termination=True, termination_height=0.15, episode_length=10.0,
               root_pos_termination_dist=0.6, root_rot_termination_angle=1.309, pose_termination_dist=ptd)
    wc = (1 if root_h else 0) + 12 + 6 * (J - 1) + D + 3 * 4
    obs = torch.full((n, 3 + wc + 2 * J + J + 5), SENT, device="cuda")
    out = dict(char_obs=obs[:, 3:3 + wc], tar_contacts=obs[:, 3 + wc:3 + wc + 2 * J],
               char_contacts=obs[:, 3 + wc + 2 * J:3 + wc + 3 * J], reward=torch.empty(n, 5, device="cuda"),
               done=torch.empty(n, dtype=torch.int32, device="cuda"), joint_rot=torch.empty(n, J - 1, 4, device="cuda"))
    simd = dict(root_pos=sim[0], root_rot=sim[1], root_vel=sim[2], root_ang_vel=sim[3], dof_pos=dof_pos, dof_vel=sim[5],
                body_pos=body_pos, contact_force=forces, time=tm, env_offsets=off, char_contacts=char_contacts)
    refd = dict(refv, tar_contacts=big["contacts"][:, 1:])
    plan = ops.SimStepPlan(gpu_model.c_model(), simd, refd, kid, jw, dw, hfd, out, cfg=cfg, contact_body_ids=allowed)
    plan.launch()
    # stand-alone operators
    jr = gpu_model.dof_to_rot(dof_pos)
    want_obs = ops.char_obs(sim[0], sim[1], sim[2], sim[3], jr, sim[5], body_pos, global_obs, root_h, key_body_ids=kid)
    want_rew = ops.deepmimic_reward((sim[0], sim[1], sim[2], sim[3], jr, sim[5], body_pos),
                                    tuple(refv[k] for k in names) + (refv["body_pos"],), jw, dw, track_h, track_root,
                                    key_body_ids=kid)
    want_done = ops.done_flags(tm, 10.0, sim[1], body_pos, refv["root_rot"], refv["body_pos"], forces, allowed, pose, ptd,
                               True, track_root, 0.6, 1.309, hf=hfd, env_offsets=off, termination_height=0.15)
    assert_close(out["joint_rot"], jr, what="joint_rot")
    assert_close(out["char_obs"], want_obs, atol=2e-6, what="char_obs")
    assert_close(out["reward"], want_rew, atol=2e-6, what="reward")
    assert torch.equal(out["done"], want_done)
    assert torch.equal(out["tar_contacts"], big["contacts"][:, 1:].reshape(n, -1))
    assert torch.equal(out["char_contacts"], char_contacts)
    assert (obs[:, :3] == SENT).all() and (obs[:, -5:] == SENT).all()
    # the two halves of the step (PARC_SIM_STEP_PRE beside the query, _POST behind it) == the single launch: observation,
    # contact blocks and episode flags bit for bit, the reward terms to fp32 rounding (two instantiations of the same
    # expressions; the compiler may contract their FMAs differently)
    single = {k: v.clone() for k, v in out.items()}
    whole_row = obs.clone()
    obs.fill_(SENT); out["reward"].fill_(SENT); out["done"].fill_(-9); out["joint_rot"].fill_(SENT)
    pre = ops.SimStepPlan(gpu_model.c_model(), simd, refd, kid, jw, dw, hfd, out, cfg=cfg, contact_body_ids=allowed, phase=1)
    post = ops.SimStepPlan(gpu_model.c_model(), simd, refd, kid, jw, dw, hfd, out, cfg=cfg, contact_body_ids=allowed, phase=2)
    pre.launch()
    assert (out["reward"] == SENT).all() and (out["done"] == -9).all() and (out["tar_contacts"] == SENT).all()
    assert torch.equal(out["char_obs"], single["char_obs"]) and torch.equal(out["joint_rot"], single["joint_rot"])
    post.launch()
    for k in single:
        if k == "reward":
            assert_close(out[k], single[k], rtol=1e-6, atol=1e-7, what="reward terms of the split step")
        else:
            assert torch.equal(out[k], single[k]), k
    assert torch.equal(obs, whole_row)


def test_query_accepts_any_leading_shape_and_int32_ids(golden_lib):
    """The reference indexes its tables with whatever shape / integer dtype the ids have; so does the mirror."""
    gen = torch.Generator().manual_seed(6)
    ids = torch.randint(0, 3, (4, 5), generator=gen).cuda()
    times = (torch.rand(4, 5, generator=gen) * 3.0).cuda()
    flat = golden_lib.calc_motion_frame(ids.reshape(-1), times.reshape(-1))
    shaped = golden_lib.calc_motion_frame(ids.to(torch.int32), times)
    assert shaped[0].shape == (4, 5, 3) and shaped[4].shape == (4, 5, 14, 4) and shaped[6].shape == (4, 5, 15)
    for a, b in zip(shaped, flat):
        assert torch.equal(a.reshape(b.shape), b)
    scalar = golden_lib.calc_motion_frame(torch.tensor(1, device="cuda"), torch.tensor(0.7, device="cuda"))
    assert scalar[0].shape == (3,) and scalar[4].shape == (14, 4)
    one = golden_lib.calc_motion_frame(torch.tensor([1], device="cuda"), torch.tensor([0.7], device="cuda"))
    assert torch.equal(scalar[1], one[1][0])
    g = golden_lib.get_motion_frame(ids, torch.randint(0, 30, (4, 5), generator=gen).cuda())
    assert g[5].shape == (4, 5, 28)
    i0, i1, bl = golden_lib._calc_frame_blend(ids, times)
    assert i0.shape == (4, 5) and i0.dtype == torch.int64 and bl.shape == (4, 5)


@pytest.mark.parametrize("n", [500, 6000])
def test_fused_obs_any_template_size(golden_lib, O, oracle_tables, n):
    """The fused sweep's pass structure (whole passes without bounds checks, a partial last pass, templates shorter
    than one pass or longer than the shared-memory staging) for template sizes around every boundary, in the
    one-wave instantiation (n = 500: 448-sample passes) and the large-batch one (n = 6000: 112-sample passes)."""
    from parc_b200 import ops
    t = _civ_terrain()
    gen = torch.Generator().manual_seed(77 + n)
    ids = torch.randint(0, 3, (n,), generator=gen)
    times = torch.rand(n, generator=gen) * oracle_tables.lengths[ids]
    ref = O.calc_motion_frame(oracle_tables, ids, times)
    ot = O.Terrain(hf=t.hf.cpu(), min_point=t.min_point.cpu(), dxdy=t.dxdy.cpu())
    heading = O.calc_heading(ref[1])
    for P in (1, 15, 16, 100, 112, 113, 200, 224, 441, 448, 449, 500, 900, 1100):
        tmpl = (torch.rand(P, 2, generator=gen) * 2 - 1) * 3.0
        out = golden_lib.calc_motion_frame_fk_obs(ids.cuda(), times.cuda(), hf_desc=t.hf_desc(), obs_tmpl=tmpl.cuda())
        want = O.ray_obs(ot, ref[0], heading, tmpl)
        got = out["obs"].cpu()
        assert got.shape == (n, P)
        mism = got != want
        # a mismatch is legitimate only where sin/cos rounding moved a sample across a cell border
        assert float(mism.float().mean()) < 2e-3, (P, float(mism.float().mean()))
        assert torch.isfinite(got).all() and (got.abs() <= 3.0).all()
