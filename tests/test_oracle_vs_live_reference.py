"""The oracle against the REAL reference, live, on seeded inputs that are NOT the golden ones.

Runs only where the reference tree is mounted (the authoring container: /root/reference); skipped elsewhere --
the GPU box has neither the tree nor any need for it (its tests use the committed golden vectors).  Every
comparison is torch.equal: the oracle restates the reference's algorithm in the reference's own arithmetic
(torch fp32 on CPU), so identical inputs must give identical bits.
"""
import os
import pickle

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import parc_oracle as O
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref():
    ref_shim.activate()
    import anim.kin_char_model as kcm
    import anim.motion_lib as mlib
    import envs.ig_char_env as ice
    import envs.ig_parkour.mgdm_dm_util as dm
    import util.terrain_util as tu
    import util.torch_util as tt
    km = kcm.KinCharModel("cpu")
    km.load_char_file(os.path.join(ref_shim.REFERENCE_ROOT, "data/assets/humanoid.xml"))
    return dict(kcm=kcm, mlib=mlib, ice=ice, dm=dm, tu=tu, tt=tt, km=km)


@pytest.fixture(scope="module")
def model():
    return O.CharModel.from_npz(os.path.join(GOLDEN, "humanoid_model.npz"))


def _quats(gen, *shape):
    q = torch.randn(*shape, 4, generator=gen)
    return q / q.norm(dim=-1, keepdim=True)


def test_quaternion_helpers_live(ref):
    tt = ref["tt"]
    g = torch.Generator().manual_seed(101)
    a, b = _quats(g, 500), _quats(g, 500)
    v = torch.randn(500, 3, generator=g)
    e = torch.randn(500, 3, generator=g) * torch.rand(500, 1, generator=g) * 3.5
    e[:5] = 0.0
    t = torch.rand(500, generator=g)
    b[:20] = a[:20]                                       # |cos| >= 1 branch
    b[20:40] = tt.quat_mul(a[20:40], tt.exp_map_to_quat(torch.randn(20, 3, generator=g) * 5e-4))   # sin < 1e-3 branch
    assert torch.equal(O.quat_mul(a, b), tt.quat_mul(a, b))
    assert torch.equal(O.quat_rotate(a, v), tt.quat_rotate(a, v))
    assert torch.equal(O.exp_map_to_quat(e), tt.exp_map_to_quat(e))
    assert torch.equal(O.quat_to_exp_map(a), tt.quat_to_exp_map(a))
    assert torch.equal(O.quat_diff_angle(a, b), tt.quat_diff_angle(a, b))
    assert torch.equal(O.slerp(a, b, t), tt.slerp(a, b, t))
    assert torch.equal(O.calc_heading(a), tt.calc_heading(a))
    assert torch.equal(O.heading_inverse_quat(a), tt.calc_heading_quat_inv(a))
    assert torch.equal(O.quat_to_tan_norm(a), tt.quat_to_tan_norm(a))
    ang = torch.randn(500, generator=g) * 4
    assert torch.equal(O.rotate_2d(v[:, :2], ang), tt.rotate_2d_vec(v[:, :2], ang))


def test_fk_and_dof_conversions_live(ref, model):
    km = ref["km"]
    g = torch.Generator().manual_seed(102)
    dof = (torch.rand(300, 28, generator=g) * 2 - 1) * 1.5
    dof[:3] = 0.0
    jr = km.dof_to_rot(dof)
    assert torch.equal(O.dof_to_rot(model, dof), jr)
    assert torch.equal(O.rot_to_dof(model, jr), km.rot_to_dof(jr))
    rp, rq = torch.randn(300, 3, generator=g), _quats(g, 300)
    bp, br = km.forward_kinematics(rp, rq, jr)
    obp, obr = O.forward_kinematics(model, rp, rq, jr)
    assert torch.equal(obp, bp) and torch.equal(obr, br)
    # arbitrary leading dims, as the reference accepts
    bp2, _ = km.forward_kinematics(rp.view(10, 30, 3), rq.view(10, 30, 4), jr.view(10, 30, 14, 4))
    assert torch.equal(O.forward_kinematics(model, rp.view(10, 30, 3), rq.view(10, 30, 4), jr.view(10, 30, 14, 4))[0], bp2)


def test_motion_lib_tables_and_queries_live(ref, model, tmp_path):
    rng = np.random.default_rng(103)
    clips, lines = [], ["motions:"]
    for c in range(5):
        n = int(rng.integers(2, 90))
        fps = float(rng.choice([24.0, 30.0, 60.0]))
        fr = np.zeros((n, 34), np.float32)
        t = np.arange(n)[:, None] / fps
        fr[:, 0:3] = np.cumsum(rng.normal(scale=0.03, size=(n, 3)), axis=0) + rng.uniform(-5, 5, 3)
        fr[:, 3:6] = 0.4 * np.sin(t * rng.uniform(0.5, 2, 3) + rng.uniform(0, 6, 3)) + 0.05
        fr[:, 6:] = 0.7 * np.sin(t * rng.uniform(0.5, 3, 28) + rng.uniform(0, 6, 28)) + 0.01
        ct = (rng.uniform(size=(n, 15)) < 0.4).astype(np.float32)
        loop = "WRAP" if c % 2 else "CLAMP"
        w = float(rng.uniform(0.2, 2.0))
        path = str(tmp_path / f"c{c}.pkl")
        with open(path, "wb") as f:
            pickle.dump({"frames": fr, "contacts": ct, "fps": fps, "loop_mode": loop}, f)
        lines += [f"- file: {path}", f"  weight: {w}"]
        clips.append(O.Clip(fr, ct, fps, O.WRAP if loop == "WRAP" else O.CLAMP, w))
    y = str(tmp_path / "lib.yaml")
    open(y, "w").write("\n".join(lines) + "\n")
    lib = ref["mlib"].MotionLib(y, ref["km"], "cpu", init_type="motion_file", contact_info=True)
    tb = O.build_tables(model, clips)
    for a, b in (("_frame_root_pos", "root_pos"), ("_frame_root_rot", "root_rot"), ("_frame_joint_rot", "joint_rot"),
                 ("_frame_root_vel", "root_vel"), ("_frame_root_ang_vel", "root_ang_vel"), ("_frame_dof_vel", "dof_vel"),
                 ("_frame_contacts", "contacts"), ("_motion_lengths", "lengths"), ("_motion_start_idx", "start_idx"),
                 ("_motion_root_pos_delta", "root_pos_delta"), ("_motion_weights", "weights")):
        assert torch.equal(getattr(lib, a), getattr(tb, b)), a
    g = torch.Generator().manual_seed(104)
    ids = torch.randint(0, 5, (3000,), generator=g)
    times = (torch.rand(3000, generator=g) * 5 - 1.5) * tb.lengths[ids]
    times[:50] = (torch.arange(50) / 30.0)
    i0, i1, bl = lib._calc_frame_blend(ids, times)
    o0, o1, ob = O.frame_blend(tb, ids, times)
    assert torch.equal(i0, o0) and torch.equal(i1, o1) and torch.equal(bl, ob)
    for a, b in zip(lib.calc_motion_frame(ids, times), O.calc_motion_frame(tb, ids, times)):
        assert torch.equal(a, b)
    fi = torch.minimum(torch.randint(0, 90, (3000,), generator=g), tb.num_frames[ids] - 1)
    for a, b in zip(lib.get_motion_frame(ids, fi), O.get_motion_frame(tb, ids, fi)):
        assert torch.equal(a, b)
    assert torch.equal(lib.calc_motion_phase(ids, times), O.motion_phase(tb, ids, times))


def test_heightfield_and_sdf_live(ref):
    tu = ref["tu"]
    g = torch.Generator().manual_seed(105)
    t = tu.SubTerrain("t", 23, 17, 0.4, 0.25, -1.3, 2.1, device="cpu")
    t.hf[...] = torch.rand(23, 17, generator=g) * 2 - 0.5
    ot = O.Terrain(hf=t.hf.clone(), min_point=t.min_point.clone(), dxdy=t.dxdy.clone())
    xy = torch.randn(4000, 2, generator=g) * 6
    xy[:8] = torch.tensor([[-1.3 + 0.2, 2.1], [-1.3 + 0.6, 2.1 + 0.125], [1e9, -1e9], [float("nan"), 0.0],
                           [float("inf"), 0.0], [-1.3 - 0.2, 2.1 - 0.125], [7.5, 6.1], [7.7, 6.225]])
    assert torch.equal(O.grid_index(ot, xy), t.get_grid_index(xy))
    assert torch.equal(O.hf_sample(ot, xy), tu.get_local_hf_from_terrain(xy, t))
    pts = torch.randn(3, 200, 3, generator=g) * torch.tensor([4.0, 3.0, 1.5]) + torch.tensor([2.0, 4.0, 0.5])
    hfs = torch.rand(3, 9, 7, generator=g) * 1.5
    mc = torch.randn(3, 2, generator=g)
    dxdy = torch.tensor([0.4, 0.4])
    for inv in (True, False):
        assert torch.equal(O.points_hf_sdf(pts, hfs, mc, dxdy, base_z=-10.0, inverted=inv),
                           tu.points_hf_sdf(pts, hfs, mc, dxdy, base_z=-10.0, inverted=inv))


def test_step_assembly_live(ref, model):
    ice, dm = ref["ice"], ref["dm"]
    g = torch.Generator().manual_seed(106)
    n, S = 200, 3
    st = lambda: [torch.randn(n, 3, generator=g), _quats(g, n), torch.randn(n, 3, generator=g),
                  torch.randn(n, 3, generator=g), _quats(g, n, 14), torch.randn(n, 28, generator=g)]
    a, b = st(), st()
    ka, kb = torch.randn(n, 4, 3, generator=g), torch.randn(n, 4, 3, generator=g)
    for gl in (False, True):
        assert torch.equal(O.compute_char_obs(*a, ka, gl, not gl), ice.compute_char_obs(*a, ka, gl, not gl))
        tp, tr, tj, tk = torch.randn(n, S, 3, generator=g), _quats(g, n, S), _quats(g, n, S, 14), torch.randn(n, S, 4, 3, generator=g)
        assert torch.equal(O.compute_tar_obs(a[0], a[1], tp, tr, tj, tk, gl, gl),
                           dm.compute_tar_obs(a[0], a[1], tp.clone(), tr.clone(), tj, tk.clone(), gl, gl))
    jw, dw = torch.rand(14, generator=g), torch.rand(28, generator=g)
    for tr_ in (True, False):
        assert torch.equal(O.compute_deepmimic_reward(*a, ka, *b, kb, jw, dw, not tr_, tr_),
                           dm.compute_deepmimic_reward(*a, ka, *b, kb, jw, dw, not tr_, tr_))
    bp, tbp = torch.randn(n, 15, 3, generator=g) * 0.3, torch.randn(n, 15, 3, generator=g) * 0.3
    tm = torch.rand(n, generator=g) * 12
    cf = torch.randn(n, 15, 3, generator=g) * (torch.rand(n, 15, 1, generator=g) < 0.04)
    th = torch.randn(n, 15, generator=g) * 0.3 - 0.5
    ptd = torch.rand(14, generator=g) * 0.6 + 1.2
    tm[:20] = 0.0                                          # first-step envs never fail
    cids = torch.tensor([3, 9])
    done_in = torch.zeros(n, dtype=torch.int)
    seen = set()
    for pose in (True, False):
        for track in (True, False):
            want = dm.compute_done(done_in, tm, 10.0, a[1], bp, a[0], b[1], tbp, cf, cids, th, pose, ptd, False, True, track,
                                   1.0, 2.8)
            got = O.compute_done(done_in, tm, 10.0, a[1], bp, b[1], tbp, cf, cids, th, pose, ptd, True, track, 1.0, 2.8)
            assert torch.equal(got, want)
            seen |= set(want.tolist())
    assert seen == {0, 1, 3}


def test_mirror_host_helpers_equal_the_live_reference(ref):
    """The two module-level helpers SURVEY 8(a) names next to the kernels -- `motion_lib.calc_phase` (:527-538) and
    `terrain_util.points_boxes_sdf` (:1777-1804) -- are plain torch in the mirror too: same bits as the reference."""
    from parc_b200.anim import motion_lib as mirror_mlib
    from parc_b200.util import terrain_util as mirror_tu
    gen = torch.Generator().manual_seed(91)
    t = torch.rand(4000, generator=gen) * 30.0 - 4.0
    length = torch.rand(4000, generator=gen) * 8.0 + 0.3
    loop = torch.randint(0, 2, (4000,), generator=gen)
    assert torch.equal(mirror_mlib.calc_phase(t, length, loop), ref["mlib"].calc_phase(t.clone(), length, loop))
    p, c = torch.randn(3, 40, 3, generator=gen), torch.randn(3, 17, 3, generator=gen)
    h = torch.rand(3, 17, 3, generator=gen) + 0.05
    assert torch.equal(mirror_tu.points_boxes_sdf(p, c, h), ref["tu"].points_boxes_sdf(p, c, h))
    assert torch.equal(mirror_tu.points_boxes_sdf(p[0], c[0], h[0]), ref["tu"].points_boxes_sdf(p[0], c[0], h[0]))
