"""CPU oracle for PARC's batched kinematic motion-query path.

TEST INFRASTRUCTURE ONLY -- see `oracle/__init__.py`.  Never imported by the
product (`parc_b200/`).

What this is: a restatement, on `torch` CPU tensors in fp32, of the reference
algorithm for  MotionLib frame query -> KinCharModel FK -> nearest-cell
heightfield observation -> body-point penetration / contact loss.  Every
function cites the reference `file:line` it follows (paths relative to the
reference root).  The op ORDER of the reference is kept wherever rounding can
change a discrete decision (frame index, grid index, slerp branch), so that
on CPU this oracle is bit-identical to the imported reference; that is checked
by `oracle/make_golden.py` (authoring container) and re-checked against the
committed vectors in `tests/golden/` by `tests/test_oracle_golden.py`.

Parity status: the reference ships NO tests or golden vectors for this path
(SURVEY.md section 4), so the pin is "outputs of the reference itself run in the
authoring container" (fixtures + generating script committed).

Conventions: quaternions are xyzw, z is up, `hf[ix, iy]` is x-major and cell
(ix, iy) is centred on `min_point + (ix, iy) * dxdy`.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import torch

F32 = torch.float32

# joint type codes -- anim/kin_char_model.py:11-15
ROOT, HINGE, SPHERICAL, FIXED = 0, 1, 2, 3
# loop modes -- anim/motion_lib.py:11-13
CLAMP, WRAP = 0, 1


# --------------------------------------------------------------------------
# quaternion / exp-map primitives (util/torch_util.py)
# --------------------------------------------------------------------------
def vec_normalize(x, eps: float = 1e-9):
    """util/torch_util.py:9-12"""
    return x / x.norm(p=2, dim=-1).clamp(min=eps).unsqueeze(-1)


def wrap_angle(x):
    """util/torch_util.py:4-7 (normalize_angle)"""
    return torch.atan2(torch.sin(x), torch.cos(x))


def quat_mul(a, b):
    """Hamilton product in the reference's 8-multiply form -- util/torch_util.py:40-58."""
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    t_ww = (az + ax) * (bx + by)
    t_yy = (aw - ay) * (bw + bz)
    t_zz = (aw + ay) * (bw - bz)
    t_xx = t_ww + t_yy + t_zz
    half = 0.5 * (t_xx + (az - ax) * (bx - by))
    w = half - t_ww + (az - ay) * (by - bz)
    x = half - t_xx + (ax + aw) * (bx + bw)
    y = half - t_yy + (aw - ax) * (by + bz)
    z = half - t_zz + (az + ay) * (bw - bx)
    return torch.stack([x, y, z, w], dim=-1)


def quat_rotate(q, v):
    """v + w*t + q_v x t, t = 2 q_v x v -- util/torch_util.py:60-66."""
    qv = q[..., :3]
    qw = q[..., 3:]
    t = 2 * torch.cross(qv, v, dim=-1)
    return v + qw * t + torch.cross(qv, t, dim=-1)


def quat_conj(q):
    """util/torch_util.py:29-31"""
    return torch.cat([-q[..., :3], q[..., 3:]], dim=-1)


def quat_w_positive(q):
    """util/torch_util.py:33-38 (quat_pos)"""
    neg = (q[..., 3:] < 0).float()
    return (1 - 2 * neg) * q


def quat_delta(q0, q1):
    """q1 * conj(q0) -- util/torch_util.py:422-425 (quat_diff)."""
    return quat_mul(q1, quat_conj(q0))


def axis_angle_to_quat(axis, angle):
    """util/torch_util.py:311-317"""
    half = (angle / 2).unsqueeze(-1)
    xyz = vec_normalize(axis) * half.sin()
    return vec_normalize(torch.cat([xyz, half.cos()], dim=-1))


def exp_map_to_axis_angle(e):
    """util/torch_util.py:394-412.  NB: gradient at e == 0 is NaN in the reference too."""
    angle = torch.norm(e, dim=-1)
    axis = e / angle.unsqueeze(-1)
    angle = wrap_angle(angle)
    fallback = torch.zeros_like(e)
    fallback[..., -1] = 1
    keep = torch.abs(angle) > 1e-5
    angle = torch.where(keep, angle, torch.zeros_like(angle))
    axis = torch.where(keep.unsqueeze(-1), axis, fallback)
    return axis, angle


def exp_map_to_quat(e):
    """util/torch_util.py:414-419"""
    axis, angle = exp_map_to_axis_angle(e)
    return axis_angle_to_quat(axis, angle)


def quat_to_axis_angle(q):
    """util/torch_util.py:68-88"""
    q = quat_w_positive(q)
    length = torch.norm(q[..., 0:3], dim=-1, p=2)
    angle = 2.0 * torch.atan2(length, q[..., 3])
    axis = q[..., 0:3] / length.unsqueeze(-1)
    fallback = torch.zeros_like(axis)
    fallback[..., -1] = 1
    keep = length > 1e-5
    angle = torch.where(keep, angle, torch.zeros_like(angle))
    axis = torch.where(keep.unsqueeze(-1), axis, fallback)
    return axis, angle


def quat_to_exp_map(q):
    """util/torch_util.py:346-351 (+ :329-334)"""
    axis, angle = quat_to_axis_angle(q)
    return angle.unsqueeze(-1) * axis


def quat_diff_angle(q0, q1):
    """util/torch_util.py:427-431"""
    return quat_to_axis_angle(quat_delta(q0, q1))[1]


def slerp(q0, q1, t):
    """util/torch_util.py:443-468.  Not renormalised; the two `where`s are ordered."""
    c = torch.sum(q0 * q1, dim=-1)
    q1 = torch.where((c < 0).unsqueeze(-1), -q1, q1)
    c = torch.abs(c).unsqueeze(-1)
    theta = torch.acos(c)
    s = torch.sqrt(1.0 - c * c)
    if t.dim() == q0.dim() - 1:
        t = t.unsqueeze(-1)
    ra = torch.sin((1 - t) * theta) / s
    rb = torch.sin(t * theta) / s
    out = ra * q0 + rb * q1
    out = torch.where(torch.abs(s) < 0.001, 0.5 * q0 + 0.5 * q1, out)
    out = torch.where(torch.abs(c) >= 1, q0, out)
    return out


def calc_heading(q):
    """atan2 of the rotated x axis -- util/torch_util.py:470-479."""
    ex = torch.zeros_like(q[..., 0:3])
    ex[..., 0] = 1
    d = quat_rotate(q, ex)
    return torch.atan2(d[..., 1], d[..., 0])


def rotate_2d(vec, angle):
    """util/torch_util.py:619-631"""
    x, y = vec[..., 0], vec[..., 1]
    c, s = torch.cos(angle), torch.sin(angle)
    return torch.stack([x * c - y * s, x * s + y * c], dim=-1)


# --------------------------------------------------------------------------
# character model (anim/kin_char_model.py)
# --------------------------------------------------------------------------
@dataclass
class CharModel:
    """Plain-array view of KinCharModel (anim/kin_char_model.py:147-178)."""
    body_names: List[str]
    parents: List[int]                 # [J], -1 for the root
    local_trans: torch.Tensor          # [J,3]
    local_rot: torch.Tensor            # [J,4] xyzw
    joint_type: List[int]              # [J]
    joint_axis: torch.Tensor           # [J,3] (zeros unless HINGE)
    dof_idx: List[int]                 # [J] start of this joint's DoFs
    dof_dim: List[int]                 # [J]
    body_points: List[torch.Tensor] = field(default_factory=list)  # 15 x [P_b,3]

    @property
    def num_bodies(self):
        return len(self.parents)

    @property
    def dof_size(self):
        return int(sum(self.dof_dim))

    @staticmethod
    def from_npz(path):
        z = np.load(path, allow_pickle=False)
        counts = z["body_point_counts"].tolist()
        pts = torch.from_numpy(z["body_points"].astype(np.float32))
        split, o = [], 0
        for c in counts:
            split.append(pts[o:o + c].clone())
            o += c
        return CharModel(
            body_names=[str(s) for s in z["body_names"].tolist()],
            parents=z["parents"].astype(np.int64).tolist(),
            local_trans=torch.from_numpy(z["local_translation"].astype(np.float32)),
            local_rot=torch.from_numpy(z["local_rotation"].astype(np.float32)),
            joint_type=z["joint_type"].astype(np.int64).tolist(),
            joint_axis=torch.from_numpy(z["joint_axis"].astype(np.float32)),
            dof_idx=z["dof_idx"].astype(np.int64).tolist(),
            dof_dim=z["dof_dim"].astype(np.int64).tolist(),
            body_points=split,
        )


def joint_dof_to_rot(model: CharModel, j: int, jd):
    """Joint.dof_to_rot -- anim/kin_char_model.py:57-77."""
    shape = list(jd.shape[:-1]) + [4]
    rot = torch.zeros(shape, dtype=jd.dtype)
    jt = model.joint_type[j]
    if jt == HINGE:
        axis = torch.broadcast_to(model.joint_axis[j], rot[..., 0:3].shape)
        rot[:] = axis_angle_to_quat(axis, jd.squeeze(-1))
    elif jt == SPHERICAL:
        rot[:] = exp_map_to_quat(jd)
    else:  # ROOT / FIXED
        rot[..., -1] = 1
    return rot


def dof_to_rot(model: CharModel, dof):
    """KinCharModel.dof_to_rot -- anim/kin_char_model.py:478-491.  [...,D] -> [...,J-1,4]"""
    J = model.num_bodies
    out = torch.zeros(list(dof.shape[:-1]) + [J - 1, 4], dtype=dof.dtype)
    for j in range(1, J):
        jd = dof[..., model.dof_idx[j]:model.dof_idx[j] + model.dof_dim[j]]
        out[..., j - 1, :] = joint_dof_to_rot(model, j, jd)
    return out


def rot_to_dof(model: CharModel, rot):
    """KinCharModel.rot_to_dof -- anim/kin_char_model.py:493-507 (+ Joint.rot_to_dof :79-100)."""
    J = model.num_bodies
    dof = torch.zeros(list(rot.shape[:-2]) + [model.dof_size], dtype=rot.dtype)
    for j in range(1, J):
        d = model.dof_dim[j]
        if d == 0:
            continue
        jr = rot[..., j - 1, :]
        if model.joint_type[j] == HINGE:
            axis, angle = quat_to_axis_angle(jr)
            flip = torch.sum(model.joint_axis[j] * axis, dim=-1) < 0
            angle = torch.where(flip, -angle, angle)
            jd = angle.unsqueeze(-1)
        else:
            jd = quat_to_exp_map(jr)
        dof[..., model.dof_idx[j]:model.dof_idx[j] + d] = jd
    return dof


def forward_kinematics(model: CharModel, root_pos, root_rot, joint_rot):
    """KinCharModel.forward_kinematics -- anim/kin_char_model.py:509-541."""
    J = model.num_bodies
    pos = [None] * J
    rot = [None] * J
    pos[0], rot[0] = root_pos, root_rot
    for j in range(1, J):
        p = model.parents[j]
        lt = torch.broadcast_to(model.local_trans[j], pos[p].shape)
        lr = torch.broadcast_to(model.local_rot[j], rot[p].shape)
        pos[j] = pos[p] + quat_rotate(rot[p], lt)
        rot[j] = quat_mul(rot[p], quat_mul(lr, joint_rot[..., j - 1, :]))
    return torch.stack(pos, dim=-2), torch.stack(rot, dim=-2)


def dof_velocity(model: CharModel, jr0, jr1, dt):
    """KinCharModel.compute_dof_vel -- anim/kin_char_model.py:552-581."""
    out = torch.zeros(list(jr0.shape[:-2]) + [model.dof_size], dtype=jr0.dtype)
    d = quat_mul(quat_conj(jr0), jr1)
    d = vec_normalize(quat_w_positive(d))          # quat_normalize, util/torch_util.py:438-441
    for j in range(1, model.num_bodies):
        jt = model.joint_type[j]
        if jt not in (HINGE, SPHERICAL):
            continue
        v = quat_to_exp_map(d[..., j - 1, :]) / dt
        if jt == HINGE:
            v = torch.sum(model.joint_axis[j] * v, dim=-1, keepdim=True)
        out[..., model.dof_idx[j]:model.dof_idx[j] + model.dof_dim[j]] = v
    return out


def frame_dof_velocity(model: CharModel, joint_rot, dt):
    """KinCharModel.compute_frame_dof_vel -- anim/kin_char_model.py:543-550."""
    v = dof_velocity(model, joint_rot[..., :-1, :, :], joint_rot[..., 1:, :, :], dt)
    return torch.cat([v, v[..., -1:, :]], dim=-2)


# --------------------------------------------------------------------------
# body surface samples (util/geom_util.py:725-870)
# --------------------------------------------------------------------------
def icosahedron_vertices(radius):
    """Stand-in for trimesh.creation.icosphere(subdivisions=0) -- util/geom_util.py:741-749.
    trimesh is absent here; see oracle/ref_shim.py for the vertex order used."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array(
        [[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0],
         [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
         [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return torch.from_numpy(v * radius).to(F32)


# --------------------------------------------------------------------------
# MotionLib tables + query (anim/motion_lib.py)
# --------------------------------------------------------------------------
@dataclass
class Clip:
    frames: np.ndarray                 # [F, 6+D] root pos, root exp-map, joint dofs
    contacts: Optional[np.ndarray]     # [F, J] or None
    fps: float = 30.0
    loop_mode: int = CLAMP
    weight: float = 1.0


@dataclass
class FrameTables:
    """The flat per-frame tables of MotionLib -- anim/motion_lib.py:349-375."""
    root_pos: torch.Tensor
    root_rot: torch.Tensor
    joint_rot: torch.Tensor
    root_vel: torch.Tensor
    root_ang_vel: torch.Tensor
    dof_vel: torch.Tensor
    contacts: torch.Tensor
    frames: torch.Tensor
    num_frames: torch.Tensor           # i64 [M]
    start_idx: torch.Tensor            # i64 [M]
    lengths: torch.Tensor              # f32 [M]
    loop_modes: torch.Tensor           # i32 [M]
    root_pos_delta: torch.Tensor       # f32 [M,3]
    weights: torch.Tensor              # f32 [M]
    fps: torch.Tensor
    dt: torch.Tensor


def extract_pose(model: CharModel, frames):
    """MotionLib._extract_frame_data -- anim/motion_lib.py:405-423."""
    fr = torch.as_tensor(frames, dtype=F32)
    root_pos = fr[..., 0:3].clone()
    root_rot = exp_map_to_quat(fr[..., 3:6].clone())
    joint_rot = quat_w_positive(dof_to_rot(model, fr[..., 6:].clone()))
    return root_pos, root_rot, joint_rot


def build_tables(model: CharModel, clips: Sequence[Clip]) -> FrameTables:
    """MotionLib._load_motions minus file I/O -- anim/motion_lib.py:204-380."""
    acc = {k: [] for k in ("root_pos", "root_rot", "joint_rot", "root_vel", "root_ang_vel",
                           "dof_vel", "contacts", "frames", "delta")}
    nfr, lens, loops, wts, fpss, dts = [], [], [], [], [], []
    for c in clips:
        fps = c.fps
        dt = 1.0 / fps
        n = c.frames.shape[0]
        root_pos, root_rot, joint_rot = extract_pose(model, c.frames)
        delta = root_pos[-1] - root_pos[0]
        delta[..., -1] = 0.0
        root_vel = torch.zeros_like(root_pos)
        root_vel[:-1] = fps * (root_pos[1:] - root_pos[:-1])
        root_vel[-1] = root_vel[-2]
        root_ang_vel = torch.zeros_like(root_pos)
        root_ang_vel[:-1] = fps * quat_to_exp_map(quat_delta(root_rot[:-1], root_rot[1:]))
        root_ang_vel[-1] = root_ang_vel[-2]
        acc["root_pos"].append(root_pos)
        acc["root_rot"].append(root_rot)
        acc["joint_rot"].append(joint_rot)
        acc["root_vel"].append(root_vel)
        acc["root_ang_vel"].append(root_ang_vel)
        acc["dof_vel"].append(frame_dof_velocity(model, joint_rot, dt))
        acc["delta"].append(delta)
        acc["frames"].append(torch.as_tensor(c.frames, dtype=F32))
        if c.contacts is None:
            acc["contacts"].append(torch.zeros(n, model.num_bodies, dtype=F32))
        else:
            acc["contacts"].append(torch.as_tensor(c.contacts, dtype=F32))
        nfr.append(n)
        lens.append(1.0 / fps * (n - 1))      # python double, cast once below (:275, :355)
        loops.append(c.loop_mode)
        wts.append(c.weight)
        fpss.append(fps)
        dts.append(dt)
    num_frames = torch.tensor(nfr, dtype=torch.long)
    shifted = num_frames.roll(1)
    shifted[0] = 0
    w = torch.tensor(wts, dtype=F32)
    w = w / w.sum()
    return FrameTables(
        root_pos=torch.cat(acc["root_pos"]), root_rot=torch.cat(acc["root_rot"]),
        joint_rot=torch.cat(acc["joint_rot"]), root_vel=torch.cat(acc["root_vel"]),
        root_ang_vel=torch.cat(acc["root_ang_vel"]), dof_vel=torch.cat(acc["dof_vel"]),
        contacts=torch.cat(acc["contacts"]), frames=torch.cat(acc["frames"]),
        num_frames=num_frames, start_idx=shifted.cumsum(0),
        lengths=torch.tensor(lens, dtype=F32), loop_modes=torch.tensor(loops, dtype=torch.int),
        root_pos_delta=torch.stack(acc["delta"]), weights=w,
        fps=torch.tensor(fpss, dtype=F32), dt=torch.tensor(dts, dtype=F32))


def motion_phase(tb: FrameTables, ids, times):
    """calc_phase -- anim/motion_lib.py:527-538 (via :74-78)."""
    phase = times / tb.lengths[ids]
    wrap = tb.loop_modes[ids] == WRAP
    phase = torch.where(wrap, phase - torch.floor(phase), phase)
    return torch.clip(phase, 0.0, 1.0)


def frame_blend(tb: FrameTables, ids, times):
    """MotionLib._calc_frame_blend -- anim/motion_lib.py:443-456."""
    n = tb.num_frames[ids]
    phase = motion_phase(tb, ids, times)
    i0 = (phase * (n - 1)).long()
    i1 = torch.min(i0 + 1, n - 1)
    blend = phase * (n - 1) - i0
    start = tb.start_idx[ids]
    return i0 + start, i1 + start, blend


def loop_offset(tb: FrameTables, ids, times):
    """MotionLib._calc_loop_offset -- anim/motion_lib.py:458-475."""
    wrap = (tb.loop_modes[ids] == WRAP).unsqueeze(-1)
    cycles = torch.floor(times / tb.lengths[ids]).unsqueeze(-1)
    return torch.where(wrap, cycles * tb.root_pos_delta[ids], torch.zeros(ids.shape[0], 3))


def calc_motion_frame(tb: FrameTables, ids, times):
    """MotionLib.calc_motion_frame -- anim/motion_lib.py:80-112.
    Returns (root_pos, root_rot, root_vel, root_ang_vel, joint_rot, dof_vel, contacts)."""
    i0, i1, blend = frame_blend(tb, ids, times)
    b = blend.unsqueeze(-1)
    root_pos = (1.0 - b) * tb.root_pos[i0] + b * tb.root_pos[i1]
    root_rot = slerp(tb.root_rot[i0], tb.root_rot[i1], blend)
    joint_rot = slerp(tb.joint_rot[i0], tb.joint_rot[i1], b)
    root_pos = root_pos + loop_offset(tb, ids, times)
    contacts = (1.0 - b) * tb.contacts[i0] + b * tb.contacts[i1]
    return (root_pos, root_rot, tb.root_vel[i0], tb.root_ang_vel[i0], joint_rot,
            tb.dof_vel[i0], contacts)


def get_motion_frame(tb: FrameTables, ids, frame_idxs):
    """MotionLib.get_motion_frame -- anim/motion_lib.py:114-131."""
    i = tb.start_idx[ids] + frame_idxs
    return (tb.root_pos[i], tb.root_rot[i], tb.root_vel[i], tb.root_ang_vel[i],
            tb.joint_rot[i], tb.dof_vel[i], tb.contacts[i])


# --------------------------------------------------------------------------
# heightfield sampling (util/terrain_util.py, util/geom_util.py, envs/ig_parkour/mgdm_dm_util.py)
# --------------------------------------------------------------------------
@dataclass
class Terrain:
    """SubTerrain's sampled fields -- util/terrain_util.py:21-39."""
    hf: torch.Tensor            # [X,Y] f32
    min_point: torch.Tensor     # [2] f32
    dxdy: torch.Tensor          # [2] f32

    @property
    def dims(self):
        return torch.tensor(list(self.hf.shape), dtype=torch.int64)


def grid_coord(t: Terrain, xy):
    """(p - min) / dxdy before rounding (exposed so tests can identify cell-border cases)."""
    return (xy - t.min_point) / t.dxdy


def grid_index(t: Terrain, xy):
    """SubTerrain.get_grid_index -- util/terrain_util.py:113-126 (round half-to-even, clamp)."""
    idx = torch.round(grid_coord(t, xy)).to(torch.int64)
    return torch.clamp(idx, torch.zeros_like(t.dims), t.dims - 1)


def hf_sample(t: Terrain, xy):
    """get_local_hf_from_terrain -- util/terrain_util.py:1329-1346 (== :128-130)."""
    g = grid_index(t, xy)
    return t.hf[g[..., 0], g[..., 1]]


def cone_template(dx, num_neg, num_pos, rays_neg, rays_pos, ray_angle):
    """get_xy_points_cone -- util/geom_util.py:249-270.  -> [(rays)*(pts), 2], ray-major."""
    xs = torch.linspace(-dx * num_neg, dx * num_pos, num_neg + num_pos + 1, dtype=F32)
    base = torch.stack([xs, torch.zeros_like(xs)], dim=-1)
    rays = []
    for i in range(rays_neg + 1 + rays_pos):
        ang = torch.ones(base.shape[0], dtype=F32) * (-ray_angle * (rays_neg - i))
        rays.append(rotate_2d(base, ang))
    return torch.cat(rays, dim=0)


def grid_template(dx, dy, nx_neg, nx_pos, ny_neg, ny_pos):
    """get_xy_grid_points centred on 0 -- util/geom_util.py:210-221.  -> [X,Y,2]."""
    c = torch.zeros(2, dtype=F32)
    xs = torch.linspace(c[0] - dx * nx_neg, c[0] + dx * nx_pos, nx_neg + nx_pos + 1)
    ys = torch.linspace(c[1] - dy * ny_neg, c[1] + dy * ny_pos, ny_neg + ny_pos + 1)
    gx, gy = torch.meshgrid(xs, ys, indexing="ij")
    return torch.stack([gx, gy], dim=-1)


def ray_obs_points(root_pos, heading, tmpl):
    """World xy of every template point -- envs/ig_parkour/mgdm_dm_util.py:163-167."""
    n, p = root_pos.shape[0], tmpl.shape[0]
    h = heading.unsqueeze(-1).expand(-1, p)
    return rotate_2d(tmpl.unsqueeze(0).expand(n, -1, -1), h) + root_pos[..., 0:2].unsqueeze(1)


def ray_obs(t: Terrain, root_pos, heading, tmpl, min_h=-3.0, max_h=3.0):
    """RefCharEnv._refresh_ray_obs_hfs -- envs/ig_parkour/mgdm_dm_util.py:158-179."""
    xy = ray_obs_points(root_pos, heading, tmpl)
    z = hf_sample(t, xy.reshape(-1, 2)).view(root_pos.shape[0], tmpl.shape[0])
    return torch.clamp(z - root_pos[..., 2].unsqueeze(-1), min=min_h, max=max_h)


def grid_obs(t: Terrain, center_xy, heading, tmpl):
    """sample_hf_z_on_terrain -- util/terrain_util.py:2049-2082.  tmpl [X,Y,2] -> [B,X,Y]."""
    c = center_xy.unsqueeze(1).unsqueeze(1)
    h = heading.unsqueeze(1).unsqueeze(1)
    return hf_sample(t, rotate_2d(tmpl, h) + c)


# --------------------------------------------------------------------------
# point <-> heightfield SDF and the body-point losses
# --------------------------------------------------------------------------
def sd_box(p, half):
    """util/geom_util.py:122-143"""
    q = torch.abs(p) - half
    outside = torch.norm(torch.clamp(q, min=0.0), dim=-1)
    inside = torch.clamp(torch.max(q, dim=-1)[0], max=0.0)
    return outside + inside


def hf_cell_boxes(hf, min_center, dxdy, base_z, inverted):
    """Box centres / half extents of every cell -- util/terrain_util.py:1855-1881.
    hf [B,X,Y], min_center [B,2] -> centres [B,M,3], halfdims [B,M,3]."""
    B, X, Y = hf.shape
    xs = torch.linspace(0.0, (X - 1.0) * dxdy[0].item(), X)
    ys = torch.linspace(0.0, (Y - 1.0) * dxdy[1].item(), Y)
    gx, gy = torch.meshgrid(xs, ys, indexing="ij")
    gx = gx.unsqueeze(0) + min_center[..., 0].unsqueeze(1).unsqueeze(1)
    gy = gy.unsqueeze(0) + min_center[..., 1].unsqueeze(1).unsqueeze(1)
    if inverted:
        top = -base_z
        cz, hz = (hf + top) / 2.0, (top - hf) / 2.0
    else:
        cz, hz = (hf + base_z) / 2.0, (hf - base_z) / 2.0
    centres = torch.stack([gx, gy, cz], dim=-1).view(B, X * Y, 3)
    hxy = (dxdy / 2.0).view(1, 1, 2).expand(B, X * Y, 2)
    half = torch.cat([hxy, hz.reshape(B, X * Y, 1)], dim=-1)
    return centres, half


def points_hf_sdf(points, hf, min_center, dxdy, base_z=-10.0, inverted=True, chunk=256):
    """util/terrain_util.py:1835-1893 (+ points_boxes_sdf :1777-1804).  points [B,N,3] -> [B,N].
    Evaluated in chunks over N purely to bound memory; min over ALL cells as in the reference."""
    centres, half = hf_cell_boxes(hf, min_center, dxdy, base_z, inverted)
    outs = []
    for s in range(0, points.shape[1], chunk):
        p = points[:, s:s + chunk]
        rel = p.unsqueeze(2) - centres.unsqueeze(1)
        sd = sd_box(rel, half.unsqueeze(1).expand_as(rel))
        outs.append(torch.min(sd, dim=-1)[0])
    out = torch.cat(outs, dim=1)
    return out * -1.0 if inverted else out


def body_world_points(body_pos, body_rot, pts_b, b):
    """quat_rotate(body_rot, local) + body_pos -- tools/procgen/mdm_path.py:80-86."""
    r = body_rot[..., b, :].unsqueeze(-2)
    p = body_pos[..., b, :].unsqueeze(-2)
    shape = [1] * (r.dim() - 2) + list(pts_b.shape)
    return quat_rotate(r, pts_b.view(shape)) + p


def pen_contact_terms(body_pos, body_rot, contacts, body_points, hf, min_point, dxdy, base_z):
    """Shared core of compute_motion_loss / motion_terrain_contact_loss.
    body_pos [B,F,J,3], body_rot [B,F,J,4], contacts [B,F,J]; one terrain.  -> (pen[B], contact[B])
    tools/procgen/mdm_path.py:79-110 ; tools/motion_opt/motion_optimization.py:241-272."""
    B, F = body_pos.shape[0], body_pos.shape[1]
    hfb = hf.unsqueeze(0).expand(B, -1, -1)
    mpb = min_point.unsqueeze(0).expand(B, -1)
    pen = torch.zeros(B, dtype=F32)
    con = torch.zeros(B, dtype=F32)
    for b in range(body_pos.shape[2]):
        P = body_points[b].shape[0]
        wp = body_world_points(body_pos, body_rot, body_points[b], b).reshape(B, -1, 3)
        neg = torch.clamp(points_hf_sdf(wp, hfb, mpb, dxdy, base_z=base_z, inverted=True), max=0.0)
        pen = pen + torch.sum(-neg, dim=-1)
        posd = torch.clamp(points_hf_sdf(wp, hfb, mpb, dxdy, base_z=base_z, inverted=False), min=0.0)
        closest = torch.min(posd.view(B, F, P), dim=-1)[0]
        con = con + torch.sum(closest * contacts[..., b], dim=-1)
    return pen, con


def compute_motion_loss(model: CharModel, root_pos, root_rot, joint_rot, contacts,
                        hf, min_point, dxdy, w_contact, w_pen):
    """tools/procgen/mdm_path.py:31-127.  Inputs [B,F,...]; base_z = min(hf) - 10."""
    body_pos, body_rot = forward_kinematics(model, root_pos, root_rot, joint_rot)
    base_z = torch.min(hf).item() - 10.0
    B = root_pos.shape[0]
    hfb = hf.unsqueeze(0).expand(B, -1, -1)
    mpb = min_point.unsqueeze(0).expand(B, -1)
    F_ = root_pos.shape[1]
    pen = 0.0
    con = 0.0
    for b in range(model.num_bodies):
        P = model.body_points[b].shape[0]
        wp = body_world_points(body_pos, body_rot, model.body_points[b], b).reshape(B, -1, 3)
        neg = torch.clamp(points_hf_sdf(wp, hfb, mpb, dxdy, base_z=base_z, inverted=True), max=0.0)
        pen = pen + torch.sum(-neg, dim=-1) * w_pen
        posd = torch.clamp(points_hf_sdf(wp, hfb, mpb, dxdy, base_z=base_z, inverted=False), min=0.0)
        closest = torch.min(posd.view(B, F_, P), dim=-1)[0]
        con = con + torch.sum(closest * contacts[..., b], dim=-1) * w_contact
    return {"total_loss": con + pen, "contact_loss": con, "pen_loss": pen}


def motion_opt_pen_contact(model: CharModel, tgt_root_pos, tgt_root_rot_expmap, tgt_joint_dof,
                           contacts, hf, min_point, dxdy, w_penetration, w_contact):
    """The FK front-end + penetration + contact terms of motion_terrain_contact_loss
    -- tools/motion_opt/motion_optimization.py:203-213, :241-272, :381-383 (base_z = -10).
    Inputs are the optimiser's leaves [F,3], [F,3], [F,D]; returns (weighted loss, pen, contact)."""
    rq = exp_map_to_quat(tgt_root_rot_expmap)
    jr = dof_to_rot(model, tgt_joint_dof)
    body_pos, body_rot = forward_kinematics(model, tgt_root_pos, rq, jr)
    pen, con = pen_contact_terms(body_pos.unsqueeze(0), body_rot.unsqueeze(0), contacts.unsqueeze(0),
                                 model.body_points, hf, min_point, dxdy, -10.0)
    pen, con = pen[0], con[0]
    return w_penetration * pen + w_contact * con, pen, con


def world_body_points(body_pos, body_rot, body_points):
    """All bodies' surface points in world space, concatenated body-major exactly as the reference's callers do:
    util/terrain_util.py:1918-1936 (motion_frames_hf_sdf_loss), diffusion/mdm.py:1006-1020 (compute_point_hf_sdf).
    body_pos [B,F,J,3], body_rot [B,F,J,4] -> [B, sum_b F*P_b, 3]."""
    B = body_pos.shape[0]
    out = []
    for b in range(body_pos.shape[2]):
        cur = body_points[b].unsqueeze(0).unsqueeze(0)
        out.append((quat_rotate(body_rot[..., b, :].unsqueeze(2), cur) + body_pos[..., b, :].unsqueeze(2)).view(B, -1, 3))
    return torch.cat(out, dim=1)


def hf_collision_loss(sdf):
    """0.5 * sum(clamp(sdf, max=0)^2) -- diffusion/mdm.py:735 (training loss), :1493 (guidance)."""
    return 0.5 * torch.sum(torch.square(torch.clamp(sdf, max=0.0)), dim=-1)


def motion_frames_hf_sdf_loss(model: CharModel, motion_frames, body_points, hf, min_center, dxdy,
                              interior_distance=True):
    """util/terrain_util.py:1895-1949.  motion_frames [B,S,6+D] (root pos | root exp-map | DoFs), hf [B,X,Y],
    min_center [B,2] -> (loss [B], world points [B,N,3], sdf [B,N]); base_z = -10."""
    D = model.dof_size
    root_pos = motion_frames[..., 0:3]
    rq = exp_map_to_quat(motion_frames[..., 3:6])
    jr = quat_w_positive(dof_to_rot(model, motion_frames[..., 6:6 + D]))
    body_pos, body_rot = forward_kinematics(model, root_pos, rq, jr)
    pts = world_body_points(body_pos, body_rot, body_points)
    sdf = points_hf_sdf(pts, hf, min_center, dxdy, base_z=-10.0, inverted=interior_distance)
    if interior_distance:
        loss = 0.5 * torch.sum(torch.square(torch.clamp(sdf, max=0.0)), dim=-1)
    else:
        loss = 0.5 * torch.sum(torch.square(torch.clamp(sdf, min=0.0)), dim=-1)
    return loss, pts, sdf


# --------------------------------------------------------------------------
# contact labelling + heightfield masks
# --------------------------------------------------------------------------
def frames_fk(model: CharModel, frames):
    """Front end shared by the labelling functions: zmotion_editing_tools/motion_edit_lib.py:665-670."""
    rq = exp_map_to_quat(frames[..., 3:6])
    jr = dof_to_rot(model, frames[..., 6:6 + model.dof_size])
    return forward_kinematics(model, frames[..., 0:3], rq, jr)


def box_corners(body_pos, body_rot, half, offset):
    """8 corners of a body-attached box -- util/geom_util.py:80-111.  [N,3],[N,4] -> [N,8,3]."""
    signs = torch.tensor([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1],
                          [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=F32)
    pts = signs * half + offset
    return quat_rotate(body_rot.unsqueeze(1).expand(-1, 8, -1), pts.unsqueeze(0).expand(body_pos.shape[0], -1, -1)) \
        + body_pos.unsqueeze(1)


def foot_contacts_and_pen(model: CharModel, frames, t: Terrain, feet, contact_eps=0.04):
    """compute_hf_foot_contacts_and_correct_pen -- zmotion_editing_tools/motion_edit_lib.py:654-706.
    feet: list of (body_id, half[3], offset[3]).  -> (updated frames, contacts [F,J], pen_correction [F])."""
    F_ = frames.shape[0]
    bp, br = frames_fk(model, frames)
    contacts = torch.zeros(F_, model.num_bodies, dtype=F32)
    corr = torch.zeros(F_)
    for b, half, off in feet:
        pts = box_corners(bp[:, b], br[:, b], torch.tensor(half, dtype=F32), torch.tensor(off, dtype=F32))
        h = hf_sample(t, pts[..., 0:2])
        contacts[:, b] = torch.any(pts[..., 2] < h + contact_eps, dim=-1).float()
        corr = torch.min(corr, torch.min(pts[..., 2] - h, dim=-1)[0])
    out = frames.clone()
    out[:, 2] -= corr
    return out, contacts, corr


def hand_contacts(model: CharModel, frames, t: Terrain, hands, contact_eps=0.04):
    """compute_motion_terrain_hand_contacts -- zmotion_editing_tools/motion_edit_lib.py:708-747.
    hands: list of (body_id, radius)."""
    bp, _ = frames_fk(model, frames)
    contacts = torch.zeros(frames.shape[0], model.num_bodies, dtype=F32)
    base_z = torch.min(t.hf).item() - 10.0
    for b, radius in hands:
        sd = points_hf_sdf(bp[:, b].unsqueeze(0), t.hf.unsqueeze(0), t.min_point.unsqueeze(0), t.dxdy, base_z, False)
        contacts[:, b] = ((sd[0] - radius) < contact_eps).float()    # sdRoundBox = sdBox - r, min commutes
    return contacts


def hf_mask_inds(model: CharModel, frames, t: Terrain):
    """compute_hf_mask_inds -- util/terrain_util.py:1951-1997, vectorised per frame (the reference's scalar
    loops compute exactly a per-frame unique() and a running per-cell min)."""
    bp, br = frames_fk(model, frames)
    X, Y = t.hf.shape
    min_h = torch.full((X, Y), 99999.9999, dtype=F32)
    inds = []
    for f in range(frames.shape[0]):
        pts = torch.cat([quat_rotate(br[f, b].unsqueeze(0), model.body_points[b]) + bp[f, b]
                         for b in range(model.num_bodies)], dim=0)
        g = grid_index(t, pts[:, 0:2])
        flat = g[:, 0] * Y + g[:, 1]
        min_h = min_h.view(-1).scatter_reduce(0, flat, pts[:, 2], reduce="amin", include_self=True).view(X, Y)
        inds.append(torch.unique(g, dim=0))
    return inds, min_h


# --------------------------------------------------------------------------
# the remaining motion_terrain_contact_loss terms
# --------------------------------------------------------------------------
def motion_terrain_contact_loss_full(model: CharModel, tgt_root_pos, tgt_root_rot, tgt_joint_dof, src_root_pos,
                                     src_root_rot_quat, src_joint_rot, src_body_vels, src_body_rot_vels, contacts, hf,
                                     min_point, dxdy, w, max_jerk, body_constraints=None, geom0=None):
    """tools/motion_opt/motion_optimization.py:183-395, every term.  `w`: dict of the nine weights;
    body_constraints: per body a list of (start, end, point[3]); geom0: per body (type, offset[3], dims) of its
    first geom (type 0 = BOX, 1 = SPHERE as anim/kin_char_model.py:102-107).  -> (loss, dict of term tensors)."""
    root_pos_loss = torch.sum(torch.square(tgt_root_pos - src_root_pos))
    rq = exp_map_to_quat(tgt_root_rot)
    root_rot_loss = torch.sum(torch.square(quat_diff_angle(rq, src_root_rot_quat)))
    jr = dof_to_rot(model, tgt_joint_dof)
    joint_rot_loss = torch.sum(torch.square(quat_diff_angle(jr, src_joint_rot)))
    bp, br = forward_kinematics(model, tgt_root_pos, rq, jr)
    vels = bp[1:] - bp[:-1]
    vel_err_sq = torch.square(vels - src_body_vels)
    rot_vels = quat_diff_angle(br[1:], br[:-1])
    rot_vel_err_sq = torch.square(rot_vels - src_body_rot_vels)
    smoothness = torch.sum(vel_err_sq) + torch.sum(rot_vel_err_sq)
    change = torch.clamp(torch.min(torch.cat([contacts[1:].unsqueeze(-1), contacts[:-1].unsqueeze(-1)], dim=-1), dim=-1)[0], min=0.0)
    pen, con = pen_contact_terms(bp.unsqueeze(0), br.unsqueeze(0), contacts.unsqueeze(0), model.body_points, hf,
                                 min_point, dxdy, -10.0)
    pen, con = pen[0], con[0]
    if w["w_contact"] == 0.0:
        con = 0.0
    bc_loss = 0.0
    if body_constraints is not None:
        for b in range(model.num_bodies):
            for (s_, e_, point) in body_constraints[b]:
                gtype, goff, gdims = geom0[b]
                point = torch.as_tensor(point, dtype=F32)
                if gtype == 1:      # SPHERE
                    centre = quat_rotate(br[:, b], torch.as_tensor(goff, dtype=F32).unsqueeze(0)) + bp[:, b]
                    diff = torch.norm(point.unsqueeze(0) - centre[s_:e_ + 1], dim=-1) - torch.as_tensor(gdims, dtype=F32)
                    bc_loss = bc_loss + torch.sum(torch.abs(diff))
                elif gtype == 0:    # BOX: the 18 sole points (first z slice of the box samples)
                    radius = torch.norm(torch.as_tensor(gdims, dtype=F32)) * 1.25
                    pts = quat_rotate(br[:, b].unsqueeze(1), model.body_points[b].unsqueeze(0)) + bp[:, b].unsqueeze(1)
                    sole = pts[s_:e_ + 1, 0:18].reshape(-1, 3)
                    diff = torch.norm(point.unsqueeze(0) - sole, dim=-1) - radius
                    bc_loss = bc_loss + torch.sum(torch.clamp(diff, min=0.0))
                else:
                    continue
                vel_err_sq = vel_err_sq.clone()
                vel_err_sq[s_:e_ + 1, b] *= 0.0
                rot_vel_err_sq = rot_vel_err_sq.clone()
                rot_vel_err_sq[s_:e_ + 1, b] *= 0.0
    if w["w_sliding"] != 0.0:
        c, c2 = 0.03, 0.0009
        sliding = torch.sum((torch.sqrt(torch.sum(vel_err_sq, dim=-1) + c2) - c) * change) \
            + torch.sum((torch.sqrt(rot_vel_err_sq + c2) - c) * change)
    else:
        sliding = 0.0
    acc = vels[1:] - vels[:-1]
    jerk_mag = torch.norm(acc[1:] - acc[:-1], dim=-1)
    jerk = torch.sum(torch.clamp(jerk_mag - max_jerk * ((1.0 / 30.0) ** 3), min=0.0))
    loss = w["w_root_pos"] * root_pos_loss + w["w_root_rot"] * root_rot_loss + w["w_joint_rot"] * joint_rot_loss \
        + w["w_smoothness"] * smoothness + w["w_penetration"] * pen + w["w_contact"] * con + w["w_sliding"] * sliding \
        + w["w_body_constraints"] * bc_loss + w["w_jerk"] * jerk
    return loss, dict(root_pos=root_pos_loss, root_rot=root_rot_loss, joint_rot=joint_rot_loss, smoothness=smoothness,
                      penetration=pen, contact=con, sliding=sliding, body_constraint=bc_loss, jerk=jerk)


def motion_contact_optimization(model: CharModel, src_frames, contacts, hf, min_point, dxdy, num_iters, step_size, w,
                                max_jerk, body_constraints=None, geom0=None):
    """The Adam loop of tools/motion_opt/motion_optimization.py:404-500 (logging omitted)."""
    D = model.dof_size
    src_root_pos, src_root_rot, src_joint_dof = src_frames[:, 0:3], src_frames[:, 3:6], src_frames[:, 6:6 + D]
    src_rq = exp_map_to_quat(src_root_rot)
    src_jr = dof_to_rot(model, src_joint_dof)
    sbp, sbr = forward_kinematics(model, src_root_pos, src_rq, src_jr)
    src_bv = sbp[1:] - sbp[:-1]
    src_brv = quat_diff_angle(sbr[1:], sbr[:-1])
    leaves = [src_root_pos.clone().requires_grad_(True), src_root_rot.clone().requires_grad_(True),
              src_joint_dof.clone().requires_grad_(True)]
    opt = torch.optim.Adam(leaves, lr=step_size)
    for _ in range(num_iters):
        opt.zero_grad()
        loss, _t = motion_terrain_contact_loss_full(model, leaves[0], leaves[1], leaves[2], src_root_pos, src_rq, src_jr,
                                                    src_bv, src_brv, contacts, hf, min_point, dxdy, w, max_jerk,
                                                    body_constraints, geom0)
        loss.backward()
        opt.step()
    return torch.cat([t.detach() for t in leaves], dim=-1)


# --------------------------------------------------------------------------
# MDM sampler terrain gather (SURVEY.md §8(f)-4 clause)
# --------------------------------------------------------------------------
def clip_hfs_from_data(terrains, hf_maxmins, hf_mask_inds, motion_ids, root_pos, root_rot, canon_root_z,
                       motion_time_indices, tmpl, num_x_neg, num_y_neg, max_h, relative_to_root):
    """diffusion/mdm_heightfield_contact_motion_sampler.py:449-474 (get_hfs_from_data, augmentation off) with its helper
    :414-447.  terrains: list of Terrain; hf_maxmins: list of [X,Y,2]; hf_mask_inds: per clip a list (frames) of int64
    [n,2] index tensors; tmpl [GX,GY,2].  -> (hfs [B,GX,GY], center_h [B], hf_maxmins [B,GX,GY,2])."""
    B = root_pos.shape[0]
    GX, GY = tmpl.shape[0], tmpl.shape[1]
    heading = calc_heading(root_rot).unsqueeze(-1).unsqueeze(-1).expand(-1, GX, GY)
    xy = rotate_2d(tmpl.unsqueeze(0).expand(B, -1, -1, -1), heading) + root_pos[:, 0:2].unsqueeze(1).unsqueeze(1)
    min_h = -max_h
    hfs, mms = [], []
    for i in range(B):
        t = terrains[int(motion_ids[i])]
        inds = grid_index(t, xy[i])
        hfs.append(t.hf[inds[..., 0], inds[..., 1]])
        sl = slice(int(motion_time_indices[i][0]), int(motion_time_indices[i][-1]) + 1)
        mask = torch.zeros(t.hf.shape, dtype=torch.bool)
        for fr in hf_mask_inds[int(motion_ids[i])][sl]:
            mask[fr[..., 0], fr[..., 1]] = True
        mask = mask.unsqueeze(-1).expand(-1, -1, 2)
        band = torch.zeros_like(hf_maxmins[int(motion_ids[i])])
        band[..., 0] = max_h * 2.0
        band[..., -1] = min_h * 2.0
        band[mask] = hf_maxmins[int(motion_ids[i])][mask]
        mms.append(band[inds[..., 0], inds[..., 1], :])
    hfs, mms = torch.stack(hfs, dim=0), torch.stack(mms, dim=0)
    center_h = hfs[:, num_x_neg, num_y_neg].clone()
    ref = canon_root_z if relative_to_root else center_h
    hfs = hfs - ref.unsqueeze(-1).unsqueeze(-1)
    mms = mms - ref.unsqueeze(-1).unsqueeze(-1).unsqueeze(-1)
    return hfs, center_h, mms


# --------------------------------------------------------------------------
# tracker step assembly: policy observation, reward, done
# --------------------------------------------------------------------------
DONE_NULL, DONE_FAIL, DONE_SUCC, DONE_TIME = 0, 1, 2, 3        # envs/base_env.py:12-16


def heading_inverse_quat(q):
    """Rotation about +z by -heading(q) -- util/torch_util.py:491-499."""
    up = torch.zeros_like(q[..., 0:3])
    up[..., 2] = 1
    return axis_angle_to_quat(up, -calc_heading(q))


def quat_to_tan_norm(q):
    """Rotated x axis followed by rotated z axis, 6 numbers -- util/torch_util.py:361-373."""
    ex = torch.zeros_like(q[..., 0:3])
    ex[..., 0] = 1
    ez = torch.zeros_like(q[..., 0:3])
    ez[..., 2] = 1
    return torch.cat([quat_rotate(q, ex), quat_rotate(q, ez)], dim=-1)


def _rotate_rows(q, v):
    """quat_rotate of v[..., k, 3] by q[..., 4] broadcast over k (the reference's repeat + flatten idiom)."""
    return quat_rotate(q.unsqueeze(-2).expand(v.shape[:-1] + (4,)), v)


def compute_char_obs(root_pos, root_rot, root_vel, root_ang_vel, joint_rot, dof_vel, key_pos, global_obs,
                     root_height_obs):
    """Proprioceptive observation of the simulated character -- envs/ig_char_env.py:582-626.
    joint_rot [N,J-1,4]; key_pos [N,K,3] or an empty tensor.  Layout:
    [root_h (opt)] root tan-norm 6 | root_vel 3 | root_ang_vel 3 | joint tan-norm 6(J-1) | dof_vel D | key_pos 3K."""
    hinv = heading_inverse_quat(root_rot)
    if global_obs:
        parts = [quat_to_tan_norm(root_rot), root_vel, root_ang_vel]
    else:
        parts = [quat_to_tan_norm(quat_mul(hinv, root_rot)), quat_rotate(hinv, root_vel), quat_rotate(hinv, root_ang_vel)]
    parts.append(quat_to_tan_norm(joint_rot).reshape(joint_rot.shape[0], -1))
    parts.append(dof_vel)
    if key_pos.numel() > 0:
        rel = key_pos - root_pos.unsqueeze(-2)
        if not global_obs:
            rel = _rotate_rows(hinv, rel)
        parts.append(rel.reshape(rel.shape[0], -1))
    if root_height_obs:
        parts = [root_pos[:, 2:3]] + parts
    return torch.cat(parts, dim=-1)


def compute_tar_obs(ref_root_pos, ref_root_rot, tar_root_pos, tar_root_rot, joint_rot, tar_key_pos, global_obs,
                    global_tar_root_h_obs):
    """Future-target observation -- envs/ig_parkour/mgdm_dm_util.py:462-518.  ref_* [N,.] is the character the
    targets are expressed against; tar_root_pos [N,S,3], tar_root_rot [N,S,4], joint_rot [N,S,J-1,4],
    tar_key_pos [N,S,K,3] or empty.  Returns [N,S, 3 + 6 + 6(J-1) + 3K]."""
    pos_obs = tar_root_pos - ref_root_pos.unsqueeze(-2)
    has_keys = tar_key_pos.numel() > 0
    if has_keys:
        tar_key_pos = tar_key_pos - tar_root_pos.unsqueeze(-2)
    if not global_obs:
        hinv = heading_inverse_quat(ref_root_rot)
        pos_obs = _rotate_rows(hinv, pos_obs)
        tar_root_rot = quat_mul(hinv.unsqueeze(-2).expand(tar_root_rot.shape), tar_root_rot)
        if has_keys:
            hk = hinv.unsqueeze(-2).unsqueeze(-2).expand(tar_key_pos.shape[:-1] + (4,))
            tar_key_pos = quat_rotate(hk, tar_key_pos) + pos_obs.unsqueeze(2)
    if global_tar_root_h_obs:
        pos_obs = pos_obs.clone()
        pos_obs[..., 2] = tar_root_pos[..., 2]
    parts = [pos_obs, quat_to_tan_norm(tar_root_rot), quat_to_tan_norm(joint_rot).reshape(joint_rot.shape[:2] + (-1,))]
    if has_keys:
        parts.append(tar_key_pos.reshape(tar_key_pos.shape[:2] + (-1,)))
    return torch.cat(parts, dim=-1)


def _to_heading_frame(root_rot, root_vel, root_ang_vel, key_pos):
    """envs/ig_parkour/mgdm_dm_util.py:304-326 (convert_to_local)"""
    hinv = heading_inverse_quat(root_rot)
    kp = _rotate_rows(hinv, key_pos) if key_pos.numel() > 0 else key_pos
    return quat_mul(hinv, root_rot), quat_rotate(hinv, root_vel), quat_rotate(hinv, root_ang_vel), kp


def compute_deepmimic_reward(root_pos, root_rot, root_vel, root_ang_vel, joint_rot, dof_vel, key_pos,
                             tar_root_pos, tar_root_rot, tar_root_vel, tar_root_ang_vel, tar_joint_rot, tar_dof_vel,
                             tar_key_pos, joint_rot_err_w, dof_err_w, track_root_h, track_root):
    """The five exponentiated tracking terms [N,5] = (pose, vel, root pose, root vel, key pos) --
    envs/ig_parkour/mgdm_dm_util.py:328-397."""
    ang = quat_diff_angle(joint_rot, tar_joint_rot)
    pose_err = torch.sum(joint_rot_err_w * ang * ang, dim=-1)
    dv = tar_dof_vel - dof_vel
    vel_err = torch.sum(dof_err_w * dv * dv, dim=-1)
    dp = (tar_root_pos - root_pos).clone()
    if not track_root:
        dp[..., 0:2] = 0
    if not track_root_h:
        dp[..., 2] = 0
    root_pos_err = torch.sum(dp * dp, dim=-1)
    has_keys = key_pos.numel() > 0
    if has_keys:
        key_pos = key_pos - root_pos.unsqueeze(-2)
        tar_key_pos = tar_key_pos - tar_root_pos.unsqueeze(-2)
    if not track_root:
        root_rot, root_vel, root_ang_vel, key_pos = _to_heading_frame(root_rot, root_vel, root_ang_vel, key_pos)
        tar_root_rot, tar_root_vel, tar_root_ang_vel, tar_key_pos = _to_heading_frame(tar_root_rot, tar_root_vel,
                                                                                      tar_root_ang_vel, tar_key_pos)
    rr = quat_diff_angle(root_rot, tar_root_rot)
    root_rot_err = rr * rr
    d = tar_root_vel - root_vel
    root_vel_err = torch.sum(d * d, dim=-1)
    d = tar_root_ang_vel - root_ang_vel
    root_ang_vel_err = torch.sum(d * d, dim=-1)
    if has_keys:
        d = tar_key_pos - key_pos
        key_pos_err = torch.sum(torch.sum(d * d, dim=-1), dim=-1)
    else:
        key_pos_err = torch.zeros([0])
    return torch.stack([torch.exp(-0.25 * pose_err), torch.exp(-0.01 * vel_err),
                        torch.exp(-5.0 * (root_pos_err + 0.1 * root_rot_err)),
                        torch.exp(-1.0 * (root_vel_err + 0.1 * root_ang_vel_err)),
                        torch.exp(-10.0 * key_pos_err)], dim=1)


def termination_heights(t: Terrain, body_pos, env_offsets, termination_height):
    """Terrain height under every body plus the margin -- envs/ig_parkour/mgdm_dm_util.py:208-210."""
    xy = body_pos[..., 0:2] + env_offsets[:, 0:2].unsqueeze(1)
    return hf_sample(t, xy) + termination_height


def compute_done(done_buf, time, ep_len, root_rot, body_pos, tar_root_rot, tar_body_pos, contact_force,
                 contact_body_ids, term_heights, pose_termination, pose_termination_dist, enable_early_termination,
                 track_root, root_pos_termination_dist, root_rot_termination_angle):
    """Episode flags -- envs/ig_parkour/mgdm_dm_util.py:399-460 (char_root_pos and global_obs are unused there).
    contact_body_ids: bodies ALLOWED to touch the ground; pose_termination_dist [J-1]."""
    done = torch.full_like(done_buf, DONE_NULL)
    done[time >= ep_len] = DONE_TIME
    if enable_early_termination:
        failed = torch.zeros(done.shape, dtype=torch.bool)
        if contact_body_ids.shape[0] > 0:
            force = contact_force.detach().clone()
            force[:, contact_body_ids, :] = 0
            touched = torch.any(torch.any(torch.abs(force) > 0.1, dim=-1), dim=-1)
            low = body_pos[..., 2] < term_heights
            low[:, contact_body_ids] = False
            failed = failed | (touched & torch.any(low, dim=-1))
        if pose_termination:
            rp = body_pos[..., 0:1, :]
            trp = tar_body_pos[..., 0:1, :]
            d = (tar_body_pos[..., 1:, :] - trp) - (body_pos[..., 1:, :] - rp)
            pose_fail = torch.any(torch.sum(d * d, dim=-1) > pose_termination_dist * pose_termination_dist, dim=-1)
            if track_root:
                d = rp - trp
                pose_fail = pose_fail | (torch.sum(d * d, dim=-1).squeeze(-1)
                                         > root_pos_termination_dist * root_pos_termination_dist)
                pose_fail = pose_fail | (torch.abs(quat_diff_angle(root_rot, tar_root_rot)) > root_rot_termination_angle)
            failed = failed | pose_fail
        failed = failed & (time > 1e-5)
        done[failed] = DONE_FAIL
    return done
