"""Quaternion / exp-map helpers (xyzw) used on the HOST side of parc_b200.

These run at load time (building the frame tables on CPU before they are packed and uploaded) and in
the thin parts of the API that are not on the batched hot path.  The hot path itself -- frame query,
slerp, FK, heightfield sampling, body-point losses -- runs in libparc_b200's CUDA kernels, not here.

Names and argument meaning follow the reference's `util/torch_util.py` (file:line cited per function)
so callers can switch imports; rounding-relevant operation order is kept so that tables built here
are bit-identical to the reference's on the same device.
"""
from __future__ import annotations

import torch


def normalize_angle(x):
    """Wrap to (-pi, pi] via atan2(sin, cos).  Ref util/torch_util.py:4-7."""
    return torch.atan2(torch.sin(x), torch.cos(x))


def normalize(x, eps: float = 1e-9):
    """x / max(|x|, eps).  Ref util/torch_util.py:9-12."""
    length = torch.linalg.vector_norm(x, ord=2, dim=-1).clamp(min=eps)
    return x / length.unsqueeze(-1)


quat_unit = normalize  # ref :24-27


def quat_conjugate(q):
    """Ref util/torch_util.py:29-31."""
    return torch.cat([-q[..., :3], q[..., 3:]], dim=-1)


def quat_pos(q):
    """Flip sign so that w >= 0.  Ref util/torch_util.py:33-38."""
    flip = (q[..., 3:] < 0).to(torch.float32)
    return (1 - 2 * flip) * q


def quat_mul(a, b):
    """Hamilton product, the reference's 8-multiply arrangement.  Ref util/torch_util.py:40-58."""
    assert a.shape == b.shape
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    p_ww = (az + ax) * (bx + by)
    p_yy = (aw - ay) * (bw + bz)
    p_zz = (aw + ay) * (bw - bz)
    p_xx = p_ww + p_yy + p_zz
    h = 0.5 * (p_xx + (az - ax) * (bx - by))
    return torch.stack([h - p_xx + (ax + aw) * (bx + bw),
                        h - p_yy + (aw - ax) * (by + bz),
                        h - p_zz + (az + ay) * (bw - bx),
                        h - p_ww + (az - ay) * (by - bz)], dim=-1)


def quat_rotate(q, v):
    """Rotate v by q.  Ref util/torch_util.py:60-66."""
    qv, qw = q[..., :3], q[..., 3:]
    t = 2 * torch.cross(qv, v, dim=-1)
    return v + qw * t + torch.cross(qv, t, dim=-1)


def _z_axis_like(v):
    z = torch.zeros_like(v)
    z[..., -1] = 1
    return z


def quat_to_axis_angle(q):
    """Ref util/torch_util.py:68-88."""
    q = quat_pos(q)
    vec = q[..., 0:3]
    length = torch.linalg.vector_norm(vec, ord=2, dim=-1)
    angle = 2.0 * torch.atan2(length, q[..., 3])
    axis = vec / length.unsqueeze(-1)
    ok = length > 1e-5
    angle = torch.where(ok, angle, torch.zeros_like(angle))
    axis = torch.where(ok.unsqueeze(-1), axis, _z_axis_like(axis))
    return axis, angle


def axis_angle_to_quat(axis, angle):
    """Ref util/torch_util.py:311-317."""
    half = (angle / 2).unsqueeze(-1)
    return quat_unit(torch.cat([normalize(axis) * half.sin(), half.cos()], dim=-1))


def axis_angle_to_exp_map(axis, angle):
    """Ref util/torch_util.py:329-334."""
    return angle.unsqueeze(-1) * axis


def quat_to_exp_map(q):
    """Ref util/torch_util.py:346-351."""
    axis, angle = quat_to_axis_angle(q)
    return axis_angle_to_exp_map(axis, angle)


def exp_map_to_axis_angle(exp_map):
    """Ref util/torch_util.py:394-412."""
    angle = torch.linalg.vector_norm(exp_map, dim=-1)
    axis = exp_map / angle.unsqueeze(-1)
    angle = normalize_angle(angle)
    ok = torch.abs(angle) > 1e-5
    angle = torch.where(ok, angle, torch.zeros_like(angle))
    axis = torch.where(ok.unsqueeze(-1), axis, _z_axis_like(exp_map))
    return axis, angle


def exp_map_to_quat(exp_map):
    """Ref util/torch_util.py:414-419."""
    return axis_angle_to_quat(*exp_map_to_axis_angle(exp_map))


def quat_diff(q0, q1):
    """q1 * conj(q0).  Ref util/torch_util.py:422-425."""
    return quat_mul(q1, quat_conjugate(q0))


def quat_diff_angle(q0, q1):
    """Ref util/torch_util.py:427-431."""
    return quat_to_axis_angle(quat_diff(q0, q1))[1]


def quat_normalize(q):
    """Positive-w unit quaternion.  Ref util/torch_util.py:438-441."""
    return quat_unit(quat_pos(q))


def slerp(q0, q1, t):
    """Host-side slerp with the reference's branch rules (util/torch_util.py:443-468).  The batched
    query path uses the CUDA implementation in csrc/parc_common.cuh instead."""
    assert t.dim() <= q0.dim()
    c = torch.sum(q0 * q1, dim=-1)
    q1 = torch.where((c < 0).unsqueeze(-1), -q1, q1)
    c = torch.abs(c).unsqueeze(-1)
    theta = torch.acos(c)
    s = torch.sqrt(1.0 - c * c)
    if t.dim() == q0.dim() - 1:
        t = t.unsqueeze(-1)
    out = (torch.sin((1 - t) * theta) / s) * q0 + (torch.sin(t * theta) / s) * q1
    out = torch.where(torch.abs(s) < 0.001, 0.5 * q0 + 0.5 * q1, out)
    return torch.where(torch.abs(c) >= 1, q0, out)


def calc_heading(q):
    """Yaw of the rotated x axis.  Ref util/torch_util.py:470-479."""
    assert q.shape[-1] == 4
    ex = torch.zeros_like(q[..., 0:3])
    ex[..., 0] = 1
    d = quat_rotate(q, ex)
    return torch.atan2(d[..., 1], d[..., 0])


def calc_heading_quat(q):
    """Ref util/torch_util.py:481-489."""
    return axis_angle_to_quat(_z_axis_like(q[..., 0:3]), calc_heading(q))


def calc_heading_quat_inv(q):
    """Ref util/torch_util.py:491-499."""
    return axis_angle_to_quat(_z_axis_like(q[..., 0:3]), -calc_heading(q))


def rotate_2d_vec(vec, angle):
    """Ref util/torch_util.py:619-631."""
    x, y = vec[..., 0], vec[..., 1]
    c, s = torch.cos(angle), torch.sin(angle)
    return torch.stack([x * c - y * s, x * s + y * c], dim=-1)


def quat_to_tan_norm(q):
    """The 6-number rotation encoding of the policy observations: rotated x axis, then rotated z axis.
    Ref util/torch_util.py:361-373.  (The step kernels of csrc/tracker_step.cu compute this in place.)"""
    ex = torch.zeros_like(q[..., 0:3])
    ex[..., 0] = 1
    return torch.cat([quat_rotate(q, ex), quat_rotate(q, _z_axis_like(q[..., 0:3]))], dim=-1)


def quat_abs(x):
    """Ref util/torch_util.py:433-436."""
    return x.norm(p=2, dim=-1)


def quat_inv(q):
    """Conjugate as the inverse of a unit quaternion.  Ref util/torch_util.py:597-607."""
    return torch.cat([-q[..., 0:3], q[..., 3:4]], dim=-1)


def quat_multiply(q1, q2):
    """Plain 16-multiply Hamilton product q1 * q2 (the reference keeps it beside the 8-multiply `quat_mul`).
    Ref util/torch_util.py:577-595."""
    x1, y1, z1, w1 = q1.unbind(-1)
    x2, y2, z2, w2 = q2.unbind(-1)
    return torch.stack((w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                        w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2), dim=-1)


def heading_to_quat(heading):
    """Rotation about +z by `heading`.  Ref util/torch_util.py:319-326."""
    axis = torch.zeros(list(heading.shape) + [3], dtype=torch.float32, device=heading.device)
    axis[..., 2] = 1
    return axis_angle_to_quat(axis, heading)
