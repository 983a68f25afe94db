"""motion_terrain_contact_loss: the per-clip objective of the kinematic motion optimiser.

Drop-in for the reference's `tools/motion_opt/motion_optimization.py::motion_terrain_contact_loss`
(:183-395): same arguments, returns (loss, losses-dict keyed by LossType).  The penetration and contact
terms (:241-272) -- the part that dominates the reference's 21 s / iteration -- run forward AND backward
in the fused kernel of csrc/body_loss.cu, entered through `ops.body_loss`; FK, DoF->quaternion and
exp-map->quaternion are the CUDA operators of csrc/fk.cu with hand-written VJPs.  The remaining cheap
regularisers (tracking, smoothness, sliding, jerk, body constraints) are elementwise torch expressions
over the kernels' outputs (SURVEY.md section 8(f) row 2 lists fusing them as the next step).
"""
from __future__ import annotations

import enum
from typing import List

import torch

from ... import ops
from ...anim import kin_char_model
from ...util import geom_util, torch_util
from ..procgen.mdm_path import body_points_desc


class LossType(enum.Enum):
    ROOT_POS_LOSS = 0
    ROOT_ROT_LOSS = 1
    JOINT_ROT_LOSS = 2
    SMOOTHNESS_LOSS = 3
    PENETRATION_LOSS = 4
    CONTACT_LOSS = 5
    SLIDING_LOSS = 6
    BODY_CONSTRAINT_LOSS = 7
    JERK_LOSS = 8
    LOOPING_LOSS = 9


class BodyConstraint:
    start_frame_idx = 0
    end_frame_idx = 0
    constraint_point = None


def pen_contact_loss(tgt_root_pos, tgt_root_rot_quat, tgt_joint_rot, contacts, terrain, body_points, char_model,
                     w_penetration: float, w_contact: float, base_z: float = -10.0):
    """The two heightfield terms for one clip [F,...]: returns (weighted sum, pen, contact) with
    pen / contact detached (they are only logged by the reference, :366-373)."""
    tb = ops.make_terrain_batch(terrain.hf, terrain.min_point, terrain.dxdy.detach().cpu().tolist(), base_z=base_z)
    pts = body_points_desc(char_model, body_points)
    total, pen, con = ops.body_loss(char_model.c_model(), pts, tb, tgt_root_pos.unsqueeze(0),
                                    tgt_root_rot_quat.unsqueeze(0), tgt_joint_rot.unsqueeze(0),
                                    contacts.unsqueeze(0), w_penetration, w_contact)
    return total[0], pen[0], con[0]


def motion_terrain_contact_loss(tgt_root_pos, tgt_root_rot, tgt_joint_dof, src_root_pos, src_root_rot_quat,
                                src_joint_rot, src_body_vels, src_body_rot_vels, contacts, terrain,
                                body_points: List[torch.Tensor], char_model, w_root_pos: float, w_root_rot: float,
                                w_joint_rot: float, w_smoothness: float, w_penetration: float, w_contact: float,
                                w_sliding: float, w_body_constraints: float, w_jerk: float, body_constraints: list,
                                max_jerk: float):
    root_pos_loss = torch.sum(torch.square(tgt_root_pos - src_root_pos))

    tgt_root_rot_quat = ops.exp_map_to_quat(tgt_root_rot)
    root_rot_loss = torch.sum(torch.square(torch_util.quat_diff_angle(tgt_root_rot_quat, src_root_rot_quat)))

    tgt_joint_rot = char_model.dof_to_rot(tgt_joint_dof)
    joint_rot_loss = torch.sum(torch.square(torch_util.quat_diff_angle(tgt_joint_rot, src_joint_rot)))

    tgt_body_pos, tgt_body_rot = char_model.forward_kinematics(tgt_root_pos, tgt_root_rot_quat, tgt_joint_rot)

    tgt_body_vels = tgt_body_pos[1:] - tgt_body_pos[:-1]
    body_vel_err_sq = torch.square(tgt_body_vels - src_body_vels)
    tgt_body_rot_vels = torch_util.quat_diff_angle(tgt_body_rot[1:], tgt_body_rot[:-1])
    body_rot_vel_err_sq = torch.square(tgt_body_rot_vels - src_body_rot_vels)
    smoothness_loss = torch.sum(body_vel_err_sq) + torch.sum(body_rot_vel_err_sq)

    frame_change_in_contact = torch.clamp(torch.minimum(contacts[1:], contacts[:-1]), min=0.0)

    # --- heightfield terms: fused CUDA forward + backward (:241-272) ---
    hf_total, penetration_loss, contact_loss = pen_contact_loss(
        tgt_root_pos, tgt_root_rot_quat, tgt_joint_rot, contacts, terrain, body_points, char_model,
        w_penetration, w_contact if w_contact != 0.0 else 0.0)
    contact_logged = contact_loss if w_contact != 0.0 else 0.0

    # --- body constraints (:286-332) ---
    body_constraint_loss = 0.0
    if body_constraints is not None:
        for b in range(char_model.get_num_joints()):
            if len(body_constraints[b]) == 0:
                continue
            geom0 = char_model.get_geoms(b)[0]
            curr_rot = tgt_body_rot[:, b]
            curr_pos = tgt_body_pos[:, b]
            for c in body_constraints[b]:
                s, e = c.start_frame_idx, c.end_frame_idx
                if geom0._shape_type == kin_char_model.GeomType.SPHERE:
                    centre = torch_util.quat_rotate(curr_rot, geom0._offset.unsqueeze(0)) + curr_pos
                    diff = geom_util.sdSphere(c.constraint_point.unsqueeze(0), centre[s:e + 1], geom0._dims)
                    body_constraint_loss = body_constraint_loss + torch.sum(torch.abs(diff))
                elif geom0._shape_type == kin_char_model.GeomType.BOX:
                    radius = torch.norm(geom0._dims) * 1.25
                    sole = torch_util.quat_rotate(curr_rot[s:e + 1].unsqueeze(1), body_points[b][0:18].unsqueeze(0)) \
                        + curr_pos[s:e + 1].unsqueeze(1)
                    diff = geom_util.sdSphere(c.constraint_point.unsqueeze(0), sole.reshape(-1, 3), radius)
                    body_constraint_loss = body_constraint_loss + torch.sum(torch.clamp(diff, min=0.0))
                else:
                    continue
                body_vel_err_sq = body_vel_err_sq.clone()
                body_vel_err_sq[s:e + 1, b] *= 0.0
                body_rot_vel_err_sq = body_rot_vel_err_sq.clone()
                body_rot_vel_err_sq[s:e + 1, b] *= 0.0

    # --- sliding: pseudo-Huber on contact bodies (:339-346) ---
    if w_sliding != 0.0:
        c, c2 = 0.03, 0.0009
        sliding_loss = torch.sum((torch.sqrt(torch.sum(body_vel_err_sq, dim=-1) + c2) - c) * frame_change_in_contact) \
            + torch.sum((torch.sqrt(body_rot_vel_err_sq + c2) - c) * frame_change_in_contact)
    else:
        sliding_loss = 0.0

    # --- jerk (:348-354) ---
    acc = tgt_body_vels[1:] - tgt_body_vels[:-1]
    jerk_mag = torch.norm(acc[1:] - acc[:-1], dim=-1)
    dt = 1.0 / 30.0
    jerk_loss = torch.sum(torch.clamp(jerk_mag - max_jerk * (dt ** 3), min=0.0))

    as_float = lambda v: v.item() if isinstance(v, torch.Tensor) else v
    losses = {
        LossType.ROOT_POS_LOSS: root_pos_loss.item(),
        LossType.ROOT_ROT_LOSS: root_rot_loss.item(),
        LossType.JOINT_ROT_LOSS: joint_rot_loss.item(),
        LossType.SMOOTHNESS_LOSS: smoothness_loss.item(),
        LossType.PENETRATION_LOSS: penetration_loss.item(),
        LossType.CONTACT_LOSS: as_float(contact_logged),
        LossType.SLIDING_LOSS: as_float(sliding_loss),
        LossType.JERK_LOSS: jerk_loss.item(),
        LossType.BODY_CONSTRAINT_LOSS: as_float(body_constraint_loss),
    }
    loss = w_root_pos * root_pos_loss + w_root_rot * root_rot_loss + w_joint_rot * joint_rot_loss \
        + w_smoothness * smoothness_loss + hf_total + w_sliding * sliding_loss \
        + w_body_constraints * body_constraint_loss + w_jerk * jerk_loss
    return loss, losses
