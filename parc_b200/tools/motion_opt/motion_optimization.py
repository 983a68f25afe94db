"""Kinematic motion optimisation against the terrain: the per-clip objective and its Adam loop.

Drop-in for the reference's `tools/motion_opt/motion_optimization.py`: `motion_terrain_contact_loss`
(:183-395) and `motion_contact_optimization` (:404-500), same arguments and returns.

* The penetration and contact terms (:241-272) -- the part that dominates the reference's 21 s / iteration --
  run forward AND backward in the fused kernel of csrc/body_loss.cu (`ops.body_loss`); FK, DoF->quaternion and
  exp-map->quaternion are the CUDA operators of csrc/fk.cu with hand-written VJPs.
* The remaining cheap regularisers (tracking, smoothness, sliding, jerk, body constraints) are elementwise torch
  expressions over the kernels' outputs.
* `motion_contact_optimization` captures ONE whole iteration (kernels + regularisers + backward + Adam step) in a
  CUDA graph and replays it `num_iters` times; the host only synchronises on logging iterations.  The reference
  calls `.item()` on seven loss terms every iteration.
"""
from __future__ import annotations

import enum
import time
from typing import List, Optional

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


def _terrain_desc(terrain, base_z=-10.0):
    """Host-side preparation (one tiny D2H of dxdy): kept out of the captured iteration."""
    return ops.make_terrain_batch(terrain.hf, terrain.min_point, terrain.dxdy.detach().cpu().tolist(), base_z=base_z)


def pen_contact_loss(tgt_root_pos, tgt_root_rot_quat, tgt_joint_rot, contacts, terrain, body_points, char_model,
                     w_penetration: float, w_contact: float, base_z: float = -10.0, _tb=None, _pts=None):
    """The two heightfield terms for one clip [F,...]: returns (weighted sum, pen, contact) with
    pen / contact detached (they are only logged by the reference, :366-373)."""
    tb = _tb if _tb is not None else _terrain_desc(terrain, base_z)
    pts = _pts if _pts is not None else body_points_desc(char_model, body_points)
    total, pen, con = ops.body_loss(char_model.c_model(), pts, tb, tgt_root_pos.unsqueeze(0),
                                    tgt_root_rot_quat.unsqueeze(0), tgt_joint_rot.unsqueeze(0),
                                    contacts.unsqueeze(0), w_penetration, w_contact)
    return total[0], pen[0], con[0]


def _loss_terms(tgt_root_pos, tgt_root_rot, tgt_joint_dof, src_root_pos, src_root_rot_quat, src_joint_rot,
                src_body_vels, src_body_rot_vels, contacts, body_points, char_model, w, body_constraints, max_jerk,
                tb, pts):
    """All terms as tensors (no host synchronisation, graph-capturable).  `w` = dict of the nine weights.
    Returns (weighted loss, dict LossType -> 0-dim tensor or python number)."""
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
    w_contact = w["w_contact"]
    hf_total, penetration_loss, contact_loss = pen_contact_loss(
        tgt_root_pos, tgt_root_rot_quat, tgt_joint_rot, contacts, None, body_points, char_model, w["w_penetration"],
        w_contact, _tb=tb, _pts=pts)
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
                # a constrained body does not also pay for sliding / smoothness on those frames
                body_vel_err_sq = body_vel_err_sq.clone()
                body_vel_err_sq[s:e + 1, b] *= 0.0
                body_rot_vel_err_sq = body_rot_vel_err_sq.clone()
                body_rot_vel_err_sq[s:e + 1, b] *= 0.0

    # --- sliding: pseudo-Huber on contact bodies (:339-346) ---
    if w["w_sliding"] != 0.0:
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

    terms = {
        LossType.ROOT_POS_LOSS: root_pos_loss, LossType.ROOT_ROT_LOSS: root_rot_loss,
        LossType.JOINT_ROT_LOSS: joint_rot_loss, LossType.SMOOTHNESS_LOSS: smoothness_loss,
        LossType.PENETRATION_LOSS: penetration_loss, LossType.CONTACT_LOSS: contact_logged,
        LossType.SLIDING_LOSS: sliding_loss, LossType.JERK_LOSS: jerk_loss,
        LossType.BODY_CONSTRAINT_LOSS: body_constraint_loss,
    }
    loss = w["w_root_pos"] * root_pos_loss + w["w_root_rot"] * root_rot_loss + w["w_joint_rot"] * joint_rot_loss \
        + w["w_smoothness"] * smoothness_loss + hf_total + w["w_sliding"] * sliding_loss \
        + w["w_body_constraints"] * body_constraint_loss + w["w_jerk"] * jerk_loss
    return loss, terms


def _to_floats(terms):
    return {k: (v.item() if isinstance(v, torch.Tensor) else v) for k, v in terms.items()}


def motion_terrain_contact_loss(tgt_root_pos, tgt_root_rot, tgt_joint_dof, src_root_pos, src_root_rot_quat,
                                src_joint_rot, src_body_vels, src_body_rot_vels, contacts, terrain,
                                body_points: List[torch.Tensor], char_model, w_root_pos: float, w_root_rot: float,
                                w_joint_rot: float, w_smoothness: float, w_penetration: float, w_contact: float,
                                w_sliding: float, w_body_constraints: float, w_jerk: float, body_constraints: list,
                                max_jerk: float):
    """-> (loss tensor, dict LossType -> float).  Ref :183-395."""
    w = dict(w_root_pos=w_root_pos, w_root_rot=w_root_rot, w_joint_rot=w_joint_rot, w_smoothness=w_smoothness,
             w_penetration=w_penetration, w_contact=w_contact, w_sliding=w_sliding,
             w_body_constraints=w_body_constraints, w_jerk=w_jerk)
    loss, terms = _loss_terms(tgt_root_pos, tgt_root_rot, tgt_joint_dof, src_root_pos, src_root_rot_quat,
                              src_joint_rot, src_body_vels, src_body_rot_vels, contacts, body_points, char_model, w,
                              body_constraints, max_jerk, _terrain_desc(terrain), body_points_desc(char_model, body_points))
    return loss, _to_floats(terms)


class _TextLogger:
    """Minimal stand-in for the reference's WandbLogger (wandb / tensorboard plumbing is out of scope): prints the
    table and, if asked, appends it to `log_file`."""

    def __init__(self, log_file: Optional[str]):
        self._rows = []
        self._file = open(log_file, "w") if log_file is not None else None

    def log(self, key, val):
        self._rows.append((key, val))

    def flush(self, quiet=False):
        line = "  ".join(f"{k}={v:.6g}" if isinstance(v, float) else f"{k}={v}" for k, v in self._rows)
        if not quiet:
            print(line)
        if self._file is not None:
            self._file.write(line + "\n")
        self._rows = []

    def close(self):
        if self._file is not None:
            self._file.close()


def motion_contact_optimization(src_frames: torch.Tensor, contacts: torch.Tensor, body_points: list, terrain,
                                char_model, num_iters: int, step_size: float, w_root_pos: float, w_root_rot: float,
                                w_joint_rot: float, w_smoothness: float, w_penetration: float, w_contact: float,
                                w_sliding: float, w_body_constraints: float, w_jerk: float, body_constraints: list,
                                max_jerk: float, exp_name: str = "", use_wandb: bool = False, log_file: str = None,
                                use_cuda_graph: bool = True, quiet: bool = False):
    """Adam over (root_pos, root exp-map, joint DoFs) so that the clip stops penetrating / floating above the
    terrain while staying close to `src_frames`.  -> optimised frames [F, 6+D].  Ref :404-500.
    `use_wandb` is accepted and ignored (logging back ends are out of scope)."""
    start_time = time.time()
    D = char_model.get_dof_size()
    src_root_pos = src_frames[:, 0:3]
    src_root_rot = src_frames[:, 3:6]
    src_joint_dof = src_frames[:, 6:6 + D]
    with torch.no_grad():
        src_root_rot_quat = ops.exp_map_to_quat(src_root_rot)
        src_joint_rot = char_model.dof_to_rot(src_joint_dof)
        src_body_pos, src_body_rot = char_model.forward_kinematics(src_root_pos, src_root_rot_quat, src_joint_rot)
        src_body_vels = src_body_pos[1:] - src_body_pos[:-1]
        src_body_rot_vels = torch_util.quat_diff_angle(src_body_rot[1:], src_body_rot[:-1])

    leaves = [src_root_pos.clone().requires_grad_(True), src_root_rot.clone().requires_grad_(True),
              src_joint_dof.clone().requires_grad_(True)]
    w = dict(w_root_pos=w_root_pos, w_root_rot=w_root_rot, w_joint_rot=w_joint_rot, w_smoothness=w_smoothness,
             w_penetration=w_penetration, w_contact=w_contact, w_sliding=w_sliding,
             w_body_constraints=w_body_constraints, w_jerk=w_jerk)
    tb = _terrain_desc(terrain)
    pts = body_points_desc(char_model, body_points)
    # capturable=True keeps Adam's step counter on the device so optimizer.step() can live inside a CUDA graph
    optimizer = torch.optim.Adam(leaves, lr=step_size, capturable=use_cuda_graph)

    def iteration():
        optimizer.zero_grad(set_to_none=False)
        loss, terms = _loss_terms(leaves[0], leaves[1], leaves[2], src_root_pos, src_root_rot_quat, src_joint_rot,
                                  src_body_vels, src_body_rot_vels, contacts, body_points, char_model, w,
                                  body_constraints, max_jerk, tb, pts)
        loss.backward()
        optimizer.step()
        return loss, terms

    logger = _TextLogger(log_file)
    log_iter_stride = 25

    def log(it, loss, terms):
        logger.log("Iteration", it)
        logger.log("Time (min)", (time.time() - start_time) / 60.0)
        logger.log("TOTAL WEIGHTED LOSS", loss.item())
        for key, val in _to_floats(terms).items():
            logger.log(key.name, val)
        logger.flush(quiet)

    it = 0
    if use_cuda_graph and num_iters > 3:
        # two eager iterations on a side stream (allocator warm-up, Adam state creation), then capture one
        side = torch.cuda.Stream(device=src_frames.device)
        side.wait_stream(torch.cuda.current_stream(src_frames.device))
        with torch.cuda.stream(side):
            for _ in range(2):
                loss, terms = iteration()
                if it % log_iter_stride == 0:
                    log(it, loss, terms)
                it += 1
        torch.cuda.current_stream(src_frames.device).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            g_loss, g_terms = iteration()
        while it < num_iters:                     # capture records but does not execute: replay #1 is iteration 2
            graph.replay()
            if it % log_iter_stride == 0:
                log(it, g_loss, g_terms)
            it += 1
    else:
        while it < num_iters:
            loss, terms = iteration()
            if it % log_iter_stride == 0:
                log(it, loss, terms)
            it += 1
    logger.close()
    return torch.cat([t.detach() for t in leaves], dim=-1)
