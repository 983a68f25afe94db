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

from ... import _lib, ops
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


WEIGHT_KEYS = ("w_root_pos", "w_root_rot", "w_joint_rot", "w_smoothness", "w_penetration", "w_contact", "w_sliding",
               "w_body_constraints", "w_jerk")
_DT = 1.0 / 30.0          # the reference hard-codes the frame time of the jerk term (:355)


def _constraint_records(body_constraints, char_model, body_points):
    """python BodyConstraint lists (one list per body) -> ParcBodyConstraint records, as :286-332 reads them: only
    bodies whose first geom is a sphere or a box take part; a box body uses its first 18 surface points (the sole)."""
    recs = []
    if body_constraints is None:
        return recs
    for b in range(char_model.get_num_joints()):
        if len(body_constraints[b]) == 0:
            continue
        geom0 = char_model.get_geoms(b)[0]
        if geom0._shape_type == kin_char_model.GeomType.SPHERE:
            shape, radius = _lib.PARC_CONSTRAINT_SPHERE, float(geom0._dims.reshape(-1)[0].item())
            offset = geom0._offset.detach().cpu().tolist()
        elif geom0._shape_type == kin_char_model.GeomType.BOX:
            shape, radius = _lib.PARC_CONSTRAINT_BOX, float((torch.norm(geom0._dims) * 1.25).item())
            offset = [0.0, 0.0, 0.0]
            assert body_points[b].shape[0] >= 18, "a box-constrained body needs its 18 sole points first (:320)"
        else:
            continue
        for c in body_constraints[b]:
            r = _lib.ParcBodyConstraint()
            r.body, r.start_frame, r.end_frame, r.shape = b, int(c.start_frame_idx), int(c.end_frame_idx), shape
            pt = c.constraint_point.detach().cpu().tolist()
            for k in range(3):
                r.point[k] = float(pt[k])
                r.offset[k] = float(offset[k])
            r.radius = radius
            recs.append(r)
    return recs


def source_constants(src_frames, char_model):
    """(src_root_rot_quat [F,4], src_joint_rot [F,J-1,4], src_body_vels [F-1,J,3], src_body_rot_vels [F-1,J]) of a
    source clip [F, 6+D] -- motion_optimization.py:428-436 -- computed by the same device code the objective runs on
    the leaves (include/parc_b200.h: parc_motion_opt_source), so "target == source" has errors of exactly 0."""
    import ctypes as C
    model = char_model.c_model()
    J, D = model.num_bodies, model.dof_size
    src = _lib.f32c(src_frames.detach()[:, 0:6 + D])
    _lib.require_cuda(src)
    F, dev = int(src.shape[0]), src.device
    z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)
    rq, jr, bp, br = z(F, 4), z(F, J - 1, 4), z(F, J, 3), z(F, J, 4)
    bv, brv = z(max(F - 1, 0), J, 3), z(max(F - 1, 0), J)
    with torch.cuda.device(dev):
        rc = _lib.load().parc_motion_opt_source(src.data_ptr(), F, C.byref(model), rq.data_ptr(), jr.data_ptr(),
                                                bp.data_ptr(), br.data_ptr(), bv.data_ptr(), brv.data_ptr(),
                                                _lib.stream_ptr(dev))
    _lib.LAUNCHES[0] += 1
    _lib.check(rc, "parc_motion_opt_source")
    return rq, jr, bv, brv


class MotionOptPlan:
    """Everything one clip's objective needs, resident on the device, with the argument block of the C entry points
    (include/parc_b200.h: ParcMotionOptArgs) prebuilt: `loss_grad()` is three launches, `iteration()` four, each ONE
    ctypes call with no host synchronisation -- so a whole iteration can be captured in a CUDA graph.

    frames [F, 6+D] holds the leaves (root position | root exp-map | joint DoFs) and is updated in place."""

    def __init__(self, frames, src_root_pos, src_root_rot_quat, src_joint_rot, src_body_vels, src_body_rot_vels,
                 contacts, terrain, body_points, char_model, w: dict, body_constraints, max_jerk: float,
                 step_size: float = 0.0, with_adam: bool = False, terrain_desc=None):
        dev = frames.device
        _lib.require_cuda(frames, src_root_pos, src_root_rot_quat, src_joint_rot, contacts)
        self.model = char_model.c_model()
        J, D = self.model.num_bodies, self.model.dof_size
        F = int(frames.shape[0])
        assert frames.shape == (F, 6 + D) and frames.dtype == torch.float32 and frames.is_contiguous()
        f32 = lambda t, shape: _lib.f32c(t.detach()).reshape(shape)
        self.frames = frames
        self.w = {k: float(w[k]) for k in WEIGHT_KEYS}
        self._src = (f32(src_root_pos, (F, 3)), f32(src_root_rot_quat, (F, 4)), f32(src_joint_rot, (F, J - 1, 4)),
                     f32(src_body_vels, (max(F - 1, 0), J, 3)), f32(src_body_rot_vels, (max(F - 1, 0), J)),
                     f32(contacts, (F, J)))
        self._tb = terrain_desc if terrain_desc is not None else _terrain_desc(terrain)
        self._tbs = self._tb.c_struct(1)
        self._pts = body_points_desc(char_model, body_points)
        recs = _constraint_records(body_constraints, char_model, body_points)
        self._recs = None
        if recs:
            import ctypes as C
            arr = (_lib.ParcBodyConstraint * len(recs))(*recs)
            self._recs = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)
        self.root_rot, self.joint_rot = z(F, 4), z(F, J - 1, 4)
        self.body_pos, self.body_rot = z(F, J, 3), z(F, J, 4)
        self._g = (z(F, 3), z(F, 4), z(F, J - 1, 4))
        self.pen, self.con = z(F), z(F)
        self.grad, self.terms = z(F, 6 + D), z(F, 8)
        self.exp_avg = self.exp_avg_sq = self.step = None
        if with_adam:
            self.exp_avg, self.exp_avg_sq = z(F, 6 + D), z(F, 6 + D)
            self.step = torch.zeros(1, dtype=torch.int32, device=dev)
        a = _lib.ParcMotionOptArgs()
        a.frames, a.num_frames = frames.data_ptr(), F
        (a.src_root_pos, a.src_root_rot, a.src_joint_rot, a.src_body_vels, a.src_body_rot_vels,
         a.contacts) = [t.data_ptr() for t in self._src]
        import ctypes as C
        a.terrain = C.addressof(self._tbs)
        a.pts = self._pts.c_struct()
        a.constraints, a.num_constraints = _lib.ptr(self._recs), len(recs)
        for k in WEIGHT_KEYS:
            setattr(a, k, self.w[k])
        a.max_jerk_dt3 = float(max_jerk) * (_DT ** 3)
        a.lr, a.beta1, a.beta2, a.eps = float(step_size), 0.9, 0.999, 1e-8
        a.exp_avg, a.exp_avg_sq, a.step = _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq), _lib.ptr(self.step)
        a.root_rot, a.joint_rot, a.body_pos, a.body_rot = (self.root_rot.data_ptr(), self.joint_rot.data_ptr(),
                                                           self.body_pos.data_ptr(), self.body_rot.data_ptr())
        a.g_root_pos, a.g_root_rot, a.g_joint_rot = [t.data_ptr() for t in self._g]
        a.pen, a.con, a.grad, a.terms = (self.pen.data_ptr(), self.con.data_ptr(), self.grad.data_ptr(),
                                         self.terms.data_ptr())
        self._args = a
        self._lib = _lib.load()
        self._has_constraints = bool(recs)
        self.device = dev

    def _call(self, fn, name, launches):
        import ctypes as C
        with torch.cuda.device(self.device):
            rc = fn(C.byref(self._args), C.byref(self.model), _lib.stream_ptr(self.device))
        _lib.LAUNCHES[0] += launches - 1
        _lib.check(rc, name)

    def loss_grad(self):
        """Objective + gradient at the current `frames`: fills grad / terms / pen / con.  3 launches."""
        self._call(self._lib.parc_motion_opt_loss_grad, "parc_motion_opt_loss_grad", 3)

    def adam_step(self):
        self._call(self._lib.parc_motion_opt_adam_step, "parc_motion_opt_adam_step", 1)

    def iteration(self):
        """loss_grad + Adam update of `frames`: 4 launches, one C call, no host synchronisation."""
        self._call(self._lib.parc_motion_opt_iteration, "parc_motion_opt_iteration", 4)

    def term_tensors(self):
        """(weighted loss, dict LossType -> 0-dim tensor or python number) of the LAST loss_grad / iteration --
        a few reductions over [F] vectors, only run when somebody looks (logging iterations)."""
        t = self.terms.sum(dim=0)
        pen, con = self.pen.sum(), self.con.sum()
        w = self.w
        terms = {
            LossType.ROOT_POS_LOSS: t[0], LossType.ROOT_ROT_LOSS: t[1], LossType.JOINT_ROT_LOSS: t[2],
            LossType.SMOOTHNESS_LOSS: t[3], LossType.PENETRATION_LOSS: pen,
            LossType.CONTACT_LOSS: con if w["w_contact"] != 0.0 else 0.0,
            LossType.SLIDING_LOSS: t[4] if w["w_sliding"] != 0.0 else 0.0, LossType.JERK_LOSS: t[5],
            LossType.BODY_CONSTRAINT_LOSS: t[6] if self._has_constraints else 0.0,
        }
        loss = (w["w_root_pos"] * t[0] + w["w_root_rot"] * t[1] + w["w_joint_rot"] * t[2] + w["w_smoothness"] * t[3]
                + w["w_penetration"] * pen + w["w_contact"] * con + w["w_sliding"] * t[4]
                + w["w_body_constraints"] * t[6] + w["w_jerk"] * t[5])
        return loss, terms


class _MotionOptLoss(torch.autograd.Function):
    """The nine-term objective as ONE differentiable operator: forward runs MotionOptPlan.loss_grad (value and
    gradient come out of the same three launches), backward scales the stored gradient."""

    @staticmethod
    def forward(ctx, tgt_root_pos, tgt_root_rot, tgt_joint_dof, plan_args):
        frames = torch.cat([tgt_root_pos, tgt_root_rot, tgt_joint_dof], dim=-1).detach().to(torch.float32).contiguous()
        plan = MotionOptPlan(frames, *plan_args)
        plan.loss_grad()
        loss, terms = plan.term_tensors()
        ctx.save_for_backward(plan.grad)
        ctx.D = tgt_joint_dof.shape[-1]
        ctx.mark_non_differentiable(plan.terms, plan.pen, plan.con)
        return loss, plan.terms, plan.pen, plan.con

    @staticmethod
    def backward(ctx, g, _gt, _gp, _gc):
        (grad,) = ctx.saved_tensors
        gg = grad * g
        return gg[:, 0:3], gg[:, 3:6], gg[:, 6:6 + ctx.D], None


def _to_floats(terms):
    return {k: (v.item() if isinstance(v, torch.Tensor) else v) for k, v in terms.items()}


def motion_terrain_contact_loss(tgt_root_pos, tgt_root_rot, tgt_joint_dof, src_root_pos, src_root_rot_quat,
                                src_joint_rot, src_body_vels, src_body_rot_vels, contacts, terrain,
                                body_points: List[torch.Tensor], char_model, w_root_pos: float, w_root_rot: float,
                                w_joint_rot: float, w_smoothness: float, w_penetration: float, w_contact: float,
                                w_sliding: float, w_body_constraints: float, w_jerk: float, body_constraints: list,
                                max_jerk: float):
    """-> (loss tensor, dict LossType -> float); `loss.backward()` reaches tgt_root_pos / tgt_root_rot / tgt_joint_dof
    exactly as autograd does in the reference.  Three launches forward + backward together.  Ref :183-395."""
    w = dict(w_root_pos=w_root_pos, w_root_rot=w_root_rot, w_joint_rot=w_joint_rot, w_smoothness=w_smoothness,
             w_penetration=w_penetration, w_contact=w_contact, w_sliding=w_sliding,
             w_body_constraints=w_body_constraints, w_jerk=w_jerk)
    plan_args = (src_root_pos, src_root_rot_quat, src_joint_rot, src_body_vels, src_body_rot_vels, contacts, terrain,
                 body_points, char_model, w, body_constraints, max_jerk)
    loss, terms_f, pen, con = _MotionOptLoss.apply(tgt_root_pos, tgt_root_rot, tgt_joint_dof, plan_args)
    t = terms_f.detach().sum(dim=0)
    has_bc = body_constraints is not None and any(len(c) > 0 for c in body_constraints)
    terms = {
        LossType.ROOT_POS_LOSS: t[0], LossType.ROOT_ROT_LOSS: t[1], LossType.JOINT_ROT_LOSS: t[2],
        LossType.SMOOTHNESS_LOSS: t[3], LossType.PENETRATION_LOSS: pen.detach().sum(),
        LossType.CONTACT_LOSS: con.detach().sum() if w_contact != 0.0 else 0.0,
        LossType.SLIDING_LOSS: t[4] if w_sliding != 0.0 else 0.0, LossType.JERK_LOSS: t[5],
        LossType.BODY_CONSTRAINT_LOSS: t[6] if has_bc else 0.0,
    }
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

    One iteration = MotionOptPlan.iteration(): four launches (FK of the leaves; penetration / contact terms + gradient;
    every other term + the whole backward pass; the Adam update), replayed as a CUDA graph; the host synchronises
    only on logging iterations (every 25th), where the reference calls `.item()` on seven terms every iteration.
    `use_cuda_graph=False` enqueues the same four launches from Python -- results are bit-identical.
    `use_wandb` is accepted and ignored (logging back ends are out of scope)."""
    start_time = time.time()
    D = char_model.get_dof_size()
    src = src_frames[:, 0:6 + D].detach().to(torch.float32).contiguous()
    src_root_rot_quat, src_joint_rot, src_body_vels, src_body_rot_vels = source_constants(src, char_model)
    w = dict(w_root_pos=w_root_pos, w_root_rot=w_root_rot, w_joint_rot=w_joint_rot, w_smoothness=w_smoothness,
             w_penetration=w_penetration, w_contact=w_contact, w_sliding=w_sliding,
             w_body_constraints=w_body_constraints, w_jerk=w_jerk)
    frames = src.clone()
    plan = MotionOptPlan(frames, src[:, 0:3], src_root_rot_quat, src_joint_rot, src_body_vels, src_body_rot_vels,
                         contacts, terrain, body_points, char_model, w, body_constraints, max_jerk, step_size=step_size,
                         with_adam=True)
    logger = _TextLogger(log_file)
    log_iter_stride = 25

    def log(it):
        loss, terms = plan.term_tensors()
        logger.log("Iteration", it)
        logger.log("Time (min)", (time.time() - start_time) / 60.0)
        logger.log("TOTAL WEIGHTED LOSS", loss.item())
        for key, val in _to_floats(terms).items():
            logger.log(key.name, val)
        logger.flush(quiet)

    it = 0
    dev = frames.device
    if use_cuda_graph and num_iters > 3:
        # one eager iteration on a side stream (module load, first-use work), then capture the next
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            plan.iteration()
        torch.cuda.current_stream(dev).wait_stream(side)
        log(it)
        it += 1
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            plan.iteration()
        while it < num_iters:                     # capture records but does not execute: replay #1 is iteration 1
            graph.replay()
            if it % log_iter_stride == 0:
                log(it)
            it += 1
    else:
        while it < num_iters:
            plan.iteration()
            if it % log_iter_stride == 0:
                log(it)
            it += 1
    logger.close()
    return frames
