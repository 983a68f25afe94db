"""Operator layer: torch tensors in, libparc_b200 C-ABI calls, torch tensors out.

Every function here launches hand-written sm_100a kernels on the current CUDA stream through the C ABI
of include/parc_b200.h.  Tensors must live on a CUDA device -- there is no CPU path.  Differentiable
operators are `torch.autograd.Function`s whose backward calls the matching VJP kernel.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import (ParcQueryArgs, ParcBodyPoints, ParcCharModel, ParcCharState, ParcDoneSpec, ParcSimStep, ParcKeyBodies, ParcClipMeta, ParcFkOut,
                   ParcFrameOut, ParcHeightfield, ParcMotionTables, ParcObsSpec, ParcRowLayout, ParcTerrainBatch, check,
                   f32c, ptr, require_cuda, stream_ptr)


# ----------------------------------------------------------------------------------------------
# model / layout helpers (host only)
# ----------------------------------------------------------------------------------------------
def make_char_model(parents, local_trans, local_rot, joint_types, joint_axes, dof_idx) -> ParcCharModel:
    """Fill the POD the kernels take.  Arguments are python lists / nested lists (host values)."""
    J = len(parents)
    if J > _lib.PARC_MAX_BODIES:
        raise _lib.ParcLibraryError(f"character has {J} bodies; libparc_b200 supports <= {_lib.PARC_MAX_BODIES}")
    m = ParcCharModel()
    m.num_bodies = J
    depth = [0] * J
    dof = 0
    for b in range(J):
        m.parent[b] = int(parents[b])
        if b > 0:
            depth[b] = depth[int(parents[b])] + 1
        m.depth[b] = depth[b]
        jt = int(joint_types[b])
        m.joint_type[b] = jt
        m.dof_idx[b] = int(dof_idx[b])
        dof += 1 if jt == _lib.JOINT_HINGE else (3 if jt == _lib.JOINT_SPHERICAL else 0)
        for k in range(3):
            m.local_trans[b][k] = float(local_trans[b][k])
            m.joint_axis[b][k] = float(joint_axes[b][k])
        for k in range(4):
            m.local_rot[b][k] = float(local_rot[b][k])
    m.dof_size = dof
    m.max_depth = max(depth)
    check(_lib.load().parc_validate_model(C.byref(m)), "parc_validate_model")
    return m


def row_layout(model: ParcCharModel) -> ParcRowLayout:
    lay = ParcRowLayout()
    check(_lib.load().parc_row_layout(C.byref(model), C.byref(lay)), "parc_row_layout")
    return lay


def make_tree(model: ParcCharModel, device) -> torch.Tensor:
    """The compact kinematic tree the query kernel stages into shared memory, as a device byte tensor
    (include/parc_b200.h: PARC_TREE_BYTES).  Keeps the 1.4 KB model out of the kernel's parameter block."""
    host = (C.c_uint8 * _lib.PARC_TREE_BYTES)()
    check(_lib.load().parc_tree_from_model(C.byref(model), host), "parc_tree_from_model")
    return torch.frombuffer(bytearray(host), dtype=torch.uint8).to(device)


@dataclass
class PackedTables:
    """Device-resident packed frame rows + clip metadata (include/parc_b200.h: ParcMotionTables)."""
    rows: torch.Tensor        # f32 [T, row_floats]
    clips: torch.Tensor       # uint8 [M, 32] (ParcClipMeta records)
    total_frames: int
    num_clips: int
    layout: ParcRowLayout
    tree: Optional[torch.Tensor] = None     # uint8 [PARC_TREE_BYTES], made on first use from the model

    def c_struct(self, model: Optional[ParcCharModel] = None) -> ParcMotionTables:
        if self.tree is None and model is not None:
            self.tree = make_tree(model, self.rows.device)
        t = ParcMotionTables()
        t.rows = self.rows.data_ptr()
        t.clips = self.clips.data_ptr()
        t.total_frames = self.total_frames
        t.num_clips = self.num_clips
        t.row_floats = self.layout.row_floats
        t.tree = ptr(self.tree)
        return t


def build_clip_meta(num_frames, loop_modes, start_idx, lengths, root_pos_delta, device) -> torch.Tensor:
    """Host tensors -> [M,32] uint8 device tensor of ParcClipMeta records."""
    import numpy as np
    M = int(num_frames.shape[0])
    rec = np.zeros(M, dtype=np.dtype([("num_frames", "<i4"), ("loop_mode", "<i4"), ("start_idx", "<i8"),
                                      ("length", "<f4"), ("delta", "<f4", (3,))]))
    assert rec.dtype.itemsize == C.sizeof(ParcClipMeta) == 32
    rec["num_frames"] = num_frames.cpu().numpy()
    rec["loop_mode"] = loop_modes.cpu().numpy()
    rec["start_idx"] = start_idx.cpu().numpy()
    rec["length"] = lengths.cpu().numpy()
    rec["delta"] = root_pos_delta.cpu().numpy().reshape(M, 3)
    raw = torch.from_numpy(rec.view(np.uint8).reshape(M, 32).copy())
    return raw.to(device)


def pack_frames(model: ParcCharModel, root_pos, root_rot, joint_rot, contacts, root_vel, root_ang_vel,
                dof_vel) -> Tuple[torch.Tensor, ParcRowLayout]:
    """a1: interleave the per-frame tables (device tensors) into packed rows on the GPU."""
    require_cuda(root_pos, root_rot, joint_rot, root_vel, root_ang_vel, dof_vel, contacts)
    lay = row_layout(model)
    T = root_pos.shape[0]
    rows = torch.empty((T, lay.row_floats), dtype=torch.float32, device=root_pos.device)
    args = [f32c(t) if t is not None else None
            for t in (root_pos, root_rot, joint_rot, contacts, root_vel, root_ang_vel, dof_vel)]
    with torch.cuda.device(root_pos.device):
        rc = _lib.load().parc_pack_frames(*[ptr(a) for a in args], T, C.byref(model), rows.data_ptr(),
                                          stream_ptr(root_pos.device))
    check(rc, "parc_pack_frames")
    return rows, lay


# ----------------------------------------------------------------------------------------------
# heightfield descriptors
# ----------------------------------------------------------------------------------------------
def build_tables(model: ParcCharModel, frames: torch.Tensor, contacts: Optional[torch.Tensor],
                 num_frames: torch.Tensor, fps: torch.Tensor, dof_vel_dt: torch.Tensor):
    """GPU loader: raw frames [total, >= 6+D] of M concatenated clips (`num_frames` [M] int64, `fps` /
    `dof_vel_dt` [M] fp32, all CUDA) -> (packed rows [total, row_floats], layout, start_idx [M]).  One launch
    (csrc/table_build.cu); see include/parc_b200.h::parc_build_tables."""
    fr = f32c(frames)
    ct = f32c(contacts) if contacts is not None else None
    require_cuda(fr, ct, num_frames, fps, dof_vel_dt)
    assert fr.dim() == 2
    total, M = int(fr.shape[0]), int(num_frames.shape[0])
    dev = fr.device
    nf = num_frames.to(torch.int64).contiguous()
    start = torch.cumsum(nf, 0) - nf
    frame_clip = torch.repeat_interleave(torch.arange(M, dtype=torch.int32, device=dev), nf, output_size=total)
    lay = row_layout(model)
    rows = torch.empty((total, lay.row_floats), dtype=torch.float32, device=dev)
    fps_c, dt_c = f32c(fps), f32c(dof_vel_dt)
    if ct is not None:
        assert tuple(ct.shape) == (total, model.num_bodies)
    with torch.cuda.device(dev):
        rc = _lib.load().parc_build_tables(fr.data_ptr(), total, int(fr.shape[1]), ptr(ct), frame_clip.data_ptr(),
                                           start.data_ptr(), nf.data_ptr(), fps_c.data_ptr(), dt_c.data_ptr(), M,
                                           C.byref(model), rows.data_ptr(), stream_ptr(dev))
    check(rc, "parc_build_tables")
    return rows, lay, start


def unpack_row_views(rows: torch.Tensor, lay: ParcRowLayout, num_bodies: int, dof_size: int) -> dict:
    """The reference's per-frame tables as zero-copy strided views of the packed rows."""
    J, D = num_bodies, dof_size
    cf, pf = lay.contact_slot * 4, lay.pose_slots * 4
    total = rows.shape[0]
    return dict(root_pos=rows[:, 0:3], root_rot=rows[:, 4:8], joint_rot=rows[:, 8:8 + 4 * (J - 1)].view(total, J - 1, 4),
                contacts=rows[:, cf:cf + J], root_vel=rows[:, pf:pf + 3], root_ang_vel=rows[:, pf + 4:pf + 7],
                dof_vel=rows[:, pf + 8:pf + 8 + D])


@dataclass
class HeightfieldDesc:
    """Host copy of SubTerrain's scalars + the device hf tensor, so launches never sync."""
    hf: torch.Tensor
    min_x: float
    min_y: float
    dx: float
    dy: float

    def c_struct(self) -> ParcHeightfield:
        assert self.hf.is_cuda and self.hf.dtype == torch.float32 and self.hf.dim() == 2 and self.hf.is_contiguous(), \
            "heightfield must be a contiguous fp32 CUDA tensor [X, Y]"
        h = ParcHeightfield()
        h.hf = self.hf.data_ptr()
        h.dim_x, h.dim_y = int(self.hf.shape[0]), int(self.hf.shape[1])
        h.min_x, h.min_y, h.dx, h.dy = self.min_x, self.min_y, self.dx, self.dy
        return h


def _obs_struct(tmpl_xy: torch.Tensor, relative: bool, min_h: float, max_h: float) -> ParcObsSpec:
    o = ParcObsSpec()
    o.tmpl_xy = tmpl_xy.data_ptr()
    o.num_points = int(tmpl_xy.shape[0])
    o.relative = 1 if relative else 0
    o.min_h, o.max_h = float(min_h), float(max_h)
    return o


# ----------------------------------------------------------------------------------------------
# a2-a4 (+a6, +a10): frame query, optionally fused with FK and the heightmap observation
# ----------------------------------------------------------------------------------------------
def motion_query(tables: PackedTables, model: ParcCharModel, motion_ids: torch.Tensor,
                 motion_times: Optional[torch.Tensor] = None, frame_idxs: Optional[torch.Tensor] = None, *,
                 want_frame: bool = True, want_contacts: bool = True, want_index: bool = False,
                 want_fk: bool = False, hf: Optional[HeightfieldDesc] = None,
                 obs_tmpl: Optional[torch.Tensor] = None, obs_relative: bool = True, obs_min_h: float = -3.0,
                 obs_max_h: float = 3.0, out: Optional[dict] = None, fast_heading: bool = False,
                 error_flags: Optional[torch.Tensor] = None) -> dict:
    """One launch of the fused kernel.  Exactly one of motion_times (blended query,
    MotionLib.calc_motion_frame) or frame_idxs (MotionLib.get_motion_frame) must be given.
    Returns a dict of output tensors; `out` may carry preallocated tensors to reuse.
    `error_flags` (int32 [1] on the ids' device): the kernel ORs PARC_QUERY_ERR_* bits into it when a clip id / frame
    index is out of range (see `raise_query_errors`)."""
    require_cuda(motion_ids, motion_times, frame_idxs)
    dev = motion_ids.device
    # ids may carry any leading shape (the reference gathers with whatever it is handed): flatten here, restore below
    lead = tuple(motion_ids.shape)
    if len(lead) != 1:
        other = motion_times if motion_times is not None else frame_idxs
        motion_ids, other = torch.broadcast_tensors(motion_ids, other)
        lead = tuple(motion_ids.shape)
        motion_ids = motion_ids.reshape(-1)
        if motion_times is not None:
            motion_times = other.reshape(-1)
        else:
            frame_idxs = other.reshape(-1)
    N = int(motion_ids.shape[0])
    J, D = model.num_bodies, model.dof_size
    ids = motion_ids.to(torch.int64).contiguous()
    res = {} if out is None else out

    def buf(name, shape, dtype=torch.float32):
        t = res.get(name)
        if (t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype or t.device != dev
                or not t.is_contiguous()):
            t = torch.empty(shape, dtype=dtype, device=dev)
            res[name] = t
        return t

    fo = ParcFrameOut()
    if want_frame:
        fo.root_pos = buf("root_pos", (N, 3)).data_ptr()
        fo.root_rot = buf("root_rot", (N, 4)).data_ptr()
        fo.root_vel = buf("root_vel", (N, 3)).data_ptr()
        fo.root_ang_vel = buf("root_ang_vel", (N, 3)).data_ptr()
        fo.joint_rot = buf("joint_rot", (N, J - 1, 4)).data_ptr()
        fo.dof_vel = buf("dof_vel", (N, D)).data_ptr()
        if want_contacts:
            fo.contacts = buf("contacts", (N, J)).data_ptr()
    if want_index:
        fo.frame_idx0 = buf("frame_idx0", (N,), torch.int64).data_ptr()
        fo.frame_idx1 = buf("frame_idx1", (N,), torch.int64).data_ptr()
        fo.blend = buf("blend", (N,)).data_ptr()
    fk = ParcFkOut()
    if want_fk:
        fk.body_pos = buf("body_pos", (N, J, 3)).data_ptr()
        fk.body_rot = buf("body_rot", (N, J, 4)).data_ptr()
    tb = tables.c_struct(model)
    qa = ParcQueryArgs()
    qa.tables, qa.model, qa.frame = C.addressof(tb), C.addressof(model), C.addressof(fo)
    qa.motion_ids, qa.n = ids.data_ptr(), N
    if want_fk:
        qa.fk = C.addressof(fk)
    qa.flags = _lib.PARC_QUERY_FAST_HEADING if fast_heading else 0
    if error_flags is not None:
        assert error_flags.dtype == torch.int32 and error_flags.device == dev
        qa.error_flags = error_flags.data_ptr()
    keep = [tb, fo, fk, ids]
    if motion_times is not None:
        assert frame_idxs is None
        times = f32c(motion_times)
        qa.motion_times = times.data_ptr()
        keep.append(times)
        if obs_tmpl is not None:
            assert hf is not None
            require_cuda(hf.hf, obs_tmpl)
            tm = f32c(obs_tmpl)      # if this made a copy, the allocator's stream ordering keeps it valid
            hfs, obs = hf.c_struct(), _obs_struct(tm, obs_relative, obs_min_h, obs_max_h)
            qa.hf, qa.obs = C.addressof(hfs), C.addressof(obs)
            qa.obs_out = buf("obs", (N, int(tm.shape[0]))).data_ptr()
            keep += [tm, hfs, obs]
    else:
        fi = frame_idxs.to(torch.int64).contiguous()
        qa.frame_idxs = fi.data_ptr()
        keep.append(fi)
    with torch.cuda.device(dev):
        rc = _lib.load().parc_motion_query_ex(C.byref(qa), stream_ptr(dev))
    check(rc, "parc_motion_query_ex")
    del keep
    if len(lead) != 1:
        return {k: v.reshape(*lead, *v.shape[1:]) for k, v in res.items()}
    return res


QUERY_OUTPUTS = ("root_pos", "root_rot", "root_vel", "root_ang_vel", "joint_rot", "dof_vel", "contacts", "body_pos",
                 "body_rot", "obs")


def raise_query_errors(error_flags: torch.Tensor, what: str = "motion query"):
    """Host check of the device error word a query ORs its PARC_QUERY_ERR_* bits into (synchronises).  Mirrors the
    reference, whose gathers raise IndexError on a clip id / frame index outside the tables."""
    bits = int(error_flags.item())
    if bits:
        error_flags.zero_()
        msgs = []
        if bits & _lib.PARC_QUERY_ERR_CLIP_ID:
            msgs.append("motion id out of range")
        if bits & _lib.PARC_QUERY_ERR_FRAME_IDX:
            msgs.append("frame index out of range for its clip")
        raise IndexError(f"{what}: " + "; ".join(msgs))


class MotionQueryPlan:
    """A fused query(+FK)(+obs) launch with every argument struct prebuilt: `launch()` is one ctypes call
    (~2 us of host time instead of ~40 us of Python), for callers that step the same buffers every control
    tick -- the tracker's per-step query is exactly that.  Inputs are read from `ids` / `times` (write new
    values into them in place, or build one plan per input buffer); outputs land in `self.out`.

    outputs: names out of QUERY_OUTPUTS to produce (default: all the plan's options allow) -- a caller that only
        consumes e.g. ("body_pos", "obs") saves the other stores and, end to end, their read-back.
    tar_obs: dict(sim_root_pos [E,3], sim_root_rot [E,4], key_body_ids, out [E, (S-1) * W] (may be a column block of
        the policy-observation row), global_obs, global_tar_root_h) -- the step form then also writes compute_tar_obs
        of the targets (steps 1..S-1) against the simulated character's root, from the kernel's registers.
    pdl: launch with programmatic stream serialisation (include/parc_b200.h: PARC_QUERY_PDL); with
        pdl_early_inputs=True the caller guarantees ids / times / offsets are not written by the previous kernel of
        the launch stream, and the whole read side overlaps that kernel's tail.
    variant: 0 = by batch size, 1..4 force an instantiation (tuning)."""

    def __init__(self, tables: PackedTables, model: ParcCharModel, ids: torch.Tensor, times: torch.Tensor, *,
                 want_contacts: bool = True, want_fk: bool = True, hf: Optional[HeightfieldDesc] = None,
                 obs_tmpl: Optional[torch.Tensor] = None, obs_relative: bool = True, obs_min_h: float = -3.0,
                 obs_max_h: float = 3.0, out: Optional[dict] = None, time_offsets: Optional[torch.Tensor] = None,
                 root_xy_offset: Optional[torch.Tensor] = None, outputs: Optional[Tuple[str, ...]] = None,
                 fast_heading: bool = False, pdl: bool = False, pdl_early_inputs: bool = False, variant: int = 0,
                 error_flags: Optional[torch.Tensor] = None, tar_obs: Optional[dict] = None):
        require_cuda(ids, times, tables.rows, time_offsets, root_xy_offset, error_flags)
        assert ids.dtype == torch.int64 and times.dtype == torch.float32 and ids.is_contiguous() and times.is_contiguous()
        assert ids.shape == times.shape and ids.dim() == 1
        self._keep = (tables, model, ids, times, hf, obs_tmpl, time_offsets, root_xy_offset, error_flags)
        self.device = ids.device
        E = int(ids.shape[0])                    # entries (environments)
        S = 1 if time_offsets is None else int(time_offsets.shape[0])
        N, J, D = E * S, model.num_bodies, model.dof_size
        self.n = N
        self.out = {} if out is None else out
        if outputs is not None:
            unknown = set(outputs) - set(QUERY_OUTPUTS)
            assert not unknown, f"unknown outputs {sorted(unknown)}"
        want = (lambda name: True) if outputs is None else (lambda name: name in outputs)

        def buf(name, shape):
            t = self.out.get(name)
            if (t is None or tuple(t.shape) != tuple(shape) or t.device != self.device or t.dtype != torch.float32
                    or not t.is_contiguous()):
                t = torch.empty(shape, dtype=torch.float32, device=self.device)
                self.out[name] = t
            return t

        self._fo = ParcFrameOut()
        for name, shape in (("root_pos", (N, 3)), ("root_rot", (N, 4)), ("root_vel", (N, 3)), ("root_ang_vel", (N, 3)),
                            ("joint_rot", (N, J - 1, 4)), ("dof_vel", (N, D))):
            if want(name):
                setattr(self._fo, name, buf(name, shape).data_ptr())
        if want_contacts and want("contacts"):
            self._fo.contacts = buf("contacts", (N, J)).data_ptr()
        self._fk = ParcFkOut()
        if want_fk and want("body_pos"):
            self._fk.body_pos = buf("body_pos", (N, J, 3)).data_ptr()
        if want_fk and want("body_rot"):
            self._fk.body_rot = buf("body_rot", (N, J, 4)).data_ptr()
        self._tb = tables.c_struct(model)
        self._hf = self._obs = None
        qa = ParcQueryArgs()
        qa.tables, qa.model, qa.frame, qa.fk = (C.addressof(self._tb), C.addressof(model), C.addressof(self._fo),
                                               C.addressof(self._fk))
        qa.motion_ids, qa.motion_times, qa.n = ids.data_ptr(), times.data_ptr(), E
        if obs_tmpl is not None and want("obs"):
            assert hf is not None
            require_cuda(hf.hf, obs_tmpl)
            tm = f32c(obs_tmpl)
            self._keep += (tm,)
            self._hf, self._obs = hf.c_struct(), _obs_struct(tm, obs_relative, obs_min_h, obs_max_h)
            qa.hf, qa.obs = C.addressof(self._hf), C.addressof(self._obs)
            qa.obs_out = buf("obs", (E, int(tm.shape[0]))).data_ptr()
        if time_offsets is not None:
            assert time_offsets.dtype == torch.float32 and time_offsets.is_contiguous()
            qa.time_offsets, qa.num_steps = time_offsets.data_ptr(), S
        if root_xy_offset is not None:           # [E,2]: where each env's motion sits on the shared terrain
            assert root_xy_offset.dtype == torch.float32 and root_xy_offset.is_contiguous()
            assert tuple(root_xy_offset.shape) == (E, 2)
            qa.root_xy_offset = root_xy_offset.data_ptr()
        if error_flags is not None:
            assert error_flags.dtype == torch.int32 and error_flags.device == self.device
            qa.error_flags = error_flags.data_ptr()
        self.error_flags = error_flags
        qa.flags = ((_lib.PARC_QUERY_FAST_HEADING if fast_heading else 0) | (_lib.PARC_QUERY_PDL if pdl else 0)
                    | (_lib.PARC_QUERY_PDL_EARLY_INPUTS if (pdl and pdl_early_inputs) else 0))
        qa.variant = int(variant)
        if tar_obs is not None:
            # fused compute_tar_obs for steps 1..S-1 (include/parc_b200.h: ParcTarObsSpec)
            assert time_offsets is not None and S >= 2 and want_fk and self._fk.body_pos, "tar_obs needs the step form + FK"
            rp, rr = tar_obs["sim_root_pos"], tar_obs["sim_root_rot"]
            require_cuda(rp, rr, tar_obs["out"])
            assert rp.dtype == torch.float32 and rp.is_contiguous() and tuple(rp.shape) == (E, 3)
            assert rr.dtype == torch.float32 and rr.is_contiguous() and tuple(rr.shape) == (E, 4)
            kid = _key_ids(tar_obs.get("key_body_ids"), self.device)
            K = 0 if kid is None else int(kid.shape[0])
            W = 9 + 6 * (J - 1) + 3 * K
            blk, stride = _out_rows(tar_obs["out"], E, (S - 1) * W, self.device)
            ts = _lib.ParcTarObsSpec()
            ts.sim_root_pos, ts.sim_root_rot = rp.data_ptr(), rr.data_ptr()
            ts.key_body_ids, ts.num_keys = (kid.data_ptr() if K else None), K
            ts.obs_out, ts.out_env_stride = blk.data_ptr(), stride
            ts.global_obs = int(bool(tar_obs.get("global_obs", False)))
            ts.global_tar_root_h = int(bool(tar_obs.get("global_tar_root_h", False)))
            self._tar = ts
            self._keep += (rp, rr, kid, blk)
            qa.tar_obs = C.addressof(ts)
            self.tar_obs_out = blk                       # [E, (S-1) * W], row stride `stride` floats
        self._qa = qa
        self._fn = _lib.load().parc_motion_query_ex
        self._args = (C.byref(qa),)

    def launch(self, stream: Optional[int] = None) -> dict:
        """Enqueue on `stream` (a raw cudaStream_t; default = torch's current stream of the plan's
        device).  The caller is responsible for the current device being the plan's."""
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self._fn(*self._args, stream)
        _lib.LAUNCHES[0] += 1
        if rc != 0:
            check(rc, "parc_motion_query_ex")
        return self.out

    def redirect_output(self, name: str, address: int):
        """Point output `name` ("body_pos" or "obs") of the prebuilt launch at a raw device address instead of
        `self.out[name]` -- e.g. the NVSwitch multicast address of this rank's rows of a gathered tensor
        (sharding.PeerGather.direct_ptr), so that the kernel's own stores are replicated into every GPU.  The caller
        owns the memory behind the address and its lifetime."""
        if name == "body_pos":
            assert self._fk.body_pos, "the plan does not produce body_pos"
            self._fk.body_pos = address
        elif name == "obs":
            assert self._qa.obs_out, "the plan does not produce obs"
            self._qa.obs_out = address
        else:
            raise ValueError(name)

    def check_errors(self):
        """Raise IndexError if any launch since the last check saw an out-of-range clip id (synchronises)."""
        if self.error_flags is not None:
            raise_query_errors(self.error_flags)

    def capture(self) -> "torch.cuda.CUDAGraph":
        """Record the launch in a CUDA graph (one kernel node).  `replay()` then costs a graph launch, which on
        B200 reaches the SMs ~1.8 us sooner than a stream launch of the same kernel (measured: 5.6 vs 7.5 us event
        to event for an empty kernel of this grid) -- a tenth of the whole 4096-env query.  The buffers the plan was
        built over stay the inputs / outputs of every replay."""
        self._graph = capture_launches([self])
        return self._graph

    def replay(self) -> dict:
        """Re-run the captured launch on torch's current stream (call `capture()` once first)."""
        self._graph.replay()
        _lib.LAUNCHES[0] += 1
        return self.out


def capture_launches(plans) -> "torch.cuda.CUDAGraph":
    """Record `plan.launch()` of every plan, in order, into ONE CUDA graph on the plans' device (kernel nodes only;
    PDL launches become programmatic-dependency edges).  Replaying it enqueues the whole sequence with one host call --
    the tracker holds such a graph per control step, a sweep holds one per chunk."""
    dev = plans[0].device
    with torch.cuda.device(dev):
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            plans[0].launch()                                   # warm-up outside the capture (module load)
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for pl in plans:
                pl.launch()
    return graph


# ----------------------------------------------------------------------------------------------
# a6: forward kinematics (differentiable)
# ----------------------------------------------------------------------------------------------
class _ForwardKinematics(torch.autograd.Function):
    @staticmethod
    def forward(ctx, root_pos, root_rot, joint_rot, model):
        require_cuda(root_pos, root_rot, joint_rot)
        lead = root_pos.shape[:-1]
        J = model.num_bodies
        rp = f32c(root_pos).reshape(-1, 3)
        rr = f32c(root_rot).reshape(-1, 4)
        jr = f32c(joint_rot).reshape(-1, max(J - 1, 0), 4)
        n = rp.shape[0]
        body_pos = torch.empty((n, J, 3), dtype=torch.float32, device=rp.device)
        body_rot = torch.empty((n, J, 4), dtype=torch.float32, device=rp.device)
        with torch.cuda.device(rp.device):
            rc = _lib.load().parc_fk_fwd(rp.data_ptr(), rr.data_ptr(), jr.data_ptr(), n, C.byref(model),
                                         body_pos.data_ptr(), body_rot.data_ptr(), stream_ptr(rp.device))
        check(rc, "parc_fk_fwd")
        ctx.model = model
        ctx.lead = lead
        ctx.save_for_backward(rr, jr)
        return body_pos.reshape(*lead, J, 3), body_rot.reshape(*lead, J, 4)

    @staticmethod
    def backward(ctx, g_pos, g_rot):
        rr, jr = ctx.saved_tensors
        model = ctx.model
        J = model.num_bodies
        n = rr.shape[0]
        gp = f32c(g_pos).reshape(n, J, 3) if g_pos is not None else None
        gr = f32c(g_rot).reshape(n, J, 4) if g_rot is not None else None
        g_root_pos = torch.empty((n, 3), dtype=torch.float32, device=rr.device)
        g_root_rot = torch.empty((n, 4), dtype=torch.float32, device=rr.device)
        g_joint = torch.empty((n, max(J - 1, 0), 4), dtype=torch.float32, device=rr.device)
        with torch.cuda.device(rr.device):
            rc = _lib.load().parc_fk_bwd(rr.data_ptr(), jr.data_ptr(), ptr(gp), ptr(gr), n, C.byref(model),
                                         g_root_pos.data_ptr(), g_root_rot.data_ptr(), g_joint.data_ptr(),
                                         stream_ptr(rr.device))
        check(rc, "parc_fk_bwd")
        lead = ctx.lead
        return (g_root_pos.reshape(*lead, 3), g_root_rot.reshape(*lead, 4),
                g_joint.reshape(*lead, max(J - 1, 0), 4), None)


def forward_kinematics(model: ParcCharModel, root_pos, root_rot, joint_rot):
    return _ForwardKinematics.apply(root_pos, root_rot, joint_rot, model)


# ----------------------------------------------------------------------------------------------
# a8: DoF -> joint quaternions, exp-map -> quaternion (differentiable)
# ----------------------------------------------------------------------------------------------
class _DofToRot(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dof, model):
        require_cuda(dof)
        lead = dof.shape[:-1]
        d = f32c(dof).reshape(-1, model.dof_size)
        n = d.shape[0]
        J = model.num_bodies
        out = torch.empty((n, J - 1, 4), dtype=torch.float32, device=d.device)
        with torch.cuda.device(d.device):
            rc = _lib.load().parc_dof_to_rot_fwd(d.data_ptr(), n, C.byref(model), out.data_ptr(),
                                                 stream_ptr(d.device))
        check(rc, "parc_dof_to_rot_fwd")
        ctx.model, ctx.lead = model, lead
        ctx.save_for_backward(d)
        return out.reshape(*lead, J - 1, 4)

    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        model = ctx.model
        n = d.shape[0]
        gq = f32c(g).reshape(n, model.num_bodies - 1, 4)
        gd = torch.zeros((n, model.dof_size), dtype=torch.float32, device=d.device)
        with torch.cuda.device(d.device):
            rc = _lib.load().parc_dof_to_rot_bwd(d.data_ptr(), gq.data_ptr(), n, C.byref(model), gd.data_ptr(),
                                                 stream_ptr(d.device))
        check(rc, "parc_dof_to_rot_bwd")
        return gd.reshape(*ctx.lead, model.dof_size), None


def dof_to_rot(model: ParcCharModel, dof):
    return _DofToRot.apply(dof, model)


def rot_to_dof(model: ParcCharModel, joint_rot: torch.Tensor) -> torch.Tensor:
    """joint_rot [...,J-1,4] -> dof [...,D]; one launch, forward only (no autograd)."""
    require_cuda(joint_rot)
    lead = joint_rot.shape[:-2]
    jr = f32c(joint_rot.detach()).reshape(-1, model.num_bodies - 1, 4)
    n = jr.shape[0]
    out = torch.zeros((n, model.dof_size), dtype=torch.float32, device=jr.device)
    with torch.cuda.device(jr.device):
        rc = _lib.load().parc_rot_to_dof(jr.data_ptr(), n, C.byref(model), out.data_ptr(), stream_ptr(jr.device))
    check(rc, "parc_rot_to_dof")
    return out.reshape(*lead, model.dof_size)


class _ExpMapToQuat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, e):
        require_cuda(e)
        lead = e.shape[:-1]
        ee = f32c(e).reshape(-1, 3)
        n = ee.shape[0]
        q = torch.empty((n, 4), dtype=torch.float32, device=ee.device)
        with torch.cuda.device(ee.device):
            rc = _lib.load().parc_exp_map_to_quat_fwd(ee.data_ptr(), n, q.data_ptr(), stream_ptr(ee.device))
        check(rc, "parc_exp_map_to_quat_fwd")
        ctx.lead = lead
        ctx.save_for_backward(ee)
        return q.reshape(*lead, 4)

    @staticmethod
    def backward(ctx, g):
        (ee,) = ctx.saved_tensors
        n = ee.shape[0]
        gq = f32c(g).reshape(n, 4)
        ge = torch.empty((n, 3), dtype=torch.float32, device=ee.device)
        with torch.cuda.device(ee.device):
            rc = _lib.load().parc_exp_map_to_quat_bwd(ee.data_ptr(), gq.data_ptr(), n, ge.data_ptr(),
                                                      stream_ptr(ee.device))
        check(rc, "parc_exp_map_to_quat_bwd")
        return ge.reshape(*ctx.lead, 3)


def exp_map_to_quat(exp_map):
    return _ExpMapToQuat.apply(exp_map)


# ----------------------------------------------------------------------------------------------
# a9-a11: heightfield sampling / observations
# ----------------------------------------------------------------------------------------------
def hf_sample(hf: HeightfieldDesc, xy: torch.Tensor, want_index: bool = False):
    """xy [...,2] -> z [...] (and int64 grid indices [...,2] if want_index)."""
    require_cuda(hf.hf, xy)
    lead = xy.shape[:-1]
    p = f32c(xy).reshape(-1, 2)
    n = p.shape[0]
    z = torch.empty((n,), dtype=torch.float32, device=p.device)
    gi = torch.empty((n, 2), dtype=torch.int64, device=p.device) if want_index else None
    h = hf.c_struct()
    with torch.cuda.device(p.device):
        rc = _lib.load().parc_hf_sample(C.byref(h), p.data_ptr(), n, z.data_ptr(), ptr(gi), stream_ptr(p.device))
    check(rc, "parc_hf_sample")
    if want_index:
        return z.reshape(lead), gi.reshape(*lead, 2)
    return z.reshape(lead)


def selftest_grid_index(min_coord: float, cell_size: float, dim: int, device="cuda") -> int:
    """Number of float inputs (out of all 2^32 bit patterns) for which the hoisted-reciprocal grid index of
    the observation loops differs from the reference form with a true IEEE division.  Expected 0."""
    cnt = torch.zeros(1, dtype=torch.int64, device=device)
    with torch.cuda.device(cnt.device):
        rc = _lib.load().parc_selftest_grid_index(float(min_coord), float(cell_size), int(dim), cnt.data_ptr(),
                                                  stream_ptr(cnt.device))
    check(rc, "parc_selftest_grid_index")
    return int(cnt.item())


def hf_obs(hf: HeightfieldDesc, tmpl_xy: torch.Tensor, root: torch.Tensor, heading: Optional[torch.Tensor], *,
           relative: bool, min_h: float = -3.0, max_h: float = 3.0, out: Optional[torch.Tensor] = None,
           root_rot: Optional[torch.Tensor] = None, root_offset: Optional[torch.Tensor] = None, plan: bool = False):
    """root [N,>=2 (3 if relative)], heading [N], tmpl [P,2] -> [N,P].  With heading=None the heading is taken
    from `root_rot` [N,4] inside the launch; `root_offset` [N,>=2 (3 if relative)] is added to the root first.
    plan=True returns a CallPlan over these buffers instead of launching."""
    require_cuda(hf.hf, tmpl_xy, root, heading, root_rot, root_offset)
    r = f32c(root)
    assert r.dim() == 2
    hd = f32c(heading).reshape(-1) if heading is not None else None
    rr = f32c(root_rot) if heading is None else None
    assert hd is not None or (rr is not None and rr.shape == (r.shape[0], 4))
    ro = f32c(root_offset) if root_offset is not None else None
    tm = f32c(tmpl_xy).reshape(-1, 2)
    n, P = r.shape[0], tm.shape[0]
    out, out_stride = _out_rows(out, n, P, r.device)
    h, o = hf.c_struct(), _obs_struct(tm, relative, min_h, max_h)
    args = (C.byref(h), C.byref(o), r.data_ptr(), int(r.shape[1]), ptr(hd), ptr(rr), ptr(ro),
            int(ro.shape[1]) if ro is not None else 0, n, out.data_ptr(), out_stride)
    if plan:
        return CallPlan("parc_hf_obs", args, r.device, (h, o, r, hd, rr, ro, tm, hf, out), out)
    with torch.cuda.device(r.device):
        rc = _lib.load().parc_hf_obs(*args, stream_ptr(r.device))
    check(rc, "parc_hf_obs")
    return out


# ----------------------------------------------------------------------------------------------
# a13-a15: point <-> heightfield SDF, fused body-point loss
# ----------------------------------------------------------------------------------------------
@dataclass
class TerrainBatchDesc:
    hf: torch.Tensor            # [B or 1, X, Y]
    min_center: torch.Tensor    # [B or 1, 2]
    x_nodes: torch.Tensor       # [X]
    y_nodes: torch.Tensor       # [Y]
    half_dx: float
    half_dy: float
    base_z: Optional[torch.Tensor] = None   # [B or 1] device, or None -> base_z_value
    base_z_value: float = -10.0

    def c_struct(self, batch: int) -> ParcTerrainBatch:
        t = ParcTerrainBatch()
        hb = int(self.hf.shape[0])
        X, Y = int(self.hf.shape[1]), int(self.hf.shape[2])
        assert hb in (1, batch) and self.min_center.shape[0] in (1, batch)
        for name in ("hf", "min_center", "x_nodes", "y_nodes", "base_z"):
            v = getattr(self, name)
            assert v is None or (v.is_cuda and v.dtype == torch.float32 and v.is_contiguous() and
                                 v.device == self.hf.device), f"terrain batch field {name}: contiguous fp32 CUDA expected"
        t.hf = self.hf.data_ptr()
        t.hf_batch_stride = X * Y if hb > 1 else 0
        t.min_center = self.min_center.data_ptr()
        t.min_center_stride = 2 if self.min_center.shape[0] > 1 else 0
        t.x_nodes, t.y_nodes = self.x_nodes.data_ptr(), self.y_nodes.data_ptr()
        t.dim_x, t.dim_y = X, Y
        t.half_dx, t.half_dy = float(self.half_dx), float(self.half_dy)
        if self.base_z is not None:
            t.base_z = self.base_z.data_ptr()
            t.base_z_stride = 1 if self.base_z.numel() > 1 else 0
        else:
            t.base_z = None
            t.base_z_stride = 0
        t.base_z_value = float(self.base_z_value)
        return t


def make_terrain_batch(hf: torch.Tensor, min_center: torch.Tensor, dxdy_host: Tuple[float, float],
                       base_z=None) -> TerrainBatchDesc:
    """hf [B,X,Y] (or [X,Y]); min_center [B,2] (or [2]); dxdy as python floats.  The node offsets are
    computed with torch.linspace on the device exactly as util/terrain_util.py:1855-1856 does."""
    require_cuda(hf, min_center)
    if hf.dim() == 2:
        hf = hf.unsqueeze(0)
    if min_center.dim() == 1:
        min_center = min_center.unsqueeze(0)
    if hf.stride(0) == 0:
        hf = hf[:1]
    if min_center.stride(0) == 0:
        min_center = min_center[:1]
    hf = f32c(hf)
    min_center = f32c(min_center)
    X, Y = hf.shape[1], hf.shape[2]
    dx, dy = float(dxdy_host[0]), float(dxdy_host[1])
    xs = torch.linspace(0.0, (X - 1.0) * dx, X, device=hf.device)
    ys = torch.linspace(0.0, (Y - 1.0) * dy, Y, device=hf.device)
    half = torch.tensor([dx, dy], dtype=torch.float32) / 2.0     # fp32 halving as in :1881
    desc = TerrainBatchDesc(hf=hf, min_center=min_center, x_nodes=xs, y_nodes=ys, half_dx=half[0].item(),
                            half_dy=half[1].item())
    if isinstance(base_z, torch.Tensor):
        desc.base_z = f32c(base_z).reshape(-1)
    elif base_z is not None:
        desc.base_z_value = float(base_z)
    return desc


def _points_hf_sdf_launch(p: torch.Tensor, terrain: "TerrainBatchDesc", inverted: bool, want_arg: bool):
    B, N = p.shape[0], p.shape[1]
    out = torch.empty((B, N), dtype=torch.float32, device=p.device)
    arg = torch.empty((B, N), dtype=torch.int32, device=p.device) if want_arg else None
    t = terrain.c_struct(B)
    with torch.cuda.device(p.device):
        rc = _lib.load().parc_points_hf_sdf(p.data_ptr(), B, N, C.byref(t), 1 if inverted else 0, out.data_ptr(),
                                            ptr(arg), stream_ptr(p.device))
    check(rc, "parc_points_hf_sdf")
    return out, arg


class _PointsHfSdf(torch.autograd.Function):
    """sdf [B,N] of points [B,N,3]; backward routes the upstream gradient through the arg-min cell's box SDF
    (parc_points_hf_sdf_bwd).  No gradient for the heightfield (the reference's callers never ask for one)."""

    @staticmethod
    def forward(ctx, points, terrain, inverted):
        p = f32c(points)
        need = ctx.needs_input_grad[0]
        out, arg = _points_hf_sdf_launch(p, terrain, inverted, need)
        if need:
            ctx.terrain, ctx.inverted = terrain, inverted
            ctx.save_for_backward(p, arg)
        return out

    @staticmethod
    def backward(ctx, g):
        p, arg = ctx.saved_tensors
        B, N = p.shape[0], p.shape[1]
        gs = f32c(g)
        gp = torch.empty_like(p)
        t = ctx.terrain.c_struct(B)
        with torch.cuda.device(p.device):
            rc = _lib.load().parc_points_hf_sdf_bwd(p.data_ptr(), B, N, C.byref(t), 1 if ctx.inverted else 0,
                                                    arg.data_ptr(), gs.data_ptr(), gp.data_ptr(), stream_ptr(p.device))
        check(rc, "parc_points_hf_sdf_bwd")
        return gp, None, None


def points_hf_sdf(points: torch.Tensor, terrain: TerrainBatchDesc, inverted: bool, want_arg: bool = False):
    """points [B,N,3] -> sdf [B,N] (exact min over all cells).  Differentiable with respect to the points; with
    want_arg=True returns (sdf, flat arg-min cell index int32 [B,N]) detached."""
    require_cuda(points)
    assert points.dim() == 3 and points.shape[-1] == 3
    if want_arg:
        return _points_hf_sdf_launch(f32c(points.detach()), terrain, inverted, True)
    return _PointsHfSdf.apply(points, terrain, inverted)


@dataclass
class BodyPointsDesc:
    points: torch.Tensor        # [S,3] device
    point_start: torch.Tensor   # int32 [J+1] device

    def c_struct(self) -> ParcBodyPoints:
        assert self.points.is_cuda and self.points.dtype == torch.float32 and self.points.is_contiguous()
        assert self.point_start.dtype == torch.int32 and self.point_start.is_contiguous()
        assert self.point_start.device == self.points.device
        b = ParcBodyPoints()
        b.points = self.points.data_ptr()
        b.point_start = self.point_start.data_ptr()
        b.num_points = int(self.points.shape[0])
        return b


def make_body_points(body_points, device) -> BodyPointsDesc:
    counts = [int(p.shape[0]) for p in body_points]
    starts = [0]
    for c in counts:
        starts.append(starts[-1] + c)
    pts = torch.cat([f32c(p.detach()).reshape(-1, 3) for p in body_points], dim=0).to(device)
    return BodyPointsDesc(points=pts.contiguous(),
                          point_start=torch.tensor(starts, dtype=torch.int32, device=device))


class _BodyPointsWorld(torch.autograd.Function):
    @staticmethod
    def forward(ctx, body_pos, body_rot, pts):
        require_cuda(body_pos, body_rot)
        bp, br = f32c(body_pos), f32c(body_rot)
        assert bp.dim() == 4 and br.dim() == 4 and bp.shape[:3] == br.shape[:3] and bp.shape[3] == 3 and br.shape[3] == 4
        B, F, J = bp.shape[0], bp.shape[1], bp.shape[2]
        S = int(pts.points.shape[0])
        out = torch.empty((B, F * S, 3), dtype=torch.float32, device=bp.device)
        c = pts.c_struct()
        with torch.cuda.device(bp.device):
            rc = _lib.load().parc_body_points_fwd(bp.data_ptr(), br.data_ptr(), B, F, J, C.byref(c), out.data_ptr(),
                                                  stream_ptr(bp.device))
        check(rc, "parc_body_points_fwd")
        ctx.pts, ctx.dims = pts, (B, F, J)
        ctx.save_for_backward(br)
        return out

    @staticmethod
    def backward(ctx, g):
        (br,) = ctx.saved_tensors
        B, F, J = ctx.dims
        gg = f32c(g)
        gp = torch.empty((B, F, J, 3), dtype=torch.float32, device=br.device)
        gr = torch.empty((B, F, J, 4), dtype=torch.float32, device=br.device)
        c = ctx.pts.c_struct()
        with torch.cuda.device(br.device):
            rc = _lib.load().parc_body_points_bwd(br.data_ptr(), gg.data_ptr(), B, F, J, C.byref(c), gp.data_ptr(),
                                                  gr.data_ptr(), stream_ptr(br.device))
        check(rc, "parc_body_points_bwd")
        return gp, gr, None


def body_points_world(body_pos: torch.Tensor, body_rot: torch.Tensor, pts: BodyPointsDesc) -> torch.Tensor:
    """body_pos [B,F,J,3], body_rot [B,F,J,4] -> world surface points [B, F*S, 3] in the reference's body-major
    layout (body b's block = [F, P_b] points starting at F * point_start[b]); differentiable."""
    return _BodyPointsWorld.apply(body_pos, body_rot, pts)


def _body_loss_launch(model, pts, terrain, root_pos, root_rot, joint_rot, contacts, w_pen, w_contact, want_grad):
    B, F = root_pos.shape[0], root_pos.shape[1]
    J = model.num_bodies
    dev = root_pos.device
    pen = torch.empty((B, F), dtype=torch.float32, device=dev)
    con = torch.empty((B, F), dtype=torch.float32, device=dev)
    g_rp = g_rr = g_jr = None
    if want_grad:
        g_rp = torch.empty((B, F, 3), dtype=torch.float32, device=dev)
        g_rr = torch.empty((B, F, 4), dtype=torch.float32, device=dev)
        g_jr = torch.empty((B, F, J - 1, 4), dtype=torch.float32, device=dev)
    t = terrain.c_struct(B)
    bp = pts.c_struct()
    with torch.cuda.device(dev):
        rc = _lib.load().parc_body_loss(root_pos.data_ptr(), root_rot.data_ptr(), joint_rot.data_ptr(),
                                        contacts.data_ptr(), B, F, C.byref(model), C.byref(bp), C.byref(t),
                                        float(w_pen), float(w_contact), pen.data_ptr(), con.data_ptr(), ptr(g_rp),
                                        ptr(g_rr), ptr(g_jr), stream_ptr(dev))
    check(rc, "parc_body_loss")
    return pen, con, g_rp, g_rr, g_jr


class _BodyLoss(torch.autograd.Function):
    """total[b] = w_pen * pen[b] + w_contact * contact[b]; pen / contact are returned detached."""

    @staticmethod
    def forward(ctx, root_pos, root_rot, joint_rot, contacts, model, pts, terrain, w_pen, w_contact):
        require_cuda(root_pos, root_rot, joint_rot, contacts)
        rp, rr, jr, ct = f32c(root_pos), f32c(root_rot), f32c(joint_rot), f32c(contacts)
        need = any(ctx.needs_input_grad[:3])
        pen_bf, con_bf, g_rp, g_rr, g_jr = _body_loss_launch(model, pts, terrain, rp, rr, jr, ct, w_pen, w_contact,
                                                            need)
        pen, con = pen_bf.sum(dim=-1), con_bf.sum(dim=-1)
        total = w_pen * pen + w_contact * con
        if need:
            ctx.save_for_backward(g_rp, g_rr, g_jr)
        ctx.mark_non_differentiable(pen, con)
        return total, pen, con

    @staticmethod
    def backward(ctx, g_total, _g_pen, _g_con):
        g_rp, g_rr, g_jr = ctx.saved_tensors
        s = g_total.reshape(-1, 1, 1)
        return (g_rp * s, g_rr * s, g_jr * s.unsqueeze(-1), None, None, None, None, None, None)


def body_loss(model: ParcCharModel, pts: BodyPointsDesc, terrain: TerrainBatchDesc, root_pos, root_rot, joint_rot,
              contacts, w_pen: float, w_contact: float):
    """[B,F,...] pose + contacts -> (total[B], pen[B], contact[B]); differentiable wrt the pose."""
    return _BodyLoss.apply(root_pos, root_rot, joint_rot, contacts, model, pts, terrain, w_pen, w_contact)


# ----------------------------------------------------------------------------------------------
# dataset sweep -- raw frames -> FK, contact labels, heightfield masks
# ----------------------------------------------------------------------------------------------
def frames_fk(model: ParcCharModel, frames: torch.Tensor, want_rot: bool = False):
    """frames [..., >= 6+D] (root_pos | root exp-map | joint DoFs) -> body_pos [...,J,3], body_rot [...,J,4]
    (and root_rot [...,4], joint_rot [...,J-1,4] if want_rot).  One launch; forward only."""
    require_cuda(frames)
    lead = frames.shape[:-1]
    fr = f32c(frames).reshape(-1, frames.shape[-1])
    n, J = fr.shape[0], model.num_bodies
    dev = fr.device
    bp = torch.empty((n, J, 3), dtype=torch.float32, device=dev)
    br = torch.empty((n, J, 4), dtype=torch.float32, device=dev)
    rr = torch.empty((n, 4), dtype=torch.float32, device=dev) if want_rot else None
    jr = torch.empty((n, J - 1, 4), dtype=torch.float32, device=dev) if want_rot else None
    with torch.cuda.device(dev):
        rc = _lib.load().parc_frames_fk(fr.data_ptr(), n, int(fr.shape[1]), C.byref(model), ptr(rr), ptr(jr),
                                        bp.data_ptr(), br.data_ptr(), stream_ptr(dev))
    check(rc, "parc_frames_fk")
    out = (bp.reshape(*lead, J, 3), br.reshape(*lead, J, 4))
    if want_rot:
        out += (rr.reshape(*lead, 4), jr.reshape(*lead, J - 1, 4))
    return out


def make_key_bodies(feet, hands) -> ParcKeyBodies:
    """feet: list of (body_id, half_extents[3], offset[3]); hands: list of (body_id, radius)."""
    k = ParcKeyBodies()
    assert len(feet) <= _lib.PARC_MAX_KEY_BODIES and len(hands) <= _lib.PARC_MAX_KEY_BODIES
    k.num_feet, k.num_hands = len(feet), len(hands)
    for i, (b, half, off) in enumerate(feet):
        k.foot_body[i] = int(b)
        for c in range(3):
            k.foot_half[i][c] = float(half[c])
            k.foot_offset[i][c] = float(off[c])
    for i, (b, r) in enumerate(hands):
        k.hand_body[i] = int(b)
        k.hand_radius[i] = float(r)
    return k


def clip_label(model: ParcCharModel, pts: Optional[BodyPointsDesc], terrain: TerrainBatchDesc, keys: ParcKeyBodies,
               frames: torch.Tensor, contact_eps: float = 0.04, *, want_contacts=True, want_body_hf=False,
               want_masks=False, want_fk=False, min_body_heights_init: float = 99999.9999) -> dict:
    """frames [B,F,>=6+D] with one terrain per clip (or one shared) -> dict of label tensors; see
    include/parc_b200.h::parc_clip_label.  One launch."""
    require_cuda(frames)
    fr = f32c(frames)
    assert fr.dim() == 3
    B, F, J = fr.shape[0], fr.shape[1], model.num_bodies
    dev = fr.device
    X, Y = int(terrain.hf.shape[1]), int(terrain.hf.shape[2])
    W = (X * Y + 31) // 32
    out = {}
    if want_contacts:
        out["contacts"] = torch.empty((B, F, J), dtype=torch.float32, device=dev)
        out["pen_correction"] = torch.empty((B, F), dtype=torch.float32, device=dev)
    if want_body_hf:
        out["body_hf"] = torch.empty((B, F, J), dtype=torch.float32, device=dev)
    if want_masks:
        assert pts is not None
        out["frame_mask_bits"] = torch.empty((B, F, W), dtype=torch.int32, device=dev)
        out["min_body_heights"] = torch.full((B, X, Y), min_body_heights_init, dtype=torch.float32, device=dev)
    if want_fk:
        out["body_pos"] = torch.empty((B, F, J, 3), dtype=torch.float32, device=dev)
        out["body_rot"] = torch.empty((B, F, J, 4), dtype=torch.float32, device=dev)
    t = terrain.c_struct(B)
    bp = pts.c_struct() if pts is not None else ParcBodyPoints()
    with torch.cuda.device(dev):
        rc = _lib.load().parc_clip_label(fr.data_ptr(), B, F, int(fr.shape[2]), C.byref(model), C.byref(bp), C.byref(t),
                                         C.byref(keys), float(contact_eps), ptr(out.get("contacts")),
                                         ptr(out.get("pen_correction")), ptr(out.get("body_hf")),
                                         ptr(out.get("frame_mask_bits")), ptr(out.get("min_body_heights")),
                                         ptr(out.get("body_pos")), ptr(out.get("body_rot")), stream_ptr(dev))
    check(rc, "parc_clip_label")
    return out


def unpack_frame_masks(bits: torch.Tensor, X: int, Y: int) -> torch.Tensor:
    """[..., W] int32 bit words -> [..., X, Y] bool (bit ix*Y+iy)."""
    shifts = torch.arange(32, device=bits.device, dtype=torch.int32)
    b = ((bits.unsqueeze(-1) >> shifts) & 1).to(torch.bool)
    return b.reshape(*bits.shape[:-1], -1)[..., :X * Y].reshape(*bits.shape[:-1], X, Y)


# ----------------------------------------------------------------------------------------------
# tracker step assembly: policy observation, reward, done -- one launch each
# ----------------------------------------------------------------------------------------------
class CallPlan:
    """One C entry point with its argument list prebuilt over fixed buffers: `launch()` is a single ctypes call.
    Returned by the operators that accept `plan=True`; `result` is what the eager call would have returned."""

    def __init__(self, name: str, args: tuple, device, keep, result):
        self._fn, self._name, self._args = getattr(_lib.load(), name), name, args
        self.device, self._keep, self.result = device, keep, result

    def launch(self, stream: Optional[int] = None):
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self._fn(*self._args, stream)
        _lib.LAUNCHES[0] += 1
        if rc != 0:
            check(rc, self._name)
        return self.result


def _out_rows(out: Optional[torch.Tensor], n: int, width: int, device) -> Tuple[torch.Tensor, int]:
    """(out, row stride in floats): a fresh dense [n,width] buffer, or the caller's -- which may be a column slice
    of a wider row-major buffer (e.g. a block of the policy-observation row): unit inner stride required."""
    if out is None:
        return torch.empty((n, width), dtype=torch.float32, device=device), width
    assert out.dtype == torch.float32 and tuple(out.shape) == (n, width) and out.is_cuda
    assert width <= 1 or out.stride(1) == 1
    stride = int(out.stride(0)) if n > 1 else max(int(out.stride(0)), width)
    assert stride >= width
    return out, stride


def _has(t: Optional[torch.Tensor]) -> bool:
    return t is not None and t.numel() > 0


def _rows(t: torch.Tensor, lead: int = 1):
    """(tensor, k): a zero-copy row-strided view if `t` is fp32 with contiguous inner dims and its leading stride
    is k whole rows (rows = the first `lead` dims flattened; e.g. step 0 of [n,S,...] has k = S); otherwise a
    contiguous fp32 copy with k = 1 (lead == 1) or k = shape[1] (lead == 2)."""
    inner = 1
    for d in t.shape[lead:]:
        inner *= int(d)
    ok = t.dtype == torch.float32 and t.shape[0] > 0 and inner > 0 and t.stride(0) % inner == 0
    if ok:
        probe = t[0] if lead == 1 else t[0, 0]
        ok = probe.is_contiguous() if probe.dim() > 0 else True
    if ok and lead == 2:
        ok = t.shape[1] == 1 or t.stride(1) == inner
    if ok:
        k = t.stride(0) // inner
        if k >= (1 if lead == 1 else int(t.shape[1])):
            return t, int(k)
    t = f32c(t)
    return t, (1 if lead == 1 else int(t.shape[1]))


def _key_ids(key_body_ids, device) -> Optional[torch.Tensor]:
    if key_body_ids is None:
        return None
    ids = torch.as_tensor(key_body_ids, device=device).to(torch.int32).contiguous()
    return ids


def _char_state(root_pos, root_rot, root_vel, root_ang_vel, joint_rot, dof_vel, key_pos, key_body_ids=None):
    """-> (ParcCharState, tensors kept alive, n, J-1, D, K).  Row-strided fp32 views are passed without a copy when
    all six arrays share the same row stride.  With `key_body_ids` (int tensor [K]), `key_pos` is the character's
    body positions [n,J,3] and the key bodies are picked inside the kernel."""
    views = [_rows(t) for t in (root_pos, root_rot, root_vel, root_ang_vel, joint_rot, dof_vel)]
    body_view = _rows(key_pos) if (key_body_ids is not None and _has(key_pos)) else None
    strides = {k for _, k in views} | ({body_view[1]} if body_view else set())
    if len(strides) == 1:
        keep = [t for t, _ in views]
        stride = strides.pop()
        kp = body_view[0] if body_view else None
    else:
        keep = [f32c(t) for t in (root_pos, root_rot, root_vel, root_ang_vel, joint_rot, dof_vel)]
        stride = 1
        kp = f32c(key_pos) if body_view else None
    ids = None
    if body_view:
        ids = _key_ids(key_body_ids, keep[0].device)
    elif _has(key_pos):
        kp = f32c(key_pos)
    keep += [kp, ids]
    require_cuda(*keep)
    n = keep[0].shape[0]
    jm1, d = keep[4].shape[-2], keep[5].shape[-1]
    assert keep[0].shape == (n, 3) and keep[1].shape == (n, 4) and keep[2].shape == (n, 3) and keep[3].shape == (n, 3)
    assert keep[4].shape == (n, jm1, 4) and keep[5].shape == (n, d)
    st = ParcCharState()
    (st.root_pos, st.root_rot, st.root_vel, st.root_ang_vel, st.joint_rot, st.dof_vel) = [ptr(t) for t in keep[:6]]
    st.key_pos, st.key_body_ids = ptr(kp), ptr(ids)
    st.env_stride = stride
    if ids is not None:
        k = int(ids.shape[0])
        assert kp.shape[0] == n and kp.shape[-1] == 3
        st.num_bodies = int(kp.shape[1])
    else:
        k = kp.shape[-2] if kp is not None else 0
        assert k == 0 or kp.shape == (n, k, 3)
        st.num_bodies = 0
    return st, keep, n, jm1, d, k


def char_obs(root_pos, root_rot, root_vel, root_ang_vel, joint_rot, dof_vel, key_pos, global_obs: bool,
             root_height_obs: bool, key_body_ids=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """compute_char_obs (envs/ig_char_env.py:582-626) -> [n, W]; key_pos [n,K,3] or empty/None -- or, with
    `key_body_ids`, the body positions [n,J,3] to pick the key bodies from.  `out` may be a column block of a
    wider observation buffer.  One launch."""
    st, keep, n, jm1, d, k = _char_state(root_pos, root_rot, root_vel, root_ang_vel, joint_rot, dof_vel, key_pos,
                                         key_body_ids)
    dev = keep[0].device
    out, out_stride = _out_rows(out, n, (1 if root_height_obs else 0) + 12 + 6 * jm1 + d + 3 * k, dev)
    with torch.cuda.device(dev):
        rc = _lib.load().parc_char_obs(C.byref(st), n, jm1, d, k, int(bool(global_obs)), int(bool(root_height_obs)),
                                       out.data_ptr(), out_stride, stream_ptr(dev))
    check(rc, "parc_char_obs")
    return out


def tar_obs(ref_root_pos, ref_root_rot, tar_root_pos, tar_root_rot, tar_joint_rot, tar_key_pos, global_obs: bool,
            global_tar_root_h_obs: bool, key_body_ids=None, out: Optional[torch.Tensor] = None, plan: bool = False):
    """compute_tar_obs (envs/ig_parkour/mgdm_dm_util.py:462-518): targets [n,S,...] -> [n,S,W].  The target arrays
    may be step slices of larger [n,S_total,...] buffers (no copy); with `key_body_ids`, `tar_key_pos` is the
    targets' body positions [n,S,J,3].  One launch."""
    rp, rr = f32c(ref_root_pos), f32c(ref_root_rot)
    views = [_rows(t, lead=2) for t in (tar_root_pos, tar_root_rot, tar_joint_rot)]
    use_ids = key_body_ids is not None and _has(tar_key_pos)
    kview = _rows(tar_key_pos, lead=2) if use_ids else None
    strides = {k for _, k in views} | ({kview[1]} if kview else set())
    if len(strides) == 1:
        tp, tr, tj = (t for t, _ in views)
        stride = strides.pop()
        tk = kview[0] if kview else None
    else:
        tp, tr, tj = (f32c(t) for t in (tar_root_pos, tar_root_rot, tar_joint_rot))
        stride = int(tp.shape[1])
        tk = f32c(tar_key_pos) if kview else None
    ids = _key_ids(key_body_ids, tp.device) if use_ids else None
    if not use_ids:
        tk = f32c(tar_key_pos) if _has(tar_key_pos) else None
    require_cuda(rp, rr, tp, tr, tj, tk)
    n, S, jm1 = tp.shape[0], tp.shape[1], tj.shape[-2]
    assert rp.shape == (n, 3) and rr.shape == (n, 4) and tp.shape == (n, S, 3) and tr.shape == (n, S, 4)
    assert tj.shape == (n, S, jm1, 4)
    if ids is not None:
        k, nb = int(ids.shape[0]), int(tk.shape[2])
        assert tk.shape == (n, S, nb, 3)
    else:
        k, nb = (tk.shape[-2] if tk is not None else 0), 0
        assert k == 0 or tk.shape == (n, S, k, 3)
    W = 9 + 6 * jm1 + 3 * k
    if out is None:
        flat, out_stride = torch.empty((n, S * W), dtype=torch.float32, device=tp.device), S * W
    else:                                     # [n, S*W] (possibly a column block of a wider buffer) or [n, S, W]
        flat, out_stride = _out_rows(out.view(n, S * W) if out.dim() == 3 else out, n, S * W, tp.device)
    args = (rp.data_ptr(), rr.data_ptr(), tp.data_ptr(), tr.data_ptr(), tj.data_ptr(), ptr(tk), n, S, jm1, k,
            int(bool(global_obs)), int(bool(global_tar_root_h_obs)), stride, ptr(ids), nb, flat.data_ptr(), out_stride)
    result = flat.view(n, S, W) if out is None else out
    if plan:
        return CallPlan("parc_tar_obs", args, tp.device, (rp, rr, tp, tr, tj, tk, ids, flat), result)
    with torch.cuda.device(tp.device):
        rc = _lib.load().parc_tar_obs(*args, stream_ptr(tp.device))
    check(rc, "parc_tar_obs")
    return result


def deepmimic_reward(sim: tuple, tar: tuple, joint_rot_err_w, dof_err_w, track_root_h: bool, track_root: bool,
                     key_body_ids=None):
    """compute_deepmimic_reward (envs/ig_parkour/mgdm_dm_util.py:328-397).  sim / tar are 7-tuples
    (root_pos, root_rot, root_vel, root_ang_vel, joint_rot, dof_vel, key_pos) -> [n,5]; with `key_body_ids` the
    7th entries are body positions [n,J,3].  Row-strided views (step 0 of [n,S,...]) are read in place.  One launch."""
    if not _has(sim[6]) or not _has(tar[6]) or (key_body_ids is not None and len(key_body_ids) == 0):
        raise ValueError("compute_deepmimic_reward needs key bodies (the reference fails to stack its terms without)")
    s, keep_s, n, jm1, d, k = _char_state(*sim, key_body_ids=key_body_ids)
    t, keep_t, n2, jm2, d2, k2 = _char_state(*tar, key_body_ids=key_body_ids)
    assert (n, jm1, d, k) == (n2, jm2, d2, k2)
    jw, dw = f32c(joint_rot_err_w), f32c(dof_err_w)
    require_cuda(jw, dw)
    assert jw.shape == (jm1,) and dw.shape == (d,)
    dev = keep_s[0].device
    out = torch.empty((n, 5), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().parc_deepmimic_reward(C.byref(s), C.byref(t), n, jm1, d, k, jw.data_ptr(), dw.data_ptr(),
                                               int(bool(track_root_h)), int(bool(track_root)), out.data_ptr(),
                                               stream_ptr(dev))
    check(rc, "parc_deepmimic_reward")
    return out


def done_flags(time, ep_len: float, root_rot, body_pos, tar_root_rot, tar_body_pos, contact_force,
               contact_body_ids, pose_termination: bool, pose_termination_dist, enable_early_termination: bool,
               track_root: bool, root_pos_termination_dist: float, root_rot_termination_angle: float, *,
               termination_heights: Optional[torch.Tensor] = None, hf: Optional[HeightfieldDesc] = None,
               env_offsets: Optional[torch.Tensor] = None, termination_height: float = 0.0,
               want_heights: bool = False):
    """compute_done (envs/ig_parkour/mgdm_dm_util.py:399-460) -> int32 [n].  Heights under the bodies come
    either from `termination_heights` [n,J] (the reference function's own argument) or are sampled in the same
    launch from `hf` at body xy + env_offsets[:, 0:2] plus `termination_height` (RefCharEnv.update_done,
    :205-210).  contact_body_ids: python sequence / tensor of body ids allowed to touch (host-side: it is
    configuration, and becomes a bit mask).  One launch."""
    tm, bp = f32c(time), f32c(body_pos)
    require_cuda(tm, bp)
    n, J = bp.shape[0], bp.shape[1]
    dev = bp.device
    ids = [int(i) for i in (contact_body_ids.tolist() if torch.is_tensor(contact_body_ids) else contact_body_ids)]
    ptd = f32c(pose_termination_dist) if pose_termination_dist is not None else None
    require_cuda(ptd)
    spec = _done_spec(ep_len, termination_height, root_pos_termination_dist, root_rot_termination_angle, ptd, ids, J,
                      pose_termination, enable_early_termination, track_root)
    rr = f32c(root_rot) if root_rot is not None else None
    tar_stride = 1
    trr = tbp = None
    if tar_root_rot is not None and tar_body_pos is not None:           # step 0 of [n,S,...] is read in place
        (trr, k1), (tbp, k2) = _rows(tar_root_rot), _rows(tar_body_pos)
        if k1 != k2:
            trr, tbp, k1 = f32c(tar_root_rot), f32c(tar_body_pos), 1
        tar_stride = k1
    else:
        trr = f32c(tar_root_rot) if tar_root_rot is not None else None
        tbp = f32c(tar_body_pos) if tar_body_pos is not None else None
    cf = f32c(contact_force) if contact_force is not None else None
    th = f32c(termination_heights) if termination_heights is not None else None
    eo = f32c(env_offsets) if env_offsets is not None else None
    require_cuda(rr, trr, tbp, cf, th, eo)
    if ptd is not None:
        assert ptd.shape == (J - 1,)
    hfs = hf.c_struct() if hf is not None else ParcHeightfield()
    out = torch.empty((n,), dtype=torch.int32, device=dev)
    th_out = torch.empty((n, J), dtype=torch.float32, device=dev) if want_heights else None
    with torch.cuda.device(dev):
        rc = _lib.load().parc_done(C.byref(spec), tm.data_ptr(), ptr(rr), bp.data_ptr(), ptr(trr), ptr(tbp), ptr(cf),
                                   ptr(th), C.byref(hfs) if hf is not None else None, ptr(eo),
                                   int(eo.shape[-1]) if eo is not None else 0, tar_stride, n, J, out.data_ptr(),
                                   ptr(th_out), stream_ptr(dev))
    check(rc, "parc_done")
    return (out, th_out) if want_heights else out


def _done_spec(ep_len, termination_height, root_pos_termination_dist, root_rot_termination_angle, pose_termination_dist,
               contact_body_ids, num_bodies, pose_termination, enable_early_termination, track_root):
    spec = ParcDoneSpec()
    spec.episode_length = float(ep_len)
    spec.termination_height = float(termination_height)
    spec.root_pos_termination_dist = float(root_pos_termination_dist)
    spec.root_rot_termination_angle = float(root_rot_termination_angle)
    spec.pose_termination_dist = ptr(pose_termination_dist)
    mask = 0
    for i in contact_body_ids:
        assert 0 <= int(i) < num_bodies
        mask |= 1 << int(i)
    spec.contact_body_mask = mask
    spec.has_contact_bodies = int(len(contact_body_ids) > 0)
    spec.pose_termination = int(bool(pose_termination))
    spec.enable_early_termination = int(bool(enable_early_termination))
    spec.track_root = int(bool(track_root))
    return spec


class SimStepPlan:
    """parc_sim_step over fixed buffers with every argument prebuilt (one ctypes call per step): the simulated
    character's DoF conversion, observation block, reward terms, episode flag and contact-flag blocks in ONE launch.
    `sim` = dict(root_pos, root_rot, root_vel, root_ang_vel, dof_pos, dof_vel, body_pos, contact_force, time,
    env_offsets[, char_contacts]); `ref` = dict of the reference frame as row-strided views (root_pos, root_rot,
    root_vel, root_ang_vel, joint_rot, dof_vel, body_pos[, and `tar_contacts` = contacts of steps 1..S]).
    Outputs: column blocks `char_obs` / `tar_contacts` / `char_contacts` of one observation buffer, `reward` [n,5],
    `done` [n] int32, optional `joint_rot` [n,J-1,4]."""

    def __init__(self, model: ParcCharModel, sim: dict, ref: dict, key_body_ids: torch.Tensor, joint_rot_err_w,
                 dof_err_w, hf: HeightfieldDesc, out: dict, *, cfg: dict, contact_body_ids=(), phase: int = 0):
        J, D = model.num_bodies, model.dof_size
        self.model, self.device = model, sim["root_pos"].device
        n = int(sim["root_pos"].shape[0])
        self.n = n
        keep = []

        def dense(t, shape):
            assert t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == tuple(shape), (t.shape, shape)
            require_cuda(t)
            keep.append(t)
            return t.data_ptr()

        a = ParcSimStep()
        ids = key_body_ids.to(device=self.device, dtype=torch.int32).contiguous()
        keep.append(ids)
        K = int(ids.shape[0])
        a.sim.root_pos, a.sim.root_rot = dense(sim["root_pos"], (n, 3)), dense(sim["root_rot"], (n, 4))
        a.sim.root_vel, a.sim.root_ang_vel = dense(sim["root_vel"], (n, 3)), dense(sim["root_ang_vel"], (n, 3))
        a.sim.dof_vel = dense(sim["dof_vel"], (n, D))
        a.sim.key_pos = dense(sim["body_pos"], (n, J, 3))
        a.sim.key_body_ids, a.sim.num_bodies, a.sim.env_stride = ids.data_ptr(), J, 1
        views = {k: _rows(ref[k]) for k in ("root_pos", "root_rot", "root_vel", "root_ang_vel", "joint_rot", "dof_vel",
                                            "body_pos")}
        strides = {k for _, k in views.values()}
        assert len(strides) == 1, "the reference-frame views must share one row stride"
        stride = strides.pop()
        keep += [t for t, _ in views.values()]
        rv = {k: t.data_ptr() for k, (t, _) in views.items()}
        a.ref.root_pos, a.ref.root_rot, a.ref.root_vel = rv["root_pos"], rv["root_rot"], rv["root_vel"]
        a.ref.root_ang_vel, a.ref.joint_rot, a.ref.dof_vel = rv["root_ang_vel"], rv["joint_rot"], rv["dof_vel"]
        a.ref.key_pos, a.ref.key_body_ids, a.ref.num_bodies, a.ref.env_stride = rv["body_pos"], ids.data_ptr(), J, stride
        a.ref_body_pos = rv["body_pos"]
        a.dof_pos, a.body_pos = dense(sim["dof_pos"], (n, D)), a.sim.key_pos
        a.contact_force = dense(sim["contact_force"], (n, J, 3)) if sim.get("contact_force") is not None else None
        a.time = dense(sim["time"], (n,))
        eo = sim.get("env_offsets")
        if eo is not None:
            a.env_offsets, a.offset_stride = dense(eo, (n, eo.shape[1])), int(eo.shape[1])
        jw, dw = f32c(joint_rot_err_w), f32c(dof_err_w)
        ptd = f32c(cfg["pose_termination_dist"])
        keep += [jw, dw, ptd]
        assert jw.shape == (J - 1,) and dw.shape == (D,) and ptd.shape == (J - 1,)
        a.joint_rot_err_w, a.dof_err_w = jw.data_ptr(), dw.data_ptr()
        a.done = _done_spec(cfg["episode_length"], cfg["termination_height"], cfg["root_pos_termination_dist"],
                            cfg["root_rot_termination_angle"], ptd, contact_body_ids, J, cfg["pose_termination"],
                            cfg["enable_early_termination"], cfg["track_root"])
        a.hf = hf.c_struct()
        keep.append(hf.hf)
        a.num_keys = K
        a.global_obs, a.root_height_obs = int(bool(cfg["global_obs"])), int(bool(cfg["root_height_obs"]))
        a.track_root_h, a.track_root = int(bool(cfg["track_root_h"])), int(bool(cfg["track_root"]))
        # outputs: column blocks of one observation buffer share its row stride
        width = (1 if cfg["root_height_obs"] else 0) + 12 + 6 * (J - 1) + D + 3 * K
        blk, stride_obs = _out_rows(out["char_obs"], n, width, self.device)
        a.char_obs_out, a.obs_stride = blk.data_ptr(), stride_obs
        if out.get("tar_contacts") is not None:
            tc, k2 = _rows(ref["tar_contacts"], lead=2)
            S = int(tc.shape[1])
            assert tc.shape == (n, S, J)
            b2, s2 = _out_rows(out["tar_contacts"], n, S * J, self.device)
            assert s2 == stride_obs
            keep += [tc, b2]
            a.tar_contacts, a.tar_env_stride, a.num_tar_steps, a.tar_contacts_out = tc.data_ptr(), k2, S, b2.data_ptr()
        if out.get("char_contacts") is not None:
            b3, s3 = _out_rows(out["char_contacts"], n, J, self.device)
            assert s3 == stride_obs
            keep.append(b3)
            a.char_contacts, a.char_contacts_out = dense(sim["char_contacts"], (n, J)), b3.data_ptr()
        assert out["reward"].shape == (n, 5) and out["reward"].is_contiguous() and out["done"].dtype == torch.int32
        a.reward_out, a.done_out = out["reward"].data_ptr(), out["done"].data_ptr()
        if out.get("joint_rot") is not None:
            a.joint_rot_out = dense(out["joint_rot"], (n, J - 1, 4))
        keep += [blk, out["reward"], out["done"]]
        # phase: 0 = the whole step; 1 / 2 = the half that does not / does need the reference frame (the halves meet in
        # out["joint_rot"]; include/parc_b200.h: PARC_SIM_STEP_PRE / _POST)
        assert phase in (0, 1, 2) and (phase == 0 or out.get("joint_rot") is not None)
        a.phase = int(phase)
        self._args, self._keep = a, keep
        self._fn = _lib.load().parc_sim_step

    def launch(self, stream: Optional[int] = None):
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self._fn(C.byref(self._args), self.n, C.byref(self.model), stream)
        _lib.LAUNCHES[0] += 1
        if rc != 0:
            check(rc, "parc_sim_step")
