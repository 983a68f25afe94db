"""MotionLib: packed per-frame tables for all clips + batched (clip id, time) -> frame queries.

Drop-in for the reference's `anim/motion_lib.py` (class `MotionLib`, `LoopMode`): same constructor
signature, method names, return tuples and public attributes (:22-40, :48-131, :443-475).  What differs
is underneath: the eight per-frame tables are interleaved into one row of float4 slots per frame
(include/parc_b200.h, ParcRowLayout) and a query is ONE launch of the fused sm_100a kernel in
csrc/motion_query.cu (one warp per query) instead of ~180 eager torch ops.

Loading: `_load_motions` (pickled clips) builds the tables with torch ops on the CPU, in the reference's
operation order (bit-identical tables), then uploads and packs them on the GPU; `_load_motion_frames` does the
same for host frames and builds directly on the GPU when handed CUDA frames.
"""
from __future__ import annotations

import copy
import enum
import io
import os
import pickle

import numpy as np
import torch
import yaml

from .. import ops
from ..util import torch_util
from .kin_char_model import KinCharModel


class LoopMode(enum.Enum):
    CLAMP = 0
    WRAP = 1


def extract_pose_data(frame):
    """[...,6+D] -> root_pos, root_rot (exp-map), joint_dof.  Ref anim/motion_lib.py:15-19."""
    return frame[..., 0:3], frame[..., 3:6], frame[..., 6:]


class _RefUnpickler(pickle.Unpickler):
    """Clip pickles written by the reference name its own modules (`util.terrain_util.SubTerrain`);
    resolve those to this package so the files load without the reference on sys.path."""
    _REMAP = {("util.terrain_util", "SubTerrain"): ("parc_b200.util.terrain_util", "SubTerrain")}

    def find_class(self, module, name):
        module, name = self._REMAP.get((module, name), (module, name))
        return super().find_class(module, name)


def load_clip_file(path):
    with open(path, "rb") as f:
        return _RefUnpickler(io.BytesIO(f.read())).load()


class MotionLib:
    def __init__(self, motion_input, kin_char_model: KinCharModel, device, init_type="motion_file",
                 loop_mode=None, fps=None, contact_info=False, contacts=None, build_on_device=False):
        """Reference signature (:22-40) plus `build_on_device`: False (default) builds the frame tables with host
        torch ops in the reference's operation order (bit-identical tables); True hands the raw frames to the GPU
        loader (one launch for all clips, csrc/table_build.cu) -- what a 100k-clip library wants.  CUDA frames
        passed to init_type="motion_frames" always use the GPU loader.  init_type="packed_file" opens a file
        written by `save_packed` (no per-clip work at all)."""
        self._device = device
        self._kin_char_model = kin_char_model
        self._contact_info = contact_info
        self._hf_mask_inds = None
        self._packed = None
        # table building runs on the host; the model is mirrored there once
        self._host_model = kin_char_model if torch.device(kin_char_model._device).type == "cpu" \
            else kin_char_model.get_copy("cpu")

        self._build_on_device = bool(build_on_device)
        if init_type == "motion_file":
            self._load_motions(motion_input)
        elif init_type == "packed_file":
            from . import packed_format
            packed_format.load_into(self, motion_input)
        elif init_type == "motion_frames":   # (num motions, num_frames, dofs)
            self._load_motion_frames(motion_input, loop_mode, fps, frame_contacts=contacts if contact_info else None)
        else:
            # the reference's "diffusion_file" branch calls a method that does not exist (:30-31)
            raise ValueError(f"unsupported init_type {init_type!r}")

    def save_packed(self, path):
        """Write the library as one packed binary file (anim/packed_format.py); reopen it with
        `MotionLib(path, kin_char_model, device, init_type="packed_file", contact_info=...)`."""
        from . import packed_format
        packed_format.save(self, path)

    # ------------------------------------------------------------------ simple accessors
    def num_motions(self):
        return self._motion_lengths.shape[0]

    def get_total_length(self):
        return torch.sum(self._motion_lengths).item()

    def sample_motions(self, n, motion_weights=None):
        """Ref :48-52."""
        w = self._motion_weights if motion_weights is None else motion_weights
        return torch.multinomial(w, num_samples=n, replacement=True)

    def sample_time(self, motion_ids, truncate_time=None):
        """Ref :54-63."""
        phase = torch.rand(motion_ids.shape, device=self._device)
        motion_len = self._motion_lengths[motion_ids]
        if truncate_time is not None:
            assert truncate_time >= 0.0
            motion_len = motion_len - truncate_time
        return phase * motion_len

    def get_motion_length(self, motion_ids):
        return self._motion_lengths[motion_ids]

    def get_motion_loop_mode(self, motion_ids):
        return self._motion_loop_modes[motion_ids]

    def get_motion_loop_mode_enum(self, motion_id):
        return LoopMode(self._motion_loop_modes[motion_id].item())

    def get_motion_names(self):
        assert hasattr(self, "_motion_names")
        return self._motion_names

    def calc_motion_phase(self, motion_ids, times):
        """Ref :74-78 + calc_phase :527-538 (elementwise; stays in torch on the tensors' device)."""
        phase = times / self._motion_lengths[motion_ids]
        wrap = self._motion_loop_modes[motion_ids] == LoopMode.WRAP.value
        phase = torch.where(wrap, phase - torch.floor(phase), phase)
        return torch.clip(phase, 0.0, 1.0)

    # ------------------------------------------------------------------ the hot path
    def _error_word(self, device):
        """Device int32[1] the query kernels OR their PARC_QUERY_ERR_* bits into (one per device used)."""
        words = self.__dict__.setdefault("_query_error_words", {})
        w = words.get(device)
        if w is None:
            w = words[device] = torch.zeros(1, dtype=torch.int32, device=device)
        return w

    def check_query_errors(self):
        """Raise IndexError if any query since the last check was handed a clip id / frame index outside the tables
        (the reference's gathers raise there: IndexError on CPU, a device assert on CUDA).  Synchronises.  Queries
        never read out of bounds either way: the offending entry is answered with clip 0 / the nearest valid frame.
        Set `validate_ids = True` (or PARC_B200_VALIDATE=1) to check after every call."""
        for w in self.__dict__.get("_query_error_words", {}).values():
            ops.raise_query_errors(w, "MotionLib query")

    @property
    def validate_ids(self):
        v = self.__dict__.get("_validate_ids")
        if v is None:
            import os
            v = self.__dict__["_validate_ids"] = os.environ.get("PARC_B200_VALIDATE", "0") not in ("", "0")
        return v

    @validate_ids.setter
    def validate_ids(self, value):
        self.__dict__["_validate_ids"] = bool(value)

    def _query(self, motion_ids, **kw):
        r = ops.motion_query(self._packed, self._kin_char_model.c_model(), motion_ids,
                             error_flags=self._error_word(motion_ids.device), **kw)
        if self.validate_ids:
            self.check_query_errors()
        return r

    def calc_motion_frame(self, motion_ids, motion_times):
        """(root_pos, root_rot, root_vel, root_ang_vel, joint_rot, dof_vel[, contacts]).  Ref :80-112."""
        r = self._query(motion_ids, motion_times=motion_times, want_contacts=self._contact_info)
        return self._as_tuple(r)

    def get_motion_frame(self, motion_ids, frame_idxs):
        """Integer-frame lookup, no blending.  Ref :114-131."""
        r = self._query(motion_ids, frame_idxs=frame_idxs, want_contacts=self._contact_info)
        return self._as_tuple(r)

    def _as_tuple(self, r):
        ret = [r["root_pos"], r["root_rot"], r["root_vel"], r["root_ang_vel"], r["joint_rot"], r["dof_vel"]]
        if self._contact_info:
            ret.append(r["contacts"])
        return tuple(ret)

    def _calc_frame_blend(self, motion_ids, times):
        """(frame_idx0, frame_idx1, blend), absolute table rows.  Ref :443-456."""
        r = self._query(motion_ids, motion_times=times, want_frame=False, want_index=True)
        return r["frame_idx0"], r["frame_idx1"], r["blend"]

    def calc_motion_frame_fk_obs(self, motion_ids, motion_times, hf_desc=None, obs_tmpl=None, obs_relative=True,
                                 min_obs_h=-3.0, max_obs_h=3.0, out=None, fast_heading=False):
        """Fused extension (no reference counterpart as ONE call): the frame query, the forward
        kinematics of the blended pose and the heightmap observation around it in a single launch --
        what dm_env.py:570-595 + ig_parkour_env.py:636-656 do in ~1500 eager ops.  Returns a dict.
        The observation's heading follows the reference chain (atan2 -> cos / sin); fast_heading=True takes cos / sin
        straight from the rotated x axis instead (a few instructions cheaper, ~2 ulp apart)."""
        return self._query(motion_ids, motion_times=motion_times, want_contacts=self._contact_info, want_fk=True,
                           hf=hf_desc, obs_tmpl=obs_tmpl, obs_relative=obs_relative, obs_min_h=min_obs_h,
                           obs_max_h=max_obs_h, out=out, fast_heading=fast_heading)

    def make_query_plan(self, motion_ids, motion_times, hf_desc=None, obs_tmpl=None, obs_relative=True,
                        min_obs_h=-3.0, max_obs_h=3.0, want_fk=True, out=None, time_offsets=None, root_xy_offset=None,
                        outputs=None, fast_heading=False, pdl=False, pdl_early_inputs=False, variant=0, tar_obs=None):
        """Prebuilt launch of `calc_motion_frame_fk_obs` over fixed input/output buffers: returns an
        `ops.MotionQueryPlan` whose `.launch()` costs one C call.  Update `motion_ids` / `motion_times`
        in place between launches (as the tracker does with its time buffer).

        time_offsets [S] (fp32, offsets[0] = 0) turns it into the tracker-step form: every env is queried at
        t + offsets[k] -- the reference frame plus the future targets of fetch_tar_obs_data
        (`timestep * tar_obs_steps`) -- outputs are [N * S, ...] rows in env-major order (view as [N, S, ...]) and the
        observation [N, P] is taken at the current frame only.

        root_xy_offset [N,2] (fp32): added to root x,y of every step before FK and the observation -- where each
        env's motion sits on the shared terrain (DMEnv._move_to_motion_terrain, envs/ig_parkour/dm_env.py:604-615).

        outputs / fast_heading / pdl / pdl_early_inputs / variant / tar_obs: see `ops.MotionQueryPlan`."""
        return ops.MotionQueryPlan(self._packed, self._kin_char_model.c_model(), motion_ids, motion_times,
                                   want_contacts=self._contact_info, want_fk=want_fk, hf=hf_desc, obs_tmpl=obs_tmpl,
                                   obs_relative=obs_relative, obs_min_h=min_obs_h, obs_max_h=max_obs_h, out=out,
                                   time_offsets=time_offsets, root_xy_offset=root_xy_offset, outputs=outputs,
                                   fast_heading=fast_heading, pdl=pdl, pdl_early_inputs=pdl_early_inputs,
                                   variant=variant, error_flags=self._error_word(motion_ids.device), tar_obs=tar_obs)

    def _calc_loop_offset(self, motion_ids, times):
        """floor(t / len) * root_pos_delta for WRAP clips, zero otherwise (ref :458-475).  Kept for callers that
        reach for it; `calc_motion_frame` applies the same offset inside its kernel."""
        wrap = self._motion_loop_modes[motion_ids] == LoopMode.WRAP.value
        cycles = torch.floor(times / self._motion_lengths[motion_ids]).unsqueeze(-1)
        off = cycles * self._motion_root_pos_delta[motion_ids]
        return torch.where(wrap.unsqueeze(-1), off, torch.zeros_like(off))

    def joint_rot_to_dof(self, joint_rot):
        return self._kin_char_model.rot_to_dof(joint_rot)

    def calc_motion_frame_dofs(self, motion_ids, motion_times):
        """Ref :477-489."""
        fr = self.calc_motion_frame(motion_ids, motion_times)
        parts = [fr[0], torch_util.quat_to_exp_map(fr[1]), self.joint_rot_to_dof(fr[4])]
        if self._contact_info:
            parts.append(fr[6])
        return torch.cat(parts, dim=-1)

    def get_frame_data(self, motion_id, frame_start, frame_end):
        """Ref :425-441."""
        assert frame_start < frame_end, "frame_start must be less than frame_end"
        assert motion_id < self.num_motions(), "motion_id out of range"
        s = self._motion_start_idx[motion_id] + frame_start
        e = self._motion_start_idx[motion_id] + frame_end
        ret = [self._frame_root_pos[s:e], self._frame_root_rot[s:e], self._frame_joint_rot[s:e]]
        if self._contact_info:
            ret.append(self._frame_contacts[s:e])
        return tuple(ret)

    def get_frames_for_id(self, id):
        """Ref :503-513."""
        n = self._motion_num_frames[id]
        s = self._motion_start_idx[id]
        sl = slice(s, s + n)
        root_pos, root_rot, joint_rot = self._frame_root_pos[sl], self._frame_root_rot[sl], self._frame_joint_rot[sl]
        body_pos, body_rot = self._kin_char_model.forward_kinematics(root_pos, root_rot, joint_rot)
        return root_pos, root_rot, joint_rot, body_pos, body_rot

    def maxpool_contacts(self, kernel_size):
        """Ref :491-497."""
        return torch.max_pool1d(self._frame_contacts.unsqueeze(0), kernel_size=kernel_size, stride=1,
                                padding=kernel_size // 2).squeeze(0)

    def clone(self, device):
        """Ref :515-524."""
        new = copy.copy(self)
        new._device = device
        for k, v in vars(self).items():
            if isinstance(v, torch.Tensor):
                setattr(new, k, v.to(device))
        new._kin_char_model = self._kin_char_model.get_copy(device)
        new._packed = None
        new._finalize_device_tables()
        return new

    # ------------------------------------------------------------------ loading (host side)
    def _extract_frame_data(self, frame):
        """frames -> root_pos, root_rot quat, joint_rot (w >= 0), computed on the host.  Ref :405-423."""
        fr = torch.as_tensor(frame, dtype=torch.float32).detach().cpu()
        root_pos, root_exp, joint_dof = extract_pose_data(fr)
        root_pos = root_pos.clone()
        root_rot = torch_util.exp_map_to_quat(root_exp.clone())
        joint_rot = torch_util.quat_pos(self._host_model.host_dof_to_rot(joint_dof.clone()))
        return root_pos, root_rot, joint_rot

    @staticmethod
    def _finite_diff_vels(root_pos, root_rot, fps):
        """Root linear / angular velocity by forward differences, last frame repeated.  Ref :281-288."""
        root_vel = torch.zeros_like(root_pos)
        root_vel[..., :-1, :] = fps * (root_pos[..., 1:, :] - root_pos[..., :-1, :])
        root_vel[..., -1, :] = root_vel[..., -2, :]
        root_ang_vel = torch.zeros_like(root_pos)
        drot = torch_util.quat_diff(root_rot[..., :-1, :], root_rot[..., 1:, :])
        root_ang_vel[..., :-1, :] = fps * torch_util.quat_to_exp_map(drot)
        root_ang_vel[..., -1, :] = root_ang_vel[..., -2, :]
        return root_vel, root_ang_vel

    def _load_motion_frames(self, motion_frames, loop_mode, fps, frame_contacts=None):
        """All clips share one length.  Ref :137-202 -- including its quirk of handing `fps` to
        compute_frame_dof_vel where a time step is expected (:178), which parity requires."""
        if motion_frames.dim() == 2:
            motion_frames = motion_frames.unsqueeze(0)
        if frame_contacts is not None and frame_contacts.dim() == 2:
            frame_contacts = frame_contacts.unsqueeze(0)
        if frame_contacts is not None:
            assert motion_frames.dim() == frame_contacts.dim() == 3, frame_contacts.shape
        M, F = motion_frames.shape[0], motion_frames.shape[1]
        on_device = torch.is_tensor(motion_frames) and motion_frames.is_cuda
        if (on_device or self._build_on_device) and torch.device(self._device).type == "cuda":
            # CUDA frames (e.g. a batch the MDM just generated) or an explicit request: the GPU loader builds the
            # packed rows of all clips in one launch; nothing visits the host.
            dev = self._device
            fr = torch.as_tensor(motion_frames).detach().to(device=dev, dtype=torch.float32)
            ct = None if frame_contacts is None else torch.as_tensor(frame_contacts).detach().to(device=dev,
                                                                                                 dtype=torch.float32)
            ones = torch.ones(M, dtype=torch.float32, device=dev)
            self._adopt_device_build(fr.reshape(M * F, -1), None if ct is None else ct.reshape(M * F, -1),
                                     num_frames=F * torch.ones(M, dtype=torch.long, device=dev), fps=fps * ones,
                                     dof_vel_dt=fps * ones,                 # the reference's quirk (:178), kept
                                     loop_modes=loop_mode.value * torch.ones(M, dtype=torch.int, device=dev),
                                     weights=ones.clone())
            return
        model = self._host_model
        root_pos, root_rot, joint_rot = self._extract_frame_data(motion_frames)
        delta = root_pos[:, -1] - root_pos[:, 0]
        delta[..., -1] = 0.0
        root_vel, root_ang_vel = self._finite_diff_vels(root_pos, root_rot, fps)
        dof_vel = model.compute_frame_dof_vel(joint_rot, fps)
        J = model.get_num_joints()
        host = lambda t: torch.as_tensor(t, dtype=torch.float32).detach().cpu()
        dev = self._device
        self._motion_fps = fps * torch.ones(M, dtype=torch.float32, device=dev)
        self._motion_dt = 1.0 / fps * torch.ones(M, dtype=torch.float32, device=dev)
        self._motion_num_frames = F * torch.ones(M, dtype=torch.long, device=dev)
        self._motion_lengths = (1.0 / fps * (F - 1)) * torch.ones(M, dtype=torch.float32, device=dev)
        self._motion_loop_modes = loop_mode.value * torch.ones(M, dtype=torch.int, device=dev)
        self._motion_root_pos_delta = delta.to(dev)
        self._motion_weights = torch.ones(M, dtype=torch.float32, device=dev)
        self._host_tables = dict(
            root_pos=root_pos.reshape(-1, 3), root_rot=root_rot.reshape(-1, 4),
            joint_rot=joint_rot.reshape(-1, J - 1, 4), root_vel=root_vel.reshape(-1, 3),
            root_ang_vel=root_ang_vel.reshape(-1, 3), dof_vel=dof_vel.reshape(-1, dof_vel.shape[-1]),
            contacts=None if frame_contacts is None else
            host(frame_contacts.to(torch.float32)).reshape(-1, frame_contacts.shape[-1]),
            frames=host(motion_frames.to(torch.float32)).reshape(-1, motion_frames.shape[-1]))
        self._finish_load()

    def _adopt_device_build(self, frames, contacts, num_frames, fps, dof_vel_dt, loop_modes, weights, lengths=None):
        """GPU loader: `frames` [total, 6+D] / `contacts` [total, J] of the concatenated clips and the
        per-clip meta tensors, all on the device -> packed rows (one launch); the reference's per-frame tables become
        zero-copy strided views of those rows."""
        dev = self._device
        model = self._kin_char_model.c_model()
        rows, lay, start = ops.build_tables(model, frames, contacts if self._contact_info else None, num_frames, fps,
                                            dof_vel_dt)
        self._adopt_rows(rows, lay, frames, num_frames, start, fps, loop_modes, weights, lengths=lengths)

    def _adopt_rows(self, rows, lay, frames, num_frames, start, fps, loop_modes, weights, lengths=None, delta=None):
        dev = self._device
        M = int(num_frames.shape[0])
        J, D = self._kin_char_model.get_num_joints(), self._kin_char_model.get_dof_size()
        v = ops.unpack_row_views(rows, lay, J, D)
        self._frame_root_pos, self._frame_root_rot, self._frame_joint_rot = v["root_pos"], v["root_rot"], v["joint_rot"]
        self._frame_root_vel, self._frame_root_ang_vel, self._frame_dof_vel = v["root_vel"], v["root_ang_vel"], v["dof_vel"]
        if self._contact_info:
            self._frame_contacts = v["contacts"]
        self._motion_frames = frames
        self._motion_fps = fps.to(torch.float32)
        self._motion_dt = 1.0 / self._motion_fps
        self._motion_num_frames = num_frames.to(torch.long)
        # length = 1/fps * (n-1), formed in double and rounded once, as the reference's python floats are (:275,:355)
        self._motion_lengths = lengths if lengths is not None else \
            (1.0 / fps.double() * (num_frames.double() - 1.0)).to(torch.float32)
        self._motion_loop_modes = loop_modes.to(torch.int)
        self._motion_weights = weights
        self._motion_ids = torch.arange(M, dtype=torch.long, device=dev)
        self._motion_start_idx = start
        if delta is None:
            delta = v["root_pos"][start + self._motion_num_frames - 1] - v["root_pos"][start]
            delta[..., -1] = 0.0
        self._motion_root_pos_delta = delta
        clips = ops.build_clip_meta(self._motion_num_frames, self._motion_loop_modes, self._motion_start_idx,
                                    self._motion_lengths, self._motion_root_pos_delta, dev)
        self._packed = ops.PackedTables(rows=rows, clips=clips, total_frames=int(rows.shape[0]), num_clips=M, layout=lay)

    def _load_motions(self, motion_file):
        """yaml list of clip pickles (or a single pickle).  Ref :204-380."""
        files, weights = self._fetch_motion_files(motion_file)
        acc = {k: [] for k in ("root_pos", "root_rot", "joint_rot", "root_vel", "root_ang_vel", "dof_vel",
                               "contacts", "frames", "delta")}
        meta = {k: [] for k in ("fps", "dt", "n", "len", "loop")}
        self._motion_files, self._motion_names, self._motion_extras = [], [], []
        self._terrains, self._hf_mask_inds = [], []
        D = self._host_model.get_dof_size()
        J = self._host_model.get_num_joints()
        # GPU loader: the host only unpickles and concatenates; every conversion happens in one launch afterwards
        on_device = self._build_on_device and torch.device(self._device).type == "cuda"
        for f, path in enumerate(files):
            if len(files) < 1000 or f % 500 == 0:
                print("Loading {:d}/{:d} motion files: {:s}".format(f + 1, len(files), path))
            clip = load_clip_file(path)
            fps = clip.get("fps", 30)
            if isinstance(fps, np.ndarray):
                fps = fps.item()
            frames = clip.get("frames")
            if frames is None:
                frames = np.zeros([3, 6 + D], dtype=np.float32)
            elif frames.ndim == 3 and frames.shape[0] == 1:
                frames = np.squeeze(frames)
            name = os.path.basename(os.path.splitext(path)[0])
            assert name not in self._motion_names
            self._motion_names.append(name)
            n = frames.shape[0]
            if not on_device:
                root_pos, root_rot, joint_rot = self._extract_frame_data(frames)
                delta = root_pos[-1] - root_pos[0]
                delta[..., -1] = 0.0
                root_vel, root_ang_vel = self._finite_diff_vels(root_pos, root_rot, fps)
                acc["root_pos"].append(root_pos)
                acc["root_rot"].append(root_rot)
                acc["joint_rot"].append(joint_rot)
                acc["root_vel"].append(root_vel)
                acc["root_ang_vel"].append(root_ang_vel)
                acc["dof_vel"].append(self._host_model.compute_frame_dof_vel(joint_rot, 1.0 / fps))
                acc["delta"].append(delta)
            acc["frames"].append(torch.as_tensor(np.asarray(frames), dtype=torch.float32))
            if self._contact_info:
                if "contacts" in clip:
                    c = torch.tensor(np.asarray(clip["contacts"]), dtype=torch.float32)
                    if c.dim() == 3 and c.shape[0] == 1:
                        c = torch.squeeze(c)
                else:
                    c = torch.zeros([n, J], dtype=torch.float32)
                acc["contacts"].append(c)
            meta["fps"].append(fps)
            meta["dt"].append(1.0 / fps)
            meta["n"].append(n)
            meta["len"].append(1.0 / fps * (n - 1))
            meta["loop"].append(LoopMode[clip.get("loop_mode", "CLAMP")].value)
            self._motion_files.append(path)
            self._motion_extras.append(clip.get("extra"))
            terrain = clip.get("terrain")
            if terrain is not None:
                terrain.update_old()
                terrain.to_torch(self._device)
            self._terrains.append(terrain)
            inds = clip.get("hf_mask_inds")
            if inds is not None:
                inds = [t.to(device=self._device) for t in inds]
            self._hf_mask_inds.append(inds)

        dev = self._device
        w = torch.tensor(weights, dtype=torch.float32, device=dev)
        if on_device:
            fps_t = torch.tensor(meta["fps"], dtype=torch.float32, device=dev)
            self._adopt_device_build(
                torch.cat(acc["frames"], dim=0).to(dev), torch.cat(acc["contacts"], dim=0).to(dev) if self._contact_info else None,
                num_frames=torch.tensor(meta["n"], dtype=torch.long, device=dev), fps=fps_t,
                dof_vel_dt=torch.tensor(meta["dt"], dtype=torch.float32, device=dev),
                loop_modes=torch.tensor(meta["loop"], dtype=torch.int, device=dev), weights=w / w.sum(),
                lengths=torch.tensor(meta["len"], dtype=torch.float32, device=dev))
            self._motion_dt = torch.tensor(meta["dt"], dtype=torch.float32, device=dev)
            print("Loaded {:d} motions with a total length of {:.3f}s.".format(self.num_motions(), self.get_total_length()))
            return
        self._motion_weights = w / w.sum()
        self._motion_fps = torch.tensor(meta["fps"], dtype=torch.float32, device=dev)
        self._motion_dt = torch.tensor(meta["dt"], dtype=torch.float32, device=dev)
        self._motion_num_frames = torch.tensor(meta["n"], dtype=torch.long, device=dev)
        self._motion_lengths = torch.tensor(meta["len"], dtype=torch.float32, device=dev)
        self._motion_loop_modes = torch.tensor(meta["loop"], dtype=torch.int, device=dev)
        self._motion_root_pos_delta = torch.stack(acc["delta"], dim=0).to(dev)
        self._host_tables = {k: torch.cat(acc[k], dim=0) for k in
                             ("root_pos", "root_rot", "joint_rot", "root_vel", "root_ang_vel", "dof_vel", "frames")}
        self._host_tables["contacts"] = torch.cat(acc["contacts"], dim=0) if self._contact_info else None
        self._finish_load()
        print("Loaded {:d} motions with a total length of {:.3f}s.".format(self.num_motions(), self.get_total_length()))

    def _fetch_motion_files(self, motion_file):
        """Ref :382-403."""
        if os.path.splitext(motion_file)[1] == ".yaml":
            with open(motion_file, "r") as f:
                entries = yaml.load(f, Loader=yaml.SafeLoader)["motions"]
            for e in entries:
                assert e["weight"] >= 0
            return [e["file"] for e in entries], [e["weight"] for e in entries]
        return [motion_file], [1.0]

    def _finish_load(self):
        dev = self._device
        M = self.num_motions()
        self._motion_ids = torch.arange(M, dtype=torch.long, device=dev)
        shifted = self._motion_num_frames.roll(1)
        shifted[0] = 0
        self._motion_start_idx = shifted.cumsum(0)
        h = self._host_tables
        self._frame_root_pos = h["root_pos"].to(dev)
        self._frame_root_rot = h["root_rot"].to(dev)
        self._frame_joint_rot = h["joint_rot"].to(dev)
        self._frame_root_vel = h["root_vel"].to(dev)
        self._frame_root_ang_vel = h["root_ang_vel"].to(dev)
        self._frame_dof_vel = h["dof_vel"].to(dev)
        self._motion_frames = h["frames"].to(dev)
        if h["contacts"] is not None:
            self._frame_contacts = h["contacts"].to(dev)
        del self._host_tables
        self._finalize_device_tables()

    def _finalize_device_tables(self):
        """Pack the device tables into float4 rows (GPU kernel).  On a CPU device (host-logic tests) the
        unpacked tables exist but queries are unavailable -- there is no CPU fallback."""
        if torch.device(self._device).type != "cuda":
            self._packed = None
            return
        model = self._kin_char_model.c_model()
        contacts = getattr(self, "_frame_contacts", None)
        rows, lay = ops.pack_frames(model, self._frame_root_pos, self._frame_root_rot, self._frame_joint_rot,
                                    contacts, self._frame_root_vel, self._frame_root_ang_vel, self._frame_dof_vel)
        clips = ops.build_clip_meta(self._motion_num_frames, self._motion_loop_modes, self._motion_start_idx,
                                    self._motion_lengths, self._motion_root_pos_delta, self._device)
        self._packed = ops.PackedTables(rows=rows, clips=clips, total_frames=int(rows.shape[0]),
                                        num_clips=self.num_motions(), layout=lay)


def calc_phase(times, motion_len, loop_mode):
    """Module-level twin of `MotionLib.calc_motion_phase` with the reference's signature (anim/motion_lib.py:527-538):
    phase = times / motion_len, fractional part for WRAP clips, clipped to [0, 1].  Elementwise torch on the tensors'
    device (the fused query kernel evaluates the same chain internally, bit for bit: `frame_blend` in
    csrc/motion_query.cu)."""
    phase = times / motion_len
    wrap = loop_mode == LoopMode.WRAP.value
    phase = torch.where(wrap, phase - torch.floor(phase), phase)
    return torch.clip(phase, 0.0, 1.0)

