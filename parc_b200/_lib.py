"""ctypes binding of libparc_b200.so (the C ABI declared in include/parc_b200.h).

There is deliberately no fallback: if the shared library is missing, or an operator is handed a
tensor that is not on a CUDA device, the call raises.  PyTorch is used for device memory and streams
only.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libparc_b200.so")

PARC_MAX_BODIES = 24
PARC_MAX_DOF = 96

JOINT_ROOT, JOINT_HINGE, JOINT_SPHERICAL, JOINT_FIXED = 0, 1, 2, 3
LOOP_CLAMP, LOOP_WRAP = 0, 1

c_float_p = C.POINTER(C.c_float)
c_i64_p = C.POINTER(C.c_int64)
c_i32_p = C.POINTER(C.c_int32)


class ParcCharModel(C.Structure):
    _fields_ = [
        ("num_bodies", C.c_int32), ("dof_size", C.c_int32), ("max_depth", C.c_int32), ("reserved", C.c_int32),
        ("parent", C.c_int32 * PARC_MAX_BODIES), ("depth", C.c_int32 * PARC_MAX_BODIES),
        ("joint_type", C.c_int32 * PARC_MAX_BODIES), ("dof_idx", C.c_int32 * PARC_MAX_BODIES),
        ("local_trans", (C.c_float * 3) * PARC_MAX_BODIES),
        ("local_rot", (C.c_float * 4) * PARC_MAX_BODIES),
        ("joint_axis", (C.c_float * 3) * PARC_MAX_BODIES),
    ]


class ParcRowLayout(C.Structure):
    _fields_ = [("row_floats", C.c_int32), ("pose_slots", C.c_int32), ("contact_slot", C.c_int32),
                ("vel_slot", C.c_int32), ("vel_slots", C.c_int32), ("reserved", C.c_int32 * 3)]


class ParcClipMeta(C.Structure):
    _fields_ = [("num_frames", C.c_int32), ("loop_mode", C.c_int32), ("start_idx", C.c_int64),
                ("length", C.c_float), ("root_pos_delta", C.c_float * 3)]


class ParcMotionTables(C.Structure):
    _fields_ = [("rows", C.c_void_p), ("clips", C.c_void_p), ("total_frames", C.c_int64),
                ("num_clips", C.c_int64), ("row_floats", C.c_int32), ("reserved", C.c_int32), ("tree", C.c_void_p)]


PARC_TREE_BYTES = 880


class ParcFrameOut(C.Structure):
    _fields_ = [("root_pos", C.c_void_p), ("root_rot", C.c_void_p), ("root_vel", C.c_void_p),
                ("root_ang_vel", C.c_void_p), ("joint_rot", C.c_void_p), ("dof_vel", C.c_void_p),
                ("contacts", C.c_void_p), ("frame_idx0", C.c_void_p), ("frame_idx1", C.c_void_p),
                ("blend", C.c_void_p)]


class ParcFkOut(C.Structure):
    _fields_ = [("body_pos", C.c_void_p), ("body_rot", C.c_void_p)]


class ParcHeightfield(C.Structure):
    _fields_ = [("hf", C.c_void_p), ("dim_x", C.c_int32), ("dim_y", C.c_int32), ("min_x", C.c_float),
                ("min_y", C.c_float), ("dx", C.c_float), ("dy", C.c_float)]


class ParcObsSpec(C.Structure):
    _fields_ = [("tmpl_xy", C.c_void_p), ("num_points", C.c_int32), ("relative", C.c_int32),
                ("min_h", C.c_float), ("max_h", C.c_float)]


class ParcQueryArgs(C.Structure):
    _fields_ = [("tables", C.c_void_p), ("motion_ids", C.c_void_p), ("motion_times", C.c_void_p),
                ("frame_idxs", C.c_void_p), ("n", C.c_int64), ("time_offsets", C.c_void_p),
                ("root_xy_offset", C.c_void_p), ("model", C.c_void_p), ("frame", C.c_void_p), ("fk", C.c_void_p),
                ("hf", C.c_void_p), ("obs", C.c_void_p), ("obs_out", C.c_void_p), ("error_flags", C.c_void_p),
                ("num_steps", C.c_int32), ("flags", C.c_uint32), ("variant", C.c_int32), ("reserved", C.c_int32),
                ("tar_obs", C.c_void_p)]


class ParcTarObsSpec(C.Structure):
    _fields_ = [("sim_root_pos", C.c_void_p), ("sim_root_rot", C.c_void_p), ("key_body_ids", C.c_void_p),
                ("obs_out", C.c_void_p), ("out_env_stride", C.c_int64), ("num_keys", C.c_int32),
                ("global_obs", C.c_int32), ("global_tar_root_h", C.c_int32), ("reserved", C.c_int32)]


PARC_QUERY_FAST_HEADING, PARC_QUERY_PDL, PARC_QUERY_PDL_EARLY_INPUTS = 1, 2, 4
PARC_QUERY_ERR_CLIP_ID, PARC_QUERY_ERR_FRAME_IDX = 1, 2


class ParcTerrainBatch(C.Structure):
    _fields_ = [("hf", C.c_void_p), ("hf_batch_stride", C.c_int64), ("min_center", C.c_void_p),
                ("x_nodes", C.c_void_p), ("y_nodes", C.c_void_p), ("base_z", C.c_void_p),
                ("dim_x", C.c_int32), ("dim_y", C.c_int32), ("min_center_stride", C.c_int32),
                ("base_z_stride", C.c_int32), ("half_dx", C.c_float), ("half_dy", C.c_float),
                ("base_z_value", C.c_float), ("reserved", C.c_int32)]


class ParcClipTerrain(C.Structure):
    _fields_ = [("cell_offset", C.c_int64), ("mask_offset", C.c_int64), ("dim_x", C.c_int32), ("dim_y", C.c_int32),
                ("num_frames", C.c_int32), ("reserved", C.c_int32), ("min_x", C.c_float), ("min_y", C.c_float),
                ("dx", C.c_float), ("dy", C.c_float)]


class ParcClipTerrains(C.Structure):
    _fields_ = [("clips", C.c_void_p), ("hf", C.c_void_p), ("hf_maxmin", C.c_void_p), ("mask_words", C.c_void_p),
                ("num_clips", C.c_int64), ("max_mask_words", C.c_int32), ("reserved", C.c_int32)]


class ParcClipHfQuery(C.Structure):
    _fields_ = [("motion_ids", C.c_void_p), ("root_pos", C.c_void_p), ("root_rot", C.c_void_p),
                ("canon_root_z", C.c_void_p), ("frame_lo", C.c_void_p), ("frame_hi", C.c_void_p),
                ("tmpl_xy", C.c_void_p), ("n", C.c_int64), ("grid_x", C.c_int32), ("grid_y", C.c_int32),
                ("centre_x", C.c_int32), ("centre_y", C.c_int32), ("free_max", C.c_float), ("free_min", C.c_float),
                ("hf_out", C.c_void_p), ("maxmin_out", C.c_void_p), ("centre_h_out", C.c_void_p)]


class ParcBodyPoints(C.Structure):
    _fields_ = [("points", C.c_void_p), ("point_start", C.c_void_p), ("num_points", C.c_int32),
                ("reserved", C.c_int32)]


class ParcBodyConstraint(C.Structure):
    _fields_ = [("body", C.c_int32), ("start_frame", C.c_int32), ("end_frame", C.c_int32), ("shape", C.c_int32),
                ("point", C.c_float * 3), ("radius", C.c_float), ("offset", C.c_float * 3), ("reserved", C.c_float)]


PARC_CONSTRAINT_SPHERE, PARC_CONSTRAINT_BOX = 0, 1


class ParcMotionOptArgs(C.Structure):
    _fields_ = [("frames", C.c_void_p), ("num_frames", C.c_int64), ("src_root_pos", C.c_void_p),
                ("src_root_rot", C.c_void_p), ("src_joint_rot", C.c_void_p), ("src_body_vels", C.c_void_p),
                ("src_body_rot_vels", C.c_void_p), ("contacts", C.c_void_p), ("terrain", C.c_void_p),
                ("pts", ParcBodyPoints), ("constraints", C.c_void_p), ("num_constraints", C.c_int32),
                ("reserved", C.c_int32), ("w_root_pos", C.c_float), ("w_root_rot", C.c_float),
                ("w_joint_rot", C.c_float), ("w_smoothness", C.c_float), ("w_penetration", C.c_float),
                ("w_contact", C.c_float), ("w_sliding", C.c_float), ("w_body_constraints", C.c_float),
                ("w_jerk", C.c_float), ("reserved2", C.c_float), ("max_jerk_dt3", C.c_double), ("lr", C.c_double),
                ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double), ("exp_avg", C.c_void_p),
                ("exp_avg_sq", C.c_void_p), ("step", C.c_void_p), ("root_rot", C.c_void_p), ("joint_rot", C.c_void_p),
                ("body_pos", C.c_void_p), ("body_rot", C.c_void_p), ("g_root_pos", C.c_void_p),
                ("g_root_rot", C.c_void_p), ("g_joint_rot", C.c_void_p), ("pen", C.c_void_p), ("con", C.c_void_p),
                ("grad", C.c_void_p), ("terms", C.c_void_p)]


PARC_MAX_KEY_BODIES = 4


class ParcKeyBodies(C.Structure):
    _fields_ = [("num_feet", C.c_int32), ("num_hands", C.c_int32),
                ("foot_body", C.c_int32 * PARC_MAX_KEY_BODIES), ("hand_body", C.c_int32 * PARC_MAX_KEY_BODIES),
                ("foot_half", (C.c_float * 3) * PARC_MAX_KEY_BODIES),
                ("foot_offset", (C.c_float * 3) * PARC_MAX_KEY_BODIES),
                ("hand_radius", C.c_float * PARC_MAX_KEY_BODIES)]


class ParcCharState(C.Structure):
    _fields_ = [("root_pos", C.c_void_p), ("root_rot", C.c_void_p), ("root_vel", C.c_void_p),
                ("root_ang_vel", C.c_void_p), ("joint_rot", C.c_void_p), ("dof_vel", C.c_void_p),
                ("key_pos", C.c_void_p), ("key_body_ids", C.c_void_p), ("num_bodies", C.c_int32),
                ("env_stride", C.c_int32)]


class ParcDoneSpec(C.Structure):
    _fields_ = [("episode_length", C.c_double), ("termination_height", C.c_double),
                ("root_pos_termination_dist", C.c_double), ("root_rot_termination_angle", C.c_double),
                ("pose_termination_dist", C.c_void_p), ("contact_body_mask", C.c_uint32),
                ("has_contact_bodies", C.c_int32), ("pose_termination", C.c_int32),
                ("enable_early_termination", C.c_int32), ("track_root", C.c_int32)]


class ParcSimStep(C.Structure):
    _fields_ = [("sim", ParcCharState), ("ref", ParcCharState), ("dof_pos", C.c_void_p), ("body_pos", C.c_void_p),
                ("ref_body_pos", C.c_void_p), ("contact_force", C.c_void_p), ("time", C.c_void_p),
                ("env_offsets", C.c_void_p), ("joint_rot_err_w", C.c_void_p), ("dof_err_w", C.c_void_p),
                ("tar_contacts", C.c_void_p), ("char_contacts", C.c_void_p), ("done", ParcDoneSpec),
                ("hf", ParcHeightfield), ("offset_stride", C.c_int32), ("tar_env_stride", C.c_int32),
                ("num_tar_steps", C.c_int32), ("num_keys", C.c_int32), ("global_obs", C.c_int32),
                ("root_height_obs", C.c_int32), ("track_root_h", C.c_int32), ("track_root", C.c_int32),
                ("joint_rot_out", C.c_void_p), ("char_obs_out", C.c_void_p), ("tar_contacts_out", C.c_void_p),
                ("char_contacts_out", C.c_void_p), ("reward_out", C.c_void_p), ("done_out", C.c_void_p),
                ("obs_stride", C.c_int64), ("phase", C.c_int32), ("reserved", C.c_int32)]


PARC_SIM_STEP_ALL, PARC_SIM_STEP_PRE, PARC_SIM_STEP_POST = 0, 1, 2
PARC_MAX_PEERS, PARC_MAX_PUSH_SEGMENTS = 16, 4


class ParcPeerSignals(C.Structure):
    _fields_ = [("multicast_signal", C.c_void_p), ("peer_signal", C.c_void_p * PARC_MAX_PEERS),
                ("local_signal", C.c_void_p), ("epoch", C.c_void_p), ("world", C.c_int32), ("num_slots", C.c_int32),
                ("timeout_ns", C.c_int64), ("timeout_flag", C.c_void_p), ("rank", C.c_int32), ("reserved", C.c_int32)]


class ParcPeerSegment(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst_multicast", C.c_void_p), ("dst_peer", C.c_void_p * PARC_MAX_PEERS),
                ("bytes", C.c_int64)]


PARC_DONE_NULL, PARC_DONE_FAIL, PARC_DONE_SUCC, PARC_DONE_TIME = 0, 1, 2, 3

# name -> (restype, argtypes); every symbol include/parc_b200.h declares
_V, _I64, _I32, _F = C.c_void_p, C.c_int64, C.c_int32, C.c_float
_P = C.POINTER
SIGNATURES = {
    "parc_abi_version": (C.c_int, []),
    "parc_error_string": (C.c_char_p, [C.c_int]),
    "parc_row_layout": (C.c_int, [_P(ParcCharModel), _P(ParcRowLayout)]),
    "parc_validate_model": (C.c_int, [_P(ParcCharModel)]),
    "parc_tree_from_model": (C.c_int, [_P(ParcCharModel), _V]),
    "parc_pack_frames": (C.c_int, [_V, _V, _V, _V, _V, _V, _V, _I64, _P(ParcCharModel), _V, _V]),
    "parc_motion_query": (C.c_int, [_P(ParcMotionTables), _V, _V, _I64, _P(ParcCharModel), _P(ParcFrameOut),
                                    _P(ParcFkOut), _P(ParcHeightfield), _P(ParcObsSpec), _V, _V]),
    "parc_motion_query_steps": (C.c_int, [_P(ParcMotionTables), _V, _V, _I64, _V, _I32, _V, _P(ParcCharModel),
                                          _P(ParcFrameOut), _P(ParcFkOut), _P(ParcHeightfield), _P(ParcObsSpec),
                                          _V, _V]),
    "parc_get_motion_frame": (C.c_int, [_P(ParcMotionTables), _V, _V, _I64, _P(ParcCharModel),
                                        _P(ParcFrameOut), _P(ParcFkOut), _V]),
    "parc_motion_query_ex": (C.c_int, [_P(ParcQueryArgs), _V]),
    "parc_fk_fwd": (C.c_int, [_V, _V, _V, _I64, _P(ParcCharModel), _V, _V, _V]),
    "parc_fk_bwd": (C.c_int, [_V, _V, _V, _V, _I64, _P(ParcCharModel), _V, _V, _V, _V]),
    "parc_dof_to_rot_fwd": (C.c_int, [_V, _I64, _P(ParcCharModel), _V, _V]),
    "parc_dof_to_rot_bwd": (C.c_int, [_V, _V, _I64, _P(ParcCharModel), _V, _V]),
    "parc_rot_to_dof": (C.c_int, [_V, _I64, _P(ParcCharModel), _V, _V]),
    "parc_exp_map_to_quat_fwd": (C.c_int, [_V, _I64, _V, _V]),
    "parc_exp_map_to_quat_bwd": (C.c_int, [_V, _V, _I64, _V, _V]),
    "parc_hf_sample": (C.c_int, [_P(ParcHeightfield), _V, _I64, _V, _V, _V]),
    "parc_selftest_grid_index": (C.c_int, [_F, _F, _I32, _V, _V]),
    "parc_hf_obs": (C.c_int, [_P(ParcHeightfield), _P(ParcObsSpec), _V, _I32, _V, _V, _V, _I32, _I64, _V, _I64, _V]),
    "parc_clip_hf_gather": (C.c_int, [_P(ParcClipTerrains), _P(ParcClipHfQuery), _V]),
    "parc_points_hf_sdf": (C.c_int, [_V, _I64, _I64, _P(ParcTerrainBatch), _I32, _V, _V, _V]),
    "parc_points_hf_sdf_bwd": (C.c_int, [_V, _I64, _I64, _P(ParcTerrainBatch), _I32, _V, _V, _V, _V]),
    "parc_body_points_fwd": (C.c_int, [_V, _V, _I64, _I64, _I32, _P(ParcBodyPoints), _V, _V]),
    "parc_body_points_bwd": (C.c_int, [_V, _V, _I64, _I64, _I32, _P(ParcBodyPoints), _V, _V, _V]),
    "parc_frames_fk": (C.c_int, [_V, _I64, _I32, _P(ParcCharModel), _V, _V, _V, _V, _V]),
    "parc_clip_label": (C.c_int, [_V, _I64, _I64, _I32, _P(ParcCharModel), _P(ParcBodyPoints), _P(ParcTerrainBatch),
                                  _P(ParcKeyBodies), _F, _V, _V, _V, _V, _V, _V, _V, _V]),
    "parc_body_loss": (C.c_int, [_V, _V, _V, _V, _I64, _I64, _P(ParcCharModel), _P(ParcBodyPoints),
                                 _P(ParcTerrainBatch), _F, _F, _V, _V, _V, _V, _V, _V]),
    "parc_build_tables": (C.c_int, [_V, _I64, _I32, _V, _V, _V, _V, _V, _V, _I64, _P(ParcCharModel), _V, _V]),
    "parc_motion_opt_source": (C.c_int, [_V, _I64, _P(ParcCharModel), _V, _V, _V, _V, _V, _V, _V]),
    "parc_motion_opt_loss_grad": (C.c_int, [_P(ParcMotionOptArgs), _P(ParcCharModel), _V]),
    "parc_motion_opt_adam_step": (C.c_int, [_P(ParcMotionOptArgs), _P(ParcCharModel), _V]),
    "parc_motion_opt_iteration": (C.c_int, [_P(ParcMotionOptArgs), _P(ParcCharModel), _V]),
    "parc_char_obs": (C.c_int, [_P(ParcCharState), _I64, _I32, _I32, _I32, _I32, _I32, _V, _I64, _V]),
    "parc_tar_obs": (C.c_int, [_V, _V, _V, _V, _V, _V, _I64, _I32, _I32, _I32, _I32, _I32, _I32, _V, _I32, _V, _I64,
                               _V]),
    "parc_deepmimic_reward": (C.c_int, [_P(ParcCharState), _P(ParcCharState), _I64, _I32, _I32, _I32, _V, _V, _I32,
                                        _I32, _V, _V]),
    "parc_sim_step": (C.c_int, [_P(ParcSimStep), _I64, _P(ParcCharModel), _V]),
    "parc_peer_barrier": (C.c_int, [_P(ParcPeerSignals), _I32, _V]),
    "parc_peer_push": (C.c_int, [_P(ParcPeerSegment), _I32, _P(ParcPeerSignals), _I32, _V]),
    "parc_done": (C.c_int, [_P(ParcDoneSpec), _V, _V, _V, _V, _V, _V, _V, _P(ParcHeightfield), _V, _I32, _I32, _I64,
                            _I32, _V, _V, _V]),
}

_lib: Optional[C.CDLL] = None


class ParcLibraryError(RuntimeError):
    pass


def load() -> C.CDLL:
    """dlopen libparc_b200.so and bind every declared symbol.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ParcLibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or parc_b200/csrc/build.sh -- there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


# number of kernel-launching C-ABI calls made by this process (bench.py reports it as gpu_launches)
LAUNCHES = [0]
_HOST_ONLY = {"parc_validate_model", "parc_row_layout", "parc_tree_from_model"}


def check(rc: int, what: str):
    if what not in _HOST_ONLY:
        LAUNCHES[0] += 1
    if rc != 0:
        msg = load().parc_error_string(rc).decode()
        raise ParcLibraryError(f"{what} failed: {msg} (code {rc})")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def require_cuda(*tensors: torch.Tensor):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise ParcLibraryError("parc_b200 operators need CUDA tensors (no CPU fallback); got a "
                                   f"{t.device} tensor")


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def f32c(t: torch.Tensor) -> torch.Tensor:
    """fp32 + contiguous (no copy when already so)."""
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    return t.contiguous()
