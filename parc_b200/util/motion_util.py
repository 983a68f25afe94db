"""MotionFrames: the struct-of-tensors the loss functions take (reference util/motion_util.py:6-222).
Only the container and the helpers the kinematic path uses."""
from __future__ import annotations

import torch

from . import torch_util

_FIELDS = ("root_pos", "root_rot", "joint_rot", "body_pos", "body_rot", "contacts")


class MotionFrames:
    def __init__(self, root_pos=None, root_rot=None, joint_rot=None, body_pos=None, body_rot=None, contacts=None):
        self.root_pos, self.root_rot, self.joint_rot = root_pos, root_rot, joint_rot
        self.body_pos, self.body_rot, self.contacts = body_pos, body_rot, contacts

    def _map(self, fn):
        return MotionFrames(**{k: (None if getattr(self, k) is None else fn(getattr(self, k))) for k in _FIELDS})

    def init_blank_frames(self, char_model, history_length: int, batch_size=1):
        dev, J = char_model._device, char_model.get_num_joints()
        z = lambda *s: torch.zeros(size=[batch_size, history_length, *s], dtype=torch.float32, device=dev)
        self.root_pos, self.root_rot, self.joint_rot = z(3), z(4), z(J - 1, 4)
        self.root_rot[..., 3] = 1.0
        self.joint_rot[..., 3] = 1.0
        self.body_pos, self.body_rot, self.contacts = z(J, 3), z(J, 4), z(J)

    def get_mlib_format(self, char_model):
        frames = torch.cat([self.root_pos, torch_util.quat_to_exp_map(self.root_rot),
                            char_model.rot_to_dof(self.joint_rot)], dim=-1)
        return frames, self.contacts

    def get_slice(self, in_slice):
        return self._map(lambda t: t[:, in_slice])

    def unsqueeze(self, dim):
        return self._map(lambda t: t.unsqueeze(dim))

    def squeeze(self, dim):
        return self._map(lambda t: t.squeeze(dim))

    def expand_first_dim(self, b):
        return self._map(lambda t: t.expand(b, *t.shape[1:]))

    def get_copy(self, new_device=None):
        return self._map(lambda t: t.clone() if new_device is None else t.clone().to(new_device))


def cat_motion_frames(motion_frames_list):
    """Concatenate [B, F_i, ...] MotionFrames along time; a field is kept iff the first element has it.
    Ref util/motion_util.py:145-193."""
    first = motion_frames_list[0]
    assert first.root_pos.dim() == 3
    return MotionFrames(**{k: (None if getattr(first, k) is None else
                               torch.cat([getattr(m, k) for m in motion_frames_list], dim=1)) for k in _FIELDS})


def motion_frames_from_mlib_format(mlib_motion_frames, char_model, contacts=None):
    """[..., 6+D] raw frames -> MotionFrames with FK filled in (one launch on CUDA frames: exp-map / DoF
    conversion + FK fused, csrc/dataset_sweep.cu).  Ref util/motion_util.py:195-222."""
    root_pos = mlib_motion_frames[..., 0:3]
    if mlib_motion_frames.is_cuda and not mlib_motion_frames.requires_grad:
        from .. import ops
        bp, br, rr, jr = ops.frames_fk(char_model.c_model(), mlib_motion_frames, want_rot=True)
        return MotionFrames(root_pos=root_pos, root_rot=rr, joint_rot=jr, body_pos=bp, body_rot=br, contacts=contacts)
    root_rot = torch_util.exp_map_to_quat(mlib_motion_frames[..., 3:6])
    joint_rot = char_model.dof_to_rot(mlib_motion_frames[..., 6:])
    body_pos, body_rot = char_model.forward_kinematics(root_pos, root_rot, joint_rot)
    return MotionFrames(root_pos=root_pos, root_rot=root_rot, joint_rot=joint_rot, body_pos=body_pos, body_rot=body_rot,
                        contacts=contacts)
