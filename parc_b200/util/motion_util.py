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
