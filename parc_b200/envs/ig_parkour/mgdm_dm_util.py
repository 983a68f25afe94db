"""Tracker-side callers of the kinematic-query path.

Drop-in for the two functions of the reference's `envs/ig_parkour/mgdm_dm_util.py` that sit on the
path: `RefCharEnv._refresh_ray_obs_hfs` (:158-179) -- here the free function `refresh_ray_obs_hfs` --
and `fetch_tar_obs_data` (:279-302).  The Isaac Gym environment classes around them are out of scope.
"""
from __future__ import annotations

import torch

from ... import ops


def refresh_ray_obs_hfs(char_root_pos_xyz, char_heading, ray_xy_points, terrain, min_obs_h, max_obs_h, out=None):
    """clamp(hf(R(heading) * ray_xy_points + root_xy) - root_z, min_obs_h, max_obs_h) -> [N, P].
    One launch (csrc/heightfield.cu) instead of ~39 eager ops and [N*P] int64 index tensors."""
    return ops.hf_obs(terrain.hf_desc(), ray_xy_points, char_root_pos_xyz, char_heading, relative=True,
                      min_h=min_obs_h, max_h=max_obs_h, out=out)


def fetch_tar_obs_data(motion_ids, motion_times, mlib, timestep, tar_obs_steps):
    """Future target frames: ids tiled x len(tar_obs_steps), t + timestep * steps (:279-302)."""
    n = motion_ids.shape[0]
    num_steps = tar_obs_steps.shape[0]
    assert num_steps > 0
    times = (motion_times.unsqueeze(-1) + timestep * tar_obs_steps).flatten()
    ids = torch.broadcast_to(motion_ids.unsqueeze(-1), (n, num_steps)).flatten()
    fr = mlib.calc_motion_frame(ids, times)
    root_pos, root_rot, joint_rot, contacts = fr[0], fr[1], fr[4], fr[6]
    return (root_pos.reshape(n, num_steps, 3), root_rot.reshape(n, num_steps, 4),
            joint_rot.reshape(n, num_steps, joint_rot.shape[-2], 4), contacts.reshape(n, num_steps, -1))
