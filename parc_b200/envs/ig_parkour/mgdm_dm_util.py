"""Tracker-side callers of the kinematic-query path.

Drop-in for the functions of the reference's `envs/ig_parkour/mgdm_dm_util.py` that sit on the
path: `RefCharEnv._refresh_ray_obs_hfs` (:158-179) -- here the free function `refresh_ray_obs_hfs` --
and `fetch_tar_obs_data` (:279-302); and for the step assembly that follows it:
`compute_tar_obs` (:462-518), `compute_deepmimic_obs` (:520-553), `compute_deepmimic_reward` (:328-397),
`compute_done` (:399-460) and `RefCharEnv.update_done` (:205-230) -- here the free function `update_done`.
The Isaac Gym environment classes around them are out of scope.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from ... import ops
from .. import ig_char_env


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


def compute_tar_obs(ref_root_pos, ref_root_rot, tar_root_pos, tar_root_rot, joint_rot, tar_key_pos, global_obs,
                    global_tar_root_h_obs):
    """Future targets [N,S,...] expressed against the character -> [N,S, 3 + 6 + 6(J-1) + 3K].  One launch."""
    return ops.tar_obs(ref_root_pos, ref_root_rot, tar_root_pos, tar_root_rot, joint_rot, tar_key_pos, global_obs,
                       global_tar_root_h_obs)


def compute_deepmimic_obs(root_pos, root_rot, root_vel, root_ang_vel, joint_rot, dof_vel, key_pos, global_obs,
                          root_height_obs, enable_tar_obs, tar_root_pos, tar_root_rot, tar_joint_rot, tar_key_pos):
    """OrderedDict(char_obs [, tar_obs]) as in :520-553: two launches."""
    obs = OrderedDict()
    obs["char_obs"] = ig_char_env.compute_char_obs(root_pos, root_rot, root_vel, root_ang_vel, joint_rot, dof_vel,
                                                   key_pos, global_obs, root_height_obs)
    if enable_tar_obs:
        obs["tar_obs"] = compute_tar_obs(root_pos, root_rot, tar_root_pos, tar_root_rot, tar_joint_rot, tar_key_pos,
                                         global_obs, False)
    return obs


def compute_deepmimic_reward(root_pos, root_rot, root_vel, root_ang_vel, joint_rot, dof_vel, key_pos,
                             tar_root_pos, tar_root_rot, tar_root_vel, tar_root_ang_vel, tar_joint_rot, tar_dof_vel,
                             tar_key_pos, joint_rot_err_w, dof_err_w, track_root_h, track_root):
    """[N,5] = exp(-scale * err) for pose, vel, root pose, root vel, key pos.  One launch.  Raises ValueError
    without key bodies (the reference raises there too: it stacks a [0] tensor with [N] ones)."""
    return ops.deepmimic_reward((root_pos, root_rot, root_vel, root_ang_vel, joint_rot, dof_vel, key_pos),
                                (tar_root_pos, tar_root_rot, tar_root_vel, tar_root_ang_vel, tar_joint_rot, tar_dof_vel,
                                 tar_key_pos), joint_rot_err_w, dof_err_w, track_root_h, track_root)


_BODY_ID_CACHE = {}


def _host_ids(contact_body_ids):
    """The allowed-contact body ids are configuration; read them back once per tensor, not per step."""
    if not torch.is_tensor(contact_body_ids):
        return tuple(int(i) for i in contact_body_ids)
    key = (contact_body_ids.data_ptr(), contact_body_ids._version, contact_body_ids.numel(), str(contact_body_ids.device))
    ids = _BODY_ID_CACHE.get(key)
    if ids is None:
        if len(_BODY_ID_CACHE) > 64:
            _BODY_ID_CACHE.clear()
        ids = _BODY_ID_CACHE[key] = tuple(int(i) for i in contact_body_ids.tolist())
    return ids


def compute_done(done_buf, time, ep_len, root_rot, body_pos, char_root_pos, tar_root_rot, tar_body_pos,
                 contact_force, contact_body_ids, termination_heights, pose_termination, pose_termination_dist,
                 global_obs, enable_early_termination, track_root, root_pos_termination_dist,
                 root_rot_termination_angle):
    """Episode flags (DoneFlags values) in done_buf's dtype.  `char_root_pos` and `global_obs` are unused, as in
    the reference.  One launch."""
    flags = ops.done_flags(time, ep_len, root_rot, body_pos, tar_root_rot, tar_body_pos, contact_force,
                           _host_ids(contact_body_ids), pose_termination, pose_termination_dist,
                           enable_early_termination, track_root, root_pos_termination_dist, root_rot_termination_angle,
                           termination_heights=termination_heights)
    return flags if done_buf.dtype == torch.int32 else flags.to(done_buf.dtype)


def update_done(done_buf, time, terrain, env_offsets, termination_height, episode_length, contact_body_ids,
                pose_termination, pose_termination_dist, global_obs, enable_early_termination, track_root,
                root_pos_termination_dist, root_rot_termination_angle, root_rot, body_pos, tar_root_rot, tar_body_pos,
                contact_force):
    """RefCharEnv.update_done (:205-230) as a free function: the terrain height under every body
    (`terrain.hf` at body xy + env_offsets[:, 0:2], nearest cell) + termination_height and compute_done, fused
    into ONE launch; writes done_buf in place and returns it."""
    flags = ops.done_flags(time, episode_length, root_rot, body_pos, tar_root_rot, tar_body_pos, contact_force,
                           _host_ids(contact_body_ids), pose_termination, pose_termination_dist,
                           enable_early_termination, track_root, root_pos_termination_dist, root_rot_termination_angle,
                           hf=terrain.hf_desc(), env_offsets=env_offsets, termination_height=termination_height)
    done_buf[:] = flags
    return done_buf
