"""Proprioceptive observation of the simulated character.

Drop-in for the free function `compute_char_obs` of the reference's `envs/ig_char_env.py` (:582-626); the
Isaac Gym environment class around it is out of scope.  One launch (csrc/tracker_step.cu) instead of ~60
eager ops.
"""
from __future__ import annotations

from .. import ops


def compute_char_obs(root_pos, root_rot, root_vel, root_ang_vel, joint_rot, dof_vel, key_pos, global_obs,
                     root_height_obs):
    """[root z] | root tan-norm | root_vel | root_ang_vel | joint tan-norms | dof_vel | key positions -> [N, W].
    `key_pos` [N,K,3] world space, or an empty tensor when the character has no key bodies."""
    return ops.char_obs(root_pos, root_rot, root_vel, root_ang_vel, joint_rot, dof_vel, key_pos, global_obs,
                        root_height_obs)
