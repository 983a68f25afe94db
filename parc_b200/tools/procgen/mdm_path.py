"""compute_motion_loss: penetration + contact loss used to rank generated motions.

Drop-in for the reference's `tools/procgen/mdm_path.py::compute_motion_loss` (:31-127): same arguments,
same dict of [B] tensors.  The reference loops over the 15 bodies and, per body, materialises
[B, F*P_b, X*Y, 3] tensors twice; here FK, the surface-point transform, the exact point/heightfield SDF
and both reductions are one launch of the fused kernel in csrc/body_loss.cu.
"""
from __future__ import annotations

import torch

from ... import ops
from ...util.motion_util import MotionFrames


class MDMPathSettings:
    """Defaults of the reference (:19-29)."""
    next_node_lookahead = 7
    rewind_num_frames = 5
    end_of_path_buffer = 2
    max_motion_length = 10.0
    path_batch_size = 16
    mdm_batch_size = 32
    top_k = 4
    w_target = 2.0
    w_contact = 0.1
    w_pen = 0.1


def body_points_desc(char_model, body_points):
    """Cache the concatenated device copy of the per-body point lists on the model."""
    cache = char_model.__dict__.setdefault("_body_points_cache", {})
    key = tuple((id(p), p.data_ptr(), tuple(p.shape), p._version) for p in body_points)
    if key not in cache:
        cache.clear()
        # the keyed tensors are kept alive next to the descriptor, so a recycled allocator address / object id can
        # never alias a stale entry
        cache[key] = (ops.make_body_points(body_points, body_points[0].device), list(body_points))
    return cache[key][0]


def compute_motion_loss(motion_frames: MotionFrames, path_nodes, terrain, char_model, body_points: list,
                        w_contact: float, w_pen: float, w_path: float = 0.0, verbose: bool = True):
    """motion_frames fields are [B, F, ...]; one terrain for the whole batch.  base_z = min(hf) - 10 as
    in the reference (:77, :92), evaluated on the device (no host round trip)."""
    root_pos, root_rot = motion_frames.root_pos, motion_frames.root_rot
    joint_rot, contacts = motion_frames.joint_rot, motion_frames.contacts
    base_z = (torch.min(terrain.hf) - 10.0).reshape(1)
    tb = ops.make_terrain_batch(terrain.hf, terrain.min_point, terrain.dxdy.detach().cpu().tolist(), base_z=base_z)
    pts = body_points_desc(char_model, body_points)
    _, pen, con = ops.body_loss(char_model.c_model(), pts, tb, root_pos, root_rot, joint_rot, contacts, 1.0, 1.0)
    contact_loss = con * w_contact
    penetration_loss = pen * w_pen
    return {"total_loss": contact_loss + penetration_loss, "contact_loss": contact_loss,
            "pen_loss": penetration_loss}
