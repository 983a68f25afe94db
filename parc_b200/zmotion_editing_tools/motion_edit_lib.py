"""Contact labelling on the heightfield.

Drop-in for the three labelling functions of the reference's `zmotion_editing_tools/motion_edit_lib.py`
that BASELINE config 5 (dataset sweep) names: `compute_hf_foot_contacts_and_correct_pen` (:654-706) and
`compute_motion_terrain_hand_contacts` (:708-747).  Same arguments and returns; FK, the box-corner /
rounded-box tests and the reductions run in one launch of `parc_clip_label` (csrc/dataset_sweep.cu).
The rest of that file (MotionData editing helpers, GUI glue) is out of scope.
"""
from __future__ import annotations

import torch

from .. import ops
from ..anim.kin_char_model import GeomType


def key_bodies(char_model, feet=("left_foot", "right_foot"), hands=("left_hand", "right_hand")):
    """The ParcKeyBodies POD for a character: box feet / sphere hands read from the first geom of each
    named body, as the reference does (`char_model._geoms[body_id][0]`, :679-686, :731-733)."""
    cache = char_model.__dict__.setdefault("_key_bodies_cache", {})
    key = (tuple(feet), tuple(hands))
    if key not in cache:
        f, h = [], []
        for name in feet:
            b = char_model.get_body_id(name)
            g = char_model.get_geoms(b)[0]
            assert g._shape_type == GeomType.BOX, "foot bodies are expected to carry a box geom"
            f.append((b, g._dims.detach().cpu().tolist(), g._offset.detach().cpu().tolist()))
        for name in hands:
            b = char_model.get_body_id(name)
            g = char_model.get_geoms(b)[0]
            h.append((b, float(g._dims.reshape(-1)[0].item())))
        cache[key] = ops.make_key_bodies(f, h)
    return cache[key]


def _terrain_batch(terrain):
    base_z = (torch.min(terrain.hf) - 10.0).reshape(1)            # :739, evaluated on the device
    return ops.make_terrain_batch(terrain.hf, terrain.min_point, terrain.dxdy.detach().cpu().tolist(), base_z=base_z)


def _feet(char_model):
    kb = key_bodies(char_model)
    return [(int(kb.foot_body[i]), list(kb.foot_half[i]), list(kb.foot_offset[i])) for i in range(kb.num_feet)]


def _hands(char_model):
    kb = key_bodies(char_model)
    return [(int(kb.hand_body[i]), float(kb.hand_radius[i])) for i in range(kb.num_hands)]


def compute_hf_foot_contacts_and_correct_pen(motion_frames, terrain, char_model, contact_eps=0.04):
    """-> (updated_motion_frames with root z lifted out of the terrain, contacts [F, J] with the two foot
    columns set).  Ref :654-706."""
    out = ops.clip_label(char_model.c_model(), None, _terrain_batch(terrain), ops.make_key_bodies(_feet(char_model), []),
                         motion_frames.unsqueeze(0), contact_eps)
    updated = motion_frames.clone()
    updated[:, 2] -= out["pen_correction"][0]
    return updated, out["contacts"][0]


def compute_motion_terrain_hand_contacts(motion_frames, terrain, char_model, contact_eps=0.04):
    """-> contacts [F, J] with the two hand columns set.  Ref :708-747."""
    out = ops.clip_label(char_model.c_model(), None, _terrain_batch(terrain), ops.make_key_bodies([], _hands(char_model)),
                         motion_frames.unsqueeze(0), contact_eps)
    return out["contacts"][0]


def label_clips(motion_frames, terrain_batch, char_model, body_points=None, contact_eps=0.04, want_masks=False,
                want_body_hf=True, want_fk=False):
    """Batched form for dataset sweeps (no single-call reference counterpart): frames [B, F, 6+D], one
    terrain per clip (`ops.TerrainBatchDesc`) -> dict(contacts, pen_correction, body_hf, frame_mask_bits,
    min_body_heights, body_pos, body_rot).  Foot AND hand contacts in the same launch."""
    from ..tools.procgen.mdm_path import body_points_desc
    pts = body_points_desc(char_model, body_points) if body_points is not None else None
    return ops.clip_label(char_model.c_model(), pts, terrain_batch, key_bodies(char_model), motion_frames, contact_eps,
                          want_masks=want_masks, want_body_hf=want_body_hf, want_fk=want_fk)
