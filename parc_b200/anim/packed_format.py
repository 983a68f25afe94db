"""On-disk packed motion library.

The reference re-reads one pickle per clip and re-derives quaternions and velocities per clip on every start
(anim/motion_lib.py:204-380).  A `.parcpack` file stores the result once: the packed float4 frame rows exactly
as the query kernel reads them, the clip records, the raw frames and the per-clip terrains.  Opening one is a
memory-map, one host->device copy per array and zero per-clip work:

    lib = MotionLib(yaml_path, kin_char_model, "cuda:0", contact_info=True)     # once
    lib.save_packed("library.parcpack")
    lib = MotionLib("library.parcpack", kin_char_model, "cuda:0", init_type="packed_file", contact_info=True)

Layout (little endian):
    bytes 0..7    magic  b"PARCPACK"
    bytes 8..11   uint32 format version (1)
    bytes 12..15  uint32 length H of the JSON header
    bytes 16..    JSON header (utf-8): library metadata + {array name: dtype, shape, byte offset}
    ...           raw arrays, each starting on a 64-byte boundary (offsets are absolute)
The header records the character's body count / DoF count / row layout and a fingerprint of the kinematic tree;
opening a file with a different character model is refused.
"""
from __future__ import annotations

import hashlib
import json
import os
from typing import Dict

import numpy as np
import torch

from .. import ops
from .._lib import ParcLibraryError

MAGIC = b"PARCPACK"
VERSION = 1
_ALIGN = 64


def model_fingerprint(kin_char_model) -> str:
    """Hash of what the packed rows depend on: tree topology, joint types / axes / DoF slots, local frames."""
    m = kin_char_model.c_model()
    J = int(m.num_bodies)
    h = hashlib.sha256()
    h.update(np.array([J, int(m.dof_size)], np.int32).tobytes())
    for name in ("parent", "joint_type", "dof_idx"):
        h.update(np.array(list(getattr(m, name))[:J], np.int32).tobytes())
    for name in ("joint_axis", "local_trans", "local_rot"):
        h.update(np.array([list(r) for r in list(getattr(m, name))[:J]], np.float32).tobytes())
    return h.hexdigest()[:32]


def write_container(path: str, meta: dict, arrays: Dict[str, np.ndarray]) -> None:
    """Generic writer: header + aligned raw arrays (atomic: written to `path + '.tmp'`, then renamed)."""
    index, blobs = {}, []
    for name, a in arrays.items():
        a = np.ascontiguousarray(a)
        if a.dtype.byteorder == ">":
            a = a.astype(a.dtype.newbyteorder("<"))
        index[name] = {"dtype": a.dtype.str, "shape": list(a.shape)}
        blobs.append((name, a))
    # offsets depend on the header length, which depends on the offsets' digits: reserve fixed-width fields
    for name in index:
        index[name]["offset"] = 0
    probe = json.dumps({"meta": meta, "arrays": index}).encode()
    header_room = len(probe) + 24 * len(index) + _ALIGN
    off = (16 + header_room + _ALIGN - 1) // _ALIGN * _ALIGN
    for name, a in blobs:
        index[name]["offset"] = off
        off = (off + a.nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
    header = json.dumps({"meta": meta, "arrays": index}).encode()
    assert len(header) <= header_room
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(MAGIC)
        f.write(np.array([VERSION, len(header)], "<u4").tobytes())
        f.write(header)
        for name, a in blobs:
            f.seek(index[name]["offset"])
            f.write(a.tobytes())
        f.truncate(max(off, f.tell()))
    os.replace(tmp, path)


def read_container(path: str):
    """-> (meta, {name: read-only numpy view of the memory-mapped file})."""
    size = os.path.getsize(path)
    if size < 16:
        raise ValueError(f"{path}: not a PARCPACK file (too short)")
    mm = np.memmap(path, dtype=np.uint8, mode="r")
    if bytes(mm[0:8]) != MAGIC:
        raise ValueError(f"{path}: not a PARCPACK file (bad magic)")
    version, hlen = (int(x) for x in np.frombuffer(bytes(mm[8:16]), "<u4"))
    if version != VERSION:
        raise ValueError(f"{path}: unsupported PARCPACK version {version} (this build reads {VERSION})")
    if 16 + hlen > size:
        raise ValueError(f"{path}: truncated header")
    head = json.loads(bytes(mm[16:16 + hlen]).decode())
    arrays = {}
    for name, d in head["arrays"].items():
        dt = np.dtype(d["dtype"])
        n = int(np.prod(d["shape"], dtype=np.int64)) * dt.itemsize
        if d["offset"] + n > size:
            raise ValueError(f"{path}: array {name!r} runs past the end of the file (truncated?)")
        arrays[name] = mm[d["offset"]:d["offset"] + n].view(dt).reshape(d["shape"])
    return head["meta"], arrays


def save(mlib, path: str) -> None:
    """Write `mlib` (a CUDA-resident MotionLib) to `path`."""
    if mlib._packed is None:
        raise ParcLibraryError("save_packed needs a CUDA-resident MotionLib (the packed rows live on the GPU)")
    kcm = mlib._kin_char_model
    lay = mlib._packed.layout
    cpu = lambda t: t.detach().cpu().numpy()
    arrays = {
        "rows": cpu(mlib._packed.rows), "frames": cpu(mlib._motion_frames.contiguous()),
        "num_frames": cpu(mlib._motion_num_frames), "start_idx": cpu(mlib._motion_start_idx),
        "lengths": cpu(mlib._motion_lengths), "loop_modes": cpu(mlib._motion_loop_modes), "fps": cpu(mlib._motion_fps),
        "dt": cpu(mlib._motion_dt), "weights": cpu(mlib._motion_weights),
        "root_pos_delta": cpu(mlib._motion_root_pos_delta),
    }
    terrains = getattr(mlib, "_terrains", None) or []
    t_meta = []
    for i, t in enumerate(terrains):
        if t is None:
            t_meta.append(None)
            continue
        arrays[f"terrain{i}.hf"] = cpu(t.hf)
        arrays[f"terrain{i}.hf_mask"] = cpu(t.hf_mask).astype(np.uint8)
        arrays[f"terrain{i}.hf_maxmin"] = cpu(t.hf_maxmin)
        t_meta.append({"name": t.terrain_name, "min_point": [float(v) for v in t.min_point.tolist()],
                       "dxdy": [float(v) for v in t.dxdy.tolist()]})
    # per-clip, per-frame body-cover cell lists (`_hf_mask_inds`, consumed by the MDM sampler,
    # diffusion/mdm_heightfield_contact_motion_sampler.py:433): ragged -> one [n,2] index array + per-frame counts
    mask_inds = getattr(mlib, "_hf_mask_inds", None) or []
    has_masks = []
    for i, inds in enumerate(mask_inds):
        if inds is None:
            has_masks.append(False)
            continue
        per = [np.asarray(t.detach().cpu() if isinstance(t, torch.Tensor) else t, dtype=np.int64).reshape(-1, 2) for t in inds]
        arrays[f"mask{i}.count"] = np.array([p.shape[0] for p in per], dtype=np.int64)
        arrays[f"mask{i}.inds"] = np.concatenate(per, axis=0) if per else np.zeros((0, 2), np.int64)
        has_masks.append(True)
    extras = []
    for e in (getattr(mlib, "_motion_extras", None) or []):
        try:
            json.dumps(e)
            extras.append(e)
        except (TypeError, ValueError):
            import warnings
            warnings.warn("save_packed: a clip's `extra` field is not JSON-serialisable and is dropped")
            extras.append(None)
    meta = {
        "hf_mask_inds": has_masks, "motion_extras": extras,
        "num_clips": int(mlib.num_motions()), "total_frames": int(mlib._packed.total_frames),
        "row_floats": int(lay.row_floats), "num_bodies": int(kcm.get_num_joints()), "dof_size": int(kcm.get_dof_size()),
        "contact_info": bool(mlib._contact_info), "model_fingerprint": model_fingerprint(kcm),
        "motion_names": list(getattr(mlib, "_motion_names", []) or []),
        "motion_files": list(getattr(mlib, "_motion_files", []) or []), "terrains": t_meta,
    }
    write_container(path, meta, arrays)


def load_into(mlib, path: str) -> None:
    """Fill a freshly constructed MotionLib (device, model and contact_info already set) from `path`."""
    from ..util.terrain_util import SubTerrain
    meta, a = read_container(path)
    kcm = mlib._kin_char_model
    dev = mlib._device
    if torch.device(dev).type != "cuda":
        raise RuntimeError("a packed library is opened straight onto a CUDA device (there is no CPU query path)")
    if meta["num_bodies"] != kcm.get_num_joints() or meta["dof_size"] != kcm.get_dof_size() \
            or meta["model_fingerprint"] != model_fingerprint(kcm):
        raise ValueError(f"{path} was packed for a different character model")
    lay = ops.row_layout(kcm.c_model())
    if int(lay.row_floats) != meta["row_floats"] or list(a["rows"].shape) != [meta["total_frames"], meta["row_floats"]]:
        raise ValueError(f"{path}: row layout does not match this build")
    if mlib._contact_info and not meta["contact_info"]:
        raise ValueError(f"{path} holds no contact flags but contact_info=True was requested")
    up = lambda name, dtype=None: torch.from_numpy(np.array(a[name])).to(device=dev, dtype=dtype)
    rows = up("rows")
    mlib._motion_names = list(meta["motion_names"])
    mlib._motion_files = list(meta["motion_files"])
    mlib._motion_extras = list(meta.get("motion_extras") or [None] * meta["num_clips"])
    mlib._motion_extras += [None] * (meta["num_clips"] - len(mlib._motion_extras))
    mlib._hf_mask_inds = []
    has_masks = meta.get("hf_mask_inds") or []
    for i in range(meta["num_clips"]):
        if i < len(has_masks) and has_masks[i]:
            flat = torch.from_numpy(np.array(a[f"mask{i}.inds"])).to(dev)
            mlib._hf_mask_inds.append(list(torch.split(flat, np.array(a[f"mask{i}.count"]).tolist(), dim=0)))
        else:
            mlib._hf_mask_inds.append(None)
    mlib._terrains = []
    for i, tm in enumerate(meta["terrains"]):
        if tm is None:
            mlib._terrains.append(None)
            continue
        hf = a[f"terrain{i}.hf"]
        t = SubTerrain(tm["name"], hf.shape[0], hf.shape[1], tm["dxdy"][0], tm["dxdy"][1], tm["min_point"][0],
                       tm["min_point"][1], device=dev)
        t.hf = up(f"terrain{i}.hf")
        t.hf_mask = up(f"terrain{i}.hf_mask").to(torch.bool)
        t.hf_maxmin = up(f"terrain{i}.hf_maxmin")
        mlib._terrains.append(t)
    mlib._adopt_rows(rows, lay, up("frames"), up("num_frames"), up("start_idx"), up("fps"), up("loop_modes"),
                     up("weights"), lengths=up("lengths"), delta=up("root_pos_delta"))
    mlib._motion_dt = up("dt")
