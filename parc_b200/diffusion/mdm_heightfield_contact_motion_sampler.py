"""The MDM training sampler's per-sample terrain gather on libparc_b200.

Drop-in for `MDMHeightfieldContactMotionSampler.get_hfs_from_data` / `get_hfs_from_data_helper` of the reference's
`diffusion/mdm_heightfield_contact_motion_sampler.py` (:414-474): every sample looks at ITS clip's own terrain through
a heading-rotated local grid around its root, picks up the height of each grid point and the (max, min) band the
augmenter may move that cell in -- the cell's own band where the character's body covers it during the sample's frame
window, the free band (2 max_h, 2 min_h) elsewhere -- and expresses both relative to the root floor or the root.
The reference does this with a Python loop over the samples; here all clips' terrains live packed on the device
(`ClipTerrainPack`) and one launch serves the batch.  The sampler class around it (motion sampling, noise, the random
box / noise augmentation with the host RNG) belongs to MDM training, which is out of scope.
"""
from __future__ import annotations

import ctypes as C
import enum
from typing import List, Optional, Sequence

import numpy as np
import torch

from .. import _lib
from ..util import geom_util


class RelativeZStyle(enum.Enum):
    RELATIVE_TO_ROOT = 0
    RELATIVE_TO_ROOT_FLOOR = 1


def mask_inds_to_bits(mask_inds: Sequence[torch.Tensor], dim_y: int, words: int) -> np.ndarray:
    """The reference's per-frame cell lists (`compute_hf_mask_inds`, util/terrain_util.py:1951-1997: one [n,2] int tensor
    per frame) -> uint32 [F, words] bit rows (bit ix * dim_y + iy), the layout parc_clip_label writes."""
    out = np.zeros((len(mask_inds), words), dtype=np.uint32)
    for t, inds in enumerate(mask_inds):
        a = np.asarray(inds.detach().cpu() if isinstance(inds, torch.Tensor) else inds).reshape(-1, 2).astype(np.int64)
        cell = a[:, 0] * dim_y + a[:, 1]
        np.bitwise_or.at(out[t], cell >> 5, (np.uint32(1) << (cell & 31).astype(np.uint32)))
    return out


class ClipTerrainPack:
    """Every clip's terrain, band and per-frame body masks concatenated on the device
    (include/parc_b200.h: ParcClipTerrains)."""

    def __init__(self, terrains, hf_mask_bits: Optional[List[Optional[torch.Tensor]]], device):
        recs, hfs, mms, words = [], [], [], []
        cell_off = word_off = 0
        max_w = 0
        for c, t in enumerate(terrains):
            X, Y = int(t.hf.shape[0]), int(t.hf.shape[1])
            W = (X * Y + 31) // 32
            r = _lib.ParcClipTerrain()
            r.cell_offset, r.dim_x, r.dim_y = cell_off, X, Y
            mp, dd = t.min_point.detach().cpu().tolist(), t.dxdy.detach().cpu().tolist()
            r.min_x, r.min_y, r.dx, r.dy = mp[0], mp[1], dd[0], dd[1]
            bits = hf_mask_bits[c] if hf_mask_bits is not None else None
            if bits is not None:
                bits = torch.as_tensor(bits).reshape(-1, W)
                r.mask_offset, r.num_frames = word_off, int(bits.shape[0])
                words.append(bits.to(torch.int32).reshape(-1) if bits.dtype != torch.int32 else bits.reshape(-1))
                word_off += int(bits.numel())
            else:
                r.mask_offset, r.num_frames = -1, 0
            hfs.append(t.hf.detach().to(torch.float32).reshape(-1))
            mms.append(t.hf_maxmin.detach().to(torch.float32).reshape(-1, 2))
            cell_off += X * Y
            max_w = max(max_w, W)
            recs.append(r)
        arr = (_lib.ParcClipTerrain * len(recs))(*recs)
        self.records = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
        self.hf = torch.cat(hfs).to(device).contiguous()
        self.hf_maxmin = torch.cat(mms).to(device).contiguous()
        self.mask_words = torch.cat([w.to(device) for w in words]).contiguous() if words else None
        self.num_clips, self.max_mask_words, self.device = len(recs), max_w, torch.device(device)

    @classmethod
    def from_motion_lib(cls, mlib, device=None):
        """Pack `mlib._terrains` and `mlib._hf_mask_inds` (anim/motion_lib.py:303-321; the masks in the reference's
        list-of-index-tensors form or already as bit rows)."""
        device = device if device is not None else mlib._device
        bits = None
        if getattr(mlib, "_hf_mask_inds", None) is not None:
            bits = []
            for t, inds in zip(mlib._terrains, mlib._hf_mask_inds):
                if inds is None:
                    bits.append(None)
                elif isinstance(inds, torch.Tensor):
                    bits.append(inds)
                else:
                    X, Y = int(t.hf.shape[0]), int(t.hf.shape[1])
                    bits.append(torch.from_numpy(mask_inds_to_bits(inds, Y, (X * Y + 31) // 32).view(np.int32)))
        return cls(mlib._terrains, bits, device)

    def c_struct(self) -> "_lib.ParcClipTerrains":
        s = _lib.ParcClipTerrains()
        s.clips, s.hf, s.hf_maxmin = self.records.data_ptr(), self.hf.data_ptr(), self.hf_maxmin.data_ptr()
        s.mask_words = _lib.ptr(self.mask_words)
        s.num_clips, s.max_mask_words = self.num_clips, self.max_mask_words
        return s


class ClipHeightfieldSampler:
    """The heightfield part of MDMHeightfieldContactMotionSampler (:76-100 the local grid, :449-474 the gather)."""

    def __init__(self, pack: ClipTerrainPack, dx: float, num_x_neg: int, num_x_pos: int, num_y_neg: int, num_y_pos: int,
                 max_h: float, relative_z_style: RelativeZStyle = RelativeZStyle.RELATIVE_TO_ROOT_FLOOR):
        self._pack = pack
        self._device = pack.device
        self._num_x_neg, self._num_y_neg = num_x_neg, num_y_neg
        self._grid_dim_x, self._grid_dim_y = num_x_neg + 1 + num_x_pos, num_y_neg + 1 + num_y_pos
        zero = torch.zeros(2, dtype=torch.float32, device=self._device)
        self._generic_heightmap = geom_util.get_xy_grid_points(zero, dx, dx, num_x_neg, num_x_pos, num_y_neg, num_y_pos)
        self._tmpl = self._generic_heightmap.reshape(-1, 2).contiguous()
        self._max_h, self._min_h = float(max_h), -float(max_h)
        self._relative_z_style = relative_z_style

    def get_hfs_from_data(self, motion_ids, ref_root_pos, ref_root_rot, canon_root_z, motion_time_indices,
                          want_maxmin: bool = True):
        """-> (hfs [B,GX,GY], center_h [B]) as the reference, plus hf_maxmins [B,GX,GY,2] (which the reference hands to
        its augmenter internally) when want_maxmin.  motion_time_indices [B,T]: frame indices of each sample inside
        its clip; the window first..last selects the body-mask frames (:429-430).  Ref :449-474."""
        _lib.require_cuda(motion_ids, ref_root_pos, ref_root_rot, canon_root_z)
        dev = ref_root_pos.device
        ids = motion_ids.to(torch.int64).contiguous()
        B = int(ids.shape[0])
        if B > 0 and (int(ids.min()) < 0 or int(ids.max()) >= self._pack.num_clips):
            raise IndexError("motion id out of range")
        rp, rr = _lib.f32c(ref_root_pos), _lib.f32c(ref_root_rot)
        mti = torch.as_tensor(motion_time_indices, device=dev)
        lo, hi = mti[:, 0].to(torch.int32).contiguous(), mti[:, -1].to(torch.int32).contiguous()
        GX, GY = self._grid_dim_x, self._grid_dim_y
        hfs = torch.empty((B, GX, GY), dtype=torch.float32, device=dev)
        mm = torch.empty((B, GX, GY, 2), dtype=torch.float32, device=dev) if want_maxmin else None
        ch = torch.empty((B,), dtype=torch.float32, device=dev)
        q = _lib.ParcClipHfQuery()
        q.motion_ids, q.root_pos, q.root_rot = ids.data_ptr(), rp.data_ptr(), rr.data_ptr()
        cz = None
        if self._relative_z_style == RelativeZStyle.RELATIVE_TO_ROOT:
            cz = _lib.f32c(canon_root_z).reshape(B)
            q.canon_root_z = cz.data_ptr()
        q.frame_lo, q.frame_hi, q.tmpl_xy, q.n = lo.data_ptr(), hi.data_ptr(), self._tmpl.data_ptr(), B
        q.grid_x, q.grid_y, q.centre_x, q.centre_y = GX, GY, self._num_x_neg, self._num_y_neg
        q.free_max, q.free_min = self._max_h * 2.0, self._min_h * 2.0
        q.hf_out, q.maxmin_out, q.centre_h_out = hfs.data_ptr(), _lib.ptr(mm), ch.data_ptr()
        t = self._pack.c_struct()
        with torch.cuda.device(dev):
            rc = _lib.load().parc_clip_hf_gather(C.byref(t), C.byref(q), _lib.stream_ptr(dev))
        _lib.check(rc, "parc_clip_hf_gather")
        if want_maxmin:
            return hfs, ch, mm
        return hfs, ch
