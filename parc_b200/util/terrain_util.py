"""SubTerrain + heightfield sampling / point-to-heightfield SDF on libparc_b200 kernels.

Drop-in for the sampling and SDF parts of the reference's `util/terrain_util.py`:
`SubTerrain` (:21-258, the fields and index helpers the query path touches), `get_local_hf_from_terrain`
(:1329-1346), `sample_hf_z_on_terrain` (:2049-2082), `points_hf_sdf` (:1835-1893).  Procedural terrain
generation, mesh conversion and the GUI helpers of that file are out of scope.
"""
from __future__ import annotations

import copy
from typing import Optional

import numpy as np
import torch

from .. import ops
from . import geom_util


class SubTerrain:
    """Axis-aligned heightfield: `hf[ix, iy]` is the height of the cell centred on
    `min_point + (ix, iy) * dxdy`.  Same fields as the reference (:21-39)."""

    def __init__(self, terrain_name="terrain", x_dim=256, y_dim=256, dx=1.0, dy=1.0, min_x=-1.0, min_y=-1.0,
                 device="cuda:0"):
        self.terrain_name = terrain_name
        self.hf = torch.zeros(size=(x_dim, y_dim), dtype=torch.float32, device=device)
        self.dims = torch.tensor([x_dim, y_dim], dtype=torch.int64, device=device)
        self.min_point = torch.tensor([min_x, min_y], dtype=torch.float32, device=device)
        self.dxdy = torch.tensor([dx, dy], dtype=torch.float32, device=device)
        self.hf_mask = torch.zeros(size=(x_dim, y_dim), dtype=torch.bool, device=device)
        self.hf_maxmin = torch.zeros(size=(x_dim, y_dim, 2), dtype=torch.float32, device=device)
        self.hf_maxmin[..., 0] = 1.0
        self.hf_maxmin[..., 1] = -1.0

    _TENSOR_FIELDS = (("hf", torch.float32), ("dims", torch.int64), ("min_point", torch.float32),
                      ("dxdy", torch.float32), ("hf_mask", torch.bool), ("hf_maxmin", torch.float32))

    # -- device / format management (reference :41-106, :211-224) --
    def set_device(self, device):
        for name, _ in self._TENSOR_FIELDS:
            setattr(self, name, getattr(self, name).to(device=device))
        self.__dict__.pop("_desc", None)

    def to_torch(self, device):
        for name, dt in self._TENSOR_FIELDS:
            v = getattr(self, name)
            v = v.to(device=device) if isinstance(v, torch.Tensor) else torch.tensor(v, dtype=dt, device=device)
            setattr(self, name, v)
        self.__dict__.pop("_desc", None)

    def to_numpy(self):
        for name, _ in self._TENSOR_FIELDS:
            setattr(self, name, getattr(self, name).detach().cpu().numpy())
        self.__dict__.pop("_desc", None)

    def numpy_copy(self):
        new = copy.deepcopy(self)
        new.to_numpy()
        return new

    def torch_copy(self):
        self.__dict__.pop("_desc", None)
        new = copy.deepcopy(self)
        for name, _ in self._TENSOR_FIELDS:
            setattr(new, name, getattr(self, name).clone())
        return new

    def update_old(self):
        """Older pickles lack hf_maxmin (reference :211-224)."""
        if hasattr(self, "hf_maxmin"):
            return
        if isinstance(self.hf, torch.Tensor):
            self.hf_maxmin = torch.zeros((self.hf.shape[0], self.hf.shape[1], 2), dtype=torch.float32,
                                         device=self.hf.device)
        else:
            self.hf_maxmin = np.zeros((self.hf.shape[0], self.hf.shape[1], 2), dtype=np.float32)
        self.hf_maxmin[..., 0] = 1.0
        self.hf_maxmin[..., 1] = -1.0

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop("_desc", None)
        return state

    # -- geometry helpers (reference :104-147, :180-190, :248-258) --
    def get_real_size(self):
        return self.dims * self.dxdy

    def get_max_point(self):
        return self.min_point + self.get_real_size() - self.dxdy

    def get_inbounds_grid_index(self, grid_ind: torch.Tensor):
        return torch.clamp(grid_ind, torch.zeros_like(self.dims), self.dims - 1)

    def round_point_to_grid_point(self, point: torch.Tensor):
        return torch.round((point - self.min_point) / self.dxdy) * self.dxdy + self.min_point

    def round_point_to_grid_index(self, point: torch.Tensor):
        return torch.round((point - self.min_point) / self.dxdy).to(dtype=torch.int64)

    def get_origin_index(self):
        return self.get_grid_index(torch.zeros_like(self.min_point))

    def get_point(self, ij):
        return self.min_point + ij * self.dxdy

    def get_xyz_point(self, grid_inds):
        z = self.hf[grid_inds[..., 0], grid_inds[..., 1]]
        return torch.cat([self.get_point(grid_inds), z.unsqueeze(-1)], dim=-1)

    def get_hf_val(self, grid_ind: torch.Tensor):
        g = self.get_inbounds_grid_index(grid_ind)
        return self.hf[g[0], g[1]]

    def set_hf_val(self, grid_ind, val: float):
        g = self.get_inbounds_grid_index(grid_ind)
        self.hf[g[0], g[1]] = val

    def get_grid_node_xy_points(self):
        xs = torch.linspace(0.0, (self.hf.shape[0] - 1.0) * self.dxdy[0].item(), self.hf.shape[0], device=self.hf.device)
        ys = torch.linspace(0.0, (self.hf.shape[1] - 1.0) * self.dxdy[1].item(), self.hf.shape[1], device=self.hf.device)
        gx, gy = torch.meshgrid(xs, ys, indexing="ij")
        return torch.stack([gx + self.min_point[0], gy + self.min_point[1]], dim=-1)

    # -- kernel-backed sampling --
    def hf_desc(self) -> "ops.HeightfieldDesc":
        """Host copies of min_point / dxdy next to the device hf, so launches never sync.  The cache is
        keyed on the tensors' identity and in-place version counters, so edits (pad(), flips, a new
        min_point, in-place height edits -- which matter when a non-contiguous hf is served from a contiguous
        copy) are picked up automatically."""
        key = (self.hf.data_ptr(), tuple(self.hf.shape), self.hf._version, self.hf.is_contiguous(),
               id(self.min_point), self.min_point._version, id(self.dxdy), self.dxdy._version)
        cached = self.__dict__.get("_desc")
        if cached is None or cached[0] != key:
            mp = self.min_point.detach().cpu().tolist()
            dd = self.dxdy.detach().cpu().tolist()
            hf = self.hf if self.hf.is_contiguous() else self.hf.contiguous()
            cached = (key, ops.HeightfieldDesc(hf=hf, min_x=mp[0], min_y=mp[1], dx=dd[0], dy=dd[1]))
            self.__dict__["_desc"] = cached
        return cached[1]

    def invalidate(self):
        self.__dict__.pop("_desc", None)

    def get_grid_index(self, point: torch.Tensor):
        """clamp(round((p - min) / dxdy), 0, dims-1) as int64.  Ref :113-126."""
        _, idx = ops.hf_sample(self.hf_desc(), point, want_index=True)
        return idx

    def get_hf_val_from_points(self, xy_points):
        """Nearest-cell height under each xy point.  Ref :128-130."""
        return ops.hf_sample(self.hf_desc(), xy_points)


def get_local_hf_from_terrain(xy_points, terrain: SubTerrain):
    """Ref util/terrain_util.py:1329-1346."""
    return ops.hf_sample(terrain.hf_desc(), xy_points)


def sample_hf_z_on_terrain(terrain: SubTerrain, center_xy: torch.Tensor, heading: torch.Tensor, dx: float, dy: float,
                           num_x_neg: int, num_x_pos: int, num_y_neg: int, num_y_pos: int):
    """Rotated rectangular grid of nearest-cell heights around each centre -> [B, X, Y].
    Ref util/terrain_util.py:2049-2082."""
    tmpl = geom_util.get_xy_grid_points(center=torch.zeros(2, dtype=torch.float32, device=center_xy.device), dx=dx,
                                        dy=dy, num_x_neg=num_x_neg, num_x_pos=num_x_pos, num_y_neg=num_y_neg,
                                        num_y_pos=num_y_pos)
    if center_xy.dim() == 3:
        assert center_xy.shape[1] == 1
        center_xy = center_xy.squeeze(1)
    if heading.dim() == 2:
        assert heading.shape[1] == 1
        heading = heading.squeeze(1)
    z = ops.hf_obs(terrain.hf_desc(), tmpl.reshape(-1, 2), center_xy, heading, relative=False)
    return z.view(center_xy.shape[0], tmpl.shape[0], tmpl.shape[1])


def points_boxes_sdf(points, box_centers, box_halfdims):
    """[B,N,3] points against [B,M,3] arbitrary axis-aligned boxes -> the full [B,N,M] table of box SDFs
    (util/terrain_util.py:1777-1804).  Kept with the reference's signature for callers that want the table itself;
    it materialises B*N*M values in plain torch.  The heightfield case -- boxes on a grid, minimum over M -- is
    `points_hf_sdf`, which never builds the table."""
    assert points.dim() == box_centers.dim() == box_halfdims.dim()
    if points.dim() == 2:
        points, box_centers, box_halfdims = points.unsqueeze(0), box_centers.unsqueeze(0), box_halfdims.unsqueeze(0)
    assert points.shape[0] == box_centers.shape[0]
    rel = points.unsqueeze(2) - box_centers.unsqueeze(1)
    return geom_util.sdBox(rel, box_halfdims.unsqueeze(1).expand_as(rel))


def points_hf_sdf(points: torch.Tensor, hf: torch.Tensor, hf_min_box_center: torch.Tensor, hf_dxdy: torch.Tensor,
                  base_z=-10.0, inverted=True, radius: Optional[float] = None):
    """Exact signed distance from each point to the union-of-boxes heightfield: min over ALL cells of
    the box SDF.  points [B,N,3], hf [B,X,Y], hf_min_box_center [B,2], hf_dxdy [2] -> [B,N].
    Differentiable with respect to `points` (the MDM's collision loss / guidance back-propagate through it,
    diffusion/mdm.py:729-737, :1484-1496); any batch size, any terrain size.
    Ref util/terrain_util.py:1835-1893."""
    assert points.dim() == 3 and hf.dim() == 3 and hf_min_box_center.dim() == 2 and hf_dxdy.dim() == 1
    terrain = ops.make_terrain_batch(hf, hf_min_box_center, hf_dxdy.detach().cpu().tolist(), base_z=base_z)
    sdf = ops.points_hf_sdf(points, terrain, inverted=inverted)
    if radius is not None:
        # sdRoundBox = sdBox - r (util/geom_util.py:113-120); the min over cells commutes with it
        assert isinstance(radius, float) and radius > 0.0
        sdf = sdf + radius if inverted else sdf - radius
    return sdf


def motion_frames_hf_sdf_loss(motion_frames, char_point_samples, hf, hf_min_box_center, hf_dxdy, char_model,
                              ret_vis_info=False, interior_distance=True):
    """Heightfield-collision loss of a batch of motions: 0.5 * sum over all body points and frames of
    clamp(sdf, max=0)^2 (interior distance; clamp(min=0) otherwise).  motion_frames [B,S,6+D] in the heightfield's
    frame, char_point_samples: list of [P_b,3] per body, hf [B,X,Y], hf_min_box_center [B,2], hf_dxdy [2] -> loss [B]
    (+ world points [B,N,3], sdf [B,N] with ret_vis_info).  Differentiable with respect to motion_frames: every stage
    -- exp-map / DoF conversion, FK, point transform, SDF -- is a CUDA operator with a hand-written VJP.
    Ref util/terrain_util.py:1895-1949."""
    from .. import ops as _ops
    from ..tools.procgen.mdm_path import body_points_desc
    from . import torch_util
    D = char_model.get_dof_size()
    root_pos = motion_frames[..., 0:3]
    root_rot_quat = _ops.exp_map_to_quat(motion_frames[..., 3:6])
    joint_rot = torch_util.quat_pos(char_model.dof_to_rot(motion_frames[..., 6:6 + D]))
    body_pos, body_rot = char_model.forward_kinematics(root_pos, root_rot_quat, joint_rot)
    pts = _ops.body_points_world(body_pos, body_rot, body_points_desc(char_model, char_point_samples))
    sdf = points_hf_sdf(pts, hf, hf_min_box_center, hf_dxdy, base_z=-10.0, inverted=interior_distance)
    if interior_distance:
        loss = 0.5 * torch.sum(torch.square(torch.clamp(sdf, max=0.0)), dim=-1)
    else:
        loss = 0.5 * torch.sum(torch.square(torch.clamp(sdf, min=0.0)), dim=-1)
    if ret_vis_info:
        return loss, pts, sdf
    return loss


# ---- heightfield masks of a motion ---------------------------------------
def _mask_sweep(motion_frames, terrain: SubTerrain, char_model, char_body_points):
    from ..tools.procgen.mdm_path import body_points_desc
    tb = ops.make_terrain_batch(terrain.hf, terrain.min_point, terrain.dxdy.detach().cpu().tolist(), base_z=-10.0)
    keys = ops.make_key_bodies([], [])
    out = ops.clip_label(char_model.c_model(), body_points_desc(char_model, char_body_points), tb, keys,
                         motion_frames.unsqueeze(0), want_contacts=False, want_masks=True)
    X, Y = int(terrain.hf.shape[0]), int(terrain.hf.shape[1])
    return ops.unpack_frame_masks(out["frame_mask_bits"][0], X, Y), out["min_body_heights"][0]


def compute_hf_mask_inds(motion_frames: torch.Tensor, terrain: SubTerrain, char_model, char_body_points):
    """Per frame, the unique grid cells any body surface point falls in, plus the per-cell minimum body
    height.  -> (list of int64 [K_t, 2] tensors sorted like torch.unique(dim=0), min_body_heights [X, Y]).
    Replaces the frame x body x point python loop of util/terrain_util.py:1951-1997 by one launch (FK, point
    transform, cell index, per-frame bit mask, atomic per-cell min) plus one nonzero()."""
    terrain.hf_mask[...] = False                                   # side effect of the reference (:1966)
    masks, min_h = _mask_sweep(motion_frames, terrain, char_model, char_body_points)
    nz = torch.nonzero(masks)                                      # [K, 3] rows (t, ix, iy), lexicographic
    counts = torch.bincount(nz[:, 0], minlength=masks.shape[0]).tolist()
    return list(torch.split(nz[:, 1:], counts)), min_h


def compute_hf_mask_from_inds(terrain: SubTerrain, mask_grid_inds):
    """Ref util/terrain_util.py:1999-2006."""
    hf_mask = torch.zeros_like(terrain.hf_mask)
    if len(mask_grid_inds) > 0:
        allc = torch.cat(list(mask_grid_inds), dim=0)
        hf_mask[allc[..., 0], allc[..., 1]] = True
    return hf_mask


def compute_hf_mask(motion_frames, terrain: SubTerrain, char_model, char_body_points):
    """Ref util/terrain_util.py:2008-2014."""
    inds, _ = compute_hf_mask_inds(motion_frames, terrain, char_model, char_body_points)
    return compute_hf_mask_from_inds(terrain, inds)


def compute_hf_extra_vals(motion_frames, terrain: SubTerrain, char_model, char_body_points, z_buf=3.0, jump_buf=0.8):
    """Fill terrain.hf_mask / terrain.hf_maxmin from a motion.  Ref util/terrain_util.py:2017-2047."""
    mask_grid_inds, min_body_heights = compute_hf_mask_inds(motion_frames, terrain, char_model, char_body_points)
    terrain.hf_mask = compute_hf_mask_from_inds(terrain, mask_grid_inds)
    max_h = torch.max(motion_frames[:, 2]).item()
    min_h = torch.min(terrain.hf).item()
    terrain.hf_maxmin[..., 0] = max_h + z_buf
    terrain.hf_maxmin[..., 1] = min_h - z_buf
    hf_vals = terrain.hf[terrain.hf_mask]
    terrain.hf_maxmin[..., 0][terrain.hf_mask] = hf_vals
    terrain.hf_maxmin[..., 1][terrain.hf_mask] = hf_vals
    jump = torch.logical_and((min_body_heights - terrain.hf) >= jump_buf, terrain.hf_mask)
    terrain.hf_maxmin[..., 0][jump] = min_body_heights[jump] - jump_buf
    terrain.hf_maxmin[..., 1][jump] = min_h - z_buf
    return mask_grid_inds
