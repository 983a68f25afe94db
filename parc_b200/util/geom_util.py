"""Static geometry used by the kinematic-query path: body surface sample points and the xy
templates of the heightmap observations.

Drop-in for the corresponding parts of the reference's `util/geom_util.py`: `get_char_point_samples`
(:788-870) with its per-primitive samplers (:725-786), `get_minimal_char_point_samples` (:873-932),
`get_xy_points_cone` (:249-270), `get_xy_grid_points` (:210-221), `sdBox`/`sdSphere` (:122-143, :167-171).
All of this is one-off setup (a few hundred points); it runs with torch ops on the model's device.
The OBB/SAT helpers of that file are out of scope.
"""
from __future__ import annotations

import math

import torch

from . import torch_util


def icosahedron_vertices(radius: float, device, dtype=torch.float32):
    """The 12 vertices the reference obtains from `trimesh.creation.icosphere(subdivisions=0, radius=r)`
    (util/geom_util.py:741-749): a regular icosahedron, vertices normalised to the sphere, in trimesh's
    order.  Emitted here directly so trimesh is not a dependency."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = torch.tensor([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0],
                      [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                      [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=torch.float64)
    v = v / torch.linalg.vector_norm(v, dim=1, keepdim=True) * radius
    return v.to(device=device, dtype=dtype)


def get_sphere_point_surface_samples(radius, device, num_subdivisions=0):
    assert num_subdivisions == 0, "only the base icosahedron is supported"
    return icosahedron_vertices(radius, device)


def get_box_point_surface_samples(box_halfdims, device, num_slices=2, dim_x=6, dim_y=3):
    """z-slice-major grid on the box: slice 0 is the bottom face (util/geom_util.py:725-739)."""
    unit = lambda n: torch.linspace(0.0, 1.0, n, device=device)
    z = unit(num_slices) * box_halfdims[2] * 2.0 - box_halfdims[2]
    x = unit(dim_x) * box_halfdims[0] * 2.0 - box_halfdims[0]
    y = unit(dim_y) * box_halfdims[1] * 2.0 - box_halfdims[1]
    gx, gy = torch.meshgrid(x, y, indexing="ij")
    n_xy = dim_x * dim_y
    pts = torch.stack([gx.reshape(1, n_xy).expand(num_slices, n_xy), gy.reshape(1, n_xy).expand(num_slices, n_xy),
                       z.reshape(num_slices, 1).expand(num_slices, n_xy)], dim=-1)
    return pts.reshape(-1, 3)


def get_capsule_point_surface_samples(capsule_length, capsule_radius, device, num_cylinder_slices=3,
                                      num_circle_points=4, num_sphere_subdivisons=0, ignore_hemispheres=True):
    """Rings on the cylinder wall of a z-aligned capsule, circle-point-major (util/geom_util.py:751-786)."""
    parts = []
    if not ignore_hemispheres:
        sph = get_sphere_point_surface_samples(capsule_radius, device, num_sphere_subdivisons)
        up = sph[sph[..., 2] > 1e-5].clone()
        up[..., 2] += capsule_length / 2.0
        lo = sph[sph[..., 2] < -1e-5].clone()
        lo[..., 2] -= capsule_length / 2.0
        parts += [up, lo]
    z = torch.linspace(0.0, 1.0, num_cylinder_slices, device=device) * capsule_length - capsule_length / 2.0
    theta = torch.linspace(0, 2 * torch.pi, num_circle_points + 1, device=device)[:-1]
    x = capsule_radius * torch.cos(theta)
    y = capsule_radius * torch.sin(theta)
    shape = (num_circle_points, num_cylinder_slices)
    ring = torch.stack([x.reshape(-1, 1).expand(shape), y.reshape(-1, 1).expand(shape),
                        z.reshape(1, -1).expand(shape)], dim=-1)
    parts.append(ring.reshape(-1, 3))
    return torch.cat(parts, dim=0)


def _capsule_frame(geom, device):
    """Rotation taking the z axis onto the capsule segment, and the segment midpoint
    (util/geom_util.py:823-836)."""
    seg = geom._dims
    centre = geom._offset + seg / 2.0
    z_axis = torch.tensor([0.0, 0.0, 1.0], device=device)
    axis = torch.cross(z_axis, seg, dim=-1)
    axis = z_axis if torch.linalg.vector_norm(axis) < 1e-5 else axis / torch.linalg.vector_norm(axis)
    angle = torch.acos(torch.dot(axis, seg))
    return torch_util.axis_angle_to_quat(axis, angle), centre


class _HostGeom:
    """A geom's tensors on the host.  The sample builders below are set-up code (a few hundred tiny ops per
    character): they run on the CPU and their results are moved to the model's device once, instead of issuing
    hundreds of one-element GPU launches (and a cuBLAS dot) per call."""

    def __init__(self, g):
        self._shape_type = g._shape_type
        self._dims = g._dims.detach().cpu() if isinstance(g._dims, torch.Tensor) else g._dims
        self._offset = g._offset.detach().cpu() if isinstance(g._offset, torch.Tensor) else g._offset
        self._radius = getattr(g, "_radius", None)
        if isinstance(self._radius, torch.Tensor):
            self._radius = self._radius.detach().cpu()


def get_char_point_samples(char_model, sphere_num_subdivisions=0, box_num_slices=2, box_dim_x=3, box_dim_y=6,
                           capsule_num_circle_points=4, capsule_num_sphere_subdivisons=0,
                           capsule_num_cylinder_slices=4):
    """Per-body lists of surface sample points in the body frame (util/geom_util.py:788-870)."""
    from ..anim.kin_char_model import GeomType
    device = torch.device("cpu")
    out = []
    for b in range(char_model.get_num_joints()):
        geoms = [_HostGeom(g) for g in char_model.get_geoms(b)]
        pts = []
        for g in geoms:
            if g._shape_type == GeomType.SPHERE:
                pts.append(get_sphere_point_surface_samples(g._dims.item(), device, sphere_num_subdivisions) + g._offset)
            elif g._shape_type == GeomType.BOX:
                pts.append(get_box_point_surface_samples(g._dims, device, num_slices=box_num_slices, dim_x=box_dim_x,
                                                         dim_y=box_dim_y) + g._offset)
            elif g._shape_type == GeomType.CAPSULE:
                rot, centre = _capsule_frame(g, device)
                local = get_capsule_point_surface_samples(
                    torch.linalg.vector_norm(g._dims).item(), g._radius, device,
                    num_cylinder_slices=capsule_num_cylinder_slices, num_circle_points=capsule_num_circle_points,
                    num_sphere_subdivisons=capsule_num_sphere_subdivisons)
                pts.append(torch_util.quat_rotate(rot.unsqueeze(0), local) + centre)
            else:  # cylinder / mesh: a single origin point, as the reference does
                pts.append(torch.zeros((1, 3), dtype=torch.float32, device=device))
        if len(geoms) == 0:
            pts.append(torch.zeros((1, 3), dtype=torch.float32, device=device))
        out.append(torch.cat(pts, dim=0).to(char_model._device))
    return out


def get_minimal_char_point_samples(char_model):
    """Cheaper sample set (sphere centre, 2 capsule points, 8 box corners): util/geom_util.py:873-932."""
    from ..anim.kin_char_model import GeomType
    device = torch.device("cpu")
    out = []
    for b in range(char_model.get_num_joints()):
        pts = []
        for g in (_HostGeom(g_) for g_ in char_model.get_geoms(b)):
            if g._shape_type == GeomType.SPHERE:
                pts.append(g._offset.clone().unsqueeze(0))
            elif g._shape_type == GeomType.CAPSULE:
                rot, centre = _capsule_frame(g, device)
                h = torch.linalg.vector_norm(g._dims).item()
                local = torch.tensor([[0.0, 0.0, h / 3.0], [0.0, 0.0, -h / 3.0]], dtype=torch.float32, device=device)
                pts.append(torch_util.quat_rotate(rot.unsqueeze(0), local) + centre)
            elif g._shape_type == GeomType.BOX:
                pts.append(get_box_point_surface_samples(g._dims, device, num_slices=2, dim_x=2, dim_y=2) + g._offset)
            else:
                assert False
        out.append(torch.cat(pts, dim=0).to(char_model._device))
    return out


def get_xy_grid_points(center, dx, dy, num_x_neg, num_x_pos, num_y_neg, num_y_pos):
    """[X, Y, 2] rectangular template (util/geom_util.py:210-221)."""
    xs = torch.linspace(center[0] - dx * num_x_neg, center[0] + dx * num_x_pos, num_x_neg + num_x_pos + 1,
                        device=center.device)
    ys = torch.linspace(center[1] - dy * num_y_neg, center[1] + dy * num_y_pos, num_y_neg + num_y_pos + 1,
                        device=center.device)
    gx, gy = torch.meshgrid(xs, ys, indexing="ij")
    return torch.stack([gx, gy], dim=-1)


def get_xy_points_cone(center, dx, num_neg, num_pos, num_rays_neg, num_rays_pos, angle_between_rays):
    """Fan of rays, ray-major, [(rays) * (points), 2] (util/geom_util.py:249-270)."""
    device = center.device
    xs = torch.linspace(-dx * num_neg, dx * num_pos, num_neg + num_pos + 1, device=device, dtype=torch.float32)
    ray = torch.stack([xs, torch.zeros_like(xs)], dim=-1)
    fan = []
    for i in range(num_rays_neg + 1 + num_rays_pos):
        ang = torch.ones(ray.shape[0], device=device, dtype=torch.float32) * (-angle_between_rays * (num_rays_neg - i))
        fan.append(torch_util.rotate_2d_vec(ray, ang))
    return torch.cat(fan, dim=0)


def sdBox(point, box_halfdims):
    """Signed distance to an origin-centred box (util/geom_util.py:122-143); host-side helper -- the
    batched point/heightfield path is ops.points_hf_sdf."""
    q = torch.abs(point) - box_halfdims
    outside = torch.linalg.vector_norm(torch.clamp(q, min=0.0), dim=-1)
    inside = torch.clamp(torch.max(q, dim=-1)[0], max=0.0)
    return outside + inside


def sdSphere(p, c, r):
    """util/geom_util.py:167-171"""
    return torch.linalg.vector_norm(p - c, dim=-1) - r
