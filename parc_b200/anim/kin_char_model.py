"""KinCharModel: MJCF character -> kinematic tree, with FK / DoF conversion on libparc_b200 kernels.

Drop-in for the reference's `anim/kin_char_model.py` on the kinematic-query path: same class / method
names, argument meaning and shapes (`forward_kinematics` :509-541, `dof_to_rot` :478-491, `rot_to_dof`
:493-507, `load_char_file` :206-449, `compute_frame_dof_vel` :543-581).  `forward_kinematics` and
`dof_to_rot` are autograd-aware CUDA operators (parc_b200/ops.py); the rest is host-side setup.
"""
from __future__ import annotations

import copy
import enum
import xml.etree.ElementTree as ET
from typing import List

import numpy as np
import torch

from .. import ops
from ..util import torch_util


class JointType(enum.Enum):
    ROOT = 0
    HINGE = 1
    SPHERICAL = 2
    FIXED = 3


_DOF_DIM = {JointType.ROOT: 0, JointType.HINGE: 1, JointType.SPHERICAL: 3, JointType.FIXED: 0}


class Joint:
    """One joint of the tree (reference: anim/kin_char_model.py:17-100)."""

    def __init__(self, name, joint_type: JointType, axis, limits=None):
        self.name = name
        self.joint_type = joint_type
        self.axis = axis
        self.dof_idx = -1
        self.limits = limits

    def get_copy(self, new_device):
        mv = lambda t: None if t is None else t.clone().to(device=new_device)
        return Joint(self.name, self.joint_type, mv(self.axis), mv(self.limits))

    def get_dof_dim(self) -> int:
        return _DOF_DIM[self.joint_type]

    def get_joint_dof(self, dof):
        return dof[..., self.dof_idx:self.dof_idx + self.get_dof_dim()]

    def set_joint_dof(self, j_dof, out_dof):
        out_dof[..., self.dof_idx:self.dof_idx + self.get_dof_dim()] = j_dof

    def dof_to_rot(self, dof):
        """Single-joint conversion with host torch ops (the batched path is KinCharModel.dof_to_rot)."""
        if self.joint_type == JointType.HINGE:
            axis = torch.broadcast_to(self.axis, list(dof.shape[:-1]) + [3])
            return torch_util.axis_angle_to_quat(axis, dof.squeeze(-1))
        if self.joint_type == JointType.SPHERICAL:
            return torch_util.exp_map_to_quat(dof)
        rot = torch.zeros(list(dof.shape[:-1]) + [4], device=dof.device, dtype=dof.dtype)
        rot[..., -1] = 1
        return rot

    def rot_to_dof(self, rot):
        if self.joint_type == JointType.HINGE:
            axis, angle = torch_util.quat_to_axis_angle(rot)
            flip = torch.sum(self.axis * axis, dim=-1) < 0
            return torch.where(flip, -angle, angle).unsqueeze(-1)
        if self.joint_type == JointType.SPHERICAL:
            return torch_util.quat_to_exp_map(rot)
        return torch.zeros(list(rot.shape[:-1]) + [0], device=rot.device, dtype=rot.dtype)


class GeomType(enum.Enum):
    BOX = 0
    SPHERE = 1
    CAPSULE = 2
    CYLINDER = 3
    MESH = 4


_GEOM_NAMES = {"box": GeomType.BOX, "sphere": GeomType.SPHERE, "capsule": GeomType.CAPSULE,
               "cylinder": GeomType.CYLINDER, "mesh": GeomType.MESH}


class Geom:
    """Collision primitive attached to a body (reference: anim/kin_char_model.py:109-140).
    `_offset`: centre (sphere/box) or segment start (capsule); `_dims`: radius / half extents /
    segment vector; `_radius`: capsule radius."""

    def __init__(self, shape_type, offset, dims, device, quat=None, radius=None, mesh_name=None):
        as_t = lambda v: v if isinstance(v, torch.Tensor) else torch.tensor(v, dtype=torch.float32, device=device)
        self._shape_type = shape_type
        self._offset = as_t(offset)
        self._dims = as_t(dims)
        self._radius = radius
        self._mesh_name = mesh_name
        self._quat = quat

    def get_copy(self, new_device):
        return Geom(self._shape_type, self._offset.clone().to(new_device), self._dims.clone().to(new_device),
                    new_device, quat=getattr(self, "_quat", None), radius=self._radius)


def _floats(text, default=None):
    if text is None:
        return default
    return np.array([float(tok) for tok in text.split()], dtype=float)


def _wxyz_to_xyzw(q):
    return np.array([q[1], q[2], q[3], q[0]], dtype=float)


class KinCharModel:
    def __init__(self, device):
        self._device = torch.device(device)
        self._c_model = None
        self._dof_index_cache = None

    # ------------------------------------------------------------------ construction
    def init(self, body_names, parent_indices, local_translation, local_rotation, joints, geoms=None):
        n = len(body_names)
        assert len(parent_indices) == n and len(local_translation) == n
        assert len(local_rotation) == n and len(joints) == n
        dev = self._device
        self._body_names = body_names
        to_t = lambda v, dt: v if isinstance(v, torch.Tensor) else torch.tensor(np.array(v), device=dev, dtype=dt)
        self._parent_indices = to_t(parent_indices, torch.long)
        self._local_translation = to_t(local_translation, torch.float32)
        self._original_local_translation = self._local_translation.clone()
        self._local_rotation = to_t(local_rotation, torch.float32)
        self._joints = joints
        self._dof_size = self._label_dof_indices(joints)
        self._name_body_map = {name: i for i, name in enumerate(body_names)}
        self._lower_dof_limits, self._upper_dof_limits = self._gather_joint_limits(joints)
        self._geoms = geoms
        self._c_model = None
        self._dof_index_cache = None

    def apply_scales_to_local_translation(self, x_scale, y_scale, z_scale):
        scale = torch.tensor([x_scale, y_scale, z_scale], dtype=torch.float32, device=self._local_translation.device)
        self._local_translation = self._original_local_translation * scale
        self._c_model = None

    def get_copy(self, new_device):
        other = KinCharModel(new_device)
        geoms = None if self._geoms is None else [[g.get_copy(new_device) for g in gs] for gs in self._geoms]
        other.init(copy.deepcopy(self._body_names), self._parent_indices.clone().to(new_device),
                   self._local_translation.clone().to(new_device), self._local_rotation.clone().to(new_device),
                   [j.get_copy(new_device) for j in self._joints], geoms)
        return other

    def load_char_file(self, char_file):
        """Parse an MJCF file: bodies in depth-first order, three stacked hinges fused into one
        spherical joint, one hinge kept, no joint = fixed (reference :206-449, :608-701)."""
        root = ET.parse(char_file).getroot()
        world = root.find("worldbody")
        assert world is not None
        first = world.find("body")
        assert first is not None
        default_joint_type = self._default_class_attr(root, "joint")
        default_geom = _GEOM_NAMES.get(self._first_default_geom_type(root), GeomType.SPHERE)
        self._default_geom_type = default_geom
        self._meshes = {}

        names, parents, trans, rots, joints, geoms = [], [], [], [], [], []
        stack = [(first, -1)]
        while stack:  # explicit DFS, children pushed in reverse to keep document order
            node, parent = stack.pop()
            index = len(names)
            names.append(node.attrib.get("name"))
            parents.append(parent)
            trans.append(_floats(node.attrib.get("pos"), np.zeros(3)))
            quat = _floats(node.attrib.get("quat"))
            rots.append(np.array([0.0, 0.0, 0.0, 1.0]) if quat is None else _wxyz_to_xyzw(quat))
            joints.append(self._build_root_joint() if index == 0
                          else self._parse_joint(names[-1], node.findall("joint"), default_joint_type))
            geoms.append([self._parse_geom(g, default_geom) for g in node.findall("geom")])
            for child in reversed(node.findall("body")):
                stack.append((child, index))
        self.init(names, parents, trans, rots, joints, geoms)

    @staticmethod
    def _default_class_attr(root, tag):
        d = root.find("default")
        if d is None:
            return None
        for sub in d.findall("default"):
            if sub.attrib.get("class") == "body":
                el = sub.find(tag)
                if el is not None:
                    return el.attrib.get("type")
        return None

    @staticmethod
    def _first_default_geom_type(root):
        d = root.find("default")
        if d is None:
            return None
        g = d.find("geom")
        if g is None:
            sub = d.find("default")
            g = None if sub is None else sub.find("geom")
        return None if g is None else g.attrib.get("type")

    def _parse_geom(self, node, default_geom):
        kind = _GEOM_NAMES.get(node.attrib.get("type"), default_geom)
        quat = _wxyz_to_xyzw(_floats(node.attrib.get("quat"), np.array([1.0, 0.0, 0.0, 0.0])))
        dev = self._device
        if kind in (GeomType.SPHERE, GeomType.BOX):
            offset = _floats(node.attrib.get("pos"), np.zeros(3))
            size = node.attrib.get("size")
            dims = np.array(0.1, dtype=float) if size is None else _floats(size)
            return Geom(kind, offset, dims, dev, quat=quat)
        if kind == GeomType.CAPSULE:
            ft = _floats(node.attrib.get("fromto"))
            return Geom(kind, ft[0:3], ft[3:6] - ft[0:3], dev, quat=quat, radius=float(node.attrib.get("size")))
        if kind == GeomType.CYLINDER:
            return Geom(kind, _floats(node.attrib.get("pos"), np.zeros(3)), _floats(node.attrib.get("size")), dev)
        return Geom(kind, _floats(node.attrib.get("pos"), np.zeros(3)), np.ones(3), dev, radius=1.0,
                    mesh_name=node.attrib.get("mesh"), quat=quat)

    def _build_root_joint(self):
        return Joint("root", JointType.ROOT, None)

    def _parse_joint(self, body_name, nodes, default_type):
        dev = self._device

        def limits_of(nd):
            rng = nd.attrib.get("range")
            assert rng is not None, "Need joint limits"
            return _floats(rng)

        def check_no_offset(nd):
            pos = _floats(nd.attrib.get("pos"))
            assert pos is None or not np.any(pos), "Joint offsets are not supported"

        if len(nodes) == 0:
            return Joint(body_name, JointType.FIXED, None)
        for nd in nodes:
            assert (nd.attrib.get("type") or default_type) == "hinge", "Unsupported joint type"
            check_no_offset(nd)
        if len(nodes) == 1:
            nd = nodes[0]
            lim = torch.from_numpy(limits_of(nd)).to(dtype=torch.float32, device=dev)
            lim *= torch.pi / 180.0
            axis = torch.tensor(_floats(nd.attrib.get("axis")), device=dev, dtype=torch.float32)
            return Joint(nd.attrib.get("name"), JointType.HINGE, axis, lim)
        assert len(nodes) == 3, "Series joints are not supported."
        lim = torch.from_numpy(np.stack([limits_of(nd) for nd in nodes])).to(dtype=torch.float32, device=dev)
        lim *= torch.pi / 180.0
        name = nodes[0].attrib.get("name")
        return Joint(name[:name.rfind("_")], JointType.SPHERICAL, None, lim)

    def _gather_joint_limits(self, joints):
        lo, hi = [], []
        for j in joints:
            if j.limits is None:
                continue
            if j.joint_type == JointType.HINGE:
                lo.append(j.limits[0:1])
                hi.append(j.limits[1:2])
            elif j.joint_type == JointType.SPHERICAL:
                lo.append(j.limits[:, 0])
                hi.append(j.limits[:, 1])
        return (torch.cat(lo) if lo else lo), (torch.cat(hi) if hi else hi)

    def _label_dof_indices(self, joints):
        idx = 0
        for j in joints:
            if j is not None:
                j.dof_idx = idx
                idx += j.get_dof_dim()
        return idx

    # ------------------------------------------------------------------ accessors
    def get_body_names(self):
        return self._body_names

    def get_joint(self, j) -> Joint:
        assert j > 0
        return self._joints[j]

    def get_parent_id(self, j):
        return self._parent_indices[j]

    def get_dof_size(self):
        return self._dof_size

    def get_joint_dof_idx(self, j):
        return self.get_joint(j).dof_idx

    def get_joint_dof_dim(self, j):
        return self.get_joint(j).get_dof_dim()

    def get_num_joints(self):
        return len(self._joints)

    def get_num_non_root_joints(self):
        return len(self._joints) - 1

    def get_body_name(self, body_id):
        return self._body_names[body_id]

    def get_body_id(self, body_name):
        assert body_name in self._name_body_map
        return self._name_body_map[body_name]

    def get_joint_id(self, body_name):
        return self.get_body_id(body_name) - 1

    def get_geoms(self, body_id) -> List[Geom]:
        return self._geoms[body_id]

    # ------------------------------------------------------------------ kernel-backed operators
    def c_model(self) -> "ops.ParcCharModel":
        """The POD handed to the kernels (built once from host copies of the tree)."""
        if self._c_model is None:
            J = self.get_num_joints()
            axes = [[0.0, 0.0, 0.0] if j.axis is None else j.axis.detach().cpu().tolist() for j in self._joints]
            self._c_model = ops.make_char_model(
                self._parent_indices.cpu().tolist(), self._local_translation.detach().cpu().tolist(),
                self._local_rotation.detach().cpu().tolist(), [j.joint_type.value for j in self._joints], axes,
                [j.dof_idx for j in self._joints])
            assert self._c_model.num_bodies == J
        return self._c_model

    def forward_kinematics(self, root_pos, root_rot, joint_rot):
        """root_pos [...,3], root_rot [...,4], joint_rot [...,J-1,4] -> body_pos [...,J,3], body_rot
        [...,J,4].  One warp per character on the GPU; differentiable.  Ref :509-541."""
        return ops.forward_kinematics(self.c_model(), root_pos, root_rot, joint_rot)

    def frames_forward_kinematics(self, motion_frames):
        """[..., 6+D] raw frames -> (body_pos, body_rot): exp_map_to_quat + dof_to_rot + FK in one launch
        (forward only).  The front end of MotionLib.get_frames_for_id / the labelling sweeps."""
        return ops.frames_fk(self.c_model(), motion_frames)

    def dof_to_rot(self, dof):
        """dof [...,D] -> joint_rot [...,J-1,4].  CUDA kernel, differentiable.  Ref :478-491."""
        return ops.dof_to_rot(self.c_model(), dof)

    def _dof_index(self, device):
        if self._dof_index_cache is None or self._dof_index_cache[0].device != torch.device(device):
            hinge_j, hinge_d, sph_j, sph_d, axes = [], [], [], [], []
            for j in range(1, self.get_num_joints()):
                jt = self._joints[j]
                if jt.joint_type == JointType.HINGE:
                    hinge_j.append(j - 1)
                    hinge_d.append(jt.dof_idx)
                    axes.append(jt.axis.detach().cpu())
                elif jt.joint_type == JointType.SPHERICAL:
                    sph_j.append(j - 1)
                    sph_d.append(jt.dof_idx)
            mk = lambda v: torch.tensor(v, dtype=torch.long, device=device)
            ax = torch.stack(axes).to(device) if axes else torch.zeros((0, 3), device=device)
            sph_cols = mk([d + k for d in sph_d for k in range(3)])
            self._dof_index_cache = (mk(hinge_j), mk(hinge_d), mk(sph_j), sph_cols, ax)
        return self._dof_index_cache

    def rot_to_dof(self, rot):
        """joint_rot [...,J-1,4] -> dof [...,D] (ref :493-507).  CUDA tensors without autograd go through
        `parc_rot_to_dof`; when gradients are needed (or at load time on the host) all joints are converted in
        one pass of torch ops instead of the reference's per-joint python loop."""
        if rot.is_cuda and not (rot.requires_grad and torch.is_grad_enabled()):
            return ops.rot_to_dof(self.c_model(), rot)              # one kernel launch (csrc/fk.cu)
        # differentiable / host-side use: the same arithmetic as torch ops on the tensor's device
        hj, hd, sj, scols, axes = self._dof_index(rot.device)
        dof = torch.zeros(list(rot.shape[:-2]) + [self._dof_size], device=rot.device, dtype=rot.dtype)
        axis, angle = torch_util.quat_to_axis_angle(rot)
        if hj.numel() > 0:
            a = angle[..., hj]
            flip = torch.sum(axes * axis[..., hj, :], dim=-1) < 0
            dof[..., hd] = torch.where(flip, -a, a)
        if sj.numel() > 0:
            em = angle[..., sj].unsqueeze(-1) * axis[..., sj, :]
            dof[..., scols] = em.reshape(*em.shape[:-2], -1)
        return dof

    # ------------------------------------------------------------------ load-time helpers (host torch)
    def compute_frame_dof_vel(self, joint_rot, dt):
        """Finite-difference DoF velocities, last frame repeated.  Ref :543-550."""
        v = self.compute_dof_vel(joint_rot[..., :-1, :, :], joint_rot[..., 1:, :, :], dt)
        return torch.cat([v, v[..., -1:, :]], dim=-2)

    def compute_dof_vel(self, joint_rot0, joint_rot1, dt):
        """Ref :552-581."""
        out = torch.zeros(list(joint_rot0.shape[:-2]) + [self._dof_size], device=joint_rot0.device,
                          dtype=joint_rot0.dtype)
        drot = torch_util.quat_normalize(torch_util.quat_mul(torch_util.quat_conjugate(joint_rot0), joint_rot1))
        for j in range(1, self.get_num_joints()):
            jt = self._joints[j]
            if jt.joint_type not in (JointType.HINGE, JointType.SPHERICAL):
                continue
            v = torch_util.quat_to_exp_map(drot[..., j - 1, :]) / dt
            if jt.joint_type == JointType.HINGE:
                v = torch.sum(jt.axis * v, dim=-1, keepdim=True)
            jt.set_joint_dof(v, out)
        return out

    def host_dof_to_rot(self, dof):
        """dof_to_rot with host torch ops on whatever device `dof` lives on.  Used only while BUILDING
        the frame tables (load time); queries go through the CUDA operator `dof_to_rot`."""
        J = self.get_num_joints()
        out = torch.zeros(list(dof.shape[:-1]) + [J - 1, 4], device=dof.device, dtype=dof.dtype)
        for j in range(1, J):
            jt = self._joints[j]
            out[..., j - 1, :] = jt.dof_to_rot(jt.get_joint_dof(dof))
        return out

    def extract_frame_data(self, motion_frames):
        """[...,6+D] -> root_pos, root_rot (quat), joint_rot.  Ref :937-943."""
        root_pos = motion_frames[..., 0:3]
        if motion_frames.is_cuda:
            root_rot = ops.exp_map_to_quat(motion_frames[..., 3:6])
            joint_rot = self.dof_to_rot(motion_frames[..., 6:6 + self._dof_size])
        else:
            root_rot = torch_util.exp_map_to_quat(motion_frames[..., 3:6])
            joint_rot = self.host_dof_to_rot(motion_frames[..., 6:6 + self._dof_size])
        return root_pos, root_rot, joint_rot

    def construct_frame_data(self, root_pos, root_rot_quat, joint_rot):
        """Ref :945-949."""
        return torch.cat([root_pos, torch_util.quat_to_exp_map(root_rot_quat), self.rot_to_dof(joint_rot)], dim=-1)

    def apply_joint_dof_limits(self, joint_dofs: torch.Tensor):
        """Ref :951-961."""
        assert joint_dofs.dim() in (1, 2), "unsupported"
        return torch.clamp(joint_dofs, min=self._lower_dof_limits, max=self._upper_dof_limits)
