// Axis-angle / exp-map -> quaternion conversions (and their VJPs) shared by fk.cu and dataset_sweep.cu.
// Reference: util/torch_util.py:311-317 (axis_angle_to_quat), :394-419 (exp_map_to_quat),
// anim/kin_char_model.py:57-77 (Joint.dof_to_rot).
#pragma once

#include "parc_common.cuh"

namespace parc {

// ------------------------------------------------------------------------------------------------
// axis-angle / exp-map -> quaternion, with VJPs
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 normalize4(const float4& u, float& nrm_out) {
  const float n = sqrtf(u.x * u.x + u.y * u.y + u.z * u.z + u.w * u.w);
  const float d = fmaxf(n, 1e-9f);                  // util/torch_util.py:12 clamp(min=eps)
  nrm_out = d;
  const float r = __frcp_rn(d);                     // value path: one reciprocal instead of four divisions
  return make_float4(u.x * r, u.y * r, u.z * r, u.w * r);
}

__device__ __forceinline__ float3 normalize3(const float3& a, float& nrm_out) {
  const float n = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
  const float d = fmaxf(n, 1e-9f);
  nrm_out = d;
  const float r = __frcp_rn(d);
  return make_float3(a.x * r, a.y * r, a.z * r);
}

// util/torch_util.py:311-317
__device__ __forceinline__ float4 axis_angle_to_quat(const float3& axis, float angle) {
  const float th = angle / 2.0f;
  float an, un;
  const float3 n = normalize3(axis, an);
  const float s = sinf(th), c = cosf(th);
  return normalize4(make_float4(n.x * s, n.y * s, n.z * s, c), un);
}

// VJP wrt (axis, angle) of axis_angle_to_quat for upstream g.
__device__ __forceinline__ void axis_angle_to_quat_vjp(const float3& axis, float angle, const float4& g,
                                                       float3& g_axis, float& g_angle) {
  const float th = angle / 2.0f;
  float an, un;
  const float3 n = normalize3(axis, an);
  const float s = sinf(th), c = cosf(th);
  const float4 u = make_float4(n.x * s, n.y * s, n.z * s, c);
  const float4 q = normalize4(u, un);
  // q = u / |u|
  const float qg = q.x * g.x + q.y * g.y + q.z * g.z + q.w * g.w;
  const float4 gu = make_float4((g.x - q.x * qg) / un, (g.y - q.y * qg) / un, (g.z - q.z * qg) / un,
                                (g.w - q.w * qg) / un);
  const float g_th = (gu.x * n.x + gu.y * n.y + gu.z * n.z) * c - gu.w * s;
  g_angle = 0.5f * g_th;
  // n = axis / |axis|
  const float3 gn = make_float3(gu.x * s, gu.y * s, gu.z * s);
  const float ngn = n.x * gn.x + n.y * gn.y + n.z * gn.z;
  g_axis = make_float3((gn.x - n.x * ngn) / an, (gn.y - n.y * ngn) / an, (gn.z - n.z * ngn) / an);
}

// normalize_angle(a) = atan2(sin a, cos a) for a >= 0 (util/torch_util.py:4-7): the identity on [0, pi] up to
// an ulp, so the three transcendental calls are only paid for angles beyond pi.
__device__ __forceinline__ float wrap_angle_nonneg(float a) {
  return a <= 3.14159265f ? a : atan2f(sinf(a), cosf(a));
}

// util/torch_util.py:394-412: exp-map -> (axis, angle) with the reference's wrap and small-angle mask
__device__ __forceinline__ void exp_map_to_axis_angle(const float3& e, float3& axis, float& ang) {
  const float a = sqrtf(e.x * e.x + e.y * e.y + e.z * e.z);
  const float r = 1.0f / a;                          // inf for a == 0: masked below, like the reference's NaN axis
  axis = make_float3(e.x * r, e.y * r, e.z * r);
  ang = wrap_angle_nonneg(a);
  if (!(fabsf(ang) > 1e-5f)) {
    ang = 0.0f;
    axis = make_float3(0.0f, 0.0f, 1.0f);
  }
}

// util/torch_util.py:414-419
__device__ __forceinline__ float4 exp_map_to_quat(const float3& e) {
  float3 axis;
  float ang;
  exp_map_to_axis_angle(e, axis, ang);
  return axis_angle_to_quat(axis, ang);
}

// VJP of exp_map_to_quat.  At e == 0 the reference's autograd yields NaN (0/0 behind torch.where,
// SURVEY F8d); here the masked branch returns a zero gradient instead (documented divergence).
__device__ __forceinline__ float3 exp_map_to_quat_vjp(const float3& e, const float4& g) {
  const float a = sqrtf(e.x * e.x + e.y * e.y + e.z * e.z);
  const float ang = wrap_angle_nonneg(a);
  if (!(fabsf(ang) > 1e-5f)) return make_float3(0.0f, 0.0f, 0.0f);
  const float3 axis = make_float3(e.x / a, e.y / a, e.z / a);
  float3 g_axis;
  float g_ang;
  axis_angle_to_quat_vjp(axis, ang, g, g_axis, g_ang);
  // ang = atan2(sin a, cos a): d ang / d a = (cos^2 + sin^2) / (sin^2 + cos^2) = 1
  float g_a = g_ang;
  // axis = e / a
  g_a -= (g_axis.x * e.x + g_axis.y * e.y + g_axis.z * e.z) / (a * a);
  return make_float3(g_axis.x / a + g_a * e.x / a, g_axis.y / a + g_a * e.y / a, g_axis.z / a + g_a * e.z / a);
}


// ---- small helpers shared by the step-assembly kernels (tracker_step.cu) and the query kernel's fused target
// observation (motion_query.cu) ----
__device__ __forceinline__ float3 ld3(const float* __restrict__ p) {
  return make_float3(__ldg(p), __ldg(p + 1), __ldg(p + 2));
}
__device__ __forceinline__ float4 ld4(const float* __restrict__ p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void st3(float* __restrict__ p, const float3& v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

// util/torch_util.py:491-499: rotation about +z by -heading(q)
__device__ __forceinline__ float4 heading_inverse_quat(const float4& q) {
  return axis_angle_to_quat(make_float3(0.0f, 0.0f, 1.0f), -calc_heading(q));
}

// util/torch_util.py:361-373: rotated x axis, then rotated z axis.  quat_rotate (v + w t + q_v x t, t = 2 q_v x v)
// written out for v = e_x and v = e_z: the terms the general form multiplies by an exact 0 are dropped -- they
// contribute exact zeros for finite q -- which is a third of the instructions.
__device__ __forceinline__ void store_tan_norm(float* __restrict__ o, const float4& q) {
  // v = e_x: q_v x v = (0, z, -y), t = (0, 2z, -2y), q_v x t = (y ty... ) = (-2y^2 - 2z^2, 2xy, 2xz)
  {
    const float ty = 2.0f * q.z, tz = -2.0f * q.y;
    st3(o, make_float3(1.0f + (q.y * tz - q.z * ty), q.w * ty + (-q.x * tz), q.w * tz + (q.x * ty)));
  }
  // v = e_z: q_v x v = (y, -x, 0), t = (2y, -2x, 0), q_v x t = (-z ty.. ) = (2xz, 2yz, -2x^2 - 2y^2)
  {
    const float tx = 2.0f * q.y, ty = -2.0f * q.x;
    st3(o + 3, make_float3(q.w * tx + (-q.z * ty), q.w * ty + (q.z * tx), 1.0f + (q.x * ty - q.y * tx)));
  }
}


// Joint j's quaternion from the pose's DoF vector (anim/kin_char_model.py:57-77).  Hinge and spherical joints
// (and the root's exp-map) share ONE axis_angle_to_quat call site so that lanes holding different joint types
// do not serialise two copies of the transcendental chain.
__device__ __forceinline__ float4 joint_dof_to_quat(int joint_type, const float* __restrict__ d, const float* axis) {
  float3 ax = make_float3(0.0f, 0.0f, 1.0f);
  float ang = 0.0f;
  const bool rotates = joint_type == PARC_JOINT_HINGE || joint_type == PARC_JOINT_SPHERICAL;
  if (joint_type == PARC_JOINT_HINGE) {
    ax = make_float3(axis[0], axis[1], axis[2]);
    ang = d[0];
  } else if (joint_type == PARC_JOINT_SPHERICAL) {
    exp_map_to_axis_angle(make_float3(d[0], d[1], d[2]), ax, ang);
  }
  const float4 q = axis_angle_to_quat(ax, ang);
  return rotates ? q : make_float4(0.f, 0.f, 0.f, 1.f);
}

// Lane b (0..J-1) of a warp turns its slice of one raw frame into (pos, rot) inputs of fk_warp:
// lane 0 -> root position + exp_map_to_quat(root exp-map); lane j >= 1 -> joint j's quaternion.
__device__ __forceinline__ void frame_to_lane_pose(const ParcCharModel& m, const float* __restrict__ fr, int lane,
                                                   float3& pos, float4& rot) {
  pos = make_float3(0.f, 0.f, 0.f);
  // every lane funnels into the same joint_dof_to_quat call: the root's exp-map is a "spherical joint"
  int jt = PARC_JOINT_FIXED;
  float dd[3] = {0.f, 0.f, 0.f};
  const float* axis = m.joint_axis[0];
  if (lane == 0) {
    pos = make_float3(__ldg(fr), __ldg(fr + 1), __ldg(fr + 2));
    jt = PARC_JOINT_SPHERICAL;
    dd[0] = __ldg(fr + 3); dd[1] = __ldg(fr + 4); dd[2] = __ldg(fr + 5);
  } else if (lane < m.num_bodies) {
    jt = m.joint_type[lane];
    axis = m.joint_axis[lane];
    const float* d = fr + 6 + m.dof_idx[lane];
    if (jt == PARC_JOINT_HINGE) dd[0] = __ldg(d);
    else if (jt == PARC_JOINT_SPHERICAL) { dd[0] = __ldg(d); dd[1] = __ldg(d + 1); dd[2] = __ldg(d + 2); }
  }
  rot = joint_dof_to_quat(jt, dd, axis);
}

}  // namespace parc
