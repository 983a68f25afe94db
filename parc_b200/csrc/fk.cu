// Forward kinematics (forward + VJP) and DoF -> quaternion conversion (forward + VJP), sm_100a.
//
// FK: one warp per character, lane b owns body b; the chain is resolved level by level with warp
// shuffles (fk_warp in parc_common.cuh).  The VJP recomputes the forward pass in registers (cheaper
// than re-reading body_rot from HBM) and then walks the bodies in reverse index order, each body
// pushing its contribution into its parent's lane by shuffle -- no atomics, deterministic.
//
// Reference: anim/kin_char_model.py:509-541 (FK), :478-491 + :57-77 (dof_to_rot),
// util/torch_util.py:311-317, :394-419 (axis-angle / exp-map -> quaternion).
#include "parc_common.cuh"
#include "parc_rotations.cuh"

namespace parc {

__global__ void __launch_bounds__(256) exp_map_fwd_kernel(const float* __restrict__ e, int64_t n,
                                                          float* __restrict__ q) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float3 v = make_float3(e[i * 3], e[i * 3 + 1], e[i * 3 + 2]);
    reinterpret_cast<float4*>(q)[i] = exp_map_to_quat(v);
  }
}

__global__ void __launch_bounds__(256) exp_map_bwd_kernel(const float* __restrict__ e, const float* __restrict__ g,
                                                          int64_t n, float* __restrict__ ge) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float3 v = make_float3(e[i * 3], e[i * 3 + 1], e[i * 3 + 2]);
    const float4 gq = reinterpret_cast<const float4*>(g)[i];
    const float3 r = exp_map_to_quat_vjp(v, gq);
    ge[i * 3] = r.x; ge[i * 3 + 1] = r.y; ge[i * 3 + 2] = r.z;
  }
}

// One thread per (pose, joint).  dof [N,D] -> joint_rot [N,J-1,4].
__global__ void __launch_bounds__(256)
dof_to_rot_fwd_kernel(const float* __restrict__ dof, int64_t n, const __grid_constant__ ParcCharModel m,
                      float* __restrict__ jr) {
  const int Jm1 = m.num_bodies - 1;
  const int64_t total = n * Jm1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = i / Jm1;
    const int j = (int)(i - f * Jm1) + 1;
    const float* d = dof + f * m.dof_size + m.dof_idx[j];
    float4 q = make_float4(0.f, 0.f, 0.f, 1.f);
    const int jt = m.joint_type[j];
    if (jt == PARC_JOINT_HINGE) {
      q = axis_angle_to_quat(make_float3(m.joint_axis[j][0], m.joint_axis[j][1], m.joint_axis[j][2]), d[0]);
    } else if (jt == PARC_JOINT_SPHERICAL) {
      q = exp_map_to_quat(make_float3(d[0], d[1], d[2]));
    }
    reinterpret_cast<float4*>(jr)[i] = q;
  }
}

__global__ void __launch_bounds__(256)
dof_to_rot_bwd_kernel(const float* __restrict__ dof, const float* __restrict__ gjr, int64_t n,
                      const __grid_constant__ ParcCharModel m, float* __restrict__ gdof) {
  const int Jm1 = m.num_bodies - 1;
  const int64_t total = n * Jm1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = i / Jm1;
    const int j = (int)(i - f * Jm1) + 1;
    const int jt = m.joint_type[j];
    if (jt != PARC_JOINT_HINGE && jt != PARC_JOINT_SPHERICAL) continue;
    const float* d = dof + f * m.dof_size + m.dof_idx[j];
    float* o = gdof + f * m.dof_size + m.dof_idx[j];
    const float4 g = reinterpret_cast<const float4*>(gjr)[i];
    if (jt == PARC_JOINT_HINGE) {
      float3 ga;
      float gang;
      axis_angle_to_quat_vjp(make_float3(m.joint_axis[j][0], m.joint_axis[j][1], m.joint_axis[j][2]), d[0], g, ga,
                             gang);
      o[0] = gang;
    } else {
      const float3 r = exp_map_to_quat_vjp(make_float3(d[0], d[1], d[2]), g);
      o[0] = r.x; o[1] = r.y; o[2] = r.z;
    }
  }
}

// joint_rot [N,J-1,4] -> dof [N,D]: KinCharModel.rot_to_dof (anim/kin_char_model.py:493-507) with
// Joint.rot_to_dof (:79-100) and quat_to_axis_angle / quat_to_exp_map (util/torch_util.py:68-88, :346-351).
// One thread per (pose, joint); forward only.
__global__ void __launch_bounds__(256)
rot_to_dof_kernel(const float* __restrict__ jr, int64_t n, const __grid_constant__ ParcCharModel m,
                  float* __restrict__ dof) {
  const int Jm1 = m.num_bodies - 1;
  const int64_t total = n * Jm1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = i / Jm1;
    const int j = (int)(i - f * Jm1) + 1;
    const int jt = m.joint_type[j];
    if (jt != PARC_JOINT_HINGE && jt != PARC_JOINT_SPHERICAL) continue;
    float4 q = reinterpret_cast<const float4*>(jr)[i];
    if (q.w < 0.0f) { q.x = -q.x; q.y = -q.y; q.z = -q.z; q.w = -q.w; }        // quat_pos
    const float len = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z);
    float angle = 2.0f * atan2f(len, q.w);
    float3 axis = make_float3(q.x / len, q.y / len, q.z / len);
    if (!(len > 1e-5f)) { angle = 0.0f; axis = make_float3(0.0f, 0.0f, 1.0f); }
    float* o = dof + f * m.dof_size + m.dof_idx[j];
    if (jt == PARC_JOINT_HINGE) {
      const float d = m.joint_axis[j][0] * axis.x + m.joint_axis[j][1] * axis.y + m.joint_axis[j][2] * axis.z;
      o[0] = d < 0.0f ? -angle : angle;
    } else {
      o[0] = angle * axis.x; o[1] = angle * axis.y; o[2] = angle * axis.z;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// FK forward / VJP, one warp per character, lane b = body b
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PARC_CTA_THREADS)
fk_fwd_kernel(const float* __restrict__ root_pos, const float* __restrict__ root_rot,
              const float* __restrict__ joint_rot, int64_t n, const __grid_constant__ ParcCharModel model_param,
              float* __restrict__ body_pos, float* __restrict__ body_rot) {
  __shared__ ParcCharModel sm;
  stage_model(&sm, model_param);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int J = sm.num_bodies;
  const LaneBody lb = load_lane_body(sm, lane, 0);
  const int max_depth = sm.max_depth;
  const int64_t warp0 = (int64_t)blockIdx.x * PARC_WARPS_PER_CTA + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * PARC_WARPS_PER_CTA;
  for (int64_t q = warp0; q < n; q += nwarps) {
    float3 pos = make_float3(0.f, 0.f, 0.f);
    float4 rot = make_float4(0.f, 0.f, 0.f, 1.f);
    if (lane == 0) {
      pos = make_float3(__ldg(root_pos + q * 3), __ldg(root_pos + q * 3 + 1), __ldg(root_pos + q * 3 + 2));
      rot = __ldg(reinterpret_cast<const float4*>(root_rot) + q);
    } else if (lane < J) {
      rot = __ldg(reinterpret_cast<const float4*>(joint_rot) + q * (J - 1) + (lane - 1));
    }
    fk_warp(lb, max_depth, pos, rot);
    if (lane < J) {
      if (body_pos) {
        float* o = body_pos + (q * J + lane) * 3;
        o[0] = pos.x; o[1] = pos.y; o[2] = pos.z;
      }
      if (body_rot) reinterpret_cast<float4*>(body_rot)[q * J + lane] = rot;
    }
  }
}

__global__ void __launch_bounds__(PARC_CTA_THREADS)
fk_bwd_kernel(const float* __restrict__ root_rot, const float* __restrict__ joint_rot,
              const float* __restrict__ g_body_pos, const float* __restrict__ g_body_rot, int64_t n,
              const __grid_constant__ ParcCharModel model_param, float* __restrict__ g_root_pos,
              float* __restrict__ g_root_rot, float* __restrict__ g_joint_rot) {
  __shared__ ParcCharModel sm;
  stage_model(&sm, model_param);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int J = sm.num_bodies;
  const LaneBody lb = load_lane_body(sm, lane, 0);
  const int max_depth = sm.max_depth;
  const int64_t warp0 = (int64_t)blockIdx.x * PARC_WARPS_PER_CTA + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * PARC_WARPS_PER_CTA;
  for (int64_t q = warp0; q < n; q += nwarps) {
    float3 pos = make_float3(0.f, 0.f, 0.f);
    float4 rot = make_float4(0.f, 0.f, 0.f, 1.f);
    if (lane == 0) rot = __ldg(reinterpret_cast<const float4*>(root_rot) + q);
    else if (lane < J) rot = __ldg(reinterpret_cast<const float4*>(joint_rot) + q * (J - 1) + (lane - 1));
    float4 prot, local;
    fk_warp_keep(lb, max_depth, pos, rot, prot, local);

    float3 gp = make_float3(0.f, 0.f, 0.f);
    float4 gr = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < J) {
      if (g_body_pos) {
        const float* s = g_body_pos + (q * J + lane) * 3;
        gp = make_float3(__ldg(s), __ldg(s + 1), __ldg(s + 2));
      }
      if (g_body_rot) gr = __ldg(reinterpret_cast<const float4*>(g_body_rot) + q * J + lane);
    }
    float4 gj;
    fk_warp_vjp(lb, J, lane, prot, local, gp, gr, gj);
    if (lane == 0) {
      if (g_root_pos) { g_root_pos[q * 3] = gp.x; g_root_pos[q * 3 + 1] = gp.y; g_root_pos[q * 3 + 2] = gp.z; }
      if (g_root_rot) reinterpret_cast<float4*>(g_root_rot)[q] = gr;
    } else if (lane < J) {
      if (g_joint_rot) reinterpret_cast<float4*>(g_joint_rot)[q * (J - 1) + (lane - 1)] = gj;
    }
  }
}

// Body surface points in world space, in the layout the reference's callers build with their per-body loop
// (util/terrain_util.py:1918-1936, diffusion/mdm.py:1006-1020): for batch entry i, body b, frame f, point k of the
// body's P_b points:  out[i, F * start_b + f * P_b + k] = rotate(body_rot[i,f,b], local_k) + body_pos[i,f,b].
// One thread per output point.
__global__ void __launch_bounds__(256)
body_points_fwd_kernel(const float* __restrict__ body_pos, const float* __restrict__ body_rot, int64_t batch,
                       int64_t frames, int J, const __grid_constant__ ParcBodyPoints pts, float* __restrict__ out) {
  const int S = pts.num_points;
  const int64_t per = frames * S, total = batch * per;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t bi = i / per;
    const int64_t r = i - bi * per;                 // position inside the batch entry's [F * S] block
    // body of this slot: the last b with F * start_b <= r (J <= 24: a short linear search)
    int b = 0;
    while (b + 1 < J && (int64_t)__ldg(pts.point_start + b + 1) * frames <= r) ++b;
    const int s0 = __ldg(pts.point_start + b), pb = __ldg(pts.point_start + b + 1) - s0;
    const int64_t rr = r - (int64_t)s0 * frames;
    const int64_t f = rr / pb;
    const int k = (int)(rr - f * pb);
    const int64_t q = (bi * frames + f) * J + b;
    const float4 rot = __ldg(reinterpret_cast<const float4*>(body_rot) + q);
    const float* lp = pts.points + (size_t)(s0 + k) * 3;
    const float3 w = quat_rotate(rot, make_float3(__ldg(lp), __ldg(lp + 1), __ldg(lp + 2)));
    out[i * 3] = w.x + __ldg(body_pos + q * 3);
    out[i * 3 + 1] = w.y + __ldg(body_pos + q * 3 + 1);
    out[i * 3 + 2] = w.z + __ldg(body_pos + q * 3 + 2);
  }
}

// VJP: one thread per (batch entry, frame, body) sums its points' upstream gradients in point order (deterministic).
__global__ void __launch_bounds__(256)
body_points_bwd_kernel(const float* __restrict__ body_rot, const float* __restrict__ g_out, int64_t batch,
                       int64_t frames, int J, const __grid_constant__ ParcBodyPoints pts,
                       float* __restrict__ g_body_pos, float* __restrict__ g_body_rot) {
  const int S = pts.num_points;
  const int64_t total = batch * frames * J;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(q % J);
    const int64_t bf = q / J;
    const int64_t bi = bf / frames, f = bf - bi * frames;
    const int s0 = __ldg(pts.point_start + b), pb = __ldg(pts.point_start + b + 1) - s0;
    const float4 rot = __ldg(reinterpret_cast<const float4*>(body_rot) + q);
    const float* g = g_out + (bi * frames * S + (int64_t)s0 * frames + f * pb) * 3;
    float3 gp = make_float3(0.f, 0.f, 0.f);
    float4 gr = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < pb; ++k) {
      const float3 gk = make_float3(__ldg(g + k * 3), __ldg(g + k * 3 + 1), __ldg(g + k * 3 + 2));
      const float* lp = pts.points + (size_t)(s0 + k) * 3;
      const float4 c = quat_rotate_vjp_q(rot, make_float3(__ldg(lp), __ldg(lp + 1), __ldg(lp + 2)), gk);
      gp.x += gk.x; gp.y += gk.y; gp.z += gk.z;
      gr.x += c.x; gr.y += c.y; gr.z += c.z; gr.w += c.w;
    }
    if (g_body_pos) { g_body_pos[q * 3] = gp.x; g_body_pos[q * 3 + 1] = gp.y; g_body_pos[q * 3 + 2] = gp.z; }
    if (g_body_rot) reinterpret_cast<float4*>(g_body_rot)[q] = gr;
  }
}

static int warp_grid(int64_t n) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = (n + PARC_WARPS_PER_CTA - 1) / PARC_WARPS_PER_CTA;
  const int64_t cap = (int64_t)sms * 8;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

static int flat_grid(int64_t total) {
  int64_t b = (total + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  return (int)(b > 0 ? b : 1);
}

}  // namespace parc

using namespace parc;

extern "C" int parc_fk_fwd(const float* root_pos, const float* root_rot, const float* joint_rot, int64_t n,
                           const ParcCharModel* model, float* body_pos, float* body_rot, void* stream) {
  if (!model) return PARC_E_NULL;
  int rc = parc_validate_model(model);
  if (rc) return rc;
  if (n < 0) return PARC_E_SIZE;
  if (n == 0) return PARC_OK;
  if (!root_pos || !root_rot) return PARC_E_NULL;
  if (model->num_bodies > 1 && !joint_rot) return PARC_E_NULL;
  if (!aligned16(root_rot) || !aligned16(joint_rot) || !aligned16(body_rot)) return PARC_E_ALIGN;
  if (n == 0) return PARC_OK;
  fk_fwd_kernel<<<warp_grid(n), PARC_CTA_THREADS, 0, (cudaStream_t)stream>>>(root_pos, root_rot, joint_rot, n,
                                                                             *model, body_pos, body_rot);
  return check_launch();
}

extern "C" int parc_fk_bwd(const float* root_rot, const float* joint_rot, const float* g_body_pos,
                           const float* g_body_rot, int64_t n, const ParcCharModel* model, float* g_root_pos,
                           float* g_root_rot, float* g_joint_rot, void* stream) {
  if (!model) return PARC_E_NULL;
  int rc = parc_validate_model(model);
  if (rc) return rc;
  if (n < 0) return PARC_E_SIZE;
  if (n == 0) return PARC_OK;
  if (!root_rot) return PARC_E_NULL;
  if (model->num_bodies > 1 && !joint_rot) return PARC_E_NULL;
  if (!aligned16(root_rot) || !aligned16(joint_rot) || !aligned16(g_body_rot) || !aligned16(g_root_rot) ||
      !aligned16(g_joint_rot))
    return PARC_E_ALIGN;
  if (n == 0) return PARC_OK;
  fk_bwd_kernel<<<warp_grid(n), PARC_CTA_THREADS, 0, (cudaStream_t)stream>>>(
      root_rot, joint_rot, g_body_pos, g_body_rot, n, *model, g_root_pos, g_root_rot, g_joint_rot);
  return check_launch();
}

extern "C" int parc_dof_to_rot_fwd(const float* dof, int64_t n, const ParcCharModel* model, float* joint_rot,
                                   void* stream) {
  if (!model) return PARC_E_NULL;
  int rc = parc_validate_model(model);
  if (rc) return rc;
  if (n < 0) return PARC_E_SIZE;
  if (n == 0) return PARC_OK;
  if (!dof || !joint_rot) return PARC_E_NULL;
  if (!aligned16(joint_rot)) return PARC_E_ALIGN;
  if (n == 0 || model->num_bodies < 2) return PARC_OK;
  dof_to_rot_fwd_kernel<<<flat_grid(n * (model->num_bodies - 1)), 256, 0, (cudaStream_t)stream>>>(dof, n, *model,
                                                                                                 joint_rot);
  return check_launch();
}

extern "C" int parc_dof_to_rot_bwd(const float* dof, const float* g_joint_rot, int64_t n,
                                   const ParcCharModel* model, float* g_dof, void* stream) {
  if (!model) return PARC_E_NULL;
  int rc = parc_validate_model(model);
  if (rc) return rc;
  if (n < 0) return PARC_E_SIZE;
  if (n == 0) return PARC_OK;
  if (!dof || !g_joint_rot || !g_dof) return PARC_E_NULL;
  if (!aligned16(g_joint_rot)) return PARC_E_ALIGN;
  if (n == 0 || model->num_bodies < 2) return PARC_OK;
  dof_to_rot_bwd_kernel<<<flat_grid(n * (model->num_bodies - 1)), 256, 0, (cudaStream_t)stream>>>(
      dof, g_joint_rot, n, *model, g_dof);
  return check_launch();
}

extern "C" int parc_exp_map_to_quat_fwd(const float* exp_map, int64_t n, float* quat, void* stream) {
  if (n < 0) return PARC_E_SIZE;
  if (n == 0) return PARC_OK;
  if (!exp_map || !quat) return PARC_E_NULL;
  if (!aligned16(quat)) return PARC_E_ALIGN;
  if (n == 0) return PARC_OK;
  exp_map_fwd_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(exp_map, n, quat);
  return check_launch();
}

extern "C" int parc_exp_map_to_quat_bwd(const float* exp_map, const float* g_quat, int64_t n, float* g_exp_map,
                                        void* stream) {
  if (n < 0) return PARC_E_SIZE;
  if (n == 0) return PARC_OK;
  if (!exp_map || !g_quat || !g_exp_map) return PARC_E_NULL;
  if (!aligned16(g_quat)) return PARC_E_ALIGN;
  if (n == 0) return PARC_OK;
  exp_map_bwd_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(exp_map, g_quat, n, g_exp_map);
  return check_launch();
}

extern "C" int parc_rot_to_dof(const float* joint_rot, int64_t n, const ParcCharModel* model, float* dof_out,
                               void* stream) {
  if (!model) return PARC_E_NULL;
  int rc = parc_validate_model(model);
  if (rc) return rc;
  if (n < 0) return PARC_E_SIZE;
  if (n == 0 || model->num_bodies < 2 || model->dof_size == 0) return PARC_OK;
  if (!joint_rot || !dof_out) return PARC_E_NULL;
  if (!aligned16(joint_rot)) return PARC_E_ALIGN;
  rot_to_dof_kernel<<<flat_grid(n * (model->num_bodies - 1)), 256, 0, (cudaStream_t)stream>>>(joint_rot, n, *model,
                                                                                             dof_out);
  return check_launch();
}

static int check_body_points(const ParcBodyPoints* pts, int64_t batch, int64_t frames, int32_t num_bodies) {
  if (!pts) return PARC_E_NULL;
  if (batch < 0 || frames < 0 || num_bodies < 1 || num_bodies > PARC_MAX_BODIES || pts->num_points < 0) return PARC_E_SIZE;
  if (pts->num_points > 0 && (!pts->points || !pts->point_start)) return PARC_E_NULL;
  return PARC_OK;
}

extern "C" int parc_body_points_fwd(const float* body_pos, const float* body_rot, int64_t batch, int64_t frames,
                                    int32_t num_bodies, const ParcBodyPoints* pts, float* points_out, void* stream) {
  int rc = check_body_points(pts, batch, frames, num_bodies);
  if (rc) return rc;
  const int64_t total = batch * frames * pts->num_points;
  if (total == 0) return PARC_OK;
  if (!body_pos || !body_rot || !points_out) return PARC_E_NULL;
  if (!aligned16(body_rot)) return PARC_E_ALIGN;
  body_points_fwd_kernel<<<flat_grid(total), 256, 0, (cudaStream_t)stream>>>(body_pos, body_rot, batch, frames,
                                                                           num_bodies, *pts, points_out);
  return check_launch();
}

extern "C" int parc_body_points_bwd(const float* body_rot, const float* g_points, int64_t batch, int64_t frames,
                                    int32_t num_bodies, const ParcBodyPoints* pts, float* g_body_pos,
                                    float* g_body_rot, void* stream) {
  int rc = check_body_points(pts, batch, frames, num_bodies);
  if (rc) return rc;
  const int64_t total = batch * frames * num_bodies;
  if (total == 0) return PARC_OK;
  if (!body_rot || !g_points) return PARC_E_NULL;
  if (!aligned16(body_rot) || !aligned16(g_body_rot)) return PARC_E_ALIGN;
  body_points_bwd_kernel<<<flat_grid(total), 256, 0, (cudaStream_t)stream>>>(body_rot, g_points, batch, frames,
                                                                           num_bodies, *pts, g_body_pos, g_body_rot);
  return check_launch();
}
