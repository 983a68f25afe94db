// Kinematic motion optimisation against the terrain, one Adam iteration in four launches (sm_100a).
//
// Reference: tools/motion_opt/motion_optimization.py:183-395 (motion_terrain_contact_loss: tracking, smoothness,
// pseudo-Huber sliding, jerk and body-constraint terms around the penetration / contact terms) and :404-500
// (motion_contact_optimization: torch.optim.Adam over root position, root exp-map and joint DoFs).  The reference
// builds the objective from ~250 eager torch ops per iteration and back-propagates through them with autograd;
// here an iteration is
//   1  parc_frames_fk          leaves [F, 6+D] -> root / joint quaternions, body positions / rotations (scratch)
//   2  body_loss_kernel        penetration + contact terms and their gradient (csrc/body_loss.cu, one launch)
//   3  motion_opt_grad_kernel  every other term AND the whole backward pass: one warp per frame, lane = body;
//                              the temporal stencils (velocity: 2 frames, jerk: 4 frames) are evaluated in gather
//                              form -- each frame recomputes the stencil terms it takes part in -- so the
//                              gradient needs no atomics and is bit-reproducible; then the FK VJP by warp shuffles and
//                              the DoF / exp-map VJPs in the owning lane; writes d loss / d leaves and the per-frame
//                              partial sums of every term
//   4  motion_opt_adam_kernel  torch.optim.Adam's update (default betas / eps semantics, bias correction from a
//                              device-resident step counter so the launch sequence can be replayed as a CUDA graph)
// Sub-gradient conventions follow autograd: clamp passes the gradient at its bound, norm'(0) = 0, abs'(0) = 0,
// torch.where masks route no gradient through the untaken branch.
#include "parc_common.cuh"
#include "parc_internal.h"
#include "parc_rotations.cuh"

namespace parc {

#define OPT_WARPS 4
#define OPT_THREADS (OPT_WARPS * 32)
#define OPT_TERMS 8   // per-frame partial sums: root_pos, root_rot, joint_rot, smoothness, sliding, jerk, constraint, pad

struct OptParams {
  ParcMotionOptArgs a;
  int J, D, stride;       // bodies, DoFs, floats per frame row (6 + D)
  float max_jerk;         // max_jerk * dt^3, rounded to fp32 as the reference's python float meets the tensor
};

__device__ __forceinline__ float3 ldg3(const float* __restrict__ p) {
  return make_float3(__ldg(p), __ldg(p + 1), __ldg(p + 2));
}
__device__ __forceinline__ float4 ldg4(const float* __restrict__ p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ float3 operator-(const float3& a, const float3& b) {
  return make_float3(a.x - b.x, a.y - b.y, a.z - b.z);
}
__device__ __forceinline__ float3 operator+(const float3& a, const float3& b) {
  return make_float3(a.x + b.x, a.y + b.y, a.z + b.z);
}
__device__ __forceinline__ float3 operator*(float s, const float3& a) { return make_float3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float dot3(const float3& a, const float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float4 add4(const float4& a, const float4& b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 scale4(float s, const float4& a) { return make_float4(s * a.x, s * a.y, s * a.z, s * a.w); }

// quat_diff_angle(q0, q1) (util/torch_util.py:422-431 with quat_to_axis_angle :68-88) and the gradients of the angle
// with respect to both quaternions.  angle = 2 atan2(|v|, w) of d = quat_pos(q1 * conj(q0)), 0 (no gradient) when
// |v| <= 1e-5.
// The VALUE path is one non-inlined routine: the optimiser's source constants (body rotation velocities of the source
// clip) and the objective's own rotation velocities must come out of the SAME instructions, so that "target == source"
// gives an error of exactly 0 -- as it does in the reference, where both are the same torch ops.  Two inlined copies
// could contract FMAs differently; a last-bit error would be turned into a full +-lr step by Adam's normalisation.
struct QuatDiff {
  float4 d;      // quat_pos(q1 * conj(q0))
  float len;     // |d.xyz|
  float sign;    // the sign quat_pos applied
  float angle;
};
__device__ __noinline__ QuatDiff quat_diff_value(float4 q0, float4 q1) {
  QuatDiff r;
  float4 d = quat_mul(q1, quat_conj(q0));
  r.sign = d.w < 0.0f ? -1.0f : 1.0f;
  r.d = scale4(r.sign, d);
  r.len = sqrtf(r.d.x * r.d.x + r.d.y * r.d.y + r.d.z * r.d.z);
  r.angle = r.len > 1e-5f ? 2.0f * atan2f(r.len, r.d.w) : 0.0f;
  return r;
}

__device__ __forceinline__ float quat_diff_angle_grad(const float4& q0, const float4& q1, float4& g0, float4& g1) {
  const QuatDiff r = quat_diff_value(q0, q1);
  g0 = g1 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!(r.len > 1e-5f)) return 0.0f;
  const float den = r.len * r.len + r.d.w * r.d.w;
  const float gl = 2.0f * r.d.w / den, gw = -2.0f * r.len / den;
  const float il = r.sign * gl / r.len;
  const float4 gd = make_float4(il * r.d.x, il * r.d.y, il * r.d.z, r.sign * gw);      // d angle / d (q1 * conj(q0))
  g1 = quat_mul_plain(gd, q0);
  g0 = quat_conj(quat_mul_plain(quat_conj(q1), gd));
  return r.angle;
}

__device__ __forceinline__ float warp_sum_all(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(PARC_FULL_MASK, v, o);
  return v;
}

// One warp per frame f, lane b = body b.
__global__ void __launch_bounds__(OPT_THREADS)
motion_opt_grad_kernel(const __grid_constant__ OptParams p, const __grid_constant__ ParcCharModel model_param) {
  __shared__ ParcCharModel sm;
  stage_model(&sm, model_param);
  __syncthreads();
  const ParcMotionOptArgs& a = p.a;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int J = p.J;
  const int64_t F = a.num_frames;
  const LaneBody lb = load_lane_body(sm, lane, 0);
  const bool has = lane < J;
  const int b = has ? lane : 0;
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.step) *a.step += 1;      // the Adam launch that follows reads t = step
  const float c2 = 0.0009f, c1 = 0.03f;

  for (int64_t f = (int64_t)blockIdx.x * OPT_WARPS + warp; f < F; f += (int64_t)gridDim.x * OPT_WARPS) {
    // ---- forward kinematics of frame f again (keeps the parent rotation / local rotation the VJP needs) ----
    float3 pos = make_float3(0.f, 0.f, 0.f);
    float4 rot = make_float4(0.f, 0.f, 0.f, 1.f);
    if (lane == 0) {
      pos = ldg3(a.frames + f * p.stride);
      rot = ldg4(a.root_rot + f * 4);
    } else if (has) {
      rot = ldg4(a.joint_rot + (f * (J - 1) + (lane - 1)) * 4);
    }
    const float4 own_q = rot;                       // root quaternion (lane 0) / joint quaternion (lane >= 1)
    float4 prot, local;
    fk_warp_keep(lb, sm.max_depth, pos, rot, prot, local);
    // the terms below read this frame's body transform from the same scratch arrays as its neighbours' (written by
    // parc_frames_fk), so that every difference of two frames is formed from values of one provenance
    if (has) {
      pos = ldg3(a.body_pos + (f * J + b) * 3);
      rot = ldg4(a.body_rot + (f * J + b) * 4);
    }

    float t_rp = 0.f, t_rr = 0.f, t_jr = 0.f, t_sm = 0.f, t_sl = 0.f, t_jk = 0.f, t_bc = 0.f;
    float3 gP = make_float3(0.f, 0.f, 0.f);         // d loss / d body_pos[f, b]
    float4 gQ = make_float4(0.f, 0.f, 0.f, 0.f);    // d loss / d body_rot[f, b]
    float4 g_own = make_float4(0.f, 0.f, 0.f, 0.f); // direct gradient on the root / joint quaternion
    float3 g_rp = make_float3(0.f, 0.f, 0.f);       // direct gradient on the root position (lane 0)

    // ---- tracking terms (:200-211) ----
    if (lane == 0) {
      const float3 e = ldg3(a.frames + f * p.stride) - ldg3(a.src_root_pos + f * 3);
      t_rp = e.x * e.x + e.y * e.y + e.z * e.z;
      g_rp = (2.0f * a.w_root_pos) * e;
      float4 g0, g1;
      const float ang = quat_diff_angle_grad(own_q, ldg4(a.src_root_rot + f * 4), g0, g1);
      t_rr = ang * ang;
      g_own = scale4(2.0f * a.w_root_rot * ang, g0);
    } else if (has) {
      float4 g0, g1;
      const float ang = quat_diff_angle_grad(own_q, ldg4(a.src_joint_rot + (f * (J - 1) + (lane - 1)) * 4), g0, g1);
      t_jr = ang * ang;
      g_own = scale4(2.0f * a.w_joint_rot * ang, g0);
    }

    if (has) {
      // ---- constraint mask of the velocity pairs (f-1, f) and (f, f+1): a constrained body pays no sliding there ----
      bool mask_prev = false, mask_cur = false;      // true = the pair's squared errors are multiplied by 0 (:329-332)
      for (int c = 0; c < a.num_constraints; ++c) {
        const ParcBodyConstraint& bc = a.constraints[c];
        if (bc.body != b) continue;
        if (bc.start_frame <= f - 1 && f - 1 <= bc.end_frame) mask_prev = true;
        if (bc.start_frame <= f && f <= bc.end_frame) mask_cur = true;
      }
      // ---- positions of frames f-3 .. f+3 and the velocity / acceleration / jerk stencils built from them ----
      float3 P[7];
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        const int64_t ff = f + k - 3;
        P[k] = (ff >= 0 && ff < F) ? ldg3(a.body_pos + (ff * J + b) * 3) : make_float3(0.f, 0.f, 0.f);
      }
      float3 v[6], ac[5];
#pragma unroll
      for (int k = 0; k < 6; ++k) v[k] = P[k + 1] - P[k];            // v[k] = vel[f + k - 3]
#pragma unroll
      for (int k = 0; k < 5; ++k) ac[k] = v[k + 1] - v[k];           // acc[f + k - 3]
      // jerk[f + k - 3] = acc[.+1] - acc[.], valid for 0 <= index <= F - 4; its gradient reaches P through
      // (+1, -3, +3, -1) on frames (index + 3, + 2, + 1, + 0)
      const float coef[4] = {1.0f, -3.0f, 3.0f, -1.0f};              // coefficient of P[f] in jerk[f-3], [f-2], [f-1], [f]
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int64_t ji = f + k - 3;
        if (ji < 0 || ji > F - 4) continue;
        const float3 jv = ac[k + 1] - ac[k];
        const float mag = sqrtf(jv.x * jv.x + jv.y * jv.y + jv.z * jv.z);
        const float over = mag - p.max_jerk;
        if (k == 3) t_jk += fmaxf(over, 0.0f);                        // frame f owns jerk[f]
        if (over >= 0.0f && mag > 0.0f) gP = gP + (a.w_jerk * coef[k] / mag) * jv;
      }
      // ---- smoothness + sliding on the pairs (f-1, f) and (f, f+1) (:213-222, :339-346) ----
      const float cont_f = __ldg(a.contacts + f * J + b);
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        const int64_t vi = f - 1 + side;                              // velocity index: pair (vi, vi + 1)
        if (vi < 0 || vi > F - 2) continue;
        const bool masked = side == 0 ? mask_prev : mask_cur;
        const float other = __ldg(a.contacts + (side == 0 ? f - 1 : f + 1) * J + b);
        const float fcc = fmaxf(fminf(cont_f, other), 0.0f);          // clamp(min(c[f+1], c[f]), min=0)
        // linear part
        const float3 e = v[2 + side] - ldg3(a.src_body_vels + (vi * J + b) * 3);
        const float esq = e.x * e.x + e.y * e.y + e.z * e.z;
        const float hub = sqrtf((masked ? 0.0f : esq) + c2);
        float ge = 2.0f * a.w_smoothness;
        if (!masked && a.w_sliding != 0.0f) ge += a.w_sliding * fcc / hub;
        // e = P[vi + 1] - P[vi] - src: frame f is the later frame of pair f-1 (+), the earlier one of pair f (-)
        gP = gP + (side == 0 ? ge : -ge) * e;
        // angular part: rv = quat_diff_angle(Q[vi + 1], Q[vi])
        const float4 Qa = side == 0 ? rot : ldg4(a.body_rot + ((f + 1) * J + b) * 4);     // q0 = later frame
        const float4 Qb = side == 0 ? ldg4(a.body_rot + ((f - 1) * J + b) * 4) : rot;     // q1 = earlier frame
        float4 g0, g1;
        const float rv = quat_diff_angle_grad(Qa, Qb, g0, g1);
        const float er = rv - __ldg(a.src_body_rot_vels + vi * J + b);
        const float ersq = er * er;
        const float hubr = sqrtf((masked ? 0.0f : ersq) + c2);
        float gr = 2.0f * a.w_smoothness * er;
        if (!masked && a.w_sliding != 0.0f) gr += a.w_sliding * fcc * er / hubr;
        gQ = add4(gQ, scale4(gr, side == 0 ? g0 : g1));
        if (side == 1) {                                              // frame f owns pair f
          t_sm += esq + ersq;
          if (a.w_sliding != 0.0f) t_sl += (hub - c1) * fcc + (hubr - c1) * fcc;
        }
      }
      // ---- body constraints on frame f (:286-327) ----
      for (int c = 0; c < a.num_constraints; ++c) {
        const ParcBodyConstraint& bc = a.constraints[c];
        if (bc.body != b || f < bc.start_frame || f > bc.end_frame) continue;
        const float3 cp = make_float3(bc.point[0], bc.point[1], bc.point[2]);
        if (bc.shape == PARC_CONSTRAINT_SPHERE) {
          const float3 off = make_float3(bc.offset[0], bc.offset[1], bc.offset[2]);
          const float3 d = cp - (quat_rotate(rot, off) + pos);
          const float n = sqrtf(dot3(d, d));
          const float diff = n - bc.radius;                           // sdSphere(point, centre, r)
          t_bc += fabsf(diff);
          if (n > 0.0f && diff != 0.0f) {
            const float3 gc = (-a.w_body_constraints * (diff > 0.0f ? 1.0f : -1.0f) / n) * d;   // d / d centre
            gP = gP + gc;
            gQ = add4(gQ, quat_rotate_vjp_q(rot, off, gc));
          }
        } else {                                                      // box: the 18 sole points (:320)
          const int s0 = __ldg(a.pts.point_start + b);
          for (int k = 0; k < 18; ++k) {
            const float3 lp = ldg3(a.pts.points + (size_t)(s0 + k) * 3);
            const float3 d = cp - (quat_rotate(rot, lp) + pos);
            const float n = sqrtf(dot3(d, d));
            const float diff = n - bc.radius;
            t_bc += fmaxf(diff, 0.0f);
            if (diff >= 0.0f && n > 0.0f) {
              const float3 gc = (-a.w_body_constraints / n) * d;
              gP = gP + gc;
              gQ = add4(gQ, quat_rotate_vjp_q(rot, lp, gc));
            }
          }
        }
      }
    }

    // ---- backward through the kinematic chain, then through the DoF / exp-map conversions ----
    float4 g_joint;
    fk_warp_vjp(lb, J, lane, prot, local, gP, gQ, g_joint);
    float* go = a.grad + f * p.stride;
    if (lane == 0) {
      const float3 bl = a.g_root_pos ? ldg3(a.g_root_pos + f * 3) : make_float3(0.f, 0.f, 0.f);
      go[0] = gP.x + g_rp.x + bl.x; go[1] = gP.y + g_rp.y + bl.y; go[2] = gP.z + g_rp.z + bl.z;
      float4 g = add4(gQ, g_own);
      if (a.g_root_rot) g = add4(g, ldg4(a.g_root_rot + f * 4));
      const float3 e = ldg3(a.frames + f * p.stride + 3);
      const float3 ge = exp_map_to_quat_vjp(e, g);
      go[3] = ge.x; go[4] = ge.y; go[5] = ge.z;
    } else if (has) {
      float4 g = add4(g_joint, g_own);
      if (a.g_joint_rot) g = add4(g, ldg4(a.g_joint_rot + (f * (J - 1) + (lane - 1)) * 4));
      const int jt = sm.joint_type[lane];
      const float* d = a.frames + f * p.stride + 6 + sm.dof_idx[lane];
      float* o = go + 6 + sm.dof_idx[lane];
      if (jt == PARC_JOINT_HINGE) {
        float3 ga;
        float gang;
        axis_angle_to_quat_vjp(make_float3(sm.joint_axis[lane][0], sm.joint_axis[lane][1], sm.joint_axis[lane][2]),
                               __ldg(d), g, ga, gang);
        o[0] = gang;
      } else if (jt == PARC_JOINT_SPHERICAL) {
        const float3 r = exp_map_to_quat_vjp(ldg3(d), g);
        o[0] = r.x; o[1] = r.y; o[2] = r.z;
      }
    }
    // ---- per-frame partial sums of every term (summed over frames only when somebody looks at them) ----
    t_rp = warp_sum_all(t_rp); t_rr = warp_sum_all(t_rr); t_jr = warp_sum_all(t_jr); t_sm = warp_sum_all(t_sm);
    t_sl = warp_sum_all(t_sl); t_jk = warp_sum_all(t_jk); t_bc = warp_sum_all(t_bc);
    if (lane == 0 && a.terms) {
      float* t = a.terms + f * OPT_TERMS;
      t[0] = t_rp; t[1] = t_rr; t[2] = t_jr; t[3] = t_sm; t[4] = t_sl; t[5] = t_jk; t[6] = t_bc; t[7] = 0.0f;
    }
  }
}

// torch.optim.Adam (amsgrad=False, weight_decay=0, maximize=False), single-tensor form (torch/optim/adam.py):
//   exp_avg.lerp_(grad, 1 - beta1); exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
//   step_size = lr / (1 - beta1^t); denom = exp_avg_sq.sqrt() / sqrt(1 - beta2^t) + eps
//   param.addcdiv_(exp_avg, denom, value=-step_size)
// The python-float scalars (step_size, sqrt(bias_correction2), eps) are formed in double and meet the fp32 tensors as
// their fp32 roundings, as there.
__global__ void __launch_bounds__(256)
motion_opt_adam_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m,
                       float* __restrict__ v, int64_t n, const int32_t* __restrict__ step, double lr, double beta1,
                       double beta2, double eps) {
  const int t = *step;
  const double bc1 = 1.0 - pow(beta1, (double)t), bc2 = 1.0 - pow(beta2, (double)t);
  const float step_size = (float)(lr / bc1), bc2_sqrt = (float)sqrt(bc2), epsf = (float)eps;
  const float w1 = (float)(1.0 - beta1), b2 = (float)beta2, w2 = (float)(1.0 - beta2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float g = grad[i];
    const float mi = __fmaf_rn(w1, g - m[i], m[i]);                 // lerp, weight < 0.5: start + w * (end - start)
    const float vi = __fmaf_rn(w2 * g, g, v[i] * b2);
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + epsf;
    param[i] = param[i] - step_size * (mi / denom);
  }
}

// Source-clip constants of the objective from the source's body transforms: one thread per (frame pair, body).
__global__ void __launch_bounds__(256)
motion_opt_source_kernel(const float* __restrict__ body_pos, const float* __restrict__ body_rot, int64_t F, int J,
                         float* __restrict__ vels, float* __restrict__ rot_vels) {
  const int64_t total = (F - 1) * J;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float3 v = ldg3(body_pos + (i + J) * 3) - ldg3(body_pos + i * 3);
    vels[i * 3] = v.x; vels[i * 3 + 1] = v.y; vels[i * 3 + 2] = v.z;
    rot_vels[i] = quat_diff_value(ldg4(body_rot + (i + J) * 4), ldg4(body_rot + i * 4)).angle;
  }
}

static int check_args(const ParcMotionOptArgs* a, const ParcCharModel* model, bool need_adam) {
  if (!a || !model) return PARC_E_NULL;
  int rc = parc_validate_model(model);
  if (rc) return rc;
  if (a->num_frames < 0 || a->num_constraints < 0) return PARC_E_SIZE;
  if (a->num_frames == 0) return PARC_OK;
  if (!a->frames || !a->src_root_pos || !a->src_root_rot || !a->src_joint_rot || !a->contacts || !a->terrain ||
      !a->pts.points || !a->pts.point_start || !a->root_rot || !a->joint_rot || !a->body_pos || !a->body_rot ||
      !a->g_root_pos || !a->g_root_rot || !a->g_joint_rot || !a->grad)
    return PARC_E_NULL;
  if (a->num_frames > 1 && (!a->src_body_vels || !a->src_body_rot_vels)) return PARC_E_NULL;
  if (a->num_constraints > 0 && !a->constraints) return PARC_E_NULL;
  if (!aligned16(a->src_root_rot) || !aligned16(a->src_joint_rot) || !aligned16(a->root_rot) ||
      !aligned16(a->joint_rot) || !aligned16(a->body_rot) || !aligned16(a->g_root_rot) || !aligned16(a->g_joint_rot))
    return PARC_E_ALIGN;
  if (need_adam && (!a->exp_avg || !a->exp_avg_sq || !a->step)) return PARC_E_NULL;
  return PARC_OK;
}

}  // namespace parc

using namespace parc;

extern "C" int parc_motion_opt_loss_grad(const ParcMotionOptArgs* a, const ParcCharModel* model, void* stream) {
  int rc = check_args(a, model, false);
  if (rc) return rc;
  const int64_t F = a->num_frames;
  if (F == 0) return PARC_OK;
  const int J = model->num_bodies, D = model->dof_size, stride = 6 + D;
  // 1: leaves -> quaternions + forward kinematics
  rc = parc_frames_fk(a->frames, F, stride, model, a->root_rot, a->joint_rot, a->body_pos, a->body_rot, stream);
  if (rc) return rc;
  // 2: penetration + contact terms, forward and gradient (root positions read in place from the leaf rows)
  rc = body_loss_launch(a->frames, stride, a->root_rot, a->joint_rot, a->contacts, 1, F, model, &a->pts, a->terrain,
                        a->w_penetration, a->w_contact, a->pen, a->con, a->g_root_pos, a->g_root_rot, a->g_joint_rot,
                        stream);
  if (rc) return rc;
  // 3: every other term + the whole backward pass
  OptParams p;
  p.a = *a; p.J = J; p.D = D; p.stride = stride;
  p.max_jerk = (float)a->max_jerk_dt3;
  int64_t blocks = (F + OPT_WARPS - 1) / OPT_WARPS;
  if (blocks > 148 * 8) blocks = 148 * 8;
  motion_opt_grad_kernel<<<(int)blocks, OPT_THREADS, 0, (cudaStream_t)stream>>>(p, *model);
  return check_launch();
}

extern "C" int parc_motion_opt_adam_step(const ParcMotionOptArgs* a, const ParcCharModel* model, void* stream) {
  int rc = check_args(a, model, true);
  if (rc) return rc;
  const int64_t n = a->num_frames * (6 + model->dof_size);
  if (n == 0) return PARC_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  motion_opt_adam_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(a->frames, a->grad, a->exp_avg, a->exp_avg_sq, n,
                                                                        a->step, a->lr, a->beta1, a->beta2, a->eps);
  return check_launch();
}

extern "C" int parc_motion_opt_iteration(const ParcMotionOptArgs* a, const ParcCharModel* model, void* stream) {
  int rc = parc_motion_opt_loss_grad(a, model, stream);
  if (rc) return rc;
  return parc_motion_opt_adam_step(a, model, stream);
}

extern "C" int parc_motion_opt_source(const float* src_frames, int64_t num_frames, const ParcCharModel* model,
                                      float* src_root_rot, float* src_joint_rot, float* src_body_pos,
                                      float* src_body_rot, float* src_body_vels, float* src_body_rot_vels,
                                      void* stream) {
  if (!model) return PARC_E_NULL;
  int rc = parc_validate_model(model);
  if (rc) return rc;
  if (num_frames < 0) return PARC_E_SIZE;
  if (num_frames == 0) return PARC_OK;
  if (!src_frames || !src_root_rot || !src_joint_rot || !src_body_pos || !src_body_rot) return PARC_E_NULL;
  if (num_frames > 1 && (!src_body_vels || !src_body_rot_vels)) return PARC_E_NULL;
  rc = parc_frames_fk(src_frames, num_frames, 6 + model->dof_size, model, src_root_rot, src_joint_rot, src_body_pos,
                      src_body_rot, stream);
  if (rc || num_frames == 1) return rc;
  const int64_t total = (num_frames - 1) * model->num_bodies;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  motion_opt_source_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src_body_pos, src_body_rot, num_frames,
                                                                          model->num_bodies, src_body_vels,
                                                                          src_body_rot_vels);
  return check_launch();
}
