// Tracker step assembly: what the tracking environment computes around the motion
// query every control step -- the policy observation of the simulated character, the future-target
// observation, the five DeepMimic tracking rewards and the episode flags (with the terrain-height lookup
// under every body fused in).  Each is ONE launch, one warp per env (per (env, step) for the targets),
// lane = joint / body / DoF; the reference runs each as a chain of 40-120 eager torch ops.
//
// Reference: envs/ig_char_env.py:582-626 (compute_char_obs); envs/ig_parkour/mgdm_dm_util.py:462-518
// (compute_tar_obs), :304-397 (convert_to_local, compute_deepmimic_reward), :205-230 + :399-460
// (update_done, compute_done); util/torch_util.py:361-373 (quat_to_tan_norm), :422-431 (quat_diff_angle),
// :491-499 (calc_heading_quat_inv).
//
// Arithmetic: every comparison that decides a flag (distances against thresholds, heights, forces, time)
// is computed with the explicit *_rn intrinsics in the reference's operation order, so flags are
// bit-identical unless a rotation ANGLE lands within an ulp of its threshold (atan2f vs the CPU libm).
#include "parc_common.cuh"
#include "parc_rotations.cuh"

namespace parc {

#define STEP_WARPS 4
#define STEP_THREADS (STEP_WARPS * 32)

// util/torch_util.py:422-431 with :68-88: angle of q1 * conj(q0), w made non-negative, 0 below 1e-5
__device__ __forceinline__ float quat_diff_angle(const float4& q0, const float4& q1) {
  float4 d = quat_mul(q1, quat_conj(q0));
  if (d.w < 0.0f) { d.x = -d.x; d.y = -d.y; d.z = -d.z; d.w = -d.w; }
  const float len = sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
  return len > 1e-5f ? 2.0f * atan2f(len, d.w) : 0.0f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(PARC_FULL_MASK, v, o);
  return v;
}

// Key body k of state row r: either the dense [n,K,3] array of the reference signature, or -- when key_body_ids
// is set -- body key_body_ids[k] of a [rows, num_bodies, 3] body-position array (no gather copy needed).
__device__ __forceinline__ float3 key_position(const float* __restrict__ key_pos, const int32_t* __restrict__ ids,
                                               int num_bodies, int K, int64_t dense_row, int64_t body_row, int k) {
  if (ids) return ld3(key_pos + (body_row * num_bodies + __ldg(ids + k)) * 3);
  return ld3(key_pos + (dense_row * K + k) * 3);
}

__device__ __forceinline__ float3 sub3(const float3& a, const float3& b) {
  return make_float3(a.x - b.x, a.y - b.y, a.z - b.z);
}
__device__ __forceinline__ float sq3(const float3& d) { return d.x * d.x + d.y * d.y + d.z * d.z; }

// ------------------------------------------------------------------------------------------------
// compute_char_obs: [root_h] | root tan-norm 6 | root_vel 3 | root_ang_vel 3 | joint tan-norm 6(J-1) | dof_vel D | key 3K
// ------------------------------------------------------------------------------------------------
// One env's observation row; the calling warp's lanes cooperate.  JR_IN_LANE: lane j already holds joint j's
// rotation in `my_jr` (the fused step kernel converts the DoFs in registers); otherwise it is read from s.joint_rot.
template <bool JR_IN_LANE>
__device__ __forceinline__ void char_obs_env(const ParcCharState& s, int64_t e, int Jm1, int D, int K, int global_obs,
                                             int root_height_obs, float* __restrict__ o, int lane, const float4& my_jr) {
  const int64_t r = e * s.env_stride;
  const float3 rp = ld3(s.root_pos + r * 3);
  const float4 rr = ld4(s.root_rot + r * 4);
  const float4 hinv = heading_inverse_quat(rr);
  if (root_height_obs) {
    if (lane == 0) o[0] = rp.z;
    o += 1;
  }
  if (lane == 0) {
    store_tan_norm(o, global_obs ? rr : quat_mul(hinv, rr));
  } else if (lane == 1) {
    const float3 v = ld3(s.root_vel + r * 3);
    st3(o + 6, global_obs ? v : quat_rotate(hinv, v));
  } else if (lane == 2) {
    const float3 v = ld3(s.root_ang_vel + r * 3);
    st3(o + 9, global_obs ? v : quat_rotate(hinv, v));
  }
  if (JR_IN_LANE) {
    if (lane < Jm1) store_tan_norm(o + 12 + 6 * lane, my_jr);
  } else {
    for (int j = lane; j < Jm1; j += 32) store_tan_norm(o + 12 + 6 * j, ld4(s.joint_rot + (r * Jm1 + j) * 4));
  }
  float* __restrict__ ov = o + 12 + 6 * Jm1;
  for (int d = lane; d < D; d += 32) ov[d] = __ldg(s.dof_vel + r * D + d);
  float* __restrict__ ok = ov + D;
  for (int k = lane; k < K; k += 32) {
    float3 p = sub3(key_position(s.key_pos, s.key_body_ids, s.num_bodies, K, e, r, k), rp);
    if (!global_obs) p = quat_rotate(hinv, p);
    st3(ok + 3 * k, p);
  }
}

__global__ void __launch_bounds__(STEP_THREADS)
char_obs_kernel(const __grid_constant__ ParcCharState s, int64_t n, int Jm1, int D, int K, int global_obs,
                int root_height_obs, float* __restrict__ out, int64_t W) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * STEP_WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * STEP_WARPS;
  const float4 none = make_float4(0.f, 0.f, 0.f, 1.f);
  for (int64_t e = warp0; e < n; e += nwarps)
    char_obs_env<false>(s, e, Jm1, D, K, global_obs, root_height_obs, out + e * W, lane, none);
}

// ------------------------------------------------------------------------------------------------
// compute_tar_obs: per (env, step)  root_pos_obs 3 | root tan-norm 6 | joint tan-norm 6(J-1) | key 3K
// One warp per env; the lanes stride over the env's work items so all 32 stay busy: first the S * (J-1) joint
// encodings (they do not depend on the character, so they cover the latency of its root load and of the
// atan2 / sincos chain of the heading frame), then the S * (1 + K) root / key-body items in that frame.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(STEP_THREADS)
tar_obs_kernel(const float* __restrict__ ref_root_pos, const float* __restrict__ ref_root_rot,
               const float* __restrict__ tar_root_pos, const float* __restrict__ tar_root_rot,
               const float* __restrict__ tar_joint_rot, const float* __restrict__ tar_key_pos, int64_t n, int S,
               int Jm1, int K, int global_obs, int global_tar_root_h, int tar_env_stride,
               const int32_t* __restrict__ key_body_ids, int num_bodies, float* __restrict__ out,
               int64_t out_env_stride) {
  const int lane = threadIdx.x & 31;
  const int W = 9 + 6 * Jm1 + 3 * K;
  const int joint_items = S * Jm1;            // independent of the character's frame: done first
  const int frame_items = S * (1 + K);        // root + key bodies of each step: need the heading frame
  const int64_t warp0 = (int64_t)blockIdx.x * STEP_WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * STEP_WARPS;
  for (int64_t e = warp0; e < n; e += nwarps) {
    // the character's root is requested first and consumed last: the joint encodings below hide its latency
    const float3 cp = ld3(ref_root_pos + e * 3);
    float4 cq = make_float4(0.f, 0.f, 0.f, 1.f);
    if (!global_obs) cq = ld4(ref_root_rot + e * 4);
    float* __restrict__ oe = out + e * out_env_stride;
    for (int it = lane; it < joint_items; it += 32) {
      const int st = it / Jm1;
      const int j = it - st * Jm1;
      const int64_t r = e * tar_env_stride + st;             // row of this (env, step) in the target arrays
      store_tan_norm(oe + (int64_t)st * W + 9 + 6 * j, ld4(tar_joint_rot + (r * Jm1 + j) * 4));
    }
    float4 hinv = make_float4(0.f, 0.f, 0.f, 1.f);
    if (!global_obs) hinv = heading_inverse_quat(cq);
    for (int it = lane; it < frame_items; it += 32) {
      const int st = it / (1 + K);
      const int w = it - st * (1 + K);                       // 0 = root, 1.. = key bodies
      const int64_t r = e * tar_env_stride + st;
      float* __restrict__ o = oe + (int64_t)st * W;
      const float3 tp = ld3(tar_root_pos + r * 3);
      float3 po = sub3(tp, cp);
      if (!global_obs) po = quat_rotate(hinv, po);
      if (w == 0) {
        float4 tr = ld4(tar_root_rot + r * 4);
        if (!global_obs) tr = quat_mul(hinv, tr);
        st3(o, make_float3(po.x, po.y, global_tar_root_h ? tp.z : po.z));
        store_tan_norm(o + 3, tr);
      } else {
        const int k = w - 1;
        float3 p = sub3(key_position(tar_key_pos, key_body_ids, num_bodies, K, e * S + st, r, k), tp);
        if (!global_obs) {
          p = quat_rotate(hinv, p);
          p.x += po.x; p.y += po.y; p.z += po.z;     // the rotated root offset, BEFORE the height override
        }
        st3(o + 9 + 6 * Jm1 + 3 * k, p);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// compute_deepmimic_reward: [n,5] = exp(-scale * err) for pose, vel, root pose, root vel, key pos
// ------------------------------------------------------------------------------------------------
template <bool JR_IN_LANE>
__device__ __forceinline__ void reward_env(const ParcCharState& s, const ParcCharState& t, int64_t e, int Jm1, int D, int K,
                                           const float* __restrict__ joint_w, const float* __restrict__ dof_w,
                                           int track_root_h, int track_root, float* __restrict__ o, int lane,
                                           const float4& my_jr) {
  float pose = 0.0f, vel = 0.0f, key = 0.0f;
  const int64_t rs = e * s.env_stride, rt = e * t.env_stride;
  if (JR_IN_LANE) {
    if (lane < Jm1) {
      const float a = quat_diff_angle(my_jr, ld4(t.joint_rot + (rt * Jm1 + lane) * 4));
      pose = __ldg(joint_w + lane) * a * a;
    }
  } else {
    for (int j = lane; j < Jm1; j += 32) {
      const float a = quat_diff_angle(ld4(s.joint_rot + (rs * Jm1 + j) * 4), ld4(t.joint_rot + (rt * Jm1 + j) * 4));
      pose += __ldg(joint_w + j) * a * a;
    }
  }
  for (int d = lane; d < D; d += 32) {
    const float dv = __ldg(t.dof_vel + rt * D + d) - __ldg(s.dof_vel + rs * D + d);
    vel += __ldg(dof_w + d) * dv * dv;
  }
  const float3 rp = ld3(s.root_pos + rs * 3), trp = ld3(t.root_pos + rt * 3);
  float4 rr = ld4(s.root_rot + rs * 4), trr = ld4(t.root_rot + rt * 4);
  float3 rv = ld3(s.root_vel + rs * 3), trv = ld3(t.root_vel + rt * 3);
  float3 rw = ld3(s.root_ang_vel + rs * 3), trw = ld3(t.root_ang_vel + rt * 3);
  float3 dp = sub3(trp, rp);
  if (!track_root) dp.x = dp.y = 0.0f;
  if (!track_root_h) dp.z = 0.0f;
  float4 hs = make_float4(0.f, 0.f, 0.f, 1.f), ht = hs;
  if (!track_root) {                                 // convert_to_local: each side in its OWN heading frame
    hs = heading_inverse_quat(rr);
    ht = heading_inverse_quat(trr);
    rv = quat_rotate(hs, rv); rw = quat_rotate(hs, rw); rr = quat_mul(hs, rr);
    trv = quat_rotate(ht, trv); trw = quat_rotate(ht, trw); trr = quat_mul(ht, trr);
  }
  for (int k = lane; k < K; k += 32) {
    float3 a = sub3(key_position(s.key_pos, s.key_body_ids, s.num_bodies, K, e, rs, k), rp);
    float3 b = sub3(key_position(t.key_pos, t.key_body_ids, t.num_bodies, K, e, rt, k), trp);
    if (!track_root) { a = quat_rotate(hs, a); b = quat_rotate(ht, b); }
    key += sq3(sub3(b, a));
  }
  pose = warp_sum(pose);
  vel = warp_sum(vel);
  key = warp_sum(key);
  if (lane == 0) {
    const float ra = quat_diff_angle(rr, trr);
    const float root_pose = sq3(dp) + 0.1f * (ra * ra);
    const float root_vel = sq3(sub3(trv, rv)) + 0.1f * sq3(sub3(trw, rw));
    o[0] = expf(-0.25f * pose);
    o[1] = expf(-0.01f * vel);
    o[2] = expf(-5.0f * root_pose);
    o[3] = expf(-1.0f * root_vel);
    o[4] = expf(-10.0f * key);
  }
}

__global__ void __launch_bounds__(STEP_THREADS)
deepmimic_reward_kernel(const __grid_constant__ ParcCharState s, const __grid_constant__ ParcCharState t, int64_t n,
                        int Jm1, int D, int K, const float* __restrict__ joint_w, const float* __restrict__ dof_w,
                        int track_root_h, int track_root, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * STEP_WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * STEP_WARPS;
  const float4 none = make_float4(0.f, 0.f, 0.f, 1.f);
  for (int64_t e = warp0; e < n; e += nwarps)
    reward_env<false>(s, t, e, Jm1, D, K, joint_w, dof_w, track_root_h, track_root, out + e * 5, lane, none);
}

// ------------------------------------------------------------------------------------------------
// compute_done (+ the termination-height lookup of RefCharEnv.update_done): int32 flag per env
// ------------------------------------------------------------------------------------------------
struct DoneParams {
  const float* time;
  const float* root_rot;
  const float* body_pos;
  const float* tar_root_rot;
  const float* tar_body_pos;
  const float* contact_force;
  const float* term_heights;      // [n,J] precomputed, or NULL -> sampled from hf at body xy + env_offsets
  const float* env_offsets;       // [n, offset_stride] (first two columns used) or NULL
  const float* pose_dist;         // [J-1]
  ParcHeightfield hf;
  int offset_stride;
  float termination_height;
  float ep_len, first_step_eps, force_eps, root_pos_dist_sq, root_rot_angle;
  uint32_t contact_body_mask;
  int has_contact_bodies, pose_termination, early_termination, track_root;
  int J;
  int tar_env_stride;
};

__device__ __forceinline__ void done_env(const DoneParams& p, int64_t e, int32_t* __restrict__ done,
                                         float* __restrict__ term_heights_out, int lane) {
  const int J = p.J;
  const float tm = __ldg(p.time + e);
  int flag = PARC_DONE_NULL;
  if (tm >= p.ep_len) flag = PARC_DONE_TIME;
  const bool body = lane < J;
  float3 bp = make_float3(0.f, 0.f, 0.f);
  if (body) bp = ld3(p.body_pos + (e * J + lane) * 3);
  float th = 0.0f;
  const bool need_heights = (p.early_termination && p.has_contact_bodies) || term_heights_out;
  if (body && need_heights) {
    if (p.term_heights) {
      th = __ldg(p.term_heights + e * J + lane);
    } else {
      float gx = bp.x, gy = bp.y;
      if (p.env_offsets) {
        gx = add_rn(gx, __ldg(p.env_offsets + e * p.offset_stride));
        gy = add_rn(gy, __ldg(p.env_offsets + e * p.offset_stride + 1));
      }
      // hoisted-reciprocal cell index: index-identical to the reference's true division (parc_selftest_grid_index)
      const int ix = grid_index_fast(gx, make_grid_axis(p.hf.min_x, p.hf.dx, p.hf.dim_x));
      const int iy = grid_index_fast(gy, make_grid_axis(p.hf.min_y, p.hf.dy, p.hf.dim_y));
      th = add_rn(__ldg(p.hf.hf + (size_t)ix * p.hf.dim_y + iy), p.termination_height);
    }
    if (term_heights_out) term_heights_out[e * J + lane] = th;
  }
  if (p.early_termination) {
    bool failed = false;
    if (p.has_contact_bodies) {
      const bool counted = body && !((p.contact_body_mask >> lane) & 1u);
      bool touched = false;
      if (counted) {
        const float3 f = ld3(p.contact_force + (e * J + lane) * 3);
        touched = fabsf(f.x) > p.force_eps || fabsf(f.y) > p.force_eps || fabsf(f.z) > p.force_eps;
      }
      const bool low = counted && bp.z < th;
      const bool any_touch = __any_sync(PARC_FULL_MASK, touched);
      const bool any_low = __any_sync(PARC_FULL_MASK, low);
      failed = any_touch && any_low;
    }
    if (p.pose_termination) {
      float3 tb = make_float3(0.f, 0.f, 0.f);
      if (body) tb = ld3(p.tar_body_pos + (e * p.tar_env_stride * J + lane) * 3);
      const float3 r0 = shfl3(bp, 0), t0 = shfl3(tb, 0);
      bool bad = false;
      if (body && lane > 0) {
        const float dx = sub_rn(sub_rn(tb.x, t0.x), sub_rn(bp.x, r0.x));
        const float dy = sub_rn(sub_rn(tb.y, t0.y), sub_rn(bp.y, r0.y));
        const float dz = sub_rn(sub_rn(tb.z, t0.z), sub_rn(bp.z, r0.z));
        const float d2 = add_rn(add_rn(mul_rn(dx, dx), mul_rn(dy, dy)), mul_rn(dz, dz));
        const float lim = __ldg(p.pose_dist + lane - 1);
        bad = d2 > mul_rn(lim, lim);
      } else if (lane == 0 && p.track_root) {
        const float dx = sub_rn(bp.x, tb.x), dy = sub_rn(bp.y, tb.y), dz = sub_rn(bp.z, tb.z);
        const float d2 = add_rn(add_rn(mul_rn(dx, dx), mul_rn(dy, dy)), mul_rn(dz, dz));
        const float ang = quat_diff_angle(ld4(p.root_rot + e * 4), ld4(p.tar_root_rot + e * p.tar_env_stride * 4));
        bad = d2 > p.root_pos_dist_sq || fabsf(ang) > p.root_rot_angle;
      }
      failed = failed || __any_sync(PARC_FULL_MASK, bad);
    }
    if (failed && tm > p.first_step_eps) flag = PARC_DONE_FAIL;
  }
  if (lane == 0 && done) done[e] = flag;
}

__global__ void __launch_bounds__(STEP_THREADS)
done_kernel(const __grid_constant__ DoneParams p, int64_t n, int32_t* __restrict__ done,
            float* __restrict__ term_heights_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * STEP_WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * STEP_WARPS;
  for (int64_t e = warp0; e < n; e += nwarps) done_env(p, e, done, term_heights_out, lane);
}

// ------------------------------------------------------------------------------------------------
// The simulated character's whole step in ONE launch: DoF -> joint rotations (kept in registers, lane = joint),
// the character observation block, the reward terms, the episode flag and the two contact-flag blocks of the
// policy-observation row.  Same device code as the stand-alone kernels above.
// ------------------------------------------------------------------------------------------------
struct SimStepParams {
  ParcCharState sim;              // joint_rot unused (converted from dof_pos); key_pos = body positions if key ids set
  ParcCharState ref;              // reference frame (step 0 of the query output, env_stride = S + 1)
  const float* dof_pos;           // [n, D]
  const float* joint_w;
  const float* dof_w;
  const float* tar_contacts;      // [n, tar_env_stride, J] at step 1, or nullptr
  const float* char_contacts;     // [n, J] or nullptr
  float* joint_rot_out;           // [n, J-1, 4] or nullptr
  float* char_obs_out;            // row stride obs_stride
  float* tar_contacts_out;        // row stride obs_stride, S*J floats per env, or nullptr
  float* char_contacts_out;       // row stride obs_stride, J floats per env, or nullptr
  float* reward_out;              // [n, 5]
  int32_t* done_out;              // [n]
  int64_t obs_stride;
  int tar_env_stride, num_tar_steps;
  int K, global_obs, root_height_obs, track_root_h, track_root;
  DoneParams done;
};

// PHASE: PARC_SIM_STEP_ALL = the whole step in one launch.  _PRE = the share that does not depend on the reference
// frame -- DoF conversion (joint rotations written to joint_rot_out), character observation, character contact block --
// which a caller can run BESIDE the motion query; _POST = the share that needs it -- target contact block, reward
// terms, episode flag -- reading the joint rotations _PRE stored (the same bits the single launch keeps in registers).

// 7 CTAs x 4 warps per SM = 28 resident warps: exactly one wave for 4096 envs on 148 SMs (27.7 warps per SM); at 80
// registers only 6 CTAs fit and the kernel needs a second wave (+8 us measured).
template <int PHASE>
__global__ void __launch_bounds__(STEP_THREADS, 7)
sim_step_kernel(const __grid_constant__ SimStepParams p, const __grid_constant__ ParcCharModel model_param, int64_t n) {
  const int lane = threadIdx.x & 31;
  const int J = model_param.num_bodies, Jm1 = J - 1, D = model_param.dof_size;
  // lane j (< J-1) owns joint j + 1: its five constants are read once, straight from the kernel parameter (staging
  // the whole 1.4 KB model through shared memory held 21 % of this kernel's stall samples)
  int jt = PARC_JOINT_FIXED, didx = 0;
  float axis[3] = {0.0f, 0.0f, 1.0f};
  if (PHASE != PARC_SIM_STEP_POST && lane < Jm1) {
    jt = model_param.joint_type[lane + 1];
    didx = model_param.dof_idx[lane + 1];
    axis[0] = model_param.joint_axis[lane + 1][0];
    axis[1] = model_param.joint_axis[lane + 1][1];
    axis[2] = model_param.joint_axis[lane + 1][2];
  }
  const int64_t warp0 = (int64_t)blockIdx.x * STEP_WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * STEP_WARPS;
  for (int64_t e = warp0; e < n; e += nwarps) {
    float4 jr = make_float4(0.f, 0.f, 0.f, 1.f);
    if (PHASE == PARC_SIM_STEP_POST) {
      if (lane < Jm1) jr = ld4(p.joint_rot_out + (e * Jm1 + lane) * 4);
    } else {
      float dd[3] = {0.f, 0.f, 0.f};
      const float* d = p.dof_pos + e * D + didx;
      if (jt == PARC_JOINT_HINGE) dd[0] = __ldg(d);
      else if (jt == PARC_JOINT_SPHERICAL) { dd[0] = __ldg(d); dd[1] = __ldg(d + 1); dd[2] = __ldg(d + 2); }
      jr = joint_dof_to_quat(jt, dd, axis);
      if (p.joint_rot_out && lane < Jm1) reinterpret_cast<float4*>(p.joint_rot_out)[e * Jm1 + lane] = jr;
      char_obs_env<true>(p.sim, e, Jm1, D, p.K, p.global_obs, p.root_height_obs, p.char_obs_out + e * p.obs_stride, lane, jr);
      if (p.char_contacts_out && lane < J) p.char_contacts_out[e * p.obs_stride + lane] = __ldg(p.char_contacts + e * J + lane);
    }
    if (PHASE != PARC_SIM_STEP_PRE) {
      if (p.tar_contacts_out) {
        const float* __restrict__ src = p.tar_contacts + e * p.tar_env_stride * J;
        float* __restrict__ dst = p.tar_contacts_out + e * p.obs_stride;
        for (int i = lane; i < p.num_tar_steps * J; i += 32) dst[i] = __ldg(src + i);
      }
      reward_env<true>(p.sim, p.ref, e, Jm1, D, p.K, p.joint_w, p.dof_w, p.track_root_h, p.track_root,
                       p.reward_out + e * 5, lane, jr);
      done_env(p.done, e, p.done_out, nullptr, lane);
    }
  }
}

static int warp_grid(int64_t warps) {
  int64_t b = (warps + STEP_WARPS - 1) / STEP_WARPS;
  if (b > 148 * 16) b = 148 * 16;
  return (int)(b > 0 ? b : 1);
}

static int check_state(const ParcCharState* s, int D, int K, bool need_vel) {
  if (!s) return PARC_E_NULL;
  if (!s->root_pos || !s->root_rot || !s->joint_rot) return PARC_E_NULL;
  if (need_vel && (!s->root_vel || !s->root_ang_vel || (D > 0 && !s->dof_vel))) return PARC_E_NULL;
  if (K > 0 && !s->key_pos) return PARC_E_NULL;
  if (s->env_stride < 1 || (s->key_body_ids && s->num_bodies < 1)) return PARC_E_SIZE;
  if (!aligned16(s->root_rot) || !aligned16(s->joint_rot)) return PARC_E_ALIGN;
  return PARC_OK;
}

}  // namespace parc

using namespace parc;

extern "C" int parc_char_obs(const ParcCharState* state, int64_t n, int32_t num_joint_rots, int32_t dof_size,
                             int32_t num_keys, int32_t global_obs, int32_t root_height_obs, float* obs_out,
                             int64_t out_stride, void* stream) {
  if (n < 0 || num_joint_rots < 0 || dof_size < 0 || num_keys < 0) return PARC_E_SIZE;
  const int64_t width = (root_height_obs ? 1 : 0) + 12 + 6 * (int64_t)num_joint_rots + dof_size + 3 * (int64_t)num_keys;
  if (out_stride == 0) out_stride = width;
  if (out_stride < width) return PARC_E_SIZE;
  if (n == 0) return PARC_OK;
  int rc = check_state(state, dof_size, num_keys, true);
  if (rc) return rc;
  if (!obs_out) return PARC_E_NULL;
  char_obs_kernel<<<warp_grid(n), STEP_THREADS, 0, (cudaStream_t)stream>>>(*state, n, num_joint_rots, dof_size,
                                                                         num_keys, global_obs, root_height_obs, obs_out,
                                                                         out_stride);
  return check_launch();
}

extern "C" int parc_tar_obs(const float* ref_root_pos, const float* ref_root_rot, const float* tar_root_pos,
                            const float* tar_root_rot, const float* tar_joint_rot, const float* tar_key_pos,
                            int64_t n, int32_t num_steps, int32_t num_joint_rots, int32_t num_keys,
                            int32_t global_obs, int32_t global_tar_root_h_obs, int32_t tar_env_stride,
                            const int32_t* key_body_ids, int32_t num_bodies, float* obs_out,
                            int64_t out_env_stride, void* stream) {
  if (n < 0 || num_steps < 0 || num_joint_rots < 0 || num_keys < 0) return PARC_E_SIZE;
  if (tar_env_stride < num_steps || (key_body_ids && num_bodies < 1)) return PARC_E_SIZE;
  const int64_t env_width = (int64_t)num_steps * (9 + 6 * (int64_t)num_joint_rots + 3 * (int64_t)num_keys);
  if (out_env_stride == 0) out_env_stride = env_width;
  if (out_env_stride < env_width) return PARC_E_SIZE;
  if (n == 0 || num_steps == 0) return PARC_OK;
  if (!ref_root_pos || !ref_root_rot || !tar_root_pos || !tar_root_rot || !obs_out) return PARC_E_NULL;
  if ((num_joint_rots > 0 && !tar_joint_rot) || (num_keys > 0 && !tar_key_pos)) return PARC_E_NULL;
  if (!aligned16(ref_root_rot) || !aligned16(tar_root_rot) || !aligned16(tar_joint_rot)) return PARC_E_ALIGN;
  tar_obs_kernel<<<warp_grid(n), STEP_THREADS, 0, (cudaStream_t)stream>>>(
      ref_root_pos, ref_root_rot, tar_root_pos, tar_root_rot, tar_joint_rot, tar_key_pos, n, num_steps,
      num_joint_rots, num_keys, global_obs, global_tar_root_h_obs, tar_env_stride, key_body_ids, num_bodies, obs_out,
      out_env_stride);
  return check_launch();
}

extern "C" int parc_deepmimic_reward(const ParcCharState* sim, const ParcCharState* tar, int64_t n,
                                     int32_t num_joint_rots, int32_t dof_size, int32_t num_keys,
                                     const float* joint_rot_err_w, const float* dof_err_w, int32_t track_root_h,
                                     int32_t track_root, float* reward_out, void* stream) {
  if (n < 0 || num_joint_rots < 0 || dof_size < 0 || num_keys < 0) return PARC_E_SIZE;
  if (num_keys == 0) return PARC_E_SIZE;      // the reference cannot stack an empty key term either (:395-397)
  if (n == 0) return PARC_OK;
  int rc = check_state(sim, dof_size, num_keys, true);
  if (rc) return rc;
  rc = check_state(tar, dof_size, num_keys, true);
  if (rc) return rc;
  if (!reward_out || (num_joint_rots > 0 && !joint_rot_err_w) || (dof_size > 0 && !dof_err_w)) return PARC_E_NULL;
  deepmimic_reward_kernel<<<warp_grid(n), STEP_THREADS, 0, (cudaStream_t)stream>>>(
      *sim, *tar, n, num_joint_rots, dof_size, num_keys, joint_rot_err_w, dof_err_w, track_root_h, track_root,
      reward_out);
  return check_launch();
}

// Validation + parameter block shared by parc_done and parc_sim_step.
static int build_done_params(const ParcDoneSpec* spec, const float* time, const float* root_rot, const float* body_pos,
                             const float* tar_root_rot, const float* tar_body_pos, const float* contact_force,
                             const float* term_heights, const ParcHeightfield* hf, const float* env_offsets,
                             int32_t offset_stride, int32_t tar_env_stride, int32_t num_bodies, bool want_heights_out,
                             DoneParams* out) {
  if (!spec) return PARC_E_NULL;
  if (num_bodies < 1 || num_bodies > PARC_MAX_BODIES || tar_env_stride < 1) return PARC_E_SIZE;
  if (!time || !body_pos) return PARC_E_NULL;
  const bool early = spec->enable_early_termination != 0;
  const bool fall = early && spec->has_contact_bodies;
  if (fall && !contact_force) return PARC_E_NULL;
  if ((fall || want_heights_out) && !term_heights) {
    if (!hf || !hf->hf) return PARC_E_NULL;
    if (hf->dim_x <= 0 || hf->dim_y <= 0) return PARC_E_SIZE;
    if (env_offsets && offset_stride < 2) return PARC_E_SIZE;
  }
  if (early && spec->pose_termination) {
    if (!tar_body_pos || !spec->pose_termination_dist) return PARC_E_NULL;
    if (spec->track_root && (!root_rot || !tar_root_rot)) return PARC_E_NULL;
    if (spec->track_root && (!aligned16(root_rot) || !aligned16(tar_root_rot))) return PARC_E_ALIGN;
  }
  DoneParams p{};
  p.time = time; p.root_rot = root_rot; p.body_pos = body_pos; p.tar_root_rot = tar_root_rot;
  p.tar_body_pos = tar_body_pos; p.contact_force = contact_force; p.term_heights = term_heights;
  p.env_offsets = env_offsets; p.pose_dist = spec->pose_termination_dist;
  if (hf) p.hf = *hf;
  p.offset_stride = offset_stride;
  // python-float scalars meet fp32 tensors as fp32 values; products of two scalars are formed in double first
  p.termination_height = (float)spec->termination_height;
  p.ep_len = (float)spec->episode_length;
  p.first_step_eps = (float)1e-5;
  p.force_eps = (float)0.1;
  p.root_pos_dist_sq = (float)(spec->root_pos_termination_dist * spec->root_pos_termination_dist);
  p.root_rot_angle = (float)spec->root_rot_termination_angle;
  p.contact_body_mask = spec->contact_body_mask;
  p.has_contact_bodies = spec->has_contact_bodies; p.pose_termination = spec->pose_termination;
  p.early_termination = spec->enable_early_termination; p.track_root = spec->track_root;
  p.J = num_bodies;
  p.tar_env_stride = tar_env_stride;
  *out = p;
  return PARC_OK;
}

extern "C" int parc_done(const ParcDoneSpec* spec, const float* time, const float* root_rot, const float* body_pos,
                         const float* tar_root_rot, const float* tar_body_pos, const float* contact_force,
                         const float* term_heights, const ParcHeightfield* hf, const float* env_offsets,
                         int32_t offset_stride, int32_t tar_env_stride, int64_t n, int32_t num_bodies,
                         int32_t* done_out, float* term_heights_out, void* stream) {
  if (!spec) return PARC_E_NULL;
  if (n < 0 || num_bodies < 1 || num_bodies > PARC_MAX_BODIES || tar_env_stride < 1) return PARC_E_SIZE;
  if (n == 0) return PARC_OK;
  if (!done_out && !term_heights_out) return PARC_E_NULL;
  DoneParams p;
  const int rc = build_done_params(spec, time, root_rot, body_pos, tar_root_rot, tar_body_pos, contact_force, term_heights,
                                   hf, env_offsets, offset_stride, tar_env_stride, num_bodies, term_heights_out != nullptr,
                                   &p);
  if (rc) return rc;
  done_kernel<<<warp_grid(n), STEP_THREADS, 0, (cudaStream_t)stream>>>(p, n, done_out, term_heights_out);
  return check_launch();
}

extern "C" int parc_sim_step(const ParcSimStep* a, int64_t n, const ParcCharModel* model, void* stream) {
  if (!a || !model) return PARC_E_NULL;
  int rc = parc_validate_model(model);
  if (rc) return rc;
  if (n < 0 || a->num_keys < 1 || a->num_tar_steps < 0) return PARC_E_SIZE;   // the reward needs key bodies
  if (n == 0) return PARC_OK;
  const int J = model->num_bodies, D = model->dof_size;
  ParcCharState sim = a->sim;
  sim.joint_rot = a->sim.root_rot;          // never read (the DoFs are converted in registers); keeps the checker happy
  rc = check_state(&sim, D, a->num_keys, true);
  if (rc) return rc;
  rc = check_state(&a->ref, D, a->num_keys, true);
  if (rc) return rc;
  if (!a->dof_pos || !a->char_obs_out || !a->reward_out || !a->done_out || !a->joint_rot_err_w || (D > 0 && !a->dof_err_w))
    return PARC_E_NULL;
  if (a->joint_rot_out && !aligned16(a->joint_rot_out)) return PARC_E_ALIGN;
  if ((a->tar_contacts_out && !a->tar_contacts) || (a->char_contacts_out && !a->char_contacts)) return PARC_E_NULL;
  if (a->tar_contacts_out && a->tar_env_stride < a->num_tar_steps) return PARC_E_SIZE;
  const int64_t width = (a->root_height_obs ? 1 : 0) + 12 + 6 * (int64_t)(J - 1) + D + 3 * (int64_t)a->num_keys;
  if (a->obs_stride < width) return PARC_E_SIZE;
  SimStepParams p{};
  rc = build_done_params(&a->done, a->time, a->sim.root_rot, a->body_pos, a->ref.root_rot, a->ref_body_pos, a->contact_force,
                         nullptr, &a->hf, a->env_offsets, a->offset_stride, a->ref.env_stride, J, false, &p.done);
  if (rc) return rc;
  p.sim = sim; p.ref = a->ref; p.dof_pos = a->dof_pos; p.joint_w = a->joint_rot_err_w; p.dof_w = a->dof_err_w;
  p.tar_contacts = a->tar_contacts; p.char_contacts = a->char_contacts; p.joint_rot_out = a->joint_rot_out;
  p.char_obs_out = a->char_obs_out; p.tar_contacts_out = a->tar_contacts_out; p.char_contacts_out = a->char_contacts_out;
  p.reward_out = a->reward_out; p.done_out = a->done_out; p.obs_stride = a->obs_stride;
  p.tar_env_stride = a->tar_env_stride; p.num_tar_steps = a->num_tar_steps;
  p.K = a->num_keys; p.global_obs = a->global_obs; p.root_height_obs = a->root_height_obs;
  p.track_root_h = a->track_root_h; p.track_root = a->track_root;
  if (a->phase < PARC_SIM_STEP_ALL || a->phase > PARC_SIM_STEP_POST) return PARC_E_SIZE;
  if (a->phase != PARC_SIM_STEP_ALL && !a->joint_rot_out) return PARC_E_NULL;     // the two halves meet in joint_rot_out
  if (a->phase == PARC_SIM_STEP_PRE)
    sim_step_kernel<PARC_SIM_STEP_PRE><<<warp_grid(n), STEP_THREADS, 0, (cudaStream_t)stream>>>(p, *model, n);
  else if (a->phase == PARC_SIM_STEP_POST)
    sim_step_kernel<PARC_SIM_STEP_POST><<<warp_grid(n), STEP_THREADS, 0, (cudaStream_t)stream>>>(p, *model, n);
  else
    sim_step_kernel<PARC_SIM_STEP_ALL><<<warp_grid(n), STEP_THREADS, 0, (cudaStream_t)stream>>>(p, *model, n);
  return check_launch();
}
