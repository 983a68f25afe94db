// Fused MotionLib frame query -> forward kinematics -> heightmap observation (sm_100a).
//
// One warp per (clip id, time) query.  Lane s of the warp owns float4 slot s of the packed frame row
// (include/parc_b200.h, ParcRowLayout): lane 0 = root position, lane 1 = root rotation, lanes 2..J =
// joint rotations, the next ceil(J/4) lanes = contacts.  Both key frames are fetched with one
// coalesced 16-byte load per lane; the velocity part of frame 0 with a third.  The lanes that own a
// rotation slerp it, then the same lanes run the kinematic chain with warp shuffles (fk_warp), and
// finally all 32 lanes sample the heightfield under the rotated observation template.
//
// Reference semantics: anim/motion_lib.py:80-131, :443-475, :527-538; anim/kin_char_model.py:509-541;
// envs/ig_parkour/mgdm_dm_util.py:158-179; util/terrain_util.py:113-130.
#include "parc_common.cuh"

namespace parc {

struct QueryParams {
  ParcMotionTables tb;
  const int64_t* ids;
  const float* times;       // calc_motion_frame
  const int64_t* frame_idx; // get_motion_frame
  int64_t n;
  ParcRowLayout lay;
  ParcFrameOut out;
  ParcFkOut fk;
  ParcHeightfield hf;
  ParcObsSpec obs;
  float* obs_out;
  int want_fk;
  int want_obs;
};

// anim/motion_lib.py:527-538 + :443-456.  All ops individually rounded; indices are exact.
__device__ __forceinline__ void frame_blend(const ParcClipMeta& cm, float t, int64_t& i0, int64_t& i1,
                                            float& blend, float& cycles) {
  float phase = div_rn(t, cm.length);
  cycles = floorf(phase);                         // reused by the WRAP root offset (:468-469)
  if (cm.loop_mode == PARC_LOOP_WRAP) phase = sub_rn(phase, cycles);
  // torch.clip(phase, 0, 1); a NaN time would index out of bounds in the reference -- we pin it to 0.
  phase = (phase >= 0.0f) ? fminf(phase, 1.0f) : 0.0f;
  const int n1 = cm.num_frames - 1;
  const float x = mul_rn(phase, (float)n1);
  int f0 = (int)x;                                // .long() truncation, x >= 0
  int f1 = min(f0 + 1, n1);
  blend = sub_rn(x, (float)f0);
  i0 = cm.start_idx + f0;
  i1 = cm.start_idx + f1;
}

template <bool BLEND>
__global__ void __launch_bounds__(PARC_CTA_THREADS)
motion_query_kernel(const __grid_constant__ QueryParams p, const __grid_constant__ ParcCharModel model_param) {
  __shared__ ParcCharModel sm;
  stage_model(&sm, model_param);
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int J = sm.num_bodies;
  const int D = sm.dof_size;
  const LaneBody lb = load_lane_body(sm, lane, /*lane_of_body0=*/1);
  const int max_depth = sm.max_depth;
  const int row_f4 = p.lay.row_floats >> 2;
  const int pose_slots = p.lay.pose_slots;
  const int contact_slot = p.lay.contact_slot;
  const int vel_slot = p.lay.vel_slot;
  const int vel_slots = p.lay.vel_slots;
  const float4* __restrict__ rows = reinterpret_cast<const float4*>(p.tb.rows);

  const int64_t warp0 = (int64_t)blockIdx.x * PARC_WARPS_PER_CTA + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * PARC_WARPS_PER_CTA;

  for (int64_t q = warp0; q < p.n; q += nwarps) {
    // ---- clip metadata + frame indices (warp-uniform; every lane computes the same values) ----
    int64_t id = __ldg(p.ids + q);
    if (id < 0 || id >= p.tb.num_clips) id = 0;   // reference would raise an index error
    const int4* cmp = reinterpret_cast<const int4*>(p.tb.clips + id);
    const int4 c0 = __ldg(cmp);
    const int4 c1 = __ldg(cmp + 1);
    ParcClipMeta cm;
    cm.num_frames = c0.x;
    cm.loop_mode = c0.y;
    cm.start_idx = (int64_t)(((uint64_t)(uint32_t)c0.w << 32) | (uint32_t)c0.z);
    cm.length = __int_as_float(c1.x);
    cm.root_pos_delta[0] = __int_as_float(c1.y);
    cm.root_pos_delta[1] = __int_as_float(c1.z);
    cm.root_pos_delta[2] = __int_as_float(c1.w);

    int64_t i0, i1;
    float blend = 0.0f, cycles = 0.0f;
    if (BLEND) {
      frame_blend(cm, __ldg(p.times + q), i0, i1, blend, cycles);
    } else {
      i0 = i1 = cm.start_idx + __ldg(p.frame_idx + q);
    }

    // ---- gather: one float4 per lane per key frame, plus the velocity slots of frame 0 ----
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* r0 = rows + i0 * row_f4;
    const float4 A = lane < pose_slots ? __ldg(r0 + lane) : zero4;
    float4 B = A;
    if (BLEND) {
      const float4* r1 = rows + i1 * row_f4;
      B = lane < pose_slots ? __ldg(r1 + lane) : zero4;
    }
    const float4 V = lane < vel_slots ? __ldg(r0 + vel_slot + lane) : zero4;

    // ---- blend by role ----
    float4 R = A;  // result of this lane's slot
    if (BLEND) {
      if (lane == 0) {
        R.x = lerp_rn(A.x, B.x, blend);
        R.y = lerp_rn(A.y, B.y, blend);
        R.z = lerp_rn(A.z, B.z, blend);
        if (cm.loop_mode == PARC_LOOP_WRAP) {      // anim/motion_lib.py:458-475
          R.x = add_rn(R.x, mul_rn(cycles, cm.root_pos_delta[0]));
          R.y = add_rn(R.y, mul_rn(cycles, cm.root_pos_delta[1]));
          R.z = add_rn(R.z, mul_rn(cycles, cm.root_pos_delta[2]));
        }
      } else if (lane <= J) {
        R = slerp(A, B, blend);
      } else if (lane < pose_slots) {
        R.x = lerp_rn(A.x, B.x, blend);
        R.y = lerp_rn(A.y, B.y, blend);
        R.z = lerp_rn(A.z, B.z, blend);
        R.w = lerp_rn(A.w, B.w, blend);
      }
    }

    // ---- frame outputs ----
    if (lane == 0) {
      if (p.out.root_pos) {
        float* o = p.out.root_pos + q * 3;
        o[0] = R.x; o[1] = R.y; o[2] = R.z;
      }
      if (p.out.frame_idx0) p.out.frame_idx0[q] = i0;
      if (p.out.frame_idx1) p.out.frame_idx1[q] = i1;
      if (p.out.blend) p.out.blend[q] = blend;
    } else if (lane == 1) {
      if (p.out.root_rot) reinterpret_cast<float4*>(p.out.root_rot)[q] = R;
    } else if (lane <= J) {
      if (p.out.joint_rot) reinterpret_cast<float4*>(p.out.joint_rot)[q * (J - 1) + (lane - 2)] = R;
    } else if (lane < pose_slots) {
      if (p.out.contacts) {
        const int k = (lane - contact_slot) * 4;
        float* o = p.out.contacts + q * J + k;
        if (k + 0 < J) o[0] = R.x;
        if (k + 1 < J) o[1] = R.y;
        if (k + 2 < J) o[2] = R.z;
        if (k + 3 < J) o[3] = R.w;
      }
    }
    if (lane < vel_slots) {
      const float v[4] = {V.x, V.y, V.z, V.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int k = lane * 4 + c;
        if (k < 3) {
          if (p.out.root_vel) p.out.root_vel[q * 3 + k] = v[c];
        } else if (k < 6) {
          if (p.out.root_ang_vel) p.out.root_ang_vel[q * 3 + (k - 3)] = v[c];
        } else if (k - 6 < D) {
          if (p.out.dof_vel) p.out.dof_vel[q * D + (k - 6)] = v[c];
        }
      }
    }

    if (!p.want_fk && !p.want_obs) continue;

    // root position lives in lane 0, root rotation in lane 1 (= body 0's lane)
    const float3 rp = shfl3(make_float3(R.x, R.y, R.z), 0);

    // ---- heightmap observation: issue the gathers before the FK math so they overlap ----
    if (p.want_obs) {
      const float4 rr = shfl4(R, 1);
      const float heading = calc_heading(rr);
      float sn, cs;
      sn = sinf(heading);
      cs = cosf(heading);
      const float2* __restrict__ tmpl = reinterpret_cast<const float2*>(p.obs.tmpl_xy);
      float* __restrict__ o = p.obs_out + q * p.obs.num_points;
      const int P = p.obs.num_points;
#pragma unroll 4
      for (int k = lane; k < P; k += 32) {
        const float2 w = rotate_offset_2d(__ldg(tmpl + k), cs, sn, rp.x, rp.y);
        float z = hf_lookup(p.hf, w.x, w.y);
        if (p.obs.relative) z = fminf(fmaxf(sub_rn(z, rp.z), p.obs.min_h), p.obs.max_h);
        o[k] = z;
      }
    }

    // ---- forward kinematics down the tree ----
    if (p.want_fk) {
      float3 pos = rp;
      float4 rot = R;
      fk_warp(lb, max_depth, pos, rot);
      if (lb.body >= 0) {
        if (p.fk.body_pos) {
          float* o = p.fk.body_pos + (q * J + lb.body) * 3;
          o[0] = pos.x; o[1] = pos.y; o[2] = pos.z;
        }
        if (p.fk.body_rot) reinterpret_cast<float4*>(p.fk.body_rot)[q * J + lb.body] = rot;
      }
    }
  }
}

// ---- a1: pack the reference's separate tables into rows ---------------------------------------
struct PackParams {
  const float *root_pos, *root_rot, *joint_rot, *contacts, *root_vel, *root_ang_vel, *dof_vel;
  int64_t total;
  int J, D;
  ParcRowLayout lay;
  float* rows;
};

__global__ void __launch_bounds__(256) pack_frames_kernel(const __grid_constant__ PackParams p) {
  const int rf = p.lay.row_floats;
  const int64_t total_floats = p.total * rf;
  const int pose_floats = p.lay.pose_slots * 4;
  const int contact_f = p.lay.contact_slot * 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_floats;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = i / rf;
    const int k = (int)(i - f * rf);
    float v = 0.0f;
    if (k < 3) v = p.root_pos[f * 3 + k];
    else if (k < 4) v = 0.0f;
    else if (k < 8) v = p.root_rot[f * 4 + (k - 4)];
    else if (k < contact_f) v = p.joint_rot[f * (p.J - 1) * 4 + (k - 8)];
    else if (k < pose_floats) {
      const int c = k - contact_f;
      v = (c < p.J && p.contacts) ? p.contacts[f * p.J + c] : 0.0f;
    } else {
      const int c = k - pose_floats;
      if (c < 3) v = p.root_vel[f * 3 + c];
      else if (c < 6) v = p.root_ang_vel[f * 3 + (c - 3)];
      else if (c - 6 < p.D) v = p.dof_vel[f * p.D + (c - 6)];
    }
    p.rows[i] = v;
  }
}

static int query_grid(int64_t n) {
  // one warp per query; cap the grid at a few resident waves and let the warps stride
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = (n + PARC_WARPS_PER_CTA - 1) / PARC_WARPS_PER_CTA;
  const int64_t cap = (int64_t)sms * 8;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace parc

using namespace parc;

extern "C" int parc_validate_model(const ParcCharModel* m) {
  if (!m) return PARC_E_NULL;
  const int J = m->num_bodies;
  if (J < 1 || J > PARC_MAX_BODIES) return PARC_E_MODEL;
  if (m->dof_size < 0 || m->dof_size > PARC_MAX_DOF) return PARC_E_MODEL;
  if (m->parent[0] != -1 || m->depth[0] != 0) return PARC_E_MODEL;
  int maxd = 0, dof = 0;
  for (int b = 0; b < J; ++b) {
    if (b > 0) {
      if (m->parent[b] < 0 || m->parent[b] >= b) return PARC_E_MODEL;
      if (m->depth[b] != m->depth[m->parent[b]] + 1) return PARC_E_MODEL;
    }
    if (m->depth[b] > maxd) maxd = m->depth[b];
    const int jt = m->joint_type[b];
    const int dd = jt == PARC_JOINT_HINGE ? 1 : (jt == PARC_JOINT_SPHERICAL ? 3 : 0);
    if (jt < 0 || jt > 3) return PARC_E_MODEL;
    if (dd > 0 && m->dof_idx[b] != dof) return PARC_E_MODEL;
    dof += dd;
  }
  if (dof != m->dof_size || maxd != m->max_depth) return PARC_E_MODEL;
  return PARC_OK;
}

extern "C" int parc_row_layout(const ParcCharModel* m, ParcRowLayout* out) {
  if (!m || !out) return PARC_E_NULL;
  const int rc = parc_validate_model(m);
  if (rc) return rc;
  const int J = m->num_bodies;
  out->contact_slot = J + 1;
  out->pose_slots = J + 1 + (J + 3) / 4;
  out->vel_slot = out->pose_slots;
  out->vel_slots = (6 + m->dof_size + 3) / 4;
  const int floats = (out->pose_slots + out->vel_slots) * 4;
  out->row_floats = (floats + 7) / 8 * 8;
  out->reserved[0] = out->reserved[1] = out->reserved[2] = 0;
  if (out->pose_slots > 32 || out->vel_slots > 32) return PARC_E_MODEL;
  return PARC_OK;
}

extern "C" int parc_pack_frames(const float* root_pos, const float* root_rot, const float* joint_rot,
                                const float* contacts, const float* root_vel, const float* root_ang_vel,
                                const float* dof_vel, int64_t total_frames, const ParcCharModel* model,
                                float* rows_out, void* stream) {
  if (!root_pos || !root_rot || !joint_rot || !root_vel || !root_ang_vel || !dof_vel || !rows_out || !model)
    return PARC_E_NULL;
  if (total_frames < 0) return PARC_E_SIZE;
  PackParams p;
  int rc = parc_row_layout(model, &p.lay);
  if (rc) return rc;
  if (total_frames == 0) return PARC_OK;
  p.root_pos = root_pos; p.root_rot = root_rot; p.joint_rot = joint_rot; p.contacts = contacts;
  p.root_vel = root_vel; p.root_ang_vel = root_ang_vel; p.dof_vel = dof_vel;
  p.total = total_frames; p.J = model->num_bodies; p.D = model->dof_size; p.rows = rows_out;
  const int64_t total_floats = total_frames * p.lay.row_floats;
  int64_t blocks = (total_floats + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_frames_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch();
}

static int launch_query(bool blend, const ParcMotionTables* tables, const int64_t* ids, const float* times,
                        const int64_t* frame_idx, int64_t n, const ParcCharModel* model,
                        const ParcFrameOut* frame, const ParcFkOut* fk, const ParcHeightfield* hf,
                        const ParcObsSpec* obs, float* obs_out, void* stream) {
  if (!tables || !ids || !model) return PARC_E_NULL;
  if (blend ? !times : !frame_idx) return PARC_E_NULL;
  if (!tables->rows || !tables->clips) return PARC_E_NULL;
  if (n < 0 || tables->num_clips <= 0 || tables->total_frames <= 0) return PARC_E_SIZE;
  QueryParams p;
  int rc = parc_row_layout(model, &p.lay);
  if (rc) return rc;
  if (tables->row_floats != p.lay.row_floats) return PARC_E_LAYOUT;
  if (!aligned16(tables->rows) || !aligned16(tables->clips)) return PARC_E_ALIGN;
  p.tb = *tables;
  p.ids = ids; p.times = times; p.frame_idx = frame_idx; p.n = n;
  ParcFrameOut none = {};
  p.out = frame ? *frame : none;
  if (!aligned16(p.out.root_rot) || !aligned16(p.out.joint_rot)) return PARC_E_ALIGN;
  p.want_fk = (fk && (fk->body_pos || fk->body_rot)) ? 1 : 0;
  p.fk.body_pos = p.want_fk ? fk->body_pos : nullptr;
  p.fk.body_rot = p.want_fk ? fk->body_rot : nullptr;
  if (!aligned16(p.fk.body_rot)) return PARC_E_ALIGN;
  p.want_obs = obs_out ? 1 : 0;
  p.obs_out = obs_out;
  ParcHeightfield hf0 = {};
  ParcObsSpec obs0 = {};
  p.hf = hf0; p.obs = obs0;
  if (p.want_obs) {
    if (!hf || !obs || !hf->hf || !obs->tmpl_xy) return PARC_E_NULL;
    if (hf->dim_x <= 0 || hf->dim_y <= 0 || obs->num_points < 0) return PARC_E_SIZE;
    if ((reinterpret_cast<uintptr_t>(obs->tmpl_xy) & 7u) != 0) return PARC_E_ALIGN;
    p.hf = *hf; p.obs = *obs;
  }
  if (n == 0) return PARC_OK;
  const int grid = query_grid(n);
  if (blend)
    motion_query_kernel<true><<<grid, PARC_CTA_THREADS, 0, (cudaStream_t)stream>>>(p, *model);
  else
    motion_query_kernel<false><<<grid, PARC_CTA_THREADS, 0, (cudaStream_t)stream>>>(p, *model);
  return check_launch();
}

extern "C" int parc_motion_query(const ParcMotionTables* tables, const int64_t* motion_ids,
                                 const float* motion_times, int64_t n, const ParcCharModel* model,
                                 const ParcFrameOut* frame, const ParcFkOut* fk, const ParcHeightfield* hf,
                                 const ParcObsSpec* obs, float* obs_out, void* stream) {
  return launch_query(true, tables, motion_ids, motion_times, nullptr, n, model, frame, fk, hf, obs, obs_out,
                      stream);
}

extern "C" int parc_get_motion_frame(const ParcMotionTables* tables, const int64_t* motion_ids,
                                     const int64_t* frame_idxs, int64_t n, const ParcCharModel* model,
                                     const ParcFrameOut* frame, const ParcFkOut* fk, void* stream) {
  return launch_query(false, tables, motion_ids, nullptr, frame_idxs, n, model, frame, fk, nullptr, nullptr,
                      nullptr, stream);
}
