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
#include <cstring>
#include "parc_common.cuh"
#include "parc_rotations.cuh"

namespace parc {

struct QueryParams {
  ParcMotionTables tb;
  const int64_t* ids;
  const float* times;       // calc_motion_frame
  const int64_t* frame_idx; // get_motion_frame
  const float* offsets;     // [num_steps] time offsets (tracker step form) or nullptr
  const float* xy_offset;   // [entries,2] added to root xy after the query (dm_env.py:604-615) or nullptr
  int num_steps;            // queries per (id, time) entry; query q = entry * num_steps + step
  int64_t n;                // total queries = entries * num_steps
  int64_t entries;
  ParcRowLayout lay;
  ParcFrameOut out;
  ParcFkOut fk;
  ParcHeightfield hf;
  ParcObsSpec obs;
  float* obs_out;
  int want_fk;
  int want_obs;
  int32_t* err;             // device error bits (PARC_QUERY_ERR_*) or nullptr
  int fast_heading;         // 0: heading -> atan2f -> sincosf as the reference; 1: cos/sin straight from the rotated x axis
  int pdl_early;            // 1: ids / times are safe to read before the previous kernel of the stream has finished
  ParcTarObsSpec tar;       // fused compute_tar_obs for steps >= 1 of the tracker-step form (TAROBS instantiations)
};

// Programmatic dependent launch (sm_90+): `launch_dependents` lets the next kernel of the stream start its prologue
// on SMs this grid no longer fills; `wait` blocks until the previous kernel has completed and flushed.  Both are
// no-ops for a launch without the programmatic-serialisation attribute.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// 1-D bulk copy global -> shared through the TMA unit (cp.async.bulk), completion counted on an mbarrier.  Used by the
// TMA_TMPL instantiation to stage the observation template with one instruction from one thread instead of 7 loads +
// 7 shared stores from each of the CTA's 64 threads.  dst / src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra WAIT_%=;\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// Work item -> (entry, step).  In the tracker-step form the items are ordered so that all step-0 queries -- the only
// ones that sweep the observation template -- come first and share warps with each other: a warp then either sweeps
// for both of its characters or for neither, instead of idling one half-warp through the other's sweep.
__device__ __forceinline__ void decode_item(int64_t idx, int64_t entries, int S, int64_t& entry, int& step) {
  entry = idx;
  step = 0;
  if (S > 1 && idx >= entries) {
    const int64_t r = idx - entries;
    entry = r / (S - 1);
    step = 1 + (int)(r - entry * (S - 1));
  }
}

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

// heightfield gathers kept in flight per lane: INFLIGHT * G covers the 441-point ray template in one sweep
#define PARC_TMPL_SMEM_MAX 2048  // observation templates up to this many points are staged in shared memory

__device__ __forceinline__ float shfl_g(float v, int src, int width) {
  return __shfl_sync(PARC_FULL_MASK, v, src, width);
}

// Per-query constants of the observation sweep (packed f32x2 operands).
struct ObsCtx {
  float2 cc, ss, off, neg_min, inv2, neg_d;
  unsigned hix, hiy;
  int dim_y;
};

// World cell of template point tp: the same individually-rounded operations as the reference's
// rotate_2d_vec + root offset + (p - min) / dxdy + round + clamp, in packed FMUL2 / FADD2 / FFMA2.
__device__ __forceinline__ int obs_cell(const ObsCtx& c, float2 tp) {
  const float2 t1 = __fmul2_rn(tp, c.cc);                       // (x cos, y cos)
  const float2 t2 = __fmul2_rn(tp, c.ss);                       // (x sin, y sin)
  float2 w = make_float2(sub_rn(t1.x, t2.y), add_rn(t2.x, t1.y));
  w = __fadd2_rn(w, c.off);
  const float2 v = __fadd2_rn(w, c.neg_min);                    // p - min
  const float2 q0 = __fmul2_rn(v, c.inv2);
  const float2 r = __ffma2_rn(c.neg_d, q0, v);
  const float2 qd = __ffma2_rn(r, c.inv2, q0);                  // == (p - min) / d, IEEE (see GridAxis)
  // clamp(rint(q), 0, dim-1): one saturating round-to-nearest-even conversion to u32 (negatives and NaN
  // -> 0, as the reference's clamp gives; >= 2^32 -> UINT_MAX), then an unsigned min with dim-1.
  // Not emulated: the reference's int64 wrap-around to cell 0 for q >= 2^63 (|coordinate| > 3.6e18 cells).
  const unsigned ix = min(__float2uint_rn(qd.x), c.hix);
  const unsigned iy = min(__float2uint_rn(qd.y), c.hiy);
  return (int)(ix * (unsigned)c.dim_y + iy);
}

// One GROUP of G lanes per query: G = 32 is the general layout, G = 16 packs two characters into a warp
// (the humanoid's 1 position + 15 rotation slots fill exactly 16 lanes), which halves the issue cost of
// everything except the observation sweep.  Requires J + 1 <= G.
//   group lane 0        root position (float4 slot 0)
//   group lanes 1..J    root / joint rotations (slots 1..J), later the bodies of the FK chain
//   contacts            slots J+1.. are fetched in a second pass by the first lanes of the group
//   velocities          slots vel_slot.. by the first vel_slots lanes (looped if vel_slots > G)
// Latency structure per query: [ids, times | template + tree staging] -> clip meta -> frame rows ->
// heightfield gathers; the gathers of the first INFLIGHT * G template points are all issued
// BEFORE the FK chain and consumed after it, so that round trip hides behind the FK math.
#define QUERY_WARPS_PER_CTA 2      // small CTAs: 4096 queries -> 1024 CTAs -> 6.9 per SM (1% imbalance)
#define QUERY_CTA_THREADS (QUERY_WARPS_PER_CTA * 32)

// MINB = resident CTAs per SM the register allocation must allow: 8 (<= 128 registers, whole template
// sweep in flight) for launches that fit one wave and are latency-bound; 16 (64 registers, a quarter of the
// sweep = 7 gathers in flight) for large launches, which are issue-bound and want more warps per scheduler.
// (Middle points -- whole sweep at <= 102 registers, half a sweep at <= 85 -- measured slower: profiles/README.md.)
// DEFER: every store of the first work item waits until the forward kinematics are computed and the observation
// gathers are in flight -- with an early-input PDL launch the whole read AND compute side then overlaps the previous
// kernel of the stream, and only the stores are ordered behind it (the deferred values stay in registers, so this is
// for the <= 128-register one-wave variant).
// TMA_TMPL: the observation template is staged by one cp.async.bulk (experiment behind variant 6).
// TAROBS (tracker-step form only): queries of steps >= 1 are the future targets; their observation relative to the
// SIMULATED character's root (compute_tar_obs, envs/ig_parkour/mgdm_dm_util.py:462-518) is written from the registers
// that hold the slerped rotations and the FK result -- no separate launch, no re-read of the target frames.
// MEASURED SLOWER than the separate parc_tar_obs launch in the tracker's step (58.9-60.9 us against 55.1 us per 4096-env
// step, profiles/r2_bench_tracker_step*.json): the stand-alone kernel runs beside parc_sim_step on a parallel branch
// and takes the heading frame once per env, here it sits on the critical path and is taken once per (env, step).
// Kept as an opt-in (TrackerStep(fuse_tar_obs=True)); the default is the separate launch.
template <bool BLEND, int G, int INFLIGHT, bool RELATIVE, int MINB, bool STEPFORM, bool DEFER = false,
          bool TMA_TMPL = false, bool TAROBS = false>
__global__ void __launch_bounds__(QUERY_CTA_THREADS, MINB)
motion_query_kernel(const __grid_constant__ QueryParams p) {
  __shared__ TreeSmem sm;
  __shared__ uint64_t s_bar;
  extern __shared__ __align__(16) float2 s_tmpl[];
  constexpr int GROUPS = 32 / G;
  const int lane = threadIdx.x & 31;
  const int l = lane & (G - 1);
  const int grp = lane / G;
  const int64_t first = ((int64_t)blockIdx.x * QUERY_WARPS_PER_CTA + (threadIdx.x >> 5)) * GROUPS;
  const int64_t stride = (int64_t)gridDim.x * QUERY_WARPS_PER_CTA * GROUPS;
  const int P = p.obs.num_points;
  const bool tmpl_in_smem = p.want_obs && P <= PARC_TMPL_SMEM_MAX;
  // shared copy is padded up to a whole sweep with copies of point 0, so the sweep never bounds-checks
  const int P_pad = (P + G * INFLIGHT - 1) / (G * INFLIGHT) * (G * INFLIGHT);

  // Prologue: everything that does not depend on anything else is requested up front so the cold misses
  // overlap -- this group's first id / time, the observation template, the kinematic tree.
  int64_t id_pre = 0;
  float t_pre = 0.0f;
  int64_t f_pre = 0;
  griddep_launch_dependents();
  if (!p.pdl_early) griddep_wait();     // the inputs may come from the previous kernel of the stream
  {
    const int64_t q0 = first + grp < p.n ? first + grp : p.n - 1;
    int64_t e0 = p.num_steps > 1 ? q0 / p.num_steps : q0;
    int s0 = 0;
    if (STEPFORM) decode_item(q0, p.entries, p.num_steps, e0, s0);
    id_pre = __ldg(p.ids + e0);
    if (BLEND) t_pre = __ldg(p.times + e0); else f_pre = __ldg(p.frame_idx + e0);
  }
  if (TMA_TMPL && tmpl_in_smem) {
    // whole 16-byte chunks of the template by one bulk copy; the odd tail and the padding up to a whole sweep
    // (copies of point 0) by plain loads
    const float2* __restrict__ g = reinterpret_cast<const float2*>(p.obs.tmpl_xy);
    const int bulk_pts = P & ~1;
    if (threadIdx.x == 0) {
      mbar_init(&s_bar, 1);
      bulk_g2s(s_tmpl, g, (uint32_t)bulk_pts * 8u, &s_bar);
    }
    for (int i = bulk_pts + threadIdx.x; i < P_pad; i += QUERY_CTA_THREADS) s_tmpl[i] = __ldg(g + (i < P ? i : 0));
  } else if (tmpl_in_smem) {
    // all of a thread's template loads are issued before the first shared-memory store, so the cold misses
    // overlap instead of queueing one round trip per element
    const float2* __restrict__ g = reinterpret_cast<const float2*>(p.obs.tmpl_xy);
    for (int i0 = threadIdx.x; i0 < P_pad; i0 += QUERY_CTA_THREADS * 8) {
      float2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * QUERY_CTA_THREADS;
        v[u] = __ldg(g + (i < P ? i : 0));
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * QUERY_CTA_THREADS;
        if (i < P_pad) s_tmpl[i] = v[u];
      }
    }
  }
  stage_tree_global(&sm, p.tb.tree);
  __syncthreads();
  if (TMA_TMPL && tmpl_in_smem) mbar_wait(&s_bar, 0);

  const int J = sm.num_bodies;
  const int D = sm.dof_size;
  const LaneBody lb = load_lane_body(sm, l, /*lane_of_body0=*/1);
  const int max_depth = sm.max_depth;
  int key_slot = -1;                                   // which key body this lane's body is (target observation)
  if (TAROBS && lb.body >= 0)
    for (int k = 0; k < p.tar.num_keys; ++k)
      if (__ldg(p.tar.key_body_ids + k) == lb.body) key_slot = k;
  const int tar_w = 9 + 6 * (J - 1) + 3 * p.tar.num_keys;
  const int row_f4 = p.lay.row_floats >> 2;
  const float4* __restrict__ rows = reinterpret_cast<const float4*>(p.tb.rows);
  const float2* __restrict__ tmpl = tmpl_in_smem ? s_tmpl : reinterpret_cast<const float2*>(p.obs.tmpl_xy);
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int64_t base = first; base < p.n; base += stride) {
    const int64_t qq = base + grp;
    const bool active = qq < p.n;
    // ---- clip metadata + frame indices (uniform within the group) ----
    // entry (env) and step of this work item; its output row is q = entry * num_steps + step
    // (the plain instantiation keeps the natural order q = item: its register allocation is the tuned one)
    const int64_t item = active ? qq : p.n - 1;
    int64_t entry = p.num_steps > 1 ? item / p.num_steps : item;
    int step = (int)(item - entry * p.num_steps);
    if (STEPFORM) decode_item(item, p.entries, p.num_steps, entry, step);
    const int64_t q = STEPFORM ? entry * p.num_steps + step : item;
    int64_t id = id_pre;
    float t_q = t_pre;
    int64_t f_q = f_pre;
    if (base != first) {
      id = __ldg(p.ids + entry);
      if (BLEND) t_q = __ldg(p.times + entry); else f_q = __ldg(p.frame_idx + entry);
    }
    // motion_times + timestep * tar_obs_steps (envs/ig_parkour/mgdm_dm_util.py:289-291): one fp32 add
    if (BLEND && p.offsets) t_q = add_rn(t_q, __ldg(p.offsets + step));
    // where the env's motion sits on the shared terrain (_move_to_motion_terrain, dm_env.py:604-615): root lane only
    float2 xy_off = make_float2(0.0f, 0.0f);
    if (STEPFORM && p.xy_offset && l == 0) xy_off = __ldg(reinterpret_cast<const float2*>(p.xy_offset) + entry);
    // target observation: the simulated character's root (the frame the targets are expressed in)
    const bool tar_on = TAROBS && step > 0;
    float3 cp = make_float3(0.f, 0.f, 0.f);
    float4 cq = make_float4(0.f, 0.f, 0.f, 1.f);
    if (tar_on) {
      cp = ld3(p.tar.sim_root_pos + entry * 3);
      if (!p.tar.global_obs) cq = ld4(p.tar.sim_root_rot + entry * 4);
    }
    if (id < 0 || id >= p.tb.num_clips) {         // the reference raises an IndexError / device assert here
      id = 0;
      if (p.err && l == 0) atomicOr(p.err, PARC_QUERY_ERR_CLIP_ID);
    }
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
      frame_blend(cm, t_q, i0, i1, blend, cycles);
    } else {
      if (f_q < 0 || f_q >= cm.num_frames) {      // never read outside the clip's rows
        f_q = f_q < 0 ? 0 : cm.num_frames - 1;
        if (p.err && l == 0) atomicOr(p.err, PARC_QUERY_ERR_FRAME_IDX);
      }
      i0 = i1 = cm.start_idx + f_q;
    }
    const float4* r0 = rows + i0 * row_f4;
    const float4* r1 = rows + i1 * row_f4;

    // ---- all row loads of this query are issued together ----
    const bool own = l <= J;
    const float4 A = own ? __ldg(r0 + l) : zero4;
    const float4 B = (BLEND && own) ? __ldg(r1 + l) : A;
    const int cs_ = p.lay.contact_slot + l;                  // contacts: first lanes of the group
    const bool has_c = p.out.contacts != nullptr && cs_ < p.lay.pose_slots;
    const float4 CA = has_c ? __ldg(r0 + cs_) : zero4;
    const float4 CB = (BLEND && has_c) ? __ldg(r1 + cs_) : CA;
    const bool has_v = l < p.lay.vel_slots && (p.out.root_vel || p.out.root_ang_vel || p.out.dof_vel);
    const float4 V = has_v ? __ldg(r0 + p.lay.vel_slot + l) : zero4;

    // ---- blend position (lane 0) / rotations (lanes 1..J) ----
    float4 R = A;
    if (BLEND) {
      if (l >= 1) {
        R = slerp(A, B, blend);
      } else {
        R.x = lerp_rn(A.x, B.x, blend);
        R.y = lerp_rn(A.y, B.y, blend);
        R.z = lerp_rn(A.z, B.z, blend);
        if (cm.loop_mode == PARC_LOOP_WRAP) {      // anim/motion_lib.py:458-475
          R.x = add_rn(R.x, mul_rn(cycles, cm.root_pos_delta[0]));
          R.y = add_rn(R.y, mul_rn(cycles, cm.root_pos_delta[1]));
          R.z = add_rn(R.z, mul_rn(cycles, cm.root_pos_delta[2]));
        }
      }
    }
    if (STEPFORM && p.xy_offset && l == 0) {         // pos + offset: one fp32 add per component, as there
      R.x = add_rn(R.x, xy_off.x);
      R.y = add_rn(R.y, xy_off.y);
    }
    auto store_frame = [&]() {
      if (active) {
        if (l == 0) {
          if (p.out.root_pos) {
            float* o = p.out.root_pos + q * 3;
            o[0] = R.x; o[1] = R.y; o[2] = R.z;
          }
          if (p.out.frame_idx0) p.out.frame_idx0[q] = i0;
          if (p.out.frame_idx1) p.out.frame_idx1[q] = i1;
          if (p.out.blend) p.out.blend[q] = blend;
        } else if (l == 1) {
          if (p.out.root_rot) reinterpret_cast<float4*>(p.out.root_rot)[q] = R;
        } else if (own) {
          if (p.out.joint_rot) reinterpret_cast<float4*>(p.out.joint_rot)[q * (J - 1) + (l - 2)] = R;
        }
        // contacts, lerped (anim/motion_lib.py:109)
        if (has_c) {
          float4 c = CA;
          if (BLEND) {
            c.x = lerp_rn(CA.x, CB.x, blend); c.y = lerp_rn(CA.y, CB.y, blend);
            c.z = lerp_rn(CA.z, CB.z, blend); c.w = lerp_rn(CA.w, CB.w, blend);
          }
          const int ck = l * 4;
          float* o = p.out.contacts + q * J + ck;
          o[0] = c.x;
          if (ck + 1 < J) o[1] = c.y;
          if (ck + 2 < J) o[2] = c.z;
          if (ck + 3 < J) o[3] = c.w;
        }
        // velocities of key frame 0, un-blended (anim/motion_lib.py:89-95):
        // vel slot 0 = root_vel.xyz, 1 = root_ang_vel.xyz, 2.. = dof_vel in groups of 4
        if (has_v) {
          if (l == 0) {
            if (p.out.root_vel) { float* o = p.out.root_vel + q * 3; o[0] = V.x; o[1] = V.y; o[2] = V.z; }
          } else if (l == 1) {
            if (p.out.root_ang_vel) { float* o = p.out.root_ang_vel + q * 3; o[0] = V.x; o[1] = V.y; o[2] = V.z; }
          } else if (p.out.dof_vel) {
            const int k = (l - 2) * 4;
            float* o = p.out.dof_vel + q * D + k;
            if ((D & 3) == 0) {
              *reinterpret_cast<float4*>(o) = V;
            } else {
              o[0] = V.x;
              if (k + 1 < D) o[1] = V.y;
              if (k + 2 < D) o[2] = V.z;
              if (k + 3 < D) o[3] = V.w;
            }
          }
        }
      }
    };
    // target observation, pose part (root rotation on lane 1, joint rotations on lanes 2..J): the same device
    // functions as tar_obs_kernel, fed from registers
    float4 hinv = make_float4(0.f, 0.f, 0.f, 1.f);
    if (tar_on && !p.tar.global_obs) hinv = heading_inverse_quat(cq);
    float* __restrict__ tar_o = nullptr;
    if (tar_on) tar_o = p.tar.obs_out + entry * p.tar.out_env_stride + (int64_t)(step - 1) * tar_w;
    auto store_tar_pose = [&]() {
      if (tar_on && active && own) {
        if (l >= 2) store_tan_norm(tar_o + 9 + 6 * (l - 2), R);
        else if (l == 1) store_tan_norm(tar_o + 3, p.tar.global_obs ? R : quat_mul(hinv, R));
      }
    };
    // early-input PDL launches ran everything above -- immutable tables and caller-guaranteed inputs only -- while the
    // previous kernel of the stream was still draining; nothing may be written before it has finished
    if (!DEFER) {
      if (p.pdl_early && base == first) griddep_wait();
      store_frame();
      store_tar_pose();
    }

    if (!p.want_fk && !p.want_obs) {
      if (DEFER) {
        if (p.pdl_early && base == first) griddep_wait();
        store_frame();
        store_tar_pose();
      }
      continue;
    }

    // root position lives in group lane 0, root rotation in group lane 1 (= body 0's lane)
    const float3 rp = make_float3(shfl_g(R.x, 0, G), shfl_g(R.y, 0, G), shfl_g(R.z, 0, G));
    // target observation, root offset in the character's heading frame (every lane: the key bodies add it below)
    float3 tar_po = make_float3(rp.x - cp.x, rp.y - cp.y, rp.z - cp.z);
    if (tar_on && !p.tar.global_obs) tar_po = quat_rotate(hinv, tar_po);

    // ---- heightmap observation, part 1: cells of the first G * INFLIGHT points, gathers issued ----
    ObsCtx oc;
    float z[INFLIGHT];
    const float* __restrict__ hfp = p.hf.hf;
    // observations belong to step 0 of an entry (the current frame); shuffles are done by the whole warp
    const bool grp_obs = p.want_obs && step == 0;
    float4 rr = make_float4(0.f, 0.f, 0.f, 1.f);
    if (p.want_obs) rr = make_float4(shfl_g(R.x, 1, G), shfl_g(R.y, 1, G), shfl_g(R.z, 1, G), shfl_g(R.w, 1, G));
    if (grp_obs) {
      const float3 dir = quat_rotate(rr, make_float3(1.0f, 0.0f, 0.0f));
      float sn, cs;
      if (!p.fast_heading) {
        // the reference's chain: heading = atan2(d.y, d.x) (util/torch_util.py:470-479), then cos / sin of it
        // (rotate_2d_vec, :619-631)
        sincos_reduced(atan2f(dir.y, dir.x), sn, cs);
      } else {
        // PARC_QUERY_FAST_HEADING: cos / sin taken directly from the rotated x axis d; within ~2 ulp of the chain
        const float n2 = dir.x * dir.x + dir.y * dir.y;
        sn = 0.0f; cs = (dir.x < 0.0f) ? -1.0f : 1.0f;           // atan2(0, +-0)
        if (n2 > 0.0f) {
          const float rn = rsqrtf(n2);
          cs = dir.x * rn;
          sn = dir.y * rn;
        }
      }
      const GridAxis gx = make_grid_axis(p.hf.min_x, p.hf.dx, p.hf.dim_x);
      const GridAxis gy = make_grid_axis(p.hf.min_y, p.hf.dy, p.hf.dim_y);
      oc.cc = make_float2(cs, cs); oc.ss = make_float2(sn, sn); oc.off = make_float2(rp.x, rp.y);
      oc.neg_min = make_float2(-gx.mn, -gy.mn); oc.inv2 = make_float2(gx.inv, gy.inv);
      oc.neg_d = make_float2(-gx.d, -gy.d);
      oc.hix = (unsigned)(p.hf.dim_x - 1); oc.hiy = (unsigned)(p.hf.dim_y - 1); oc.dim_y = p.hf.dim_y;
      if (tmpl_in_smem) {
#pragma unroll
        for (int u = 0; u < INFLIGHT; ++u) z[u] = __ldg(hfp + obs_cell(oc, s_tmpl[l + G * u]));
      } else {
#pragma unroll
        for (int u = 0; u < INFLIGHT; ++u) {
          const int k = l + G * u;
          z[u] = __ldg(hfp + obs_cell(oc, tmpl[k < P ? k : l]));
        }
      }
    }

    // ---- forward kinematics down the tree (shuffles stay inside the group) ----
    float3 pos = rp;
    float4 rot = R;
    if (p.want_fk) {
      float4 local = rot;
      if (lb.body > 0) local = quat_mul_plain(lb.lr, rot);
#pragma unroll 1
      for (int d = 1; d <= max_depth; ++d) {
        const float3 pp = make_float3(shfl_g(pos.x, lb.parent_lane, G), shfl_g(pos.y, lb.parent_lane, G),
                                      shfl_g(pos.z, lb.parent_lane, G));
        const float4 pr = make_float4(shfl_g(rot.x, lb.parent_lane, G), shfl_g(rot.y, lb.parent_lane, G),
                                      shfl_g(rot.z, lb.parent_lane, G), shfl_g(rot.w, lb.parent_lane, G));
        if (lb.depth == d) {
          const float3 wt = quat_rotate(pr, lb.lt);
          pos = make_float3(pp.x + wt.x, pp.y + wt.y, pp.z + wt.z);
          rot = quat_mul_plain(pr, local);
        }
      }
    }
    if (DEFER) {
      if (p.pdl_early && base == first) griddep_wait();
      store_frame();
      store_tar_pose();
    }
    if (tar_on && active) {
      if (l == 0) st3(tar_o, make_float3(tar_po.x, tar_po.y, p.tar.global_tar_root_h ? rp.z : tar_po.z));
      if (key_slot >= 0) {                        // key body relative to the target root, then + the root offset
        float3 kp = make_float3(pos.x - rp.x, pos.y - rp.y, pos.z - rp.z);
        if (!p.tar.global_obs) {
          kp = quat_rotate(hinv, kp);
          kp.x += tar_po.x; kp.y += tar_po.y; kp.z += tar_po.z;
        }
        st3(tar_o + 9 + 6 * (J - 1) + 3 * key_slot, kp);
      }
    }
    if (p.want_fk && active && lb.body >= 0) {
      if (p.fk.body_pos) {
        float* o = p.fk.body_pos + (q * J + lb.body) * 3;
        o[0] = pos.x; o[1] = pos.y; o[2] = pos.z;
      }
      if (p.fk.body_rot) reinterpret_cast<float4*>(p.fk.body_rot)[q * J + lb.body] = rot;
    }

    // ---- heightmap observation, part 2: consume the gathers; then any further points ----
    if (grp_obs && active) {
      float* __restrict__ o = p.obs_out + entry * P + l;      // lane's first output; u-th is o[G * u]
      const float root_z = rp.z, lo = p.obs.min_h, hi = p.obs.max_h;
      if (P >= G * (INFLIGHT - 1) + G) {
        // every lane is in range for the whole sweep (uniform branch): no per-element bounds predicate
#pragma unroll
        for (int u = 0; u < INFLIGHT; ++u) {
          float v = z[u];
          if (RELATIVE) v = fminf(fmaxf(sub_rn(v, root_z), lo), hi);
          o[G * u] = v;
        }
      } else if (P >= G * (INFLIGHT - 1)) {
        // only the last iteration is partial (the 441-point ray fan: 27 full iterations of 16 + 9)
#pragma unroll
        for (int u = 0; u < INFLIGHT - 1; ++u) {
          float v = z[u];
          if (RELATIVE) v = fminf(fmaxf(sub_rn(v, root_z), lo), hi);
          o[G * u] = v;
        }
        float v = z[INFLIGHT - 1];
        if (RELATIVE) v = fminf(fmaxf(sub_rn(v, root_z), lo), hi);
        if (l + G * (INFLIGHT - 1) < P) o[G * (INFLIGHT - 1)] = v;
      } else {
#pragma unroll
        for (int u = 0; u < INFLIGHT; ++u) {
          float v = z[u];
          if (RELATIVE) v = fminf(fmaxf(sub_rn(v, root_z), lo), hi);
          if (l + G * u < P) o[G * u] = v;
        }
      }
      for (int k0 = G * INFLIGHT; k0 < P; k0 += G * INFLIGHT) {
        if (tmpl_in_smem) {               // padded to whole sweeps: no bounds check on the read side
#pragma unroll
          for (int u = 0; u < INFLIGHT; ++u) z[u] = __ldg(hfp + obs_cell(oc, s_tmpl[k0 + l + G * u]));
        } else {
#pragma unroll
          for (int u = 0; u < INFLIGHT; ++u) {
            const int k = k0 + l + G * u;
            z[u] = __ldg(hfp + obs_cell(oc, tmpl[k < P ? k : l]));
          }
        }
        float* __restrict__ ok = o + k0;        // one pointer per pass: the stores below use immediate offsets
        if (k0 + G * INFLIGHT <= P) {           // a full pass: no bounds checks
#pragma unroll
          for (int u = 0; u < INFLIGHT; ++u) {
            float v = z[u];
            if (RELATIVE) v = fminf(fmaxf(sub_rn(v, root_z), lo), hi);
            ok[G * u] = v;
          }
        } else if (P - k0 >= G * (INFLIGHT - 1)) {
          // only the last sample of the pass can be out of range (the 441-point fan: 105 samples left, 7 * 16 slots)
#pragma unroll
          for (int u = 0; u < INFLIGHT - 1; ++u) {
            float v = z[u];
            if (RELATIVE) v = fminf(fmaxf(sub_rn(v, root_z), lo), hi);
            ok[G * u] = v;
          }
          float v = z[INFLIGHT - 1];
          if (RELATIVE) v = fminf(fmaxf(sub_rn(v, root_z), lo), hi);
          if (k0 + l + G * (INFLIGHT - 1) < P) ok[G * (INFLIGHT - 1)] = v;
        } else {
          const int left = P - k0 - l;          // this lane's valid samples are those with G * u < left
#pragma unroll
          for (int u = 0; u < INFLIGHT; ++u) {
            float v = z[u];
            if (RELATIVE) v = fminf(fmaxf(sub_rn(v, root_z), lo), hi);
            if (G * u < left) ok[G * u] = v;
          }
        }
      }
    }
  }
}

// grid_index_fast's packed twin used in the observation loop, one axis at a time, for the self test
__device__ __forceinline__ int grid_index_packed_form(float p, const GridAxis& a, int hi) {
  // drive obs_cell with an identity rotation and zero offset so that w == (p, p)
  ObsCtx c;
  c.cc = make_float2(1.0f, 1.0f); c.ss = make_float2(0.0f, 0.0f); c.off = make_float2(0.0f, 0.0f);
  c.neg_min = make_float2(-a.mn, -a.mn); c.inv2 = make_float2(a.inv, a.inv); c.neg_d = make_float2(-a.d, -a.d);
  c.hix = (unsigned)hi; c.hiy = 0u; c.dim_y = 1;
  // x*1 - 0*0 = x and 0*... exact; (p + 0) == p except -0 -> +0, which indexes identically
  return obs_cell(c, make_float2(p, 0.0f));
}

// Exhaustive check of grid_index_fast against the reference-form IEEE division, over every float bit
// pattern of the coordinate, for one axis description.  mismatches[0] += number of differing indices.
__global__ void __launch_bounds__(256)
selftest_grid_index_kernel(float mn, float d, int dim, unsigned long long* mismatches) {
  const GridAxis a = make_grid_axis(mn, d, dim);
  unsigned long long bad = 0;
  const uint64_t total = 1ull << 32;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const float x = __uint_as_float((uint32_t)i);
    const int ref = grid_index_1d(x, mn, d, dim);
    const float gq = div_rn(sub_rn(x, mn), d);
    const bool wraps = gq >= 9.2e18f;               // reference wraps through int64 here; sweep form clamps
    bad += (grid_index_fast(x, a) != ref) || (!wraps && grid_index_packed_form(x, a, dim - 1) != ref);
  }
  if (bad) atomicAdd(mismatches, bad);
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
      else if (c < 4) v = 0.0f;
      else if (c < 7) v = p.root_ang_vel[f * 3 + (c - 4)];
      else if (c < 8) v = 0.0f;
      else if (c - 8 < p.D) v = p.dof_vel[f * p.D + (c - 8)];
    }
    p.rows[i] = v;
  }
}

static int query_grid(int64_t n, int sms) {
  // one warp per n; cap the grid at a few resident waves and let the warps stride
  const int64_t want = (n + QUERY_WARPS_PER_CTA - 1) / QUERY_WARPS_PER_CTA;
  const int64_t cap = (int64_t)sms * 16;
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

extern "C" int parc_tree_from_model(const ParcCharModel* m, void* tree_host_out) {
  if (!m || !tree_host_out) return PARC_E_NULL;
  const int rc = parc_validate_model(m);
  if (rc) return rc;
  TreeSmem t;
  memset(&t, 0, sizeof(t));
  t.num_bodies = m->num_bodies; t.dof_size = m->dof_size; t.max_depth = m->max_depth;
  for (int b = 0; b < m->num_bodies; ++b) {
    t.parent[b] = m->parent[b];
    t.depth[b] = m->depth[b];
    for (int k = 0; k < 3; ++k) t.lt[b][k] = m->local_trans[b][k];
    for (int k = 0; k < 4; ++k) t.lr[b][k] = m->local_rot[b][k];
  }
  memcpy(tree_host_out, &t, sizeof(t));
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
  out->vel_slots = 2 + (m->dof_size + 3) / 4;
  const int floats = (out->pose_slots + out->vel_slots) * 4;
  out->row_floats = (floats + 7) / 8 * 8;
  out->reserved[0] = out->reserved[1] = out->reserved[2] = 0;
  if (m->num_bodies + 1 > 32) return PARC_E_MODEL;
  return PARC_OK;
}

extern "C" int parc_pack_frames(const float* root_pos, const float* root_rot, const float* joint_rot,
                                const float* contacts, const float* root_vel, const float* root_ang_vel,
                                const float* dof_vel, int64_t total_frames, const ParcCharModel* model,
                                float* rows_out, void* stream) {
  if (!model) return PARC_E_NULL;
  if (total_frames < 0) return PARC_E_SIZE;
  if (total_frames > 0 &&
      (!root_pos || !root_rot || !joint_rot || !root_vel || !root_ang_vel || !dof_vel || !rows_out))
    return PARC_E_NULL;
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

// Launch helper: plain <<<>>> or, for PARC_QUERY_PDL, cudaLaunchKernelEx with programmatic stream serialisation (the
// kernel may begin while the previous kernel of the stream drains; it orders itself with griddepcontrol.wait).
template <typename K>
static void launch_maybe_pdl(K kernel, int grid, int block, size_t smem, cudaStream_t st, bool pdl, const QueryParams& p) {
  if (!pdl) {
    kernel<<<grid, block, smem, st>>>(p);
    return;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3((unsigned)block, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, p);
}

extern "C" int parc_motion_query_ex(const ParcQueryArgs* a, void* stream) {
  if (!a) return PARC_E_NULL;
  const ParcMotionTables* tables = a->tables;
  const ParcCharModel* model = a->model;
  if (!tables || !model) return PARC_E_NULL;
  if (!tables->rows || !tables->clips) return PARC_E_NULL;
  const bool blend = a->frame_idxs == nullptr;
  if (!blend && a->motion_times) return PARC_E_SIZE;            // exactly one of times / frame indices
  const int num_steps = a->num_steps < 1 ? 1 : a->num_steps;
  const int64_t n_entries = a->n;
  if (n_entries < 0 || a->num_steps < 0 || tables->num_clips <= 0 || tables->total_frames <= 0) return PARC_E_SIZE;
  if (num_steps > 1 && (!a->time_offsets || !blend)) return PARC_E_NULL;
  if (!blend && a->root_xy_offset) return PARC_E_SIZE;
  if (a->variant < 0 || a->variant > 6) return PARC_E_SIZE;
  const int64_t n = n_entries * num_steps;
  if (n > 0 && (!a->motion_ids || (blend && !a->motion_times))) return PARC_E_NULL;
  QueryParams p;
  int rc = parc_row_layout(model, &p.lay);
  if (rc) return rc;
  if (tables->row_floats != p.lay.row_floats) return PARC_E_LAYOUT;
  if (!aligned16(tables->rows) || !aligned16(tables->clips)) return PARC_E_ALIGN;
  if (!tables->tree) return PARC_E_NULL;
  if (!aligned16(tables->tree)) return PARC_E_ALIGN;
  p.tb = *tables;
  p.ids = a->motion_ids; p.times = a->motion_times; p.frame_idx = a->frame_idxs; p.n = n;
  p.offsets = a->time_offsets; p.num_steps = num_steps; p.entries = n_entries;
  if ((reinterpret_cast<uintptr_t>(a->root_xy_offset) & 7u) != 0) return PARC_E_ALIGN;
  p.xy_offset = a->root_xy_offset;
  if ((reinterpret_cast<uintptr_t>(a->error_flags) & 3u) != 0) return PARC_E_ALIGN;
  p.err = a->error_flags;
  p.fast_heading = (a->flags & PARC_QUERY_FAST_HEADING) ? 1 : 0;
  const bool pdl = (a->flags & PARC_QUERY_PDL) != 0;
  p.pdl_early = (pdl && (a->flags & PARC_QUERY_PDL_EARLY_INPUTS)) ? 1 : 0;
  ParcFrameOut none = {};
  p.out = a->frame ? *a->frame : none;
  if (!aligned16(p.out.root_rot) || !aligned16(p.out.joint_rot)) return PARC_E_ALIGN;
  const ParcFkOut* fk = a->fk;
  p.want_fk = (fk && (fk->body_pos || fk->body_rot)) ? 1 : 0;
  p.fk.body_pos = p.want_fk ? fk->body_pos : nullptr;
  p.fk.body_rot = p.want_fk ? fk->body_rot : nullptr;
  if (!aligned16(p.fk.body_rot)) return PARC_E_ALIGN;
  // an empty observation template asks for nothing: the sweep would otherwise read an unstaged template
  p.want_obs = (blend && a->obs_out && !(a->obs && a->obs->num_points == 0)) ? 1 : 0;
  p.obs_out = p.want_obs ? a->obs_out : nullptr;
  ParcHeightfield hf0 = {};
  ParcObsSpec obs0 = {};
  p.hf = hf0; p.obs = obs0;
  if (p.want_obs) {
    const ParcHeightfield* hf = a->hf;
    const ParcObsSpec* obs = a->obs;
    if (!hf || !obs || !hf->hf || !obs->tmpl_xy) return PARC_E_NULL;
    if (hf->dim_x <= 0 || hf->dim_y <= 0 || obs->num_points < 0) return PARC_E_SIZE;
    if ((reinterpret_cast<uintptr_t>(obs->tmpl_xy) & 7u) != 0) return PARC_E_ALIGN;
    if ((int64_t)hf->dim_x * hf->dim_y >= (1ll << 31)) return PARC_E_SIZE;
    p.hf = *hf; p.obs = *obs;
  }
  // fused target observation (tracker-step form)
  ParcTarObsSpec tar0 = {};
  p.tar = tar0;
  const bool want_tar = a->tar_obs != nullptr;
  if (want_tar) {
    const ParcTarObsSpec* t = a->tar_obs;
    if (!blend || num_steps < 2 || !p.want_fk) return PARC_E_SIZE;      // targets are steps >= 1 and need FK
    if (t->num_keys < 0 || t->num_keys > PARC_MAX_BODIES) return PARC_E_SIZE;
    if (!t->sim_root_pos || !t->obs_out || (!t->global_obs && !t->sim_root_rot) || (t->num_keys > 0 && !t->key_body_ids))
      return PARC_E_NULL;
    if (!t->global_obs && !aligned16(t->sim_root_rot)) return PARC_E_ALIGN;
    const int64_t w = 9 + 6 * (int64_t)(model->num_bodies - 1) + 3 * (int64_t)t->num_keys;
    if (t->out_env_stride < (int64_t)(num_steps - 1) * w) return PARC_E_SIZE;
    p.tar = *t;
  }
  if (n == 0) return PARC_OK;
  if (!aligned16(p.out.dof_vel)) return PARC_E_ALIGN;
  // Group size: two characters per warp (G = 16) whenever position + rotations fit 16 lanes; it halves
  // the instruction count of everything but the observation sweep.
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const bool fits16 = model->num_bodies + 1 <= 16;
  // Regimes (measured on B200): one resident wave -> the whole 28-deep sweep in flight at <= 128 registers;
  // beyond it occupancy wins: 7 gathers in flight at 64 registers / 16 CTAs per SM beat 14 in flight at 80
  // registers / 12 CTAs (65 536 envs: 87.3 -> 85.2 us; the 7-step tracker form: 396 -> 337 us), while 4 in
  // flight (88.5 us) and 48 registers / 20 CTAs (97.7 us, spills) lose again.
  // variant: 0 = by regime; 1 = G16 / 28 in flight / 8 CTAs per SM; 2 = G16 / 7 / 16; 3 = G16 / 14 / 12; 4 = G32 / 14 / 8;
  // 5 = 1 with deferred stores -- the one-wave choice: same speed as 1 launched alone (7.9 vs 8.0 us at 4096 envs),
  // 5.4 vs 5.9 us per step in a chain of early-input PDL launches (profiles/r2_variants.json)
  int variant = a->variant;
  if (!fits16) variant = 4;
  if (variant == 0) {
    const int64_t warps16 = (n + 1) / 2;
    variant = blend ? (warps16 <= (int64_t)sms * 16 ? 5 : 2) : 3;
    // a launch so small that one character per warp still fits a single resident wave: the observation sweep is
    // then spread over twice the lanes (2048 envs: 4.9 us against 5.4 us per step)
    if (blend && p.want_obs && n <= (int64_t)sms * 16) variant = 4;
  }
  if (!blend && variant != 4) variant = 3;
  // the fused target observation exists for the two-characters-per-warp step forms 3 and 5 (the 64-register form 2
  // spills 164 bytes with it and measured slowest)
  if (want_tar) {
    if (!fits16) return PARC_E_MODEL;
    if (variant == 1 || variant == 6) variant = 5;
    if (variant == 4 || variant == 2) variant = 3;
  }
  const bool half = variant != 4;
  const int64_t warps = half ? (n + 1) / 2 : n;
  const int grid = query_grid(warps, sms);
  // 6 = 5 with the template staged by a TMA bulk copy (needs a 16-byte aligned template of >= 2 points)
  if (variant == 6 && (!p.want_obs || p.obs.num_points < 2 || p.obs.num_points > PARC_TMPL_SMEM_MAX ||
                       (reinterpret_cast<uintptr_t>(p.obs.tmpl_xy) & 15u) != 0))
    variant = 5;
  const int inflight = (variant == 1 || variant == 5 || variant == 6) ? 28 : (variant == 2 ? 7 : 14);
  size_t smem = 0;
  if (p.want_obs && p.obs.num_points <= PARC_TMPL_SMEM_MAX) {
    const int sweep = (half ? 16 : 32) * inflight;
    smem = (size_t)((p.obs.num_points + sweep - 1) / sweep * sweep) * 8;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const bool rel = p.want_obs && p.obs.relative != 0;
  // STEPFORM is a template flag: the per-entry xy offset and the step-0-first item order cost live registers that
  // spill in the 64-register large-batch variant (measured: -2..4 % at 65 536 envs when they were run-time
  // branches), so the plain one-query-per-entry call runs an instantiation without them -- whose code must stay
  // exactly the tuned one: making `step` a compile-time 0 there changed the schedule and cost 13 %.
  const bool step_form = blend && (p.num_steps > 1 || p.xy_offset);
#define PARC_LAUNCH_QUERY_D(B, GG, NF, RL, MB, DF)                                                                   \
  do {                                                                                                               \
    if (B && step_form)                                                                                              \
      launch_maybe_pdl(motion_query_kernel<B, GG, NF, RL, MB, B, DF>, grid, QUERY_CTA_THREADS, smem, st, pdl, p);    \
    else                                                                                                             \
      launch_maybe_pdl(motion_query_kernel<B, GG, NF, RL, MB, false, DF>, grid, QUERY_CTA_THREADS, smem, st, pdl, p); \
  } while (0)
#define PARC_LAUNCH_QUERY(B, GG, NF, RL, MB) PARC_LAUNCH_QUERY_D(B, GG, NF, RL, MB, false)
#define PARC_LAUNCH_QUERY_TMA(RL)                                                                                    \
  do {                                                                                                               \
    if (step_form)                                                                                                   \
      launch_maybe_pdl(motion_query_kernel<true, 16, 28, RL, 8, true, true, true>, grid, QUERY_CTA_THREADS, smem, st, pdl, p);  \
    else                                                                                                             \
      launch_maybe_pdl(motion_query_kernel<true, 16, 28, RL, 8, false, true, true>, grid, QUERY_CTA_THREADS, smem, st, pdl, p); \
  } while (0)
#define PARC_LAUNCH_QUERY_TAR(NF, RL, MB, DF)                                                                        \
  launch_maybe_pdl(motion_query_kernel<true, 16, NF, RL, MB, true, DF, false, true>, grid, QUERY_CTA_THREADS, smem, st, pdl, p)
  if (want_tar) {
    switch (variant) {
      case 3: if (rel) PARC_LAUNCH_QUERY_TAR(14, true, 12, false); else PARC_LAUNCH_QUERY_TAR(14, false, 12, false); break;
      default: if (rel) PARC_LAUNCH_QUERY_TAR(28, true, 8, true); else PARC_LAUNCH_QUERY_TAR(28, false, 8, true); break;
    }
  } else if (blend) {
    switch (variant) {
      case 1: if (rel) PARC_LAUNCH_QUERY(true, 16, 28, true, 8); else PARC_LAUNCH_QUERY(true, 16, 28, false, 8); break;
      case 2: if (rel) PARC_LAUNCH_QUERY(true, 16, 7, true, 16); else PARC_LAUNCH_QUERY(true, 16, 7, false, 16); break;
      case 3: if (rel) PARC_LAUNCH_QUERY(true, 16, 14, true, 12); else PARC_LAUNCH_QUERY(true, 16, 14, false, 12); break;
      case 5: if (rel) PARC_LAUNCH_QUERY_D(true, 16, 28, true, 8, true); else PARC_LAUNCH_QUERY_D(true, 16, 28, false, 8, true); break;
      case 6: if (rel) PARC_LAUNCH_QUERY_TMA(true); else PARC_LAUNCH_QUERY_TMA(false); break;
      default: if (rel) PARC_LAUNCH_QUERY(true, 32, 14, true, 8); else PARC_LAUNCH_QUERY(true, 32, 14, false, 8); break;
    }
  } else {
    if (half) PARC_LAUNCH_QUERY(false, 16, 14, false, 12); else PARC_LAUNCH_QUERY(false, 32, 14, false, 8);
  }
#undef PARC_LAUNCH_QUERY
#undef PARC_LAUNCH_QUERY_D
#undef PARC_LAUNCH_QUERY_TMA
#undef PARC_LAUNCH_QUERY_TAR
  return check_launch();
}

static int query_basic(const ParcMotionTables* tables, const int64_t* ids, const float* times,
                       const int64_t* frame_idx, const float* offsets, int num_steps, const float* xy_offset,
                       int64_t n_entries, const ParcCharModel* model, const ParcFrameOut* frame, const ParcFkOut* fk,
                       const ParcHeightfield* hf, const ParcObsSpec* obs, float* obs_out, void* stream) {
  ParcQueryArgs a = {};
  a.tables = tables; a.motion_ids = ids; a.motion_times = times; a.frame_idxs = frame_idx; a.n = n_entries;
  a.time_offsets = offsets; a.num_steps = num_steps; a.root_xy_offset = xy_offset; a.model = model;
  a.frame = frame; a.fk = fk; a.hf = hf; a.obs = obs; a.obs_out = obs_out;
  return parc_motion_query_ex(&a, stream);
}

extern "C" int parc_motion_query(const ParcMotionTables* tables, const int64_t* motion_ids,
                                 const float* motion_times, int64_t n, const ParcCharModel* model,
                                 const ParcFrameOut* frame, const ParcFkOut* fk, const ParcHeightfield* hf,
                                 const ParcObsSpec* obs, float* obs_out, void* stream) {
  if (n > 0 && !motion_times) return PARC_E_NULL;
  return query_basic(tables, motion_ids, motion_times, nullptr, nullptr, 1, nullptr, n, model, frame, fk, hf, obs,
                     obs_out, stream);
}

extern "C" int parc_motion_query_steps(const ParcMotionTables* tables, const int64_t* motion_ids,
                                       const float* motion_times, int64_t n, const float* time_offsets,
                                       int32_t num_steps, const float* root_xy_offset, const ParcCharModel* model,
                                       const ParcFrameOut* frame, const ParcFkOut* fk, const ParcHeightfield* hf,
                                       const ParcObsSpec* obs, float* obs_out, void* stream) {
  if (num_steps < 1) return PARC_E_SIZE;
  if (n > 0 && !motion_times) return PARC_E_NULL;
  return query_basic(tables, motion_ids, motion_times, nullptr, time_offsets, num_steps, root_xy_offset, n, model,
                     frame, fk, hf, obs, obs_out, stream);
}

extern "C" int parc_get_motion_frame(const ParcMotionTables* tables, const int64_t* motion_ids,
                                     const int64_t* frame_idxs, int64_t n, const ParcCharModel* model,
                                     const ParcFrameOut* frame, const ParcFkOut* fk, void* stream) {
  if (n > 0 && !frame_idxs) return PARC_E_NULL;
  if (!frame_idxs) {                        // n == 0: still validate the remaining arguments as a blended query of 0
    static const int64_t dummy = 0;
    frame_idxs = &dummy;
  }
  return query_basic(tables, motion_ids, nullptr, frame_idxs, nullptr, 1, nullptr, n, model, frame, fk, nullptr,
                     nullptr, nullptr, stream);
}

extern "C" int parc_selftest_grid_index(float min_coord, float cell_size, int32_t dim, uint64_t* mismatches_dev,
                                        void* stream) {
  if (!mismatches_dev) return PARC_E_NULL;
  if (dim <= 0 || !(cell_size > 0.0f)) return PARC_E_SIZE;
  selftest_grid_index_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(
      min_coord, cell_size, dim, reinterpret_cast<unsigned long long*>(mismatches_dev));
  return check_launch();
}
