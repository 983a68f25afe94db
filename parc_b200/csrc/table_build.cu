// GPU loader: raw clip frames -> packed frame rows in ONE launch.
//
// Replaces the per-clip preprocessing of MotionLib._load_motions / _load_motion_frames
// (anim/motion_lib.py:137-202, :204-380): exp-map -> root quaternion and DoF -> joint quaternions
// (extract_frame_data, :405-423, with quat_pos on the joints), forward-difference root linear / angular
// velocity with the last frame repeated (:281-288), KinCharModel.compute_frame_dof_vel
// (anim/kin_char_model.py:543-581) and the interleaving into float4 rows (parc_pack_frames), for every frame of
// every clip at once.  One group of lanes per frame (two frames per warp for the humanoid), lane = body; a group walks a
// run of consecutive frames so that a frame's difference partner (the next frame, or the previous one for a clip's last
// frame) is converted only once.
//
// Values agree with the host-built tables to fp32 rounding of the transcendental calls; the finite differences
// amplify that by fps (see tests).  Callers that need tables bit-identical to the reference keep the host build.
#include "parc_common.cuh"
#include "parc_rotations.cuh"

namespace parc {

#define BUILD_WARPS 4

struct BuildParams {
  const float* frames;            // [total, frame_stride]: root_pos 3 | root exp-map 3 | DoFs D
  const float* contacts;          // [total, J] or nullptr
  const int32_t* frame_clip;      // [total]
  const int64_t* clip_start;      // [M]
  const int64_t* clip_num_frames; // [M]
  const float* clip_fps;          // [M]
  const float* clip_dof_vel_dt;   // [M]  divisor of the DoF velocities (1/fps; `fps` itself for the reference's
                                  //      motion_frames quirk, anim/motion_lib.py:178)
  int64_t total;
  int frame_stride;
  ParcRowLayout lay;
  float* rows;
};

// lane 0: root position + exp_map_to_quat(root exp-map); lane j >= 1: quat_pos(Joint.dof_to_rot).
// Deliberately NOT inlined: a frame's pose is computed once as "this frame" and once as the neighbour's difference
// partner; one shared body guarantees both get the same bits (two inlined copies may contract FMAs differently),
// which keeps "the last frame repeats the previous velocity" exact.
__device__ __noinline__ void lane_pose(const ParcCharModel& m, const float* __restrict__ fr, int lane, float3& pos,
                                          float4& rot) {
  pos = make_float3(0.f, 0.f, 0.f);
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
  if (lane >= 1 && rot.w < 0.0f) { rot.x = -rot.x; rot.y = -rot.y; rot.z = -rot.z; rot.w = -rot.w; }   // quat_pos
}

// util/torch_util.py:346-351 over :68-88: angle * axis of the w >= 0 representative; (0,0,1) * 0 below 1e-5
__device__ __forceinline__ float3 quat_to_exp_map(float4 q) {
  if (q.w < 0.0f) { q.x = -q.x; q.y = -q.y; q.z = -q.z; q.w = -q.w; }
  const float len = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z);
  if (!(len > 1e-5f)) return make_float3(0.0f, 0.0f, 0.0f);
  const float angle = 2.0f * atan2f(len, q.w);
  return make_float3(angle * (q.x / len), angle * (q.y / len), angle * (q.z / len));
}

// G lanes per frame (G = 16: two frames per warp whenever the character has <= 16 bodies), and every group walks RUNS of
// consecutive frames: frame f's difference partner f + 1 is the next frame of the run, so each pose is converted once
// per run instead of twice per frame (a run of 16 frames costs 17 conversions instead of 32).  The (earlier, later)
// poses of the previous frame stay cached in registers; a clip's last frame finds both of its poses there.
#define BUILD_RUN 16

template <int G>
__global__ void __launch_bounds__(BUILD_WARPS * 32)
build_tables_kernel(const __grid_constant__ BuildParams p, const __grid_constant__ ParcCharModel model_param) {
  __shared__ ParcCharModel sm;
  stage_model(&sm, model_param);
  __syncthreads();
  constexpr int GROUPS = 32 / G;
  const int lane = threadIdx.x & 31;
  const int l = lane & (G - 1);
  const int J = sm.num_bodies, D = sm.dof_size;
  const int rf = p.lay.row_floats;
  const int pose_f = p.lay.pose_slots * 4, contact_f = p.lay.contact_slot * 4;
  const int64_t g0 = ((int64_t)blockIdx.x * BUILD_WARPS + (threadIdx.x >> 5)) * GROUPS + lane / G;
  const int64_t ngroups = (int64_t)gridDim.x * BUILD_WARPS * GROUPS;
  const int64_t runs = (p.total + BUILD_RUN - 1) / BUILD_RUN;
  const int jt = (l >= 1 && l < J) ? sm.joint_type[l] : PARC_JOINT_FIXED;
  const bool has_dof = jt == PARC_JOINT_HINGE || jt == PARC_JOINT_SPHERICAL;
  for (int64_t r = g0; r < runs; r += ngroups) {
    const int64_t f_begin = r * BUILD_RUN;
    const int64_t f_end = f_begin + BUILD_RUN < p.total ? f_begin + BUILD_RUN : p.total;
    int64_t ia = -1, ib = -1;                       // frames whose poses are cached in (pa, ra) / (pb, rb)
    float3 pa = make_float3(0.f, 0.f, 0.f), pb = pa;
    float4 ra = make_float4(0.f, 0.f, 0.f, 1.f), rb = ra;
    // the clip record is re-read only when the run crosses into the next clip
    int64_t start = 0, n = 0;
    float fps = 0.0f, dt = 1.0f;
    for (int64_t f = f_begin; f < f_end; ++f) {
      if (f >= start + n) {
        const int c = __ldg(p.frame_clip + f);
        start = __ldg(p.clip_start + c);
        n = __ldg(p.clip_num_frames + c);
        fps = __ldg(p.clip_fps + c);
        dt = __ldg(p.clip_dof_vel_dt + c);
      }
      const bool last = (f - start) == n - 1;
      const bool pair = n >= 2;
      // (a, b) = (earlier, later) frame of the difference: (f, f+1), or (f-1, f) for a clip's last frame
      const int64_t fa = (pair && last) ? f - 1 : f;
      const int64_t fb = pair ? fa + 1 : f;
      float3 qa_p, qb_p;
      float4 qa_r, qb_r;
      if (fa == ib) { qa_p = pb; qa_r = rb; }
      else if (fa == ia) { qa_p = pa; qa_r = ra; }
      else lane_pose(sm, p.frames + fa * p.frame_stride, l, qa_p, qa_r);
      if (fb == fa) { qb_p = qa_p; qb_r = qa_r; }
      else if (fb == ib) { qb_p = pb; qb_r = rb; }
      else lane_pose(sm, p.frames + fb * p.frame_stride, l, qb_p, qb_r);
      ia = fa; ib = fb; pa = qa_p; ra = qa_r; pb = qb_p; rb = qb_r;
      const bool mine_b = pair && last;
      const float3 pos = mine_b ? pb : pa;
      const float4 rot = mine_b ? rb : ra;
      // One difference-quaternion / exp-map chain for every lane: the root takes quat_diff(q0, q1) = q1 * conj(q0)
      // (angular velocity, anim/motion_lib.py:281-288), a joint takes quat_normalize(quat_pos(conj(q0) * q1))
      // (compute_frame_dof_vel, anim/kin_char_model.py:543-581).
      const float4 ca = quat_conj(ra);
      float4 d = quat_mul(l == 0 ? rb : ca, l == 0 ? ca : rb);
      if (l != 0) {
        float nrm;
        if (d.w < 0.0f) { d.x = -d.x; d.y = -d.y; d.z = -d.z; d.w = -d.w; }
        d = normalize4(d, nrm);
      }
      float3 e = make_float3(0.f, 0.f, 0.f);
      if (pair) e = quat_to_exp_map(d);
      float* __restrict__ row = p.rows + f * rf;
      if (l == 0) {
        row[0] = pos.x; row[1] = pos.y; row[2] = pos.z; row[3] = 0.0f;
        reinterpret_cast<float4*>(row)[1] = rot;
        float3 v = make_float3(0.f, 0.f, 0.f);
        if (pair)
          v = make_float3(mul_rn(fps, sub_rn(pb.x, pa.x)), mul_rn(fps, sub_rn(pb.y, pa.y)), mul_rn(fps, sub_rn(pb.z, pa.z)));
        reinterpret_cast<float4*>(row + pose_f)[0] = make_float4(v.x, v.y, v.z, 0.0f);
        reinterpret_cast<float4*>(row + pose_f)[1] = make_float4(fps * e.x, fps * e.y, fps * e.z, 0.0f);
      } else if (l < J) {
        reinterpret_cast<float4*>(row)[1 + l] = rot;
        if (has_dof) {
          const float3 v = make_float3(e.x / dt, e.y / dt, e.z / dt);
          float* __restrict__ o = row + pose_f + 8 + sm.dof_idx[l];
          if (jt == PARC_JOINT_HINGE) {
            o[0] = sm.joint_axis[l][0] * v.x + sm.joint_axis[l][1] * v.y + sm.joint_axis[l][2] * v.z;
          } else {
            o[0] = v.x; o[1] = v.y; o[2] = v.z;
          }
        }
      }
      // contact flags (+ zero padding of the contact slots), zero padding after the DoF velocities
      for (int k = l; k < pose_f - contact_f; k += G)
        row[contact_f + k] = (k < J && p.contacts) ? __ldg(p.contacts + f * J + k) : 0.0f;
      for (int k = pose_f + 8 + D + l; k < rf; k += G) row[k] = 0.0f;
    }
  }
}

}  // namespace parc

using namespace parc;

extern "C" int parc_build_tables(const float* frames, int64_t total_frames, int32_t frame_stride,
                                 const float* contacts, const int32_t* frame_clip, const int64_t* clip_start,
                                 const int64_t* clip_num_frames, const float* clip_fps,
                                 const float* clip_dof_vel_dt, int64_t num_clips, const ParcCharModel* model,
                                 float* rows_out, void* stream) {
  if (!model) return PARC_E_NULL;
  if (total_frames < 0 || num_clips < 0) return PARC_E_SIZE;
  BuildParams p;
  int rc = parc_row_layout(model, &p.lay);
  if (rc) return rc;
  if (frame_stride < 6 + model->dof_size) return PARC_E_SIZE;
  if (total_frames == 0) return PARC_OK;
  if (num_clips == 0) return PARC_E_SIZE;
  if (!frames || !frame_clip || !clip_start || !clip_num_frames || !clip_fps || !clip_dof_vel_dt || !rows_out)
    return PARC_E_NULL;
  if (!aligned16(rows_out)) return PARC_E_ALIGN;
  p.frames = frames; p.contacts = contacts; p.frame_clip = frame_clip; p.clip_start = clip_start;
  p.clip_num_frames = clip_num_frames; p.clip_fps = clip_fps; p.clip_dof_vel_dt = clip_dof_vel_dt;
  p.total = total_frames; p.frame_stride = frame_stride; p.rows = rows_out;
  const bool half = model->num_bodies <= 16;
  const int64_t runs = (total_frames + BUILD_RUN - 1) / BUILD_RUN;
  const int64_t groups_per_cta = BUILD_WARPS * (half ? 2 : 1);
  // at most 16 CTAs per SM, the groups stride over the runs: measured faster than one run per group (0.32 vs 0.36 ms
  // for 2048 x 265 frames) -- every CTA stages the 1.4 KB model before its first run
  int64_t ctas = (runs + groups_per_cta - 1) / groups_per_cta;
  if (ctas > 148 * 16) ctas = 148 * 16;
  if (half) build_tables_kernel<16><<<(int)ctas, BUILD_WARPS * 32, 0, (cudaStream_t)stream>>>(p, *model);
  else build_tables_kernel<32><<<(int)ctas, BUILD_WARPS * 32, 0, (cudaStream_t)stream>>>(p, *model);
  return check_launch();
}
