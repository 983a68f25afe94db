// GPU loader: raw clip frames -> packed frame rows in ONE launch.
//
// Replaces the per-clip preprocessing of MotionLib._load_motions / _load_motion_frames
// (anim/motion_lib.py:137-202, :204-380): exp-map -> root quaternion and DoF -> joint quaternions
// (extract_frame_data, :405-423, with quat_pos on the joints), forward-difference root linear / angular
// velocity with the last frame repeated (:281-288), KinCharModel.compute_frame_dof_vel
// (anim/kin_char_model.py:543-581) and the interleaving into float4 rows (parc_pack_frames), for every frame of
// every clip at once.  One warp per frame, lane = body; each lane converts its joint for the frame and for the
// frame's difference partner (the next frame, or the previous one for a clip's last frame).
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

__global__ void __launch_bounds__(BUILD_WARPS * 32)
build_tables_kernel(const __grid_constant__ BuildParams p, const __grid_constant__ ParcCharModel model_param) {
  __shared__ ParcCharModel sm;
  stage_model(&sm, model_param);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int J = sm.num_bodies, D = sm.dof_size;
  const int rf = p.lay.row_floats;
  const int pose_f = p.lay.pose_slots * 4, contact_f = p.lay.contact_slot * 4;
  const int64_t warp0 = (int64_t)blockIdx.x * BUILD_WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * BUILD_WARPS;
  for (int64_t f = warp0; f < p.total; f += nwarps) {
    const int c = __ldg(p.frame_clip + f);
    const int64_t start = __ldg(p.clip_start + c);
    const int64_t n = __ldg(p.clip_num_frames + c);
    const float fps = __ldg(p.clip_fps + c);
    const float dt = __ldg(p.clip_dof_vel_dt + c);
    const bool last = (f - start) == n - 1;
    const bool pair = n >= 2;
    // (a, b) = (earlier, later) frame of the difference: (f, f+1), or (f-1, f) for a clip's last frame
    const int64_t fa = (pair && last) ? f - 1 : f;
    const int64_t fb = pair ? fa + 1 : f;
    float3 pa, pb;
    float4 ra, rb;
    lane_pose(sm, p.frames + fa * p.frame_stride, lane, pa, ra);
    lane_pose(sm, p.frames + fb * p.frame_stride, lane, pb, rb);
    const bool mine_b = pair && last;
    const float3 pos = mine_b ? pb : pa;
    const float4 rot = mine_b ? rb : ra;
    float* __restrict__ row = p.rows + f * rf;
    if (lane == 0) {
      row[0] = pos.x; row[1] = pos.y; row[2] = pos.z; row[3] = 0.0f;
      reinterpret_cast<float4*>(row)[1] = rot;
      float3 v = make_float3(0.f, 0.f, 0.f), w = v;
      if (pair) {
        v = make_float3(mul_rn(fps, sub_rn(pb.x, pa.x)), mul_rn(fps, sub_rn(pb.y, pa.y)), mul_rn(fps, sub_rn(pb.z, pa.z)));
        const float3 e = quat_to_exp_map(quat_mul(rb, quat_conj(ra)));      // quat_diff(q0, q1) = q1 * conj(q0)
        w = make_float3(fps * e.x, fps * e.y, fps * e.z);
      }
      reinterpret_cast<float4*>(row + pose_f)[0] = make_float4(v.x, v.y, v.z, 0.0f);
      reinterpret_cast<float4*>(row + pose_f)[1] = make_float4(w.x, w.y, w.z, 0.0f);
    } else if (lane < J) {
      reinterpret_cast<float4*>(row)[1 + lane] = rot;
      const int jt = sm.joint_type[lane];
      if (jt == PARC_JOINT_HINGE || jt == PARC_JOINT_SPHERICAL) {
        float3 v = make_float3(0.f, 0.f, 0.f);
        if (pair) {
          float nrm;
          float4 d = quat_mul(quat_conj(ra), rb);
          if (d.w < 0.0f) { d.x = -d.x; d.y = -d.y; d.z = -d.z; d.w = -d.w; }
          const float3 e = quat_to_exp_map(normalize4(d, nrm));               // quat_normalize, then exp map
          v = make_float3(e.x / dt, e.y / dt, e.z / dt);
        }
        float* __restrict__ o = row + pose_f + 8 + sm.dof_idx[lane];
        if (jt == PARC_JOINT_HINGE) {
          o[0] = sm.joint_axis[lane][0] * v.x + sm.joint_axis[lane][1] * v.y + sm.joint_axis[lane][2] * v.z;
        } else {
          o[0] = v.x; o[1] = v.y; o[2] = v.z;
        }
      }
    }
    // contact flags (+ zero padding of the contact slots), zero padding after the DoF velocities
    for (int k = lane; k < pose_f - contact_f; k += 32)
      row[contact_f + k] = (k < J && p.contacts) ? __ldg(p.contacts + f * J + k) : 0.0f;
    for (int k = pose_f + 8 + D + lane; k < rf; k += 32) row[k] = 0.0f;
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
  int64_t ctas = (total_frames + BUILD_WARPS - 1) / BUILD_WARPS;
  if (ctas > 148 * 16) ctas = 148 * 16;
  build_tables_kernel<<<(int)ctas, BUILD_WARPS * 32, 0, (cudaStream_t)stream>>>(p, *model);
  return check_launch();
}
