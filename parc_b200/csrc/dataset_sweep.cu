// Dataset sweep: raw motion frames -> FK -> contact labels / penetration correction / heightfield masks.
//
// SURVEY.md section 8(f) row 1 (BASELINE config 5).  Reference:
//   zmotion_editing_tools/motion_edit_lib.py:654-706  compute_hf_foot_contacts_and_correct_pen
//   zmotion_editing_tools/motion_edit_lib.py:708-747  compute_motion_terrain_hand_contacts
//   util/terrain_util.py:1951-1997                    compute_hf_mask_inds (triple python loop)
//   util/geom_util.py:80-111                          get_box_points_batch
// and the front end every one of them repeats: exp_map_to_quat + dof_to_rot + forward_kinematics on
// frames laid out [root_pos(3) | root exp-map(3) | joint DoFs(D)].

#include "parc_common.cuh"
#include "parc_rotations.cuh"
#include "parc_sdf.cuh"

namespace parc {

// ---- frames -> FK (config 5's "FK on every frame"): G lanes per frame, two frames per warp when J <= 16 ----
template <int G>
__global__ void __launch_bounds__(PARC_CTA_THREADS)
frames_fk_kernel(const float* __restrict__ frames, int64_t n, int frame_stride,
                 const __grid_constant__ ParcCharModel model_param, float* __restrict__ root_rot_out,
                 float* __restrict__ joint_rot_out, float* __restrict__ body_pos, float* __restrict__ body_rot) {
  __shared__ ParcCharModel sm;
  stage_model(&sm, model_param);
  __syncthreads();
  constexpr int GROUPS = 32 / G;
  const int lane = threadIdx.x & 31;
  const int l = lane & (G - 1), grp = lane / G;
  const int J = sm.num_bodies;
  const LaneBody lb = load_lane_body(sm, l, 0);
  const int64_t first = ((int64_t)blockIdx.x * PARC_WARPS_PER_CTA + (threadIdx.x >> 5)) * GROUPS;
  const int64_t stride = (int64_t)gridDim.x * PARC_WARPS_PER_CTA * GROUPS;
  for (int64_t base = first; base < n; base += stride) {
    const bool active = base + grp < n;
    const int64_t q = active ? base + grp : n - 1;
    float3 pos;
    float4 rot;
    frame_to_lane_pose(sm, frames + q * frame_stride, l, pos, rot);
    if (active) {
      if (l == 0) {
        if (root_rot_out) reinterpret_cast<float4*>(root_rot_out)[q] = rot;
      } else if (l < J) {
        if (joint_rot_out) reinterpret_cast<float4*>(joint_rot_out)[q * (J - 1) + (l - 1)] = rot;
      }
    }
    fk_group(lb, sm.max_depth, G, pos, rot);
    if (active && l < J) {
      if (body_pos) { float* o = body_pos + (q * J + l) * 3; o[0] = pos.x; o[1] = pos.y; o[2] = pos.z; }
      if (body_rot) reinterpret_cast<float4*>(body_rot)[q * J + l] = rot;
    }
  }
}

// ---- per-clip labelling: ONE WARP PER FRAME ----------------------------------------------------------
// The warps of a CTA share one clip's terrain tile; inside the frame loop there is no block barrier.
//   FK               lane = body; transforms parked in the warp's shared-memory slab
//   body hf          lane = body: nearest-cell height under the body origin
//   feet             lane = foot * 8 + corner (up to 4 feet): corner z vs the height of the cell it falls in
//   hands            lane = hand: exact pruned min over cells of the solid rounded-box SDF at the body origin
//   masks            lane = surface point (10 rounds for 304 points): cell -> bit in the warp's mask words,
//                    float atomic-min into the per-cell minimum body height
#define LABEL_WARPS 8
#define LABEL_THREADS (LABEL_WARPS * 32)

struct LabelParams {
  const float* frames;          // [B, F, frame_stride]
  int64_t batch, frames_per_clip;
  int frame_stride;
  ParcBodyPoints pts;
  ParcTerrainBatch terrain;
  ParcKeyBodies keys;
  float contact_eps;
  float* contacts;              // [B,F,J]  (feet / hands written, others zero)
  float* pen_correction;        // [B,F]
  float* body_hf;               // [B,F,J]
  uint32_t* frame_mask;         // [B,F,W]  W = ceil(X*Y/32)
  float* min_body_heights;      // [B,X,Y]  caller-initialised
  float* body_pos;              // [B,F,J,3]
  float* body_rot;              // [B,F,J,4]
  int frames_per_cta;
  int mask_words;
  int minh_in_smem;             // 1: per-cell minimum body heights are combined in shared memory, flushed once per CTA
  int mask_stride;              // mask_words rounded up to a multiple of 4: keeps every warp's slab 16-byte aligned
};

// 4 CTAs per SM (<= 64 registers): the kernel waits on shared-memory reads more than on anything else, and the fourth
// CTA's 8 warps hide more of that than 8 more registers per thread buy.
__global__ void __launch_bounds__(LABEL_THREADS, 4)
clip_label_kernel(const __grid_constant__ LabelParams p, const __grid_constant__ ParcCharModel model_param) {
  extern __shared__ float smem[];
  __shared__ ParcCharModel sm;
  __shared__ float s_minmax[2];

  const int X = p.terrain.dim_x, Y = p.terrain.dim_y;
  const int S = p.pts.num_points;
  const bool want_masks = (p.frame_mask || p.min_body_heights) && S > 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // Shared memory, 16-byte aligned parts first so that body transforms and local points move as float4:
  //   s_lp4 [S] float4 local points | per warp: s_bt [J][2] float4 (pos.xyz, rot.x | rot.yzw, -) + mask words (padded)
  //   | s_hf [X*Y] | s_cx [X] | s_cy [Y] | s_body [S] | s_minh [X*Y]
  float4* s_lp4 = reinterpret_cast<float4*>(smem);
  const int slab_floats = PARC_MAX_BODIES * 8 + p.mask_stride;
  float* slabs = smem + (size_t)S * 4;
  float* slab = slabs + (size_t)warp * slab_floats;
  float4* s_bt4 = reinterpret_cast<float4*>(slab);                             // [J][2]
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(slab + PARC_MAX_BODIES * 8);  // [mask_words]
  float* s_hf = slabs + (size_t)LABEL_WARPS * slab_floats;
  float* s_cx = s_hf + X * Y;
  float* s_cy = s_cx + X;
  int* s_body = reinterpret_cast<int*>(s_cy + Y);                              // [S]
  // [X*Y] per-cell minimum of the surface-point heights over this CTA's frames
  float* s_minh = reinterpret_cast<float*>(s_body + S);
  const bool minh_smem = p.min_body_heights && p.minh_in_smem;

  const int64_t b = blockIdx.y;
  stage_model(&sm, model_param);
  stage_terrain(p.terrain, b, s_hf, s_cx, s_cy, s_minmax);
  __syncthreads();
  const float hf_max = s_minmax[1];            // bounds the solid columns from above (hand SDF pruning)
  const int J = sm.num_bodies;
  if (want_masks) {
    for (int j = threadIdx.x; j < J; j += blockDim.x) {
      const int s0 = __ldg(p.pts.point_start + j), s1 = __ldg(p.pts.point_start + j + 1);
      for (int k = s0; k < s1; ++k) s_body[k] = j;
    }
    for (int i = threadIdx.x; i < S; i += blockDim.x)
      s_lp4[i] = make_float4(__ldg(p.pts.points + 3 * i), __ldg(p.pts.points + 3 * i + 1), __ldg(p.pts.points + 3 * i + 2), 0.0f);
    if (minh_smem)
      for (int i = threadIdx.x; i < X * Y; i += blockDim.x) s_minh[i] = INFINITY;
  }
  const float* mc = p.terrain.min_center + b * p.terrain.min_center_stride;
  const float min_x = __ldg(mc), min_y = __ldg(mc + 1);
  const float dx = p.terrain.half_dx * 2.0f, dy = p.terrain.half_dy * 2.0f;   // exact: halves of fp32 values
  // hoisted-reciprocal cell index: index-identical to the reference's true division (parc_selftest_grid_index)
  const GridAxis gax = make_grid_axis(min_x, dx, X), gay = make_grid_axis(min_y, dy, Y);
  const float base = sample_base_z(p.terrain, b);
  const TileSpacing spacing = tile_spacing(s_cx, s_cy, X, Y);
  const LaneBody lb = load_lane_body(sm, lane, 0);
  const int max_depth = sm.max_depth;
  __syncthreads();

  const int64_t f_begin = (int64_t)blockIdx.x * p.frames_per_cta;
  const int64_t f_end = min(f_begin + (int64_t)p.frames_per_cta, p.frames_per_clip);
  for (int64_t f = f_begin + warp; f < f_end; f += LABEL_WARPS) {
    const int64_t q = b * p.frames_per_clip + f;
    // ---- FK (lane = body) ----
    float3 pos;
    float4 rot;
    frame_to_lane_pose(sm, p.frames + q * p.frame_stride, lane, pos, rot);
    fk_warp(lb, max_depth, pos, rot);
    if (lane < J) {
      s_bt4[lane * 2] = make_float4(pos.x, pos.y, pos.z, rot.x);
      s_bt4[lane * 2 + 1] = make_float4(rot.y, rot.z, rot.w, 0.0f);
      if (p.body_pos) { float* o = p.body_pos + (q * J + lane) * 3; o[0] = pos.x; o[1] = pos.y; o[2] = pos.z; }
      if (p.body_rot) reinterpret_cast<float4*>(p.body_rot)[q * J + lane] = rot;
      if (p.body_hf) {
        p.body_hf[q * J + lane] = s_hf[grid_index_fast(pos.x, gax) * Y + grid_index_fast(pos.y, gay)];
      }
    }
    if (want_masks && p.frame_mask)
      for (int i = lane; i < p.mask_words; i += 32) s_mask[i] = 0u;
    __syncwarp();

    if (p.contacts) {
      // ---- feet: lane = foot * 8 + corner ----
      const int foot = lane >> 3, corner = lane & 7;
      bool touch = false;
      float pen = INFINITY;
      if (foot < p.keys.num_feet) {
        const float4 ta = s_bt4[p.keys.foot_body[foot] * 2], tb4 = s_bt4[p.keys.foot_body[foot] * 2 + 1];
        const float hx = p.keys.foot_half[foot][0], hy = p.keys.foot_half[foot][1], hz = p.keys.foot_half[foot][2];
        float3 c = make_float3((corner & 1) ? hx : -hx, (corner & 2) ? hy : -hy, (corner & 4) ? hz : -hz);
        c.x += p.keys.foot_offset[foot][0]; c.y += p.keys.foot_offset[foot][1]; c.z += p.keys.foot_offset[foot][2];
        const float3 r = quat_rotate(make_float4(ta.w, tb4.x, tb4.y, tb4.z), c);
        const float3 wp = make_float3(r.x + ta.x, r.y + ta.y, r.z + ta.z);
        const float h = s_hf[grid_index_fast(wp.x, gax) * Y + grid_index_fast(wp.y, gay)];
        touch = wp.z < add_rn(h, p.contact_eps);            // box_points_z < cell_heights + contact_eps
        pen = sub_rn(wp.z, h);
      }
      const unsigned any = __ballot_sync(PARC_FULL_MASK, touch);
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) pen = fminf(pen, __shfl_xor_sync(PARC_FULL_MASK, pen, o));   // min per foot
      // pen_correction_z starts at zero (:683) and takes the min over feet
      float corr = fminf(pen, 0.0f);
      corr = fminf(corr, __shfl_xor_sync(PARC_FULL_MASK, corr, 8));
      corr = fminf(corr, __shfl_xor_sync(PARC_FULL_MASK, corr, 16));
      // ---- hands: exact min over cells of the solid rounded-box SDF at the body origin; the cells within reach are
      // split over the warp's lanes (one hand after the other), lane h keeps hand h's value ----
      float hand_sd = INFINITY;
      for (int h = 0; h < p.keys.num_hands; ++h) {
        const float4 ta = s_bt4[p.keys.hand_body[h] * 2];
        const float sol = warp_min_solid_sdf(s_hf, s_cx, s_cy, X, Y, p.terrain.half_dx, p.terrain.half_dy, base, hf_max,
                                             spacing, make_float3(ta.x, ta.y, ta.z), lane);
        if (lane == h) hand_sd = sol - p.keys.hand_radius[h];   // sdRoundBox = sdBox - r (geom_util.py:113-120)
      }
      // ---- contact row: zeros, then feet / hands ----
      float* crow = p.contacts + q * J;
      if (lane < J) crow[lane] = 0.0f;
      __syncwarp();
      if (corner == 0 && foot < p.keys.num_feet)
        crow[p.keys.foot_body[foot]] = ((any >> (foot * 8)) & 0xffu) ? 1.0f : 0.0f;
      if (lane < p.keys.num_hands) crow[p.keys.hand_body[lane]] = hand_sd < p.contact_eps ? 1.0f : 0.0f;
      if (lane == 0 && p.pen_correction) p.pen_correction[q] = corr;
    }

    // ---- masks: lane = surface point ----
    if (want_masks) {
      for (int k = lane; k < S; k += 32) {
        const int bj = s_body[k];
        const float4 ta = s_bt4[bj * 2], tb4 = s_bt4[bj * 2 + 1], l4 = s_lp4[k];
        const float3 r = quat_rotate(make_float4(ta.w, tb4.x, tb4.y, tb4.z), make_float3(l4.x, l4.y, l4.z));
        const float3 wp = make_float3(r.x + ta.x, r.y + ta.y, r.z + ta.z);
        const int cell = grid_index_fast(wp.x, gax) * Y + grid_index_fast(wp.y, gay);
        if (p.frame_mask) atomicOr(&s_mask[cell >> 5], 1u << (cell & 31));
        if (minh_smem) atomic_min_float(s_minh + cell, wp.z);
        else if (p.min_body_heights) atomic_min_float(p.min_body_heights + b * (int64_t)X * Y + cell, wp.z);
      }
      __syncwarp();
      if (p.frame_mask)
        for (int i = lane; i < p.mask_words; i += 32) p.frame_mask[q * p.mask_words + i] = s_mask[i];
    }
    __syncwarp();
  }
  if (minh_smem) {
    // one global atomic per touched cell and CTA instead of one per surface point and frame
    __syncthreads();
    float* g = p.min_body_heights + b * (int64_t)X * Y;
    for (int i = threadIdx.x; i < X * Y; i += blockDim.x) {
      const float v = s_minh[i];
      if (v < INFINITY) atomic_min_float(g + i, v);
    }
  }
}

static int warp_grid(int64_t n) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = (n + PARC_WARPS_PER_CTA - 1) / PARC_WARPS_PER_CTA;
  const int64_t cap = (int64_t)sms * 16;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace parc

using namespace parc;

extern "C" int parc_frames_fk(const float* frames, int64_t n, int32_t frame_stride, const ParcCharModel* model,
                              float* root_rot_out, float* joint_rot_out, float* body_pos, float* body_rot,
                              void* stream) {
  if (!model) return PARC_E_NULL;
  int rc = parc_validate_model(model);
  if (rc) return rc;
  if (n < 0 || frame_stride < 6 + model->dof_size) return PARC_E_SIZE;
  if (n == 0) return PARC_OK;
  if (!frames) return PARC_E_NULL;
  if (!aligned16(root_rot_out) || !aligned16(joint_rot_out) || !aligned16(body_rot)) return PARC_E_ALIGN;
  if (model->num_bodies <= 16)
    frames_fk_kernel<16><<<warp_grid((n + 1) / 2), PARC_CTA_THREADS, 0, (cudaStream_t)stream>>>(
        frames, n, frame_stride, *model, root_rot_out, joint_rot_out, body_pos, body_rot);
  else
    frames_fk_kernel<32><<<warp_grid(n), PARC_CTA_THREADS, 0, (cudaStream_t)stream>>>(
        frames, n, frame_stride, *model, root_rot_out, joint_rot_out, body_pos, body_rot);
  return check_launch();
}

extern "C" int parc_clip_label(const float* frames, int64_t batch, int64_t frames_per_clip, int32_t frame_stride,
                               const ParcCharModel* model, const ParcBodyPoints* pts,
                               const ParcTerrainBatch* terrain, const ParcKeyBodies* keys, float contact_eps,
                               float* contacts_out, float* pen_correction_out, float* body_hf_out,
                               uint32_t* frame_mask_out, float* min_body_heights, float* body_pos,
                               float* body_rot, void* stream) {
  if (!model || !pts || !terrain || !keys) return PARC_E_NULL;
  int rc = parc_validate_model(model);
  if (rc) return rc;
  if (batch < 0 || frames_per_clip < 0 || frame_stride < 6 + model->dof_size) return PARC_E_SIZE;
  if (keys->num_feet < 0 || keys->num_feet > PARC_MAX_KEY_BODIES || keys->num_hands < 0 ||
      keys->num_hands > PARC_MAX_KEY_BODIES)
    return PARC_E_SIZE;
  for (int i = 0; i < keys->num_feet; ++i)
    if (keys->foot_body[i] < 0 || keys->foot_body[i] >= model->num_bodies) return PARC_E_SIZE;
  for (int i = 0; i < keys->num_hands; ++i)
    if (keys->hand_body[i] < 0 || keys->hand_body[i] >= model->num_bodies) return PARC_E_SIZE;
  if (batch == 0 || frames_per_clip == 0) return PARC_OK;
  if (!frames || !terrain->hf || !terrain->min_center || !terrain->x_nodes || !terrain->y_nodes) return PARC_E_NULL;
  if ((frame_mask_out || min_body_heights) && (!pts->points || !pts->point_start || pts->num_points <= 0))
    return PARC_E_NULL;
  if (terrain->dim_x <= 0 || terrain->dim_y <= 0) return PARC_E_SIZE;
  if (!aligned16(body_rot)) return PARC_E_ALIGN;

  LabelParams p;
  p.frames = frames; p.batch = batch; p.frames_per_clip = frames_per_clip; p.frame_stride = frame_stride;
  p.pts = *pts; p.terrain = *terrain; p.keys = *keys; p.contact_eps = contact_eps;
  p.contacts = contacts_out; p.pen_correction = pen_correction_out; p.body_hf = body_hf_out;
  p.frame_mask = frame_mask_out; p.min_body_heights = min_body_heights; p.body_pos = body_pos; p.body_rot = body_rot;
  const int cells = terrain->dim_x * terrain->dim_y;
  p.mask_words = (cells + 31) / 32;
  p.mask_stride = (p.mask_words + 3) / 4 * 4;
  const size_t S = (frame_mask_out || min_body_heights) ? (size_t)pts->num_points : 0;
  p.pts.num_points = (int)S;
  size_t smem = ((size_t)cells + terrain->dim_x + terrain->dim_y + S * 5 +
                 (size_t)LABEL_WARPS * (PARC_MAX_BODIES * 8 + p.mask_stride)) * 4;
  if (smem > PARC_SMEM_LIMIT) return PARC_E_SIZE;       // per-clip labelling terrains are small tiles (<= ~190 x 190)
  // the per-cell minima are combined in shared memory when a second tile-sized array still fits
  p.minh_in_smem = (min_body_heights && smem + (size_t)cells * 4 <= PARC_SMEM_LIMIT) ? 1 : 0;
  if (p.minh_in_smem) smem += (size_t)cells * 4;
  static SmemOptIn opt;
  rc = ensure_dynamic_smem(clip_label_kernel, opt, smem);
  if (rc) return rc;
  int64_t rounds = (batch * frames_per_clip) / ((int64_t)148 * 8 * LABEL_WARPS);
  if (rounds < 1) rounds = 1;
  if (rounds > 8) rounds = 8;
  const int64_t fpc = rounds * LABEL_WARPS;
  p.frames_per_cta = (int)fpc;
  const int J = model->num_bodies;
  const int64_t cells64 = cells;
  for (int64_t b0 = 0; b0 < batch; b0 += PARC_GRID_Y_MAX) {       // grid.y is limited to 65 535 clips per launch
    const int64_t nb = batch - b0 < PARC_GRID_Y_MAX ? batch - b0 : PARC_GRID_Y_MAX;
    const int64_t q0 = b0 * frames_per_clip;
    p.batch = nb;
    p.frames = frames + q0 * frame_stride;
    p.terrain = *terrain;
    p.terrain.hf += b0 * terrain->hf_batch_stride;
    p.terrain.min_center += b0 * terrain->min_center_stride;
    if (p.terrain.base_z) p.terrain.base_z += b0 * terrain->base_z_stride;
    p.contacts = contacts_out ? contacts_out + q0 * J : nullptr;
    p.pen_correction = pen_correction_out ? pen_correction_out + q0 : nullptr;
    p.body_hf = body_hf_out ? body_hf_out + q0 * J : nullptr;
    p.frame_mask = frame_mask_out ? frame_mask_out + q0 * p.mask_words : nullptr;
    p.min_body_heights = min_body_heights ? min_body_heights + b0 * cells64 : nullptr;
    p.body_pos = body_pos ? body_pos + q0 * J * 3 : nullptr;
    p.body_rot = body_rot ? body_rot + q0 * J * 4 : nullptr;
    dim3 grid((unsigned)((frames_per_clip + fpc - 1) / fpc), (unsigned)nb);
    clip_label_kernel<<<grid, LABEL_THREADS, smem, (cudaStream_t)stream>>>(p, *model);
  }
  return check_launch();
}
