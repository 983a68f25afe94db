// Point <-> heightfield box SDF and the fused body-point penetration / contact loss (sm_100a).
//
// Reference: util/terrain_util.py:1835-1893 (points_hf_sdf), :1777-1804 (points_boxes_sdf),
// util/geom_util.py:122-143 (sdBox), tools/procgen/mdm_path.py:79-110 (compute_motion_loss),
// tools/motion_opt/motion_optimization.py:241-272 (motion_terrain_contact_loss pen/contact terms).
//
// The reference materialises [B,N,M,3] tensors (M = all cells) per body and lets autograd replay
// them; here every (sample, frame) is one CTA pass that never leaves the SM:
//   warp 0: FK of the frame (lane = body)          -> body transforms in shared memory
//   all   : thread = surface point; exact min over ALL cells of both box SDFs (air column for
//           penetration, solid column for contact), terrain tile staged in shared memory
//   warp 0: per-body first-index min (contact), per-body gradient sums, FK VJP -> leaf gradients
// The min is exact and tie-breaks on the first flat cell index, as torch.min does.
#include "parc_common.cuh"
#include "parc_internal.h"
#include "parc_sdf.cuh"

namespace parc {

// ------------------------------------------------------------------------------------------------
// a13 stand-alone: points [B,N,3] -> sdf [B,N]
// ------------------------------------------------------------------------------------------------
// SMEM_TILE: the sample's heightfield tile is staged in shared memory; otherwise (tiles beyond PARC_SMEM_LIMIT) it is
// read from global memory and only the cell-centre coordinates are staged.
template <bool SMEM_TILE>
__global__ void __launch_bounds__(256)
points_hf_sdf_kernel(const float* __restrict__ points, int64_t n_points, const __grid_constant__ ParcTerrainBatch t,
                     int inverted, float* __restrict__ sdf, int32_t* __restrict__ arg) {
  extern __shared__ float smem[];
  const int X = t.dim_x, Y = t.dim_y;
  float* s_cx = smem;
  float* s_cy = s_cx + X;
  float* s_hf = s_cy + Y;
  __shared__ float s_minmax[2];
  const int64_t b = blockIdx.y;
  const float* __restrict__ hfp;
  if (SMEM_TILE) {
    stage_terrain(t, b, s_hf, s_cx, s_cy, s_minmax);
    hfp = s_hf;
  } else {
    stage_terrain_global(t, b, s_cx, s_cy, s_minmax);
    hfp = t.hf + b * t.hf_batch_stride;
  }
  __syncthreads();
  TileBlocks tb = {nullptr, nullptr, 0};
  if (SMEM_TILE) {                       // per-block height ranges behind the tile
    float* s_bmax = s_hf + X * Y;
    float* s_bmin = s_bmax + sdf_blocks(X) * sdf_blocks(Y);
    stage_tile_blocks(s_hf, X, Y, s_bmax, s_bmin);
    tb.bmax = s_bmax; tb.bmin = s_bmin; tb.nby = sdf_blocks(Y);
    __syncthreads();
  }
  const float base = sample_base_z(t, b);
  const float hf_min = s_minmax[0], hf_max = s_minmax[1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_points; i += (int64_t)gridDim.x * blockDim.x) {
    const float* pp = points + (b * n_points + i) * 3;
    const float3 p = make_float3(__ldg(pp), __ldg(pp + 1), __ldg(pp + 2));
    float v;
    int a;
    if (inverted) {
      const SdfBest r = scan_cells<true, false>(hfp, s_cx, s_cy, X, Y, t.half_dx, t.half_dy, base, hf_min, hf_max, p, true, tb);
      v = -1.0f * r.inv; a = r.arg_inv;
    } else {
      const SdfBest r = scan_cells<false, true>(hfp, s_cx, s_cy, X, Y, t.half_dx, t.half_dy, base, hf_min, hf_max, p, true, tb);
      v = r.sol; a = r.arg_sol;
    }
    sdf[b * n_points + i] = v;
    if (arg) arg[b * n_points + i] = a;
  }
}

// VJP of the above with respect to the points: the min routes the gradient to the arg-min cell (first index on
// ties, recorded by the forward launch), whose box SDF has the sub-gradient of sd_box_grad; inverted negates.
// One thread per point; the cell's height / centre are read straight from global memory.
__global__ void __launch_bounds__(256)
points_hf_sdf_bwd_kernel(const float* __restrict__ points, int64_t batch, int64_t n_points,
                         const __grid_constant__ ParcTerrainBatch t, int inverted, const int32_t* __restrict__ arg,
                         const float* __restrict__ g_sdf, float* __restrict__ g_points) {
  const int64_t total = batch * n_points;
  const int Y = t.dim_y;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / n_points;
    const int cell = __ldg(arg + i);
    const int ix = cell / Y, iy = cell - ix * Y;
    const float* mc = t.min_center + b * t.min_center_stride;
    const float cx = __ldg(t.x_nodes + ix) + __ldg(mc), cy = __ldg(t.y_nodes + iy) + __ldg(mc + 1);
    const float h = __ldg(t.hf + b * t.hf_batch_stride + cell);
    const float base = sample_base_z(t, b);
    float cz, hz;
    if (inverted) { const float top = -base; cz = (h + top) * 0.5f; hz = (top - h) * 0.5f; }
    else { cz = (h + base) * 0.5f; hz = (h - base) * 0.5f; }
    const float3 p = make_float3(__ldg(points + i * 3), __ldg(points + i * 3 + 1), __ldg(points + i * 3 + 2));
    const float3 g = sd_box_grad(make_float3(p.x - cx, p.y - cy, p.z - cz), make_float3(t.half_dx, t.half_dy, hz));
    const float s = inverted ? -__ldg(g_sdf + i) : __ldg(g_sdf + i);
    g_points[i * 3] = s * g.x; g_points[i * 3 + 1] = s * g.y; g_points[i * 3 + 2] = s * g.z;
  }
}

// ------------------------------------------------------------------------------------------------
// a14/a15 fused: FK -> body points -> SDF -> pen / contact (+ gradients wrt the pose)
//
// ONE WARP PER FRAME, no block-level barrier inside the frame loop:
//   FK            lane = body (fk_warp_keep), transforms parked in the warp's shared-memory slab
//   sweep         lane = surface point (10 rounds for 304 points): world point, exact pruned SDF scan of the
//                 CTA's terrain tile, d/d(world point) of both terms -> slab
//   contact min   lane = body: first-index min over the body's points (sequential, <= 44 reads)
//   chain rule    lane = point: add the winner's contact gradient, VJP through rotate(body_rot, local) -> slab
//   body sums     lane = body: fixed-order sums of its points' 7 floats; then the FK VJP in-warp
// The warps of a CTA share one sample's terrain tile (read-only after staging).
// ------------------------------------------------------------------------------------------------
// Warps per CTA are chosen at launch.  Each warp owns a 10.6 KB slab (humanoid: 304 points x 8 floats + the body
// transforms) next to ~6 KB shared by the CTA, and shared memory is what bounds the occupancy: 10 warps = 112 KB, two
// CTAs = 20 warps per SM; 4 warps = 48 KB, four CTAs = 16 warps per SM -- the fallback when the 10-warp CTA would push
// the terrain tile (or a character with more surface points) out of shared memory.
#define LOSS_WARPS_MAX 10
#define LOSS_WARPS_MIN 4
#define LOSS_THREADS_MAX (LOSS_WARPS_MAX * 32)

struct BodyLossParams {
  const float *root_pos, *root_rot, *joint_rot, *contacts;
  int64_t root_pos_stride;      // floats between consecutive frames' root positions (3 = dense [B,F,3])
  int64_t batch, frames;
  ParcBodyPoints pts;
  ParcTerrainBatch terrain;
  float w_pen, w_contact;
  float *pen_out, *contact_out;
  float *g_root_pos, *g_root_rot, *g_joint_rot;
  int frames_per_cta;
  int want_grad;
};

// CTA_TEAM = false: one warp per frame, blockDim / 32 frames in flight per CTA, no block barrier in the frame
// loop -- the throughput form for big batches.  true: the whole CTA works on ONE frame (surface points spread
// over 128 threads, the lane = body steps on warp 0 between block barriers) -- the latency form for a single clip
// (the motion optimiser's 254 frames would otherwise occupy 254 warps of the 9 472 the GPU holds).  Both forms add
// in the same order and give identical bits.
#define LOSS_PT 8            // floats per surface point in the slab: 7 gradient slots + its penetration term
template <bool SMEM_TILE, bool CTA_TEAM>
__global__ void __launch_bounds__(LOSS_THREADS_MAX)
body_loss_kernel(const __grid_constant__ BodyLossParams p, const __grid_constant__ ParcCharModel model_param) {
  extern __shared__ float smem[];
  __shared__ ParcCharModel sm;
  __shared__ float s_minmax[2];
  const int WPF = CTA_TEAM ? (int)(blockDim.x >> 5) : 1;     // warps per frame
  const int TEAMS = CTA_TEAM ? 1 : (int)(blockDim.x >> 5);   // frames in flight per CTA

  const int X = p.terrain.dim_x, Y = p.terrain.dim_y;
  const int S = p.pts.num_points;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int team = warp / WPF, twarp = warp % WPF;      // this warp's team and its index inside it
  const int tlane = twarp * 32 + lane;                  // thread index inside the team
  float* s_cx = smem;
  float* s_cy = s_cx + X;
  int* s_body = reinterpret_cast<int*>(s_cy + Y);                    // [S] body of point
  float* s_lp = reinterpret_cast<float*>(s_body + S);                // [S][3] local points
  float* slab = s_lp + (size_t)S * 3 + (size_t)team * ((size_t)S * LOSS_PT + PARC_MAX_BODIES * 9);
  float* s_bt = slab;                                                // [J][9] pos(3) rot(4) contact(1) winner(1)
  float* s_pt = slab + PARC_MAX_BODIES * 9;                          // [S][LOSS_PT]
  float* s_tile = s_lp + (size_t)S * 3 + (size_t)TEAMS * ((size_t)S * LOSS_PT + PARC_MAX_BODIES * 9);   // [X*Y] last
  auto team_sync = [&]() {
    if (!CTA_TEAM) __syncwarp(); else __syncthreads();             // CTA_TEAM: the team is the CTA
  };

  const int64_t b = blockIdx.y;
  stage_model(&sm, model_param);
  const float* __restrict__ s_hf;       // the sample's tile: shared memory, or global for tiles beyond PARC_SMEM_LIMIT
  if (SMEM_TILE) {
    stage_terrain(p.terrain, b, s_tile, s_cx, s_cy, s_minmax);
    s_hf = s_tile;
  } else {
    stage_terrain_global(p.terrain, b, s_cx, s_cy, s_minmax);
    s_hf = p.terrain.hf + b * p.terrain.hf_batch_stride;
  }
  __syncthreads();
  const float hf_min = s_minmax[0], hf_max = s_minmax[1];
  const int J = sm.num_bodies;
  for (int j = threadIdx.x; j < J; j += blockDim.x) {
    const int s0 = __ldg(p.pts.point_start + j), s1 = __ldg(p.pts.point_start + j + 1);
    for (int k = s0; k < s1; ++k) s_body[k] = j;
  }
  for (int i = threadIdx.x; i < S * 3; i += blockDim.x) s_lp[i] = __ldg(p.pts.points + i);
  TileBlocks tb = {nullptr, nullptr, 0};
  if (SMEM_TILE) {                       // per-block height ranges behind the tile (visible after the barrier below)
    float* s_bmax = s_tile + X * Y;
    float* s_bmin = s_bmax + sdf_blocks(X) * sdf_blocks(Y);
    stage_tile_blocks(s_tile, X, Y, s_bmax, s_bmin);
    tb.bmax = s_bmax; tb.bmin = s_bmin; tb.nby = sdf_blocks(Y);
  }
  const float base = sample_base_z(p.terrain, b);
  const float hx = p.terrain.half_dx, hy = p.terrain.half_dy;
  const LaneBody lb = load_lane_body(sm, lane, 0);
  const int max_depth = sm.max_depth;
  const int my_s0 = lane < J ? __ldg(p.pts.point_start + lane) : 0;
  const int my_s1 = lane < J ? __ldg(p.pts.point_start + lane + 1) : 0;
  __syncthreads();

  const int64_t f_begin = (int64_t)blockIdx.x * p.frames_per_cta;
  const int64_t f_end = min(f_begin + (int64_t)p.frames_per_cta, p.frames);
  for (int64_t f = f_begin + team; f < f_end; f += TEAMS) {
    const int64_t q = b * p.frames + f;
    // ---- FK, lane = body (the team's first warp) ----
    float4 prot = make_float4(0.f, 0.f, 0.f, 1.f), local = prot, rot = prot;
    float3 pos = make_float3(0.f, 0.f, 0.f);
    float my_contact = 0.0f;
    if (twarp == 0) {
      if (lane == 0) {
        const float* rp = p.root_pos + q * p.root_pos_stride;
        pos = make_float3(__ldg(rp), __ldg(rp + 1), __ldg(rp + 2));
        rot = __ldg(reinterpret_cast<const float4*>(p.root_rot) + q);
      } else if (lane < J) {
        rot = __ldg(reinterpret_cast<const float4*>(p.joint_rot) + q * (J - 1) + (lane - 1));
      }
      fk_warp_keep(lb, max_depth, pos, rot, prot, local);
      if (lane < J) {
        my_contact = __ldg(p.contacts + q * J + lane);
        float* t = s_bt + lane * 9;
        t[0] = pos.x; t[1] = pos.y; t[2] = pos.z; t[3] = rot.x; t[4] = rot.y; t[5] = rot.z; t[6] = rot.w;
        t[7] = my_contact;
      }
    }
    team_sync();

    // ---- sweep, team thread = surface point ----
    for (int k = tlane; k < S; k += 32 * WPF) {
      const int bj = s_body[k];
      const float* t = s_bt + bj * 9;
      const float3 lp = make_float3(s_lp[k * 3], s_lp[k * 3 + 1], s_lp[k * 3 + 2]);
      const float4 br = make_float4(t[3], t[4], t[5], t[6]);
      const float3 r = quat_rotate(br, lp);
      const float3 wp = make_float3(r.x + t[0], r.y + t[1], r.z + t[2]);
      // A body whose contact weight is exactly 0 contributes exactly 0 to the contact term and to its
      // gradient (closest * 0), so its solid-column scan is skipped.
      // (one call for both cases: the lanes of a warp are consecutive points and may belong to bodies with and
      // without a contact weight; two instantiations would run one after the other)
      const bool sol_on = t[7] != 0.0f;
      SdfBest best = scan_cells<true, true>(s_hf, s_cx, s_cy, X, Y, hx, hy, base, hf_min, hf_max, wp, sol_on, tb);
      if (!sol_on) { best.sol = 0.0f; best.arg_sol = 0; }
      // penetration: sdf = -best.inv ; neg = min(sdf, 0) ; pen += -neg
      const float sdf_inv = -1.0f * best.inv;
      float* g7 = s_pt + (size_t)k * LOSS_PT;
      g7[7] = -fminf(sdf_inv, 0.0f);
      g7[6] = fmaxf(best.sol, 0.0f);                 // clamp(sdf_solid, min=0)
      if (p.want_grad) {
        // d pen / d wp = [sdf_inv <= 0] * grad sdBox(air cell), already weighted
        float3 gp = make_float3(0.f, 0.f, 0.f);
        if (sdf_inv <= 0.0f) {
          const float3 g = cell_grad(s_hf, s_cx, s_cy, Y, hx, hy, base, true, best.arg_inv, wp);
          gp = make_float3(p.w_pen * g.x, p.w_pen * g.y, p.w_pen * g.z);
        }
        // solid-cell gradient, used only if this point wins its body's min and the clamp passes
        float3 gs = make_float3(0.f, 0.f, 0.f);
        if (best.sol >= 0.0f && t[7] != 0.0f) gs = cell_grad(s_hf, s_cx, s_cy, Y, hx, hy, base, false, best.arg_sol, wp);
        g7[0] = gp.x; g7[1] = gp.y; g7[2] = gp.z;
        g7[3] = gs.x; g7[4] = gs.y; g7[5] = gs.z;
      }
    }
    team_sync();

    // ---- penetration sum + contact term on the team's first warp: lane-strided partial sums in point order, then
    //      the xor tree (the same order for every WPF); first-index min over each body's points ----
    if (twarp == 0) {
      float pen_local = 0.0f;
      for (int k = lane; k < S; k += 32) pen_local += s_pt[(size_t)k * LOSS_PT + 7];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) pen_local += __shfl_xor_sync(PARC_FULL_MASK, pen_local, o);
      float cterm = 0.0f;
      int win = -1;
      if (lane < J) {
        float bestv = INFINITY;
        for (int k = my_s0; k < my_s1; ++k) {
          const float v = s_pt[(size_t)k * LOSS_PT + 6];
          if (v < bestv) { bestv = v; win = k; }
        }
        cterm = bestv * my_contact;                    // closest_distances * contacts[..., b]
        s_bt[lane * 9 + 8] = __int_as_float(win);
      }
      float contact_f = 0.0f;
      for (int j = 0; j < J; ++j) contact_f += __shfl_sync(PARC_FULL_MASK, cterm, j);   // body order, as the reference
      if (lane == 0) {
        if (p.pen_out) p.pen_out[q] = pen_local;
        if (p.contact_out) p.contact_out[q] = contact_f;
      }
    }
    team_sync();
    if (!p.want_grad) continue;

    // ---- chain rule, team thread = point: add the winner's contact gradient, VJP through the body transform ----
    for (int k = tlane; k < S; k += 32 * WPF) {
      const int bj = s_body[k];
      const float* t = s_bt + bj * 9;
      float* g7 = s_pt + (size_t)k * LOSS_PT;
      float3 g = make_float3(g7[0], g7[1], g7[2]);
      if (__float_as_int(t[8]) == k) {               // this point won its body's contact min
        const float cw = p.w_contact * t[7];
        g.x += cw * g7[3]; g.y += cw * g7[4]; g.z += cw * g7[5];
      }
      const float3 lp = make_float3(s_lp[k * 3], s_lp[k * 3 + 1], s_lp[k * 3 + 2]);
      const float4 gq = quat_rotate_vjp_q(make_float4(t[3], t[4], t[5], t[6]), lp, g);
      g7[0] = g.x; g7[1] = g.y; g7[2] = g.z;
      g7[3] = gq.x; g7[4] = gq.y; g7[5] = gq.z; g7[6] = gq.w;
    }
    team_sync();

    // ---- body sums (fixed order), lane = body; then the FK VJP ----
    if (twarp == 0) {
      float3 gp = make_float3(0.f, 0.f, 0.f);
      float4 gr = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = my_s0; k < my_s1; ++k) {
        const float* g7 = s_pt + (size_t)k * LOSS_PT;
        gp.x += g7[0]; gp.y += g7[1]; gp.z += g7[2];
        gr.x += g7[3]; gr.y += g7[4]; gr.z += g7[5]; gr.w += g7[6];
      }
      float4 gj;
      fk_warp_vjp(lb, J, lane, prot, local, gp, gr, gj);
      if (lane == 0) {
        if (p.g_root_pos) { p.g_root_pos[q * 3] = gp.x; p.g_root_pos[q * 3 + 1] = gp.y; p.g_root_pos[q * 3 + 2] = gp.z; }
        if (p.g_root_rot) reinterpret_cast<float4*>(p.g_root_rot)[q] = gr;
      } else if (lane < J) {
        if (p.g_joint_rot) reinterpret_cast<float4*>(p.g_joint_rot)[q * (J - 1) + (lane - 1)] = gj;
      }
    }
    team_sync();
  }
}

static size_t nodes_smem_bytes(const ParcTerrainBatch* t) { return ((size_t)t->dim_x + t->dim_y) * sizeof(float); }
// the tile plus its per-block (max, min) heights
static size_t tile_smem_bytes(const ParcTerrainBatch* t) {
  return ((size_t)t->dim_x * t->dim_y + 2 * (size_t)sdf_blocks(t->dim_x) * sdf_blocks(t->dim_y)) * sizeof(float);
}

static int check_terrain(const ParcTerrainBatch* t) {
  if (!t || !t->hf || !t->min_center || !t->x_nodes || !t->y_nodes) return PARC_E_NULL;
  if (t->dim_x <= 0 || t->dim_y <= 0 || t->hf_batch_stride < 0) return PARC_E_SIZE;
  if ((int64_t)t->dim_x * t->dim_y >= (1ll << 31)) return PARC_E_SIZE;
  return PARC_OK;
}

// the terrain descriptor of samples [b0, ...) of a batch (launches are chunked to the grid's y limit)
static ParcTerrainBatch terrain_from(const ParcTerrainBatch& t, int64_t b0) {
  ParcTerrainBatch r = t;
  r.hf += b0 * t.hf_batch_stride;
  r.min_center += b0 * t.min_center_stride;
  if (r.base_z) r.base_z += b0 * t.base_z_stride;
  return r;
}

}  // namespace parc

using namespace parc;

extern "C" int parc_points_hf_sdf(const float* points, int64_t batch, int64_t n_points,
                                  const ParcTerrainBatch* terrain, int32_t inverted, float* sdf_out,
                                  int32_t* arg_out, void* stream) {
  if (batch < 0 || n_points < 0) return PARC_E_SIZE;
  if (batch == 0 || n_points == 0) return PARC_OK;
  if (!points || !sdf_out) return PARC_E_NULL;
  int rc = check_terrain(terrain);
  if (rc) return rc;
  // the tile goes to shared memory when it fits; larger terrains are scanned from global memory
  const bool smem_tile = nodes_smem_bytes(terrain) + tile_smem_bytes(terrain) <= PARC_SMEM_LIMIT;
  const size_t smem = nodes_smem_bytes(terrain) + (smem_tile ? tile_smem_bytes(terrain) : 0);
  if (smem > PARC_SMEM_LIMIT) return PARC_E_SIZE;     // only the cell-centre coordinates of a > 25 600-cell-wide grid
  static SmemOptIn opt_tile, opt_global;
  rc = smem_tile ? ensure_dynamic_smem(points_hf_sdf_kernel<true>, opt_tile, smem)
                 : ensure_dynamic_smem(points_hf_sdf_kernel<false>, opt_global, smem);
  if (rc) return rc;
  int64_t gx = (n_points + 255) / 256;
  if (gx > 4096) gx = 4096;
  for (int64_t b0 = 0; b0 < batch; b0 += PARC_GRID_Y_MAX) {       // grid.y is limited to 65 535 samples per launch
    const int64_t nb = batch - b0 < PARC_GRID_Y_MAX ? batch - b0 : PARC_GRID_Y_MAX;
    const ParcTerrainBatch t = terrain_from(*terrain, b0);
    dim3 grid((unsigned)gx, (unsigned)nb);
    const float* pp = points + b0 * n_points * 3;
    float* so = sdf_out + b0 * n_points;
    int32_t* ao = arg_out ? arg_out + b0 * n_points : nullptr;
    if (smem_tile) points_hf_sdf_kernel<true><<<grid, 256, smem, (cudaStream_t)stream>>>(pp, n_points, t, inverted, so, ao);
    else points_hf_sdf_kernel<false><<<grid, 256, smem, (cudaStream_t)stream>>>(pp, n_points, t, inverted, so, ao);
  }
  return check_launch();
}

extern "C" int parc_points_hf_sdf_bwd(const float* points, int64_t batch, int64_t n_points,
                                      const ParcTerrainBatch* terrain, int32_t inverted, const int32_t* arg,
                                      const float* g_sdf, float* g_points_out, void* stream) {
  if (batch < 0 || n_points < 0) return PARC_E_SIZE;
  if (batch == 0 || n_points == 0) return PARC_OK;
  if (!points || !arg || !g_sdf || !g_points_out) return PARC_E_NULL;
  int rc = check_terrain(terrain);
  if (rc) return rc;
  const int64_t total = batch * n_points;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  points_hf_sdf_bwd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(points, batch, n_points, *terrain, inverted,
                                                                          arg, g_sdf, g_points_out);
  return check_launch();
}

extern "C" int parc_body_loss(const float* root_pos, const float* root_rot, const float* joint_rot,
                              const float* contacts, int64_t batch, int64_t frames, const ParcCharModel* model,
                              const ParcBodyPoints* pts, const ParcTerrainBatch* terrain, float w_pen,
                              float w_contact, float* pen_out, float* contact_out, float* g_root_pos,
                              float* g_root_rot, float* g_joint_rot, void* stream) {
  return parc::body_loss_launch(root_pos, 3, root_rot, joint_rot, contacts, batch, frames, model, pts, terrain, w_pen,
                                w_contact, pen_out, contact_out, g_root_pos, g_root_rot, g_joint_rot, stream);
}

// The launch behind parc_body_loss; root positions may be rows of a wider array (root_pos_stride floats apart), which
// lets the motion optimiser (motion_opt.cu) read them straight out of its [F, 6+D] leaf buffer.
int parc::body_loss_launch(const float* root_pos, int64_t root_pos_stride, const float* root_rot, const float* joint_rot,
                           const float* contacts, int64_t batch, int64_t frames, const ParcCharModel* model,
                           const ParcBodyPoints* pts, const ParcTerrainBatch* terrain, float w_pen, float w_contact,
                           float* pen_out, float* contact_out, float* g_root_pos, float* g_root_rot,
                           float* g_joint_rot, void* stream) {
  if (!model || !pts) return PARC_E_NULL;
  if (root_pos_stride < 3) return PARC_E_SIZE;
  int rc = parc_validate_model(model);
  if (rc) return rc;
  if (batch < 0 || frames < 0 || pts->num_points <= 0) return PARC_E_SIZE;
  if (batch == 0 || frames == 0) return PARC_OK;
  if (!root_pos || !root_rot || !contacts || !pts->points || !pts->point_start) return PARC_E_NULL;
  if (model->num_bodies > 1 && !joint_rot) return PARC_E_NULL;
  rc = check_terrain(terrain);
  if (rc) return rc;
  if (!aligned16(root_rot) || !aligned16(joint_rot) || !aligned16(g_root_rot) || !aligned16(g_joint_rot))
    return PARC_E_ALIGN;

  BodyLossParams p;
  p.frames = frames; p.pts = *pts;
  p.w_pen = w_pen; p.w_contact = w_contact;
  p.want_grad = (g_root_pos || g_root_rot || g_joint_rot) ? 1 : 0;

  const size_t S = (size_t)pts->num_points;
  // Few frames in total (a single clip being optimised): the whole CTA works on one frame, one frame per CTA.
  // Otherwise one warp per frame, `warps` frames in flight per CTA, the terrain staging amortised over several
  // rounds when there is plenty of work.
  const bool team = batch * frames <= (int64_t)148 * 16;
  auto fixed_bytes = [&](int teams) {
    return nodes_smem_bytes(terrain) + (S * (1 + 3) + teams * (S * LOSS_PT + PARC_MAX_BODIES * 9)) * sizeof(float);
  };
  // CTA size: the large CTA while its slabs leave room for the tile in shared memory, else the small one
  int warps = LOSS_WARPS_MAX;
  if (!team && fixed_bytes(LOSS_WARPS_MAX) + tile_smem_bytes(terrain) > PARC_SMEM_LIMIT) warps = LOSS_WARPS_MIN;
  const int teams = team ? 1 : warps;
  const size_t fixed = fixed_bytes(teams);
  const bool smem_tile = fixed + tile_smem_bytes(terrain) <= PARC_SMEM_LIMIT;
  const size_t smem = fixed + (smem_tile ? tile_smem_bytes(terrain) : 0);
  if (smem > PARC_SMEM_LIMIT) return PARC_E_SIZE;      // too many surface points for one CTA's slabs
  static SmemOptIn opt[4];
  if (team) rc = smem_tile ? ensure_dynamic_smem(body_loss_kernel<true, true>, opt[0], smem)
                           : ensure_dynamic_smem(body_loss_kernel<false, true>, opt[1], smem);
  else rc = smem_tile ? ensure_dynamic_smem(body_loss_kernel<true, false>, opt[2], smem)
                      : ensure_dynamic_smem(body_loss_kernel<false, false>, opt[3], smem);
  if (rc) return rc;
  int64_t rounds = (batch * frames) / ((int64_t)148 * 16 * warps);
  if (rounds < 1) rounds = 1;
  if (rounds > 8) rounds = 8;
  const int64_t fpc = team ? 1 : rounds * warps;
  p.frames_per_cta = (int)fpc;
  const int J = model->num_bodies;
  for (int64_t b0 = 0; b0 < batch; b0 += PARC_GRID_Y_MAX) {       // grid.y is limited to 65 535 samples per launch
    const int64_t nb = batch - b0 < PARC_GRID_Y_MAX ? batch - b0 : PARC_GRID_Y_MAX;
    const int64_t q0 = b0 * frames;
    p.batch = nb;
    p.terrain = terrain_from(*terrain, b0);
    p.root_pos = root_pos + q0 * root_pos_stride; p.root_pos_stride = root_pos_stride; p.root_rot = root_rot + q0 * 4;
    p.joint_rot = joint_rot ? joint_rot + q0 * (J - 1) * 4 : nullptr;
    p.contacts = contacts + q0 * J;
    p.pen_out = pen_out ? pen_out + q0 : nullptr; p.contact_out = contact_out ? contact_out + q0 : nullptr;
    p.g_root_pos = g_root_pos ? g_root_pos + q0 * 3 : nullptr;
    p.g_root_rot = g_root_rot ? g_root_rot + q0 * 4 : nullptr;
    p.g_joint_rot = g_joint_rot ? g_joint_rot + q0 * (J - 1) * 4 : nullptr;
    dim3 grid((unsigned)((frames + fpc - 1) / fpc), (unsigned)nb);
    if (team) {
      if (smem_tile) body_loss_kernel<true, true><<<grid, warps * 32, smem, (cudaStream_t)stream>>>(p, *model);
      else body_loss_kernel<false, true><<<grid, warps * 32, smem, (cudaStream_t)stream>>>(p, *model);
    } else {
      if (smem_tile) body_loss_kernel<true, false><<<grid, warps * 32, smem, (cudaStream_t)stream>>>(p, *model);
      else body_loss_kernel<false, false><<<grid, warps * 32, smem, (cudaStream_t)stream>>>(p, *model);
    }
  }
  return check_launch();
}
