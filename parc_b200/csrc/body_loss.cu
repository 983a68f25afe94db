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

namespace parc {

struct SdfBest {
  float inv;   // min over cells of sdBox to the AIR column   (result of inverted=True is -inv)
  float sol;   // min over cells of sdBox to the SOLID column
  int arg_inv;
  int arg_sol;
};

// Exact min over ALL cells of the box SDF, with pruning that cannot change the result.
//
// For a cell whose xy footprint does not contain the point (mx > 0 or my > 0, m = max(|p - c| - half, 0))
// the SDF is sqrt(mx^2 + my^2 + mz^2) >= sqrt(mx^2 + my^2) =: bound.  A cell (or a whole row ix, bound mx)
// whose bound is STRICTLY greater than the best value found so far can neither be the minimum nor tie
// with it, so it is skipped; cells whose footprint contains the point are always evaluated (their SDF can
// be negative).  The scan is seeded with the cell under the point.  Because evaluation order is no longer
// index order, the first-index tie rule of torch.min is kept explicitly: a candidate replaces the best
// iff it is smaller, or equal with a smaller flat index.  The bound test carries a 1e-6 relative margin
// for the rounding of best^2 (evaluating too many cells is always safe).
// hf/cx/cy are shared-memory arrays; a warp's threads are points of the same body, so their skip patterns
// mostly coincide.
template <bool WANT_INV, bool WANT_SOL>
__device__ __forceinline__ void eval_cell(const float* __restrict__ hf, int cell, float mxy2, float qxy, float pz,
                                          float base, float top, SdfBest& b) {
  const float h = hf[cell];
  if (WANT_INV) {
    const float cz = (h + top) * 0.5f;
    const float hz = (top - h) * 0.5f;
    const float qz = fabsf(pz - cz) - hz;
    const float mz = fmaxf(qz, 0.0f);
    const float sd = sqrtf(mxy2 + mz * mz) + fminf(fmaxf(qxy, qz), 0.0f);
    if (sd < b.inv || (sd == b.inv && cell < b.arg_inv)) { b.inv = sd; b.arg_inv = cell; }
  }
  if (WANT_SOL) {
    const float cz = (h + base) * 0.5f;
    const float hz = (h - base) * 0.5f;
    const float qz = fabsf(pz - cz) - hz;
    const float mz = fmaxf(qz, 0.0f);
    const float sd = sqrtf(mxy2 + mz * mz) + fminf(fmaxf(qxy, qz), 0.0f);
    if (sd < b.sol || (sd == b.sol && cell < b.arg_sol)) { b.sol = sd; b.arg_sol = cell; }
  }
}

// squared pruning threshold for a current best value (negative best: every outside cell is > best)
__device__ __forceinline__ float prune_thr(float best) {
  return best < 0.0f ? 0.0f : best * best * 1.000001f;
}

// Effective squared xy-reach for the wanted modes.  Every solid column tops out at or below the tile's
// maximum height H and every air column starts at or above the tile's minimum height L, so for a cell whose
// footprint does not contain the point
//     sdf_solid >= sqrt(mxy^2 + vz_sol^2),  vz_sol = max(pz - H, base - pz, 0)
//     sdf_air   >= sqrt(mxy^2 + vz_inv^2),  vz_inv = max(L - pz, pz - top, 0)
// and the cell can be skipped when mxy^2 > best^2 - vz^2 for every wanted mode.
template <bool WANT_INV, bool WANT_SOL>
__device__ __forceinline__ float prune_thr2(const SdfBest& b, float vz_inv2, float vz_sol2) {
  return fmaxf(WANT_INV ? prune_thr(b.inv) - vz_inv2 : -1.0f, WANT_SOL ? prune_thr(b.sol) - vz_sol2 : -1.0f);
}

// clamp-then-convert so that huge / NaN intermediate values cannot overflow the int conversion
__device__ __forceinline__ int to_index(float v, int hi) {
  return (int)fminf(fmaxf(v, 0.0f), (float)hi);
}

template <bool WANT_INV, bool WANT_SOL>
__device__ __forceinline__ SdfBest scan_cells(const float* __restrict__ hf, const float* __restrict__ cx,
                                              const float* __restrict__ cy, int X, int Y, float hx, float hy,
                                              float base, float hf_min, float hf_max, float3 p) {
  SdfBest b;
  b.inv = INFINITY; b.sol = INFINITY; b.arg_inv = 0x7fffffff; b.arg_sol = 0x7fffffff;
  const float top = -base;
  // vertical lower bounds, shrunk by a relative 1e-6 so rounding in the per-cell evaluation cannot beat them
  const float vzs = fmaxf(fmaxf(p.z - hf_max, base - p.z), 0.0f) * 0.999999f;
  const float vzi = fmaxf(fmaxf(hf_min - p.z, p.z - top), 0.0f) * 0.999999f;
  const float vz_sol2 = vzs * vzs, vz_inv2 = vzi * vzi;
  // cells are evenly spaced (torch.linspace nodes): spacing from the end points
  const float sx = X > 1 ? (cx[X - 1] - cx[0]) / (float)(X - 1) : 1.0f;
  const float sy = Y > 1 ? (cy[Y - 1] - cy[0]) / (float)(Y - 1) : 1.0f;
  const float isx = 1.0f / sx, isy = 1.0f / sy;
  // seed: the cell whose centre is nearest in xy
  {
    const int ix = to_index(rintf((p.x - cx[0]) * isx), X - 1);
    const int iy = to_index(rintf((p.y - cy[0]) * isy), Y - 1);
    const float qx = fabsf(p.x - cx[ix]) - hx, qy = fabsf(p.y - cy[iy]) - hy;
    const float mx = fmaxf(qx, 0.0f), my = fmaxf(qy, 0.0f);
    eval_cell<WANT_INV, WANT_SOL>(hf, ix * Y + iy, mx * mx + my * my, fmaxf(qx, qy), p.z, base, top, b);
  }
  float thr = prune_thr2<WANT_INV, WANT_SOL>(b, vz_inv2, vz_sol2);
  // Index window that is a SUPERSET of every cell that can still matter: a cell further than r = sqrt(thr)
  // (+ its half width) from the point in x or in y is out of reach.  One extra cell of margin on each side
  // absorbs the rounding of the spacing; cells inside the window are still bound-checked one by one.
  const float r = sqrtf(fmaxf(thr, 0.0f));
  const int ix_lo = to_index(floorf((p.x - r - hx - cx[0]) * isx) - 1.0f, X - 1);
  const int ix_hi = to_index(ceilf((p.x + r + hx - cx[0]) * isx) + 1.0f, X - 1);
  const int iy_lo = to_index(floorf((p.y - r - hy - cy[0]) * isy) - 1.0f, Y - 1);
  const int iy_hi = to_index(ceilf((p.y + r + hy - cy[0]) * isy) + 1.0f, Y - 1);
  for (int ix = ix_lo; ix <= ix_hi; ++ix) {
    const float qx = fabsf(p.x - cx[ix]) - hx;
    const float mx = fmaxf(qx, 0.0f);
    const float mx2 = mx * mx;
    if (mx2 > 0.0f && mx2 > thr) continue;         // whole row out of reach
    for (int iy = iy_lo; iy <= iy_hi; ++iy) {
      const float qy = fabsf(p.y - cy[iy]) - hy;
      const float my = fmaxf(qy, 0.0f);
      const float mxy2 = mx2 + my * my;
      if (mxy2 > 0.0f && mxy2 > thr) continue;     // outside the footprint and strictly out of reach
      const float old_inv = b.inv, old_sol = b.sol;
      eval_cell<WANT_INV, WANT_SOL>(hf, ix * Y + iy, mxy2, fmaxf(qx, qy), p.z, base, top, b);
      if (b.inv < old_inv || b.sol < old_sol) thr = prune_thr2<WANT_INV, WANT_SOL>(b, vz_inv2, vz_sol2);
    }
  }
  return b;
}

__device__ __forceinline__ float sgn(float v) { return (v > 0.0f) ? 1.0f : ((v < 0.0f) ? -1.0f : 0.0f); }

// d sdBox(p - c, half) / d p for one cell, following autograd's sub-gradient conventions
// (SURVEY A10): clamp passes the gradient at the bound, norm'(0) = 0, abs'(0) = 0, max -> first index.
__device__ __forceinline__ float3 sd_box_grad(float3 d, float3 half) {
  const float qx = fabsf(d.x) - half.x, qy = fabsf(d.y) - half.y, qz = fabsf(d.z) - half.z;
  const float mq = fmaxf(qx, fmaxf(qy, qz));
  float3 g = make_float3(0.f, 0.f, 0.f);
  if (mq > 0.0f) {
    const float mx = fmaxf(qx, 0.f), my = fmaxf(qy, 0.f), mz = fmaxf(qz, 0.f);
    const float n = sqrtf(mx * mx + my * my + mz * mz);
    if (n > 0.0f) {
      g.x = mx / n * sgn(d.x);
      g.y = my / n * sgn(d.y);
      g.z = mz / n * sgn(d.z);
    }
  } else {
    if (qx >= qy && qx >= qz) g.x = sgn(d.x);
    else if (qy >= qz) g.y = sgn(d.y);
    else g.z = sgn(d.z);
  }
  return g;
}

__device__ __forceinline__ float3 cell_grad(const float* hf, const float* cx, const float* cy, int Y, float hx,
                                            float hy, float base, bool inverted, int cell, float3 p) {
  const int ix = cell / Y, iy = cell - ix * Y;
  const float h = hf[cell];
  float cz, hz;
  if (inverted) { const float top = -base; cz = (h + top) * 0.5f; hz = (top - h) * 0.5f; }
  else { cz = (h + base) * 0.5f; hz = (h - base) * 0.5f; }
  return sd_box_grad(make_float3(p.x - cx[ix], p.y - cy[iy], p.z - cz), make_float3(hx, hy, hz));
}

// Stage one sample's terrain: hf tile, absolute cell-centre coordinates (node offset + min centre, added
// in fp32 as util/terrain_util.py:1859-1860 does), and the tile's min / max height for the pruning bounds
// (s_minmax[0] = min, [1] = max; must be followed by __syncthreads()).
__device__ __forceinline__ void stage_terrain(const ParcTerrainBatch& t, int64_t b, float* s_hf, float* s_cx,
                                              float* s_cy, float* s_minmax) {
  const int X = t.dim_x, Y = t.dim_y;
  const float* hf = t.hf + b * t.hf_batch_stride;
  const float* mc = t.min_center + b * t.min_center_stride;
  if (threadIdx.x == 0) { s_minmax[0] = INFINITY; s_minmax[1] = -INFINITY; }
  __syncthreads();
  float lo = INFINITY, hi = -INFINITY;
  for (int i = threadIdx.x; i < X * Y; i += blockDim.x) {
    const float h = __ldg(hf + i);
    s_hf[i] = h;
    lo = fminf(lo, h); hi = fmaxf(hi, h);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(PARC_FULL_MASK, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(PARC_FULL_MASK, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    // float min / max through the int-ordered trick (values may be negative)
    if (lo >= 0.0f) atomicMin(reinterpret_cast<int*>(&s_minmax[0]), __float_as_int(lo));
    else atomicMax(reinterpret_cast<unsigned int*>(&s_minmax[0]), __float_as_uint(lo));
    if (hi >= 0.0f) atomicMax(reinterpret_cast<int*>(&s_minmax[1]), __float_as_int(hi));
    else atomicMin(reinterpret_cast<unsigned int*>(&s_minmax[1]), __float_as_uint(hi));
  }
  for (int i = threadIdx.x; i < X; i += blockDim.x) s_cx[i] = __ldg(t.x_nodes + i) + __ldg(mc);
  for (int i = threadIdx.x; i < Y; i += blockDim.x) s_cy[i] = __ldg(t.y_nodes + i) + __ldg(mc + 1);
}

__device__ __forceinline__ float sample_base_z(const ParcTerrainBatch& t, int64_t b) {
  return t.base_z ? __ldg(t.base_z + b * t.base_z_stride) : t.base_z_value;
}

// ------------------------------------------------------------------------------------------------
// a13 stand-alone: points [B,N,3] -> sdf [B,N]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
points_hf_sdf_kernel(const float* __restrict__ points, int64_t n_points, const __grid_constant__ ParcTerrainBatch t,
                     int inverted, float* __restrict__ sdf, int32_t* __restrict__ arg) {
  extern __shared__ float smem[];
  const int X = t.dim_x, Y = t.dim_y;
  float* s_hf = smem;
  float* s_cx = s_hf + X * Y;
  float* s_cy = s_cx + X;
  __shared__ float s_minmax[2];
  const int64_t b = blockIdx.y;
  stage_terrain(t, b, s_hf, s_cx, s_cy, s_minmax);
  __syncthreads();
  const float base = sample_base_z(t, b);
  const float hf_min = s_minmax[0], hf_max = s_minmax[1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_points; i += (int64_t)gridDim.x * blockDim.x) {
    const float* pp = points + (b * n_points + i) * 3;
    const float3 p = make_float3(__ldg(pp), __ldg(pp + 1), __ldg(pp + 2));
    float v;
    int a;
    if (inverted) {
      const SdfBest r = scan_cells<true, false>(s_hf, s_cx, s_cy, X, Y, t.half_dx, t.half_dy, base, hf_min, hf_max, p);
      v = -1.0f * r.inv; a = r.arg_inv;
    } else {
      const SdfBest r = scan_cells<false, true>(s_hf, s_cx, s_cy, X, Y, t.half_dx, t.half_dy, base, hf_min, hf_max, p);
      v = r.sol; a = r.arg_sol;
    }
    sdf[b * n_points + i] = v;
    if (arg) arg[b * n_points + i] = a;
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
#define LOSS_WARPS 4
#define LOSS_THREADS (LOSS_WARPS * 32)

struct BodyLossParams {
  const float *root_pos, *root_rot, *joint_rot, *contacts;
  int64_t batch, frames;
  ParcBodyPoints pts;
  ParcTerrainBatch terrain;
  float w_pen, w_contact;
  float *pen_out, *contact_out;
  float *g_root_pos, *g_root_rot, *g_joint_rot;
  int frames_per_cta;
  int want_grad;
};

__global__ void __launch_bounds__(LOSS_THREADS)
body_loss_kernel(const __grid_constant__ BodyLossParams p, const __grid_constant__ ParcCharModel model_param) {
  extern __shared__ float smem[];
  __shared__ ParcCharModel sm;
  __shared__ float s_minmax[2];

  const int X = p.terrain.dim_x, Y = p.terrain.dim_y;
  const int S = p.pts.num_points;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* s_hf = smem;
  float* s_cx = s_hf + X * Y;
  float* s_cy = s_cx + X;
  int* s_body = reinterpret_cast<int*>(s_cy + Y);                    // [S] body of point
  float* s_lp = reinterpret_cast<float*>(s_body + S);                // [S][3] local points
  float* slab = s_lp + (size_t)S * 3 + (size_t)warp * ((size_t)S * 7 + PARC_MAX_BODIES * 9);
  float* s_bt = slab;                                                // [J][9] pos(3) rot(4) contact(1) winner(1)
  float* s_pt = slab + PARC_MAX_BODIES * 9;                          // [S][7]

  const int64_t b = blockIdx.y;
  stage_model(&sm, model_param);
  stage_terrain(p.terrain, b, s_hf, s_cx, s_cy, s_minmax);
  __syncthreads();
  const float hf_min = s_minmax[0], hf_max = s_minmax[1];
  const int J = sm.num_bodies;
  for (int j = threadIdx.x; j < J; j += blockDim.x) {
    const int s0 = __ldg(p.pts.point_start + j), s1 = __ldg(p.pts.point_start + j + 1);
    for (int k = s0; k < s1; ++k) s_body[k] = j;
  }
  for (int i = threadIdx.x; i < S * 3; i += blockDim.x) s_lp[i] = __ldg(p.pts.points + i);
  const float base = sample_base_z(p.terrain, b);
  const float hx = p.terrain.half_dx, hy = p.terrain.half_dy;
  const LaneBody lb = load_lane_body(sm, lane, 0);
  const int max_depth = sm.max_depth;
  const int my_s0 = lane < J ? __ldg(p.pts.point_start + lane) : 0;
  const int my_s1 = lane < J ? __ldg(p.pts.point_start + lane + 1) : 0;
  __syncthreads();

  const int64_t f_begin = (int64_t)blockIdx.x * p.frames_per_cta;
  const int64_t f_end = min(f_begin + (int64_t)p.frames_per_cta, p.frames);
  for (int64_t f = f_begin + warp; f < f_end; f += LOSS_WARPS) {
    const int64_t q = b * p.frames + f;
    // ---- FK, lane = body ----
    float4 prot, local, rot = make_float4(0.f, 0.f, 0.f, 1.f);
    float3 pos = make_float3(0.f, 0.f, 0.f);
    if (lane == 0) {
      pos = make_float3(__ldg(p.root_pos + q * 3), __ldg(p.root_pos + q * 3 + 1), __ldg(p.root_pos + q * 3 + 2));
      rot = __ldg(reinterpret_cast<const float4*>(p.root_rot) + q);
    } else if (lane < J) {
      rot = __ldg(reinterpret_cast<const float4*>(p.joint_rot) + q * (J - 1) + (lane - 1));
    }
    fk_warp_keep(lb, max_depth, pos, rot, prot, local);
    float my_contact = 0.0f;
    if (lane < J) {
      my_contact = __ldg(p.contacts + q * J + lane);
      float* t = s_bt + lane * 9;
      t[0] = pos.x; t[1] = pos.y; t[2] = pos.z; t[3] = rot.x; t[4] = rot.y; t[5] = rot.z; t[6] = rot.w;
      t[7] = my_contact;
    }
    __syncwarp();

    // ---- sweep, lane = surface point ----
    float pen_local = 0.0f;
    for (int k = lane; k < S; k += 32) {
      const int bj = s_body[k];
      const float* t = s_bt + bj * 9;
      const float3 lp = make_float3(s_lp[k * 3], s_lp[k * 3 + 1], s_lp[k * 3 + 2]);
      const float4 br = make_float4(t[3], t[4], t[5], t[6]);
      const float3 r = quat_rotate(br, lp);
      const float3 wp = make_float3(r.x + t[0], r.y + t[1], r.z + t[2]);
      // A body whose contact weight is exactly 0 contributes exactly 0 to the contact term and to its
      // gradient (closest * 0), so its solid-column scan is skipped.
      SdfBest best;
      if (t[7] != 0.0f) {
        best = scan_cells<true, true>(s_hf, s_cx, s_cy, X, Y, hx, hy, base, hf_min, hf_max, wp);
      } else {
        best = scan_cells<true, false>(s_hf, s_cx, s_cy, X, Y, hx, hy, base, hf_min, hf_max, wp);
        best.sol = 0.0f; best.arg_sol = 0;
      }
      // penetration: sdf = -best.inv ; neg = min(sdf, 0) ; pen += -neg
      const float sdf_inv = -1.0f * best.inv;
      pen_local += -fminf(sdf_inv, 0.0f);
      float* g7 = s_pt + (size_t)k * 7;
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
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pen_local += __shfl_xor_sync(PARC_FULL_MASK, pen_local, o);
    __syncwarp();

    // ---- contact term, lane = body: first-index min over the body's points ----
    float cterm = 0.0f;
    int win = -1;
    if (lane < J) {
      float bestv = INFINITY;
      for (int k = my_s0; k < my_s1; ++k) {
        const float v = s_pt[(size_t)k * 7 + 6];
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
    __syncwarp();
    if (!p.want_grad) continue;

    // ---- chain rule, lane = point: add the winner's contact gradient, VJP through the body transform ----
    for (int k = lane; k < S; k += 32) {
      const int bj = s_body[k];
      const float* t = s_bt + bj * 9;
      float* g7 = s_pt + (size_t)k * 7;
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
    __syncwarp();

    // ---- body sums (fixed order), lane = body; then the FK VJP ----
    float3 gp = make_float3(0.f, 0.f, 0.f);
    float4 gr = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = my_s0; k < my_s1; ++k) {
      const float* g7 = s_pt + (size_t)k * 7;
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
    __syncwarp();
  }
}

static size_t terrain_smem_bytes(const ParcTerrainBatch* t) {
  return ((size_t)t->dim_x * t->dim_y + t->dim_x + t->dim_y) * sizeof(float);
}

static int check_terrain(const ParcTerrainBatch* t) {
  if (!t || !t->hf || !t->min_center || !t->x_nodes || !t->y_nodes) return PARC_E_NULL;
  if (t->dim_x <= 0 || t->dim_y <= 0 || t->hf_batch_stride < 0) return PARC_E_SIZE;
  return PARC_OK;
}

}  // namespace parc

using namespace parc;

extern "C" int parc_points_hf_sdf(const float* points, int64_t batch, int64_t n_points,
                                  const ParcTerrainBatch* terrain, int32_t inverted, float* sdf_out,
                                  int32_t* arg_out, void* stream) {
  if (batch < 0 || n_points < 0 || batch > 65535) return PARC_E_SIZE;
  if (batch == 0 || n_points == 0) return PARC_OK;
  if (!points || !sdf_out) return PARC_E_NULL;
  int rc = check_terrain(terrain);
  if (rc) return rc;
  const size_t smem = terrain_smem_bytes(terrain);
  if (smem > 200 * 1024) return PARC_E_SIZE;          // terrain tile must fit one SM's shared memory
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(points_hf_sdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  int64_t gx = (n_points + 255) / 256;
  if (gx > 4096) gx = 4096;
  dim3 grid((unsigned)gx, (unsigned)batch);
  points_hf_sdf_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(points, n_points, *terrain, inverted, sdf_out,
                                                                  arg_out);
  return check_launch();
}

extern "C" int parc_body_loss(const float* root_pos, const float* root_rot, const float* joint_rot,
                              const float* contacts, int64_t batch, int64_t frames, const ParcCharModel* model,
                              const ParcBodyPoints* pts, const ParcTerrainBatch* terrain, float w_pen,
                              float w_contact, float* pen_out, float* contact_out, float* g_root_pos,
                              float* g_root_rot, float* g_joint_rot, void* stream) {
  if (!model || !pts) return PARC_E_NULL;
  int rc = parc_validate_model(model);
  if (rc) return rc;
  if (batch < 0 || frames < 0 || batch > 65535 || pts->num_points <= 0) return PARC_E_SIZE;
  if (batch == 0 || frames == 0) return PARC_OK;
  if (!root_pos || !root_rot || !contacts || !pts->points || !pts->point_start) return PARC_E_NULL;
  if (model->num_bodies > 1 && !joint_rot) return PARC_E_NULL;
  rc = check_terrain(terrain);
  if (rc) return rc;
  if (!aligned16(root_rot) || !aligned16(joint_rot) || !aligned16(g_root_rot) || !aligned16(g_joint_rot))
    return PARC_E_ALIGN;
  if (batch == 0 || frames == 0) return PARC_OK;

  BodyLossParams p;
  p.root_pos = root_pos; p.root_rot = root_rot; p.joint_rot = joint_rot; p.contacts = contacts;
  p.batch = batch; p.frames = frames; p.pts = *pts; p.terrain = *terrain;
  p.w_pen = w_pen; p.w_contact = w_contact; p.pen_out = pen_out; p.contact_out = contact_out;
  p.g_root_pos = g_root_pos; p.g_root_rot = g_root_rot; p.g_joint_rot = g_joint_rot;
  p.want_grad = (g_root_pos || g_root_rot || g_joint_rot) ? 1 : 0;

  const size_t S = (size_t)pts->num_points;
  const size_t smem = terrain_smem_bytes(terrain) + (S * (1 + 3) + LOSS_WARPS * (S * 7 + PARC_MAX_BODIES * 9)) * sizeof(float);
  if (smem > 200 * 1024) return PARC_E_SIZE;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(body_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  // one warp per frame, LOSS_WARPS frames in flight per CTA; amortise the terrain staging over several
  // rounds when there is plenty of work, keep the grid wide when there is not
  int64_t rounds = (batch * frames) / ((int64_t)148 * 16 * LOSS_WARPS);
  if (rounds < 1) rounds = 1;
  if (rounds > 8) rounds = 8;
  const int64_t fpc = rounds * LOSS_WARPS;
  p.frames_per_cta = (int)fpc;
  dim3 grid((unsigned)((frames + fpc - 1) / fpc), (unsigned)batch);
  body_loss_kernel<<<grid, LOSS_THREADS, smem, (cudaStream_t)stream>>>(p, *model);
  return check_launch();
}
