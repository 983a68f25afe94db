// Exact point <-> heightfield box SDF (min over ALL cells, with result-preserving pruning), its sub-gradient,
// and the terrain-tile staging shared by body_loss.cu and dataset_sweep.cu.
// Reference: util/terrain_util.py:1835-1893 (points_hf_sdf), util/geom_util.py:122-143 (sdBox).
#pragma once

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
// for the rounding of the squares and sums (evaluating too many cells is always safe); the vertical part of the
// bound is bit-exact (mz_bound_solid / mz_bound_air).
// hf/cx/cy are shared-memory arrays; a warp's threads are points of the same body, so their skip patterns
// mostly coincide.
template <bool WANT_INV, bool WANT_SOL>
__device__ __forceinline__ void eval_cell(const float* __restrict__ hf, int cell, float mxy2, float qxy, float pz,
                                          float base, float top, SdfBest& b, bool do_inv = true, bool do_sol = true) {
  const float h = hf[cell];
  if (WANT_INV && do_inv) {
    const float cz = (h + top) * 0.5f;
    const float hz = (top - h) * 0.5f;
    const float qz = fabsf(pz - cz) - hz;
    const float mz = fmaxf(qz, 0.0f);
    const float sd = sqrtf(mxy2 + mz * mz) + fminf(fmaxf(qxy, qz), 0.0f);
    if (sd < b.inv || (sd == b.inv && cell < b.arg_inv)) { b.inv = sd; b.arg_inv = cell; }
  }
  if (WANT_SOL && do_sol) {
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

// Lower bounds of the clamped vertical part mz = max(qz, 0) of sdBox, for EVERY column of a set whose heights are
// bounded by h_ext, written as the very float expressions eval_cell evaluates: each step ((h + base) / 2, (h - base) / 2,
// pz - cz, ... - hz) is a correctly rounded monotone function of h, so qz_float(h) >= qz_float(h_ext) holds bit for
// bit -- no slack for the rounding of the evaluation is needed (a relative margin cannot cover it: cz and hz are ~5 m
// when base is -10 m, so qz carries ~1e-6 m of absolute rounding whatever its size).
//   solid column [base, h], h <= h_max: valid while the point is at or above the tallest column's centre
//   air column   [h, top],  h >= h_min: valid while the point is at or below the shortest air column's centre
__device__ __forceinline__ float mz_bound_solid(float pz, float h_max, float base) {
  const float cz = (h_max + base) * 0.5f, hz = (h_max - base) * 0.5f;
  const float d = pz - cz;
  return d >= 0.0f ? fmaxf(fabsf(d) - hz, 0.0f) : 0.0f;
}
__device__ __forceinline__ float mz_bound_air(float pz, float h_min, float top) {
  const float cz = (h_min + top) * 0.5f, hz = (top - h_min) * 0.5f;
  const float d = pz - cz;
  return d <= 0.0f ? fmaxf(fabsf(d) - hz, 0.0f) : 0.0f;
}

// clamp-then-convert so that huge / NaN intermediate values cannot overflow the int conversion
__device__ __forceinline__ int to_index(float v, int hi) {
  return (int)fminf(fmaxf(v, 0.0f), (float)hi);
}

// Per-block height range of a tile (blocks of PARC_SDF_BLOCK x PARC_SDF_BLOCK cells): the vertical part of the pruning
// bound taken per block instead of per tile.  On terrains with a few tall boxes the tile-wide maximum says nothing
// about the flat ground around a point; the block maximum does, and whole blocks drop out of the scan.
#define PARC_SDF_BLOCK 4
struct TileBlocks {
  const float* bmax;   // [nbx * nby] maximum height of each block, or nullptr: no per-block bounds
  const float* bmin;   // [nbx * nby] minimum height
  int nby;
};
__host__ __device__ __forceinline__ int sdf_blocks(int n) { return (n + PARC_SDF_BLOCK - 1) / PARC_SDF_BLOCK; }

// one axis of max(|p - c| - half, 0) for a cell: the exact expression the per-cell test uses
__device__ __forceinline__ float cell_q(float p, float c, float half) { return fabsf(p - c) - half; }

template <bool WANT_INV, bool WANT_SOL>
__device__ __forceinline__ SdfBest scan_cells(const float* __restrict__ hf, const float* __restrict__ cx,
                                              const float* __restrict__ cy, int X, int Y, float hx, float hy,
                                              float base, float hf_min, float hf_max, float3 p, bool sol_on = true,
                                              TileBlocks tb = TileBlocks{nullptr, nullptr, 0}) {
  SdfBest b;
  b.inv = INFINITY; b.sol = INFINITY; b.arg_inv = 0x7fffffff; b.arg_sol = 0x7fffffff;
  const float top = -base;
  const bool want_sol = WANT_SOL && sol_on;
  // tile-wide vertical bounds (squared)
  const float vzs = mz_bound_solid(p.z, hf_max, base), vzi = mz_bound_air(p.z, hf_min, top);
  const float vz_sol2 = vzs * vzs, vz_inv2 = vzi * vzi;
  // cells are evenly spaced (torch.linspace nodes): spacing from the end points
  const float sx = X > 1 ? (cx[X - 1] - cx[0]) / (float)(X - 1) : 1.0f;
  const float sy = Y > 1 ? (cy[Y - 1] - cy[0]) / (float)(Y - 1) : 1.0f;
  const float isx = 1.0f / sx, isy = 1.0f / sy;
  // seed: the cell whose centre is nearest in xy
  {
    const int ix = to_index(rintf((p.x - cx[0]) * isx), X - 1);
    const int iy = to_index(rintf((p.y - cy[0]) * isy), Y - 1);
    const float qx = cell_q(p.x, cx[ix], hx), qy = cell_q(p.y, cy[iy], hy);
    const float mx = fmaxf(qx, 0.0f), my = fmaxf(qy, 0.0f);
    eval_cell<WANT_INV, WANT_SOL>(hf, ix * Y + iy, mx * mx + my * my, fmaxf(qx, qy), p.z, base, top, b, true, want_sol);
  }
  // One reach per mode: a cell (block, column run) whose footprint does not contain the point is skipped for a mode when
  //     mxy^2 + mz_bound^2  >  best^2 * (1 + 1e-6)          (lim = prune_thr(best); 0 for a negative best)
  // i.e. when even the bound exceeds the best value; it is visited when EITHER mode can still use it, and each mode is
  // evaluated only where its own bound allows -- the air-column SDF of a point above ground is settled by the cell
  // under it (its value is negative), so the wide window a distant solid column needs does not drag the air
  // evaluation along.
  float lim_i = WANT_INV ? prune_thr(b.inv) : -1.0f;
  float lim_s = want_sol ? prune_thr(b.sol) : -1.0f;
  // Index window that is a SUPERSET of every cell that can still matter: a cell further than r (+ its half width)
  // from the point in x or in y is out of reach.  One extra cell of margin on each side absorbs the rounding of the
  // spacing; cells inside the window are still bound-checked one by one.
  const float r = sqrtf(fmaxf(fmaxf(lim_i - vz_inv2, lim_s - vz_sol2), 0.0f)) * 1.000001f;
  const int ix_lo = to_index(floorf((p.x - r - hx - cx[0]) * isx) - 1.0f, X - 1);
  const int ix_hi = to_index(ceilf((p.x + r + hx - cx[0]) * isx) + 1.0f, X - 1);
  const int iy_lo = to_index(floorf((p.y - r - hy - cy[0]) * isy) - 1.0f, Y - 1);
  const int iy_hi = to_index(ceilf((p.y + r + hy - cy[0]) * isy) + 1.0f, Y - 1);
  // The window is walked block by block.  Without per-block bounds a "block" is the whole window with the tile-wide
  // height range, which is the plain cell-by-cell scan.
  const bool blocked = tb.bmax != nullptr;
  const int bk = PARC_SDF_BLOCK;
  const int bx_lo = blocked ? ix_lo / bk : 0, bx_hi = blocked ? ix_hi / bk : 0;
  const int by_lo = blocked ? iy_lo / bk : 0, by_hi = blocked ? iy_hi / bk : 0;
  // out(m2, v2, lim): xy distance^2 m2 > 0 and the bound m2 + v2 beyond reach
#define PARC_OUT(m2, vi2, vs2) ((m2) > 0.0f && ((m2) + (vi2) > lim_i) && ((m2) + (vs2) > lim_s))
  for (int bx = bx_lo; bx <= bx_hi; ++bx) {
    const int x0 = blocked ? max(bx * bk, ix_lo) : ix_lo, x1 = blocked ? min(bx * bk + bk - 1, ix_hi) : ix_hi;
    // distance in x to the block's footprint = the per-cell expression of its nearest column (same bits)
    const float mbx = p.x < cx[x0] ? fmaxf(cell_q(p.x, cx[x0], hx), 0.0f)
                                   : (p.x > cx[x1] ? fmaxf(cell_q(p.x, cx[x1], hx), 0.0f) : 0.0f);
    const float mbx2 = mbx * mbx;
    if (PARC_OUT(mbx2, vz_inv2, vz_sol2)) continue;       // every column of the block row is out of reach
    for (int by = by_lo; by <= by_hi; ++by) {
      const int y0 = blocked ? max(by * bk, iy_lo) : iy_lo, y1 = blocked ? min(by * bk + bk - 1, iy_hi) : iy_hi;
      float vs2 = vz_sol2, vi2 = vz_inv2;
      if (blocked) {
        const float mby = p.y < cy[y0] ? fmaxf(cell_q(p.y, cy[y0], hy), 0.0f)
                                       : (p.y > cy[y1] ? fmaxf(cell_q(p.y, cy[y1], hy), 0.0f) : 0.0f);
        const float mb2 = mbx2 + mby * mby;
        const float vs = mz_bound_solid(p.z, tb.bmax[bx * tb.nby + by], base);
        const float vi = mz_bound_air(p.z, tb.bmin[bx * tb.nby + by], top);
        vs2 = vs * vs; vi2 = vi * vi;
        // every cell of the block is at least mb2 away in xy; its column ends at or below the block's maximum and its
        // air column starts at or above the block's minimum
        if (PARC_OUT(mb2, vi2, vs2)) continue;
      }
      for (int ix = x0; ix <= x1; ++ix) {
        const float qx = cell_q(p.x, cx[ix], hx);
        const float mx = fmaxf(qx, 0.0f);
        const float mx2 = mx * mx;
        if (PARC_OUT(mx2, vi2, vs2)) continue;            // whole column run out of reach
        for (int iy = y0; iy <= y1; ++iy) {
          const float qy = cell_q(p.y, cy[iy], hy);
          const float my = fmaxf(qy, 0.0f);
          const float mxy2 = mx2 + my * my;
          const bool inside = !(mxy2 > 0.0f);
          const bool do_inv = WANT_INV && (inside || !(mxy2 + vi2 > lim_i));
          const bool do_sol = want_sol && (inside || !(mxy2 + vs2 > lim_s));
          if (!do_inv && !do_sol) continue;               // outside the footprint and strictly out of reach
          const float old_inv = b.inv, old_sol = b.sol;
          eval_cell<WANT_INV, WANT_SOL>(hf, ix * Y + iy, mxy2, fmaxf(qx, qy), p.z, base, top, b, do_inv, do_sol);
          if (b.inv < old_inv) lim_i = prune_thr(b.inv);
          if (b.sol < old_sol) lim_s = prune_thr(b.sol);
        }
      }
    }
  }
#undef PARC_OUT
  return b;
}

// Block height ranges of the tile staged in s_hf -> s_bmax / s_bmin ([sdf_blocks(X) * sdf_blocks(Y)] each).  Call
// after the tile is visible to the CTA (a __syncthreads() behind stage_terrain) and follow with another barrier.
__device__ __forceinline__ void stage_tile_blocks(const float* __restrict__ s_hf, int X, int Y, float* s_bmax,
                                                  float* s_bmin) {
  const int nbx = sdf_blocks(X), nby = sdf_blocks(Y);
  for (int i = threadIdx.x; i < nbx * nby; i += blockDim.x) {
    const int bx = i / nby, by = i - bx * nby;
    float hi = -INFINITY, lo = INFINITY;
    for (int ix = bx * PARC_SDF_BLOCK; ix < min((bx + 1) * PARC_SDF_BLOCK, X); ++ix)
      for (int iy = by * PARC_SDF_BLOCK; iy < min((by + 1) * PARC_SDF_BLOCK, Y); ++iy) {
        const float h = s_hf[ix * Y + iy];
        hi = fmaxf(hi, h); lo = fminf(lo, h);
      }
    s_bmax[i] = hi; s_bmin[i] = lo;
  }
}

// Inverse cell spacing of a tile whose centre coordinates are staged in cx / cy (evenly spaced torch.linspace nodes):
// the same expressions scan_cells evaluates per call, hoisted for callers that scan many points of one tile.
struct TileSpacing {
  float isx, isy;
};
__device__ __forceinline__ TileSpacing tile_spacing(const float* __restrict__ cx, const float* __restrict__ cy, int X,
                                                    int Y) {
  const float sx = X > 1 ? (cx[X - 1] - cx[0]) / (float)(X - 1) : 1.0f;
  const float sy = Y > 1 ? (cy[Y - 1] - cy[0]) / (float)(Y - 1) : 1.0f;
  TileSpacing t;
  t.isx = 1.0f / sx; t.isy = 1.0f / sy;
  return t;
}

// Warp-cooperative form of scan_cells for the SOLID columns, value only (the labelling kernel thresholds it): every
// lane evaluates the seed cell, the index window that is a superset of all cells within reach of the seed's value is
// split over the 32 lanes, and the lanes' minima are combined.  The minimum of a set of floats does not depend on the
// order it is taken in, so this is the same number scan_cells returns in `sol`.  Must be called by the whole warp.
__device__ __forceinline__ float warp_min_solid_sdf(const float* __restrict__ hf, const float* __restrict__ cx,
                                                    const float* __restrict__ cy, int X, int Y, float hx, float hy,
                                                    float base, float hf_max, const TileSpacing& sp, float3 p,
                                                    int lane) {
  SdfBest b;
  b.inv = INFINITY; b.sol = INFINITY; b.arg_inv = 0x7fffffff; b.arg_sol = 0x7fffffff;
  const float top = -base;
  const float vzs = mz_bound_solid(p.z, hf_max, base);
  const float vz_sol2 = vzs * vzs;
  const float isx = sp.isx, isy = sp.isy;
  {
    const int ix = to_index(rintf((p.x - cx[0]) * isx), X - 1);
    const int iy = to_index(rintf((p.y - cy[0]) * isy), Y - 1);
    const float qx = fabsf(p.x - cx[ix]) - hx, qy = fabsf(p.y - cy[iy]) - hy;
    const float mx = fmaxf(qx, 0.0f), my = fmaxf(qy, 0.0f);
    eval_cell<false, true>(hf, ix * Y + iy, mx * mx + my * my, fmaxf(qx, qy), p.z, base, top, b);
  }
  float lim = prune_thr(b.sol);                    // skip when mxy^2 + vz^2 > best^2 (1 + 1e-6)
  const float r = sqrtf(fmaxf(lim - vz_sol2, 0.0f)) * 1.000001f;
  const int ix_lo = to_index(floorf((p.x - r - hx - cx[0]) * isx) - 1.0f, X - 1);
  const int ix_hi = to_index(ceilf((p.x + r + hx - cx[0]) * isx) + 1.0f, X - 1);
  const int iy_lo = to_index(floorf((p.y - r - hy - cy[0]) * isy) - 1.0f, Y - 1);
  const int iy_hi = to_index(ceilf((p.y + r + hy - cy[0]) * isy) + 1.0f, Y - 1);
  const int wy = iy_hi - iy_lo + 1;
  const int window = (ix_hi - ix_lo + 1) * wy;
  // c / wy through one multiplication: (c + 0.5) / wy is at least 0.5 / wy away from an integer and c < 2^24, so the
  // truncated product is the exact quotient up to one unit, which the remainder check below repairs
  const float inv_wy = 1.0f / (float)wy;
  for (int c = lane; c < window; c += 32) {
    int rx = (int)(((float)c + 0.5f) * inv_wy);
    int ry = c - rx * wy;
    if (ry < 0) { --rx; ry += wy; } else if (ry >= wy) { ++rx; ry -= wy; }
    const int ix = ix_lo + rx, iy = iy_lo + ry;
    const float qx = fabsf(p.x - cx[ix]) - hx, qy = fabsf(p.y - cy[iy]) - hy;
    const float mx = fmaxf(qx, 0.0f), my = fmaxf(qy, 0.0f);
    const float mxy2 = mx * mx + my * my;
    if (mxy2 > 0.0f && mxy2 + vz_sol2 > lim) continue;       // outside the footprint and strictly out of reach
    const float old = b.sol;
    eval_cell<false, true>(hf, ix * Y + iy, mxy2, fmaxf(qx, qy), p.z, base, top, b);
    if (b.sol < old) lim = prune_thr(b.sol);
  }
  float m = b.sol;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(PARC_FULL_MASK, m, o));
  return m;
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

// float atomic min / max through the int-ordering trick.  The branch is on the SIGN BIT, not on v >= 0: -0.0f has
// the bit pattern 0x80000000 = INT_MIN, which the signed compare would treat as the smallest value of all.
__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
  if (__float_as_int(v) >= 0) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (__float_as_int(v) >= 0) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// Absolute cell-centre coordinates of one sample's terrain (node offset + min centre, added in fp32 as
// util/terrain_util.py:1859-1860 does) -> shared memory.
__device__ __forceinline__ void stage_terrain_nodes(const ParcTerrainBatch& t, int64_t b, float* s_cx, float* s_cy) {
  const float* mc = t.min_center + b * t.min_center_stride;
  for (int i = threadIdx.x; i < t.dim_x; i += blockDim.x) s_cx[i] = __ldg(t.x_nodes + i) + __ldg(mc);
  for (int i = threadIdx.x; i < t.dim_y; i += blockDim.x) s_cy[i] = __ldg(t.y_nodes + i) + __ldg(mc + 1);
}

// Stage one sample's terrain: hf tile, cell-centre coordinates, and the tile's min / max height for the pruning
// bounds (s_minmax[0] = min, [1] = max; must be followed by __syncthreads()).
__device__ __forceinline__ void stage_terrain(const ParcTerrainBatch& t, int64_t b, float* s_hf, float* s_cx,
                                              float* s_cy, float* s_minmax) {
  const int X = t.dim_x, Y = t.dim_y;
  const float* hf = t.hf + b * t.hf_batch_stride;
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
    atomic_min_float(&s_minmax[0], lo);
    atomic_max_float(&s_minmax[1], hi);
  }
  stage_terrain_nodes(t, b, s_cx, s_cy);
}

// A terrain too large for shared memory stays in global memory (read through L1 / L2): only the cell-centre
// coordinates are staged, and the pruning runs without the tile's height range (min = -inf, max = +inf: the vertical
// bound degenerates to 0, which only makes the scan visit more cells -- never fewer than the exact result needs).
__device__ __forceinline__ void stage_terrain_global(const ParcTerrainBatch& t, int64_t b, float* s_cx, float* s_cy,
                                                     float* s_minmax) {
  if (threadIdx.x == 0) { s_minmax[0] = -INFINITY; s_minmax[1] = INFINITY; }
  stage_terrain_nodes(t, b, s_cx, s_cy);
}

__device__ __forceinline__ float sample_base_z(const ParcTerrainBatch& t, int64_t b) {
  return t.base_z ? __ldg(t.base_z + b * t.base_z_stride) : t.base_z_value;
}

}  // namespace parc
