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

// Scan every cell.  hf/cx/cy are shared-memory (or global) arrays; all threads of a warp read the
// same cell at the same time, so shared reads are broadcasts.
template <bool WANT_INV, bool WANT_SOL>
__device__ __forceinline__ SdfBest scan_cells(const float* __restrict__ hf, const float* __restrict__ cx,
                                              const float* __restrict__ cy, int X, int Y, float hx, float hy,
                                              float base, float3 p) {
  SdfBest b;
  b.inv = INFINITY; b.sol = INFINITY; b.arg_inv = 0; b.arg_sol = 0;
  const float top = -base;
  for (int ix = 0; ix < X; ++ix) {
    const float qx = fabsf(p.x - cx[ix]) - hx;
    const float mx = fmaxf(qx, 0.0f);
    const float mx2 = mx * mx;
    const float* __restrict__ col = hf + ix * Y;
#pragma unroll 4
    for (int iy = 0; iy < Y; ++iy) {
      const float qy = fabsf(p.y - cy[iy]) - hy;
      const float my = fmaxf(qy, 0.0f);
      const float mxy2 = mx2 + my * my;
      const float qxy = fmaxf(qx, qy);
      const float h = col[iy];
      if (WANT_INV) {
        const float cz = (h + top) * 0.5f;
        const float hz = (top - h) * 0.5f;
        const float qz = fabsf(p.z - cz) - hz;
        const float mz = fmaxf(qz, 0.0f);
        const float sd = sqrtf(mxy2 + mz * mz) + fminf(fmaxf(qxy, qz), 0.0f);
        if (sd < b.inv) { b.inv = sd; b.arg_inv = ix * Y + iy; }
      }
      if (WANT_SOL) {
        const float cz = (h + base) * 0.5f;
        const float hz = (h - base) * 0.5f;
        const float qz = fabsf(p.z - cz) - hz;
        const float mz = fmaxf(qz, 0.0f);
        const float sd = sqrtf(mxy2 + mz * mz) + fminf(fmaxf(qxy, qz), 0.0f);
        if (sd < b.sol) { b.sol = sd; b.arg_sol = ix * Y + iy; }
      }
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

// Stage one sample's terrain: hf tile, absolute cell-centre coordinates (node offset + min centre,
// added in fp32 as util/terrain_util.py:1859-1860 does).
__device__ __forceinline__ void stage_terrain(const ParcTerrainBatch& t, int64_t b, float* s_hf, float* s_cx,
                                              float* s_cy) {
  const int X = t.dim_x, Y = t.dim_y;
  const float* hf = t.hf + b * t.hf_batch_stride;
  const float* mc = t.min_center + b * t.min_center_stride;
  for (int i = threadIdx.x; i < X * Y; i += blockDim.x) s_hf[i] = __ldg(hf + i);
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
  const int64_t b = blockIdx.y;
  stage_terrain(t, b, s_hf, s_cx, s_cy);
  __syncthreads();
  const float base = sample_base_z(t, b);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_points; i += (int64_t)gridDim.x * blockDim.x) {
    const float* pp = points + (b * n_points + i) * 3;
    const float3 p = make_float3(__ldg(pp), __ldg(pp + 1), __ldg(pp + 2));
    float v;
    int a;
    if (inverted) {
      const SdfBest r = scan_cells<true, false>(s_hf, s_cx, s_cy, X, Y, t.half_dx, t.half_dy, base, p);
      v = -1.0f * r.inv; a = r.arg_inv;
    } else {
      const SdfBest r = scan_cells<false, true>(s_hf, s_cx, s_cy, X, Y, t.half_dx, t.half_dy, base, p);
      v = r.sol; a = r.arg_sol;
    }
    sdf[b * n_points + i] = v;
    if (arg) arg[b * n_points + i] = a;
  }
}

// ------------------------------------------------------------------------------------------------
// a14/a15 fused: FK -> body points -> SDF -> pen / contact (+ gradients wrt the pose)
// ------------------------------------------------------------------------------------------------
#define LOSS_THREADS 320

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
  __shared__ float s_bpos[PARC_MAX_BODIES][3];
  __shared__ float s_brot[PARC_MAX_BODIES][4];
  __shared__ float s_warp_sum[LOSS_THREADS / 32];
  __shared__ int s_winner[PARC_MAX_BODIES];
  __shared__ float s_contact_w[PARC_MAX_BODIES];   // w_contact * contacts[f,b] if the winner's clamp passes

  const int X = p.terrain.dim_x, Y = p.terrain.dim_y;
  const int S = p.pts.num_points;
  float* s_hf = smem;
  float* s_cx = s_hf + X * Y;
  float* s_cy = s_cx + X;
  float* s_sol = s_cy + Y;                         // [S] clamp(sdf_solid, min=0)
  float* s_g = s_sol + S;                          // [S][7]: d/d body_pos (3), d/d body_rot (4)
  int* s_body = reinterpret_cast<int*>(s_g + (size_t)S * 7);   // [S] body of point

  const int64_t b = blockIdx.y;
  stage_model(&sm, model_param);
  stage_terrain(p.terrain, b, s_hf, s_cx, s_cy);
  __syncthreads();
  const int J = sm.num_bodies;
  for (int j = threadIdx.x; j < J; j += blockDim.x) {
    const int s0 = __ldg(p.pts.point_start + j), s1 = __ldg(p.pts.point_start + j + 1);
    for (int k = s0; k < s1; ++k) s_body[k] = j;
  }
  const float base = sample_base_z(p.terrain, b);
  const float hx = p.terrain.half_dx, hy = p.terrain.half_dy;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const LaneBody lb = load_lane_body(sm, lane, 0);
  const int max_depth = sm.max_depth;
  __syncthreads();

  const int64_t f_begin = (int64_t)blockIdx.x * p.frames_per_cta;
  const int64_t f_end = min(f_begin + (int64_t)p.frames_per_cta, p.frames);
  for (int64_t f = f_begin; f < f_end; ++f) {
    const int64_t q = b * p.frames + f;
    // ---- (1) FK by warp 0 ----
    float4 prot = make_float4(0.f, 0.f, 0.f, 1.f), local = prot, rot = prot;
    float3 pos = make_float3(0.f, 0.f, 0.f);
    if (warp == 0) {
      if (lane == 0) {
        pos = make_float3(__ldg(p.root_pos + q * 3), __ldg(p.root_pos + q * 3 + 1), __ldg(p.root_pos + q * 3 + 2));
        rot = __ldg(reinterpret_cast<const float4*>(p.root_rot) + q);
      } else if (lane < J) {
        rot = __ldg(reinterpret_cast<const float4*>(p.joint_rot) + q * (J - 1) + (lane - 1));
      }
      fk_warp_keep(lb, max_depth, pos, rot, prot, local);
      if (lane < J) {
        s_bpos[lane][0] = pos.x; s_bpos[lane][1] = pos.y; s_bpos[lane][2] = pos.z;
        s_brot[lane][0] = rot.x; s_brot[lane][1] = rot.y; s_brot[lane][2] = rot.z; s_brot[lane][3] = rot.w;
      }
    }
    __syncthreads();

    // ---- (2) every surface point against every cell ----
    float pen_local = 0.0f;
    for (int k = threadIdx.x; k < S; k += blockDim.x) {
      const int bj = s_body[k];
      const float3 lp = make_float3(__ldg(p.pts.points + k * 3), __ldg(p.pts.points + k * 3 + 1),
                                    __ldg(p.pts.points + k * 3 + 2));
      const float4 br = make_float4(s_brot[bj][0], s_brot[bj][1], s_brot[bj][2], s_brot[bj][3]);
      const float3 r = quat_rotate(br, lp);
      const float3 wp = make_float3(r.x + s_bpos[bj][0], r.y + s_bpos[bj][1], r.z + s_bpos[bj][2]);
      const SdfBest best = scan_cells<true, true>(s_hf, s_cx, s_cy, X, Y, hx, hy, base, wp);
      // penetration: sdf = -best.inv ; neg = min(sdf, 0) ; pen += -neg
      const float sdf_inv = -1.0f * best.inv;
      pen_local += -fminf(sdf_inv, 0.0f);
      s_sol[k] = fmaxf(best.sol, 0.0f);
      if (p.want_grad) {
        // d pen / d wp = [sdf_inv <= 0] * grad sdBox(air cell)
        float3 gp = make_float3(0.f, 0.f, 0.f);
        if (sdf_inv <= 0.0f) {
          const float3 g = cell_grad(s_hf, s_cx, s_cy, Y, hx, hy, base, true, best.arg_inv, wp);
          gp = make_float3(p.w_pen * g.x, p.w_pen * g.y, p.w_pen * g.z);
        }
        // stash the solid-cell gradient (used only if this point wins its body's min)
        float3 gs = make_float3(0.f, 0.f, 0.f);
        if (best.sol >= 0.0f) gs = cell_grad(s_hf, s_cx, s_cy, Y, hx, hy, base, false, best.arg_sol, wp);
        float* g7 = s_g + (size_t)k * 7;
        g7[0] = gp.x; g7[1] = gp.y; g7[2] = gp.z;
        g7[3] = gs.x; g7[4] = gs.y; g7[5] = gs.z;
      }
    }
    // block sum of the penetration terms (fixed tree -> deterministic)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pen_local += __shfl_xor_sync(PARC_FULL_MASK, pen_local, o);
    if (lane == 0) s_warp_sum[warp] = pen_local;
    __syncthreads();

    // ---- (3) per-body first-index min over the body's points (contact term) ----
    float contact_f = 0.0f;
    if (warp == 0) {
      float cterm = 0.0f;
      if (lane < J) {
        const int s0 = __ldg(p.pts.point_start + lane), s1 = __ldg(p.pts.point_start + lane + 1);
        float best = INFINITY;
        int win = s0;
        for (int k = s0; k < s1; ++k) {
          const float v = s_sol[k];
          if (v < best) { best = v; win = k; }
        }
        const float c = __ldg(p.contacts + q * J + lane);
        cterm = best * c;                              // closest_distances * contacts[..., b]
        s_winner[lane] = win;
        s_contact_w[lane] = p.w_contact * c;
      }
      // sum over bodies in index order (as the reference's python loop accumulates)
      for (int j = 0; j < J; ++j) contact_f += __shfl_sync(PARC_FULL_MASK, cterm, j);
      if (lane == 0) {
        float pen_f = 0.0f;
        for (int w = 0; w < LOSS_THREADS / 32; ++w) pen_f += s_warp_sum[w];
        if (p.pen_out) p.pen_out[q] = pen_f;
        if (p.contact_out) p.contact_out[q] = contact_f;
      }
    }
    if (!p.want_grad) { __syncthreads(); continue; }
    __syncthreads();

    // ---- (4) chain d/d world-point through wp = rotate(body_rot, lp) + body_pos ----
    for (int k = threadIdx.x; k < S; k += blockDim.x) {
      const int bj = s_body[k];
      float* g7 = s_g + (size_t)k * 7;
      float3 g = make_float3(g7[0], g7[1], g7[2]);
      if (s_winner[bj] == k) {
        const float cw = s_contact_w[bj];
        g.x += cw * g7[3]; g.y += cw * g7[4]; g.z += cw * g7[5];
      }
      const float3 lp = make_float3(__ldg(p.pts.points + k * 3), __ldg(p.pts.points + k * 3 + 1),
                                    __ldg(p.pts.points + k * 3 + 2));
      const float4 br = make_float4(s_brot[bj][0], s_brot[bj][1], s_brot[bj][2], s_brot[bj][3]);
      const float4 gq = quat_rotate_vjp_q(br, lp, g);
      g7[0] = g.x; g7[1] = g.y; g7[2] = g.z;
      g7[3] = gq.x; g7[4] = gq.y; g7[5] = gq.z; g7[6] = gq.w;
    }
    __syncthreads();

    // ---- (5) per-body sums, then the FK VJP, by warp 0 ----
    if (warp == 0) {
      float3 gp = make_float3(0.f, 0.f, 0.f);
      float4 gr = make_float4(0.f, 0.f, 0.f, 0.f);
      if (lane < J) {
        const int s0 = __ldg(p.pts.point_start + lane), s1 = __ldg(p.pts.point_start + lane + 1);
        for (int k = s0; k < s1; ++k) {
          const float* g7 = s_g + (size_t)k * 7;
          gp.x += g7[0]; gp.y += g7[1]; gp.z += g7[2];
          gr.x += g7[3]; gr.y += g7[4]; gr.z += g7[5]; gr.w += g7[6];
        }
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
    __syncthreads();
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

  const size_t smem = terrain_smem_bytes(terrain) + (size_t)pts->num_points * (1 + 7 + 1) * sizeof(float);
  if (smem > 200 * 1024) return PARC_E_SIZE;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(body_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  // enough CTAs to fill the GPU a few times over, but amortise the terrain staging when there is
  // plenty of work
  int64_t fpc = (batch * frames) / (148 * 8);
  if (fpc < 1) fpc = 1;
  if (fpc > 8) fpc = 8;
  p.frames_per_cta = (int)fpc;
  dim3 grid((unsigned)((frames + fpc - 1) / fpc), (unsigned)batch);
  body_loss_kernel<<<grid, LOSS_THREADS, smem, (cudaStream_t)stream>>>(p, *model);
  return check_launch();
}
