// Nearest-cell heightfield sampling and heightmap observations with caller-supplied root/heading.
//
// Reference: util/terrain_util.py:113-130, :1329-1346 (sampling), :2049-2082 (grid observation),
// envs/ig_parkour/mgdm_dm_util.py:158-179 (ray observation).  The fused query+FK+obs kernel lives
// in motion_query.cu; these are the stand-alone operators the simulated character's observation
// (root state from the simulator, not from the motion table) goes through.
#include "parc_common.cuh"

namespace parc {

__global__ void __launch_bounds__(256)
hf_sample_kernel(const __grid_constant__ ParcHeightfield t, const float* __restrict__ xy, int64_t n,
                 float* __restrict__ z, int64_t* __restrict__ gidx) {
  const float2* __restrict__ p2 = reinterpret_cast<const float2*>(xy);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float2 p = __ldg(p2 + i);
    const int ix = grid_index_1d(p.x, t.min_x, t.dx, t.dim_x);
    const int iy = grid_index_1d(p.y, t.min_y, t.dy, t.dim_y);
    if (z) z[i] = __ldg(t.hf + (size_t)ix * t.dim_y + iy);
    if (gidx) {
      gidx[i * 2] = ix;
      gidx[i * 2 + 1] = iy;
    }
  }
}

// One warp per env: heading / root are resolved once per env, lanes stride over the template points (consecutive
// lanes write consecutive outputs); the cell index uses the hoisted-reciprocal form of parc_common.cuh, which
// parc_selftest_grid_index proves index-identical to the reference's true division.
__global__ void __launch_bounds__(128)
hf_obs_kernel(const __grid_constant__ ParcHeightfield t, const __grid_constant__ ParcObsSpec obs,
              const float* __restrict__ root, int root_stride, const float* __restrict__ heading,
              const float* __restrict__ root_rot, const float* __restrict__ root_offset, int offset_stride, int64_t n,
              float* __restrict__ out, int64_t out_stride) {
  const int P = obs.num_points;
  const int lane = threadIdx.x & 31;
  const float2* __restrict__ tmpl = reinterpret_cast<const float2*>(obs.tmpl_xy);
  const GridAxis gx = make_grid_axis(t.min_x, t.dx, t.dim_x);
  const GridAxis gy = make_grid_axis(t.min_y, t.dy, t.dim_y);
  const int64_t warp0 = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * 4;
  for (int64_t e = warp0; e < n; e += nwarps) {
    const float* r = root + e * root_stride;
    // heading given, or taken from the root rotation as the caller would (util/torch_util.py:470-479)
    const float h = heading ? __ldg(heading + e) : calc_heading(__ldg(reinterpret_cast<const float4*>(root_rot) + e));
    const float sn = sinf(h), cs = cosf(h);
    float rx = __ldg(r), ry = __ldg(r + 1), rz = obs.relative ? __ldg(r + 2) : 0.0f;
    if (root_offset) {               // env-local -> terrain coordinates (ig_parkour_env.py:640), one add per component
      rx = add_rn(rx, __ldg(root_offset + e * offset_stride));
      ry = add_rn(ry, __ldg(root_offset + e * offset_stride + 1));
      if (obs.relative) rz = add_rn(rz, __ldg(root_offset + e * offset_stride + 2));
    }
    float* __restrict__ o = out + e * out_stride;
#pragma unroll 4
    for (int k = lane; k < P; k += 32) {
      const float2 w = rotate_offset_2d(__ldg(tmpl + k), cs, sn, rx, ry);
      const int ix = grid_index_fast(w.x, gx), iy = grid_index_fast(w.y, gy);
      float z = __ldg(t.hf + (size_t)ix * t.dim_y + iy);
      if (obs.relative) z = fminf(fmaxf(sub_rn(z, rz), obs.min_h), obs.max_h);
      o[k] = z;
    }
  }
}

// The MDM training sampler's per-sample terrain gather (diffusion/mdm_heightfield_contact_motion_sampler.py:414-447
// get_hfs_from_data_helper + :449-474 the sampling grid and the relative-z shift of get_hfs_from_data): sample i
// looks at ITS clip's own terrain through a heading-rotated grid around its root.  The reference loops over the
// samples in Python (one get_grid_index + 3 advanced-index gathers + a mask rebuild from per-frame index lists each);
// here: one warp per sample.
//   1  OR the clip's per-frame cell bitmasks over the sample's frame window -> the warp's shared-memory words
//      (= compute_hf_mask_from_inds of the window, util/terrain_util.py:1999-2007)
//   2  lanes stride over the grid points: rotate + translate (rotate_2d_vec, then + root xy), nearest-cell index on
//      the clip's terrain, gather the height and -- where the window's mask is set -- the cell's (max, min) band,
//      elsewhere the "free" band (2 max_h, 2 min_h); subtract the relative-z reference (centre cell's height or the
//      caller's canon root z).
#define CLIPHF_WARPS 4
__global__ void __launch_bounds__(CLIPHF_WARPS * 32)
clip_hf_gather_kernel(const __grid_constant__ ParcClipTerrains ct, const __grid_constant__ ParcClipHfQuery q) {
  extern __shared__ uint32_t s_words[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* mask = s_words + (size_t)warp * ct.max_mask_words;
  const float2* __restrict__ tmpl = reinterpret_cast<const float2*>(q.tmpl_xy);
  const int P = q.grid_x * q.grid_y;
  const int centre = q.centre_x * q.grid_y + q.centre_y;
  for (int64_t i = (int64_t)blockIdx.x * CLIPHF_WARPS + warp; i < q.n; i += (int64_t)gridDim.x * CLIPHF_WARPS) {
    int64_t id = __ldg(q.motion_ids + i);
    if (id < 0 || id >= ct.num_clips) id = 0;                    // validated on the host side of the mirror
    const ParcClipTerrain rec = ct.clips[id];
    const int cells = rec.dim_x * rec.dim_y;
    const int W = (cells + 31) >> 5;
    // ---- 1: mask of the frame window ----
    int t0 = __ldg(q.frame_lo + i), t1 = __ldg(q.frame_hi + i);
    t0 = max(t0, 0); t1 = min(t1, rec.num_frames - 1);
    const bool want_band = q.maxmin_out != nullptr && ct.mask_words != nullptr && rec.mask_offset >= 0;
    if (want_band) {
      for (int w = lane; w < W; w += 32) {
        uint32_t acc = 0u;
        for (int t = t0; t <= t1; ++t) acc |= __ldg(ct.mask_words + rec.mask_offset + (int64_t)t * W + w);
        mask[w] = acc;
      }
    }
    __syncwarp();
    // ---- 2: the rotated grid ----
    const float4 rr = __ldg(reinterpret_cast<const float4*>(q.root_rot) + i);
    const float h = calc_heading(rr);
    const float sn = sinf(h), cs = cosf(h);
    const float rx = __ldg(q.root_pos + i * 3), ry = __ldg(q.root_pos + i * 3 + 1);
    const float* __restrict__ hf = ct.hf + rec.cell_offset;
    // relative-z reference: the centre cell's height (RELATIVE_TO_ROOT_FLOOR) or the caller's root z
    const float2 wc = rotate_offset_2d(__ldg(tmpl + centre), cs, sn, rx, ry);
    const float centre_h = __ldg(hf + (size_t)grid_index_1d(wc.x, rec.min_x, rec.dx, rec.dim_x) * rec.dim_y +
                                 grid_index_1d(wc.y, rec.min_y, rec.dy, rec.dim_y));
    const float ref_z = q.canon_root_z ? __ldg(q.canon_root_z + i) : centre_h;
    if (lane == 0 && q.centre_h_out) q.centre_h_out[i] = centre_h;
    for (int k = lane; k < P; k += 32) {
      const float2 w = rotate_offset_2d(__ldg(tmpl + k), cs, sn, rx, ry);
      const int cell = grid_index_1d(w.x, rec.min_x, rec.dx, rec.dim_x) * rec.dim_y +
                       grid_index_1d(w.y, rec.min_y, rec.dy, rec.dim_y);
      q.hf_out[i * P + k] = sub_rn(__ldg(hf + cell), ref_z);
      if (q.maxmin_out) {
        float hi = q.free_max, lo = q.free_min;
        if (want_band && ((mask[cell >> 5] >> (cell & 31)) & 1u)) {
          const float2 mm = __ldg(reinterpret_cast<const float2*>(ct.hf_maxmin) + rec.cell_offset + cell);
          hi = mm.x; lo = mm.y;
        }
        reinterpret_cast<float2*>(q.maxmin_out)[i * P + k] = make_float2(sub_rn(hi, ref_z), sub_rn(lo, ref_z));
      }
    }
    __syncwarp();
  }
}

static int flat_grid(int64_t total) {
  int64_t b = (total + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  return (int)(b > 0 ? b : 1);
}

}  // namespace parc

using namespace parc;

static int check_hf(const ParcHeightfield* hf) {
  if (!hf || !hf->hf) return PARC_E_NULL;
  if (hf->dim_x <= 0 || hf->dim_y <= 0) return PARC_E_SIZE;
  return PARC_OK;
}

extern "C" int parc_hf_sample(const ParcHeightfield* hf, const float* xy, int64_t n, float* z_out,
                              int64_t* grid_idx_out, void* stream) {
  int rc = check_hf(hf);
  if (rc) return rc;
  if (n < 0) return PARC_E_SIZE;
  if (n == 0) return PARC_OK;
  if (!xy || (!z_out && !grid_idx_out)) return PARC_E_NULL;
  if ((reinterpret_cast<uintptr_t>(xy) & 7u) != 0) return PARC_E_ALIGN;
  if (n == 0) return PARC_OK;
  hf_sample_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(*hf, xy, n, z_out, grid_idx_out);
  return check_launch();
}

extern "C" int parc_hf_obs(const ParcHeightfield* hf, const ParcObsSpec* obs, const float* root,
                           int32_t root_stride, const float* heading, const float* root_rot,
                           const float* root_offset, int32_t offset_stride, int64_t n, float* obs_out,
                           int64_t out_stride, void* stream) {
  int rc = check_hf(hf);
  if (rc) return rc;
  if (!obs) return PARC_E_NULL;
  if (n < 0 || obs->num_points < 0 || root_stride < (obs->relative ? 3 : 2)) return PARC_E_SIZE;
  if (n == 0 || obs->num_points == 0) return PARC_OK;
  if (!obs->tmpl_xy || !root || (!heading && !root_rot) || !obs_out) return PARC_E_NULL;
  if ((reinterpret_cast<uintptr_t>(obs->tmpl_xy) & 7u) != 0) return PARC_E_ALIGN;
  if (!heading && !aligned16(root_rot)) return PARC_E_ALIGN;
  if (root_offset && offset_stride < (obs->relative ? 3 : 2)) return PARC_E_SIZE;
  if (out_stride == 0) out_stride = obs->num_points;
  if (out_stride < obs->num_points) return PARC_E_SIZE;
  int64_t ctas = (n + 3) / 4;
  if (ctas > 148 * 16) ctas = 148 * 16;
  hf_obs_kernel<<<(int)ctas, 128, 0, (cudaStream_t)stream>>>(*hf, *obs, root, root_stride, heading, root_rot,
                                                            root_offset, offset_stride, n, obs_out, out_stride);
  return check_launch();
}


extern "C" int parc_clip_hf_gather(const ParcClipTerrains* terrains, const ParcClipHfQuery* query, void* stream) {
  if (!terrains || !query) return PARC_E_NULL;
  if (query->n < 0 || terrains->num_clips <= 0 || query->grid_x <= 0 || query->grid_y <= 0 ||
      terrains->max_mask_words < 0)
    return PARC_E_SIZE;
  if (query->centre_x < 0 || query->centre_x >= query->grid_x || query->centre_y < 0 || query->centre_y >= query->grid_y)
    return PARC_E_SIZE;
  if (query->n == 0) return PARC_OK;
  if (!terrains->clips || !terrains->hf || !query->motion_ids || !query->root_pos || !query->root_rot ||
      !query->tmpl_xy || !query->hf_out || !query->frame_lo || !query->frame_hi)
    return PARC_E_NULL;
  if (query->maxmin_out && !terrains->hf_maxmin) return PARC_E_NULL;
  if (!aligned16(query->root_rot)) return PARC_E_ALIGN;
  if ((reinterpret_cast<uintptr_t>(query->tmpl_xy) & 7u) != 0 || (reinterpret_cast<uintptr_t>(query->maxmin_out) & 7u) != 0 ||
      (reinterpret_cast<uintptr_t>(terrains->hf_maxmin) & 7u) != 0)
    return PARC_E_ALIGN;
  const size_t smem = (size_t)CLIPHF_WARPS * terrains->max_mask_words * sizeof(uint32_t);
  if (smem > PARC_SMEM_LIMIT) return PARC_E_SIZE;
  static SmemOptIn opt;
  int rc = ensure_dynamic_smem(clip_hf_gather_kernel, opt, smem);
  if (rc) return rc;
  int64_t ctas = (query->n + CLIPHF_WARPS - 1) / CLIPHF_WARPS;
  if (ctas > 148 * 16) ctas = 148 * 16;
  clip_hf_gather_kernel<<<(int)ctas, CLIPHF_WARPS * 32, smem, (cudaStream_t)stream>>>(*terrains, *query);
  return check_launch();
}
