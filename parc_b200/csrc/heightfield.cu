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
