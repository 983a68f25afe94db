/*
 * parc_b200.h -- C ABI of libparc_b200.so: PARC's batched kinematic motion-query path on B200 (sm_100a).
 *
 * The reference (ZhengmaoHe/PARC) has no FFI layer: its "operator interface" for this path is a set of
 * Python methods taking torch tensors.  Each entry point below names the reference method it replaces
 * (paths relative to the reference root).  INTEGRATION.md shows the ctypes stub a maintainer adds.
 *
 * Contract (all entry points)
 *   - Plain C: pointers, sizes, PODs.  No torch / C++ types cross the boundary.
 *   - OWNERSHIP: the caller allocates every input, output and workspace buffer (device memory unless a
 *     parameter says "host").  The library never allocates, frees or retains a pointer past return.
 *   - ASYNC: work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream)
 *     of the CURRENT device; the call returns without synchronising.
 *   - ERRORS: returns 0 on success, a positive cudaError_t if the launch failed, or a negative
 *     PARC_E_* code if an argument was rejected before anything was enqueued.  Nothing throws or exits.
 *   - THREADS: re-entrant; no global mutable state.  ParcCharModel is an immutable host POD passed by
 *     pointer and copied into the kernel's parameter space at launch.
 *   - dtypes follow the reference: fp32 values, int64 clip ids / frame indices, quaternions xyzw.
 */
#ifndef PARC_B200_H_
#define PARC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PARC_ABI_VERSION 3
#define PARC_MAX_BODIES 24   /* position + rotation slots of a packed row must fit one warp: J + 1 <= 32 */
#define PARC_MAX_DOF 96

/* negative = argument rejected; positive = cudaError_t */
enum {
  PARC_OK = 0,
  PARC_E_NULL = -1,       /* a required pointer is NULL */
  PARC_E_SIZE = -2,       /* negative / inconsistent size */
  PARC_E_MODEL = -3,      /* character model unsupported (J > PARC_MAX_BODIES, bad parents, ...) */
  PARC_E_ALIGN = -4,      /* pointer not aligned for vector access (16 B) */
  PARC_E_LAYOUT = -5      /* row_floats does not match the model's packed-row layout */
};

/* joint types -- anim/kin_char_model.py:11-15 */
enum { PARC_JOINT_ROOT = 0, PARC_JOINT_HINGE = 1, PARC_JOINT_SPHERICAL = 2, PARC_JOINT_FIXED = 3 };
/* loop modes -- anim/motion_lib.py:11-13 */
enum { PARC_LOOP_CLAMP = 0, PARC_LOOP_WRAP = 1 };

/* Kinematic tree: the arrays KinCharModel.init keeps (anim/kin_char_model.py:147-178).
 * parent[b] < b for b > 0 and parent[0] == -1 (the MJCF loader's DFS order guarantees it). */
typedef struct ParcCharModel {
  int32_t num_bodies;                         /* J */
  int32_t dof_size;                           /* D = sum of dof_dim */
  int32_t max_depth;                          /* longest root->leaf chain, root = 0 */
  int32_t reserved;
  int32_t parent[PARC_MAX_BODIES];
  int32_t depth[PARC_MAX_BODIES];
  int32_t joint_type[PARC_MAX_BODIES];
  int32_t dof_idx[PARC_MAX_BODIES];
  float local_trans[PARC_MAX_BODIES][3];
  float local_rot[PARC_MAX_BODIES][4];        /* xyzw */
  float joint_axis[PARC_MAX_BODIES][3];       /* hinge axis, zeros otherwise */
} ParcCharModel;

/* Packed frame row (one row per table frame, all float4 slots, 32 B-aligned stride):
 *   slot 0            root_pos.xyz, pad
 *   slot 1            root_rot xyzw
 *   slot 2 .. J       joint_rot[0..J-2] xyzw
 *   slot J+1 ..       contacts[0..J-1], zero padded to a multiple of 4
 *   vel part          slot 0 root_vel.xyz,pad; slot 1 root_ang_vel.xyz,pad; slots 2.. dof_vel(D), zero padded
 * Replaces the eight separate MotionLib tables (anim/motion_lib.py:349-375). */
typedef struct ParcRowLayout {
  int32_t row_floats;      /* stride in floats (multiple of 8) */
  int32_t pose_slots;      /* float4 slots read at BOTH key frames: 2 + (J-1) + ceil(J/4) */
  int32_t contact_slot;    /* first contact slot = J + 1 */
  int32_t vel_slot;        /* first velocity slot = pose_slots */
  int32_t vel_slots;       /* 2 + ceil(D / 4) */
  int32_t reserved[3];
} ParcRowLayout;

/* Per-clip metadata, 32 B (anim/motion_lib.py:352-358, :373-375). */
typedef struct ParcClipMeta {
  int32_t num_frames;
  int32_t loop_mode;
  int64_t start_idx;       /* exclusive cumsum of num_frames */
  float length;            /* fp32((n-1)/fps) exactly as the reference stores it */
  float root_pos_delta[3]; /* last - first root position, z zeroed */
} ParcClipMeta;

/* Device-resident compact copy of the kinematic tree (body count, parents, depths, local translations /
 * rotations): what the query kernel stages into shared memory.  Built on the host by parc_tree_from_model()
 * and uploaded once per model (PARC_TREE_BYTES bytes, 16-byte aligned).  It keeps the 1.4 KB ParcCharModel out
 * of the kernel's parameter block: measured on B200, a launch with a 1.7 KB parameter block costs 1.1 us more
 * than one with a small block, a twelfth of the whole 4096-env query. */
#define PARC_TREE_BYTES 880

typedef struct ParcMotionTables {
  const float* rows;            /* device [total_frames, row_floats] */
  const ParcClipMeta* clips;    /* device [num_clips] */
  int64_t total_frames;
  int64_t num_clips;
  int32_t row_floats;
  int32_t reserved;
  const void* tree;             /* device, PARC_TREE_BYTES from parc_tree_from_model(); required by the queries */
} ParcMotionTables;

/* Outputs of a frame query; any pointer may be NULL (= not wanted).  Shapes as the tuple returned by
 * MotionLib.calc_motion_frame (anim/motion_lib.py:106-112). */
typedef struct ParcFrameOut {
  float* root_pos;       /* [N,3] */
  float* root_rot;       /* [N,4] */
  float* root_vel;       /* [N,3] */
  float* root_ang_vel;   /* [N,3] */
  float* joint_rot;      /* [N,J-1,4] */
  float* dof_vel;        /* [N,D] */
  float* contacts;       /* [N,J] */
  int64_t* frame_idx0;   /* [N] absolute table rows, as MotionLib._calc_frame_blend returns */
  int64_t* frame_idx1;   /* [N] */
  float* blend;          /* [N] */
} ParcFrameOut;

typedef struct ParcFkOut {
  float* body_pos;       /* [N,J,3] */
  float* body_rot;       /* [N,J,4] */
} ParcFkOut;

/* SubTerrain's sampled fields (util/terrain_util.py:21-39): hf is [dim_x, dim_y] row-major. */
typedef struct ParcHeightfield {
  const float* hf;
  int32_t dim_x, dim_y;
  float min_x, min_y;
  float dx, dy;
} ParcHeightfield;

/* Heightmap observation: out[n,k] = hf(R(heading_n) * tmpl[k] + root_xy_n), then if relative != 0
 * clamp(z - root_z_n, min_h, max_h)  (envs/ig_parkour/mgdm_dm_util.py:158-179; template from
 * util/geom_util.py:249-270 or :210-221). */
typedef struct ParcObsSpec {
  const float* tmpl_xy;  /* device [num_points,2] */
  int32_t num_points;
  int32_t relative;
  float min_h, max_h;
} ParcObsSpec;

int parc_abi_version(void);
const char* parc_error_string(int code);

/* Host-only helpers: derive the packed-row layout; validate a model; fill the PARC_TREE_BYTES host image of
 * the compact tree (upload it and point ParcMotionTables.tree at the device copy). */
int parc_row_layout(const ParcCharModel* model, ParcRowLayout* out);
int parc_validate_model(const ParcCharModel* model);
int parc_tree_from_model(const ParcCharModel* model, void* tree_host_out);

/* a1: interleave the reference's per-frame tables into packed rows.  contacts may be NULL (zeros).
 * Replaces the layout built by MotionLib._load_motions (anim/motion_lib.py:349-375). */
int parc_pack_frames(const float* root_pos, const float* root_rot, const float* joint_rot,
                     const float* contacts, const float* root_vel, const float* root_ang_vel,
                     const float* dof_vel, int64_t total_frames, const ParcCharModel* model,
                     float* rows_out, void* stream);

/* a2+a3 (+a6, +a10 fused): MotionLib.calc_motion_frame (anim/motion_lib.py:80-112) for N
 * (clip id, time) queries; if `fk` is non-NULL also KinCharModel.forward_kinematics
 * (anim/kin_char_model.py:509-541) of the blended pose; if `obs_out` is non-NULL also the heightmap
 * observation [N,num_points] around the blended root with heading = calc_heading(root_rot)
 * (util/torch_util.py:470-479).  One warp per query, a single launch. */
int parc_motion_query(const ParcMotionTables* tables, const int64_t* motion_ids, const float* motion_times,
                      int64_t n, const ParcCharModel* model, const ParcFrameOut* frame,
                      const ParcFkOut* fk, const ParcHeightfield* hf, const ParcObsSpec* obs,
                      float* obs_out, void* stream);

/* Tracker-step form: every (id, time) entry e is queried at num_steps times  t_e + time_offsets[k]
 * (query q = e * num_steps + k; outputs are [n, num_steps, ...], i.e. [n * num_steps, ...] rows) -- what the
 * tracker does every control step: the reference frame (dm_env.py:570-595) plus the future targets of
 * fetch_tar_obs_data (envs/ig_parkour/mgdm_dm_util.py:279-302: ids tiled, motion_times + timestep *
 * tar_obs_steps, one fp32 add as there).  Pass time_offsets[0] = 0 for the current frame (time_offsets may be
 * NULL when num_steps == 1).  The heightmap observation, if requested, is produced for step 0 of each entry
 * only: obs_out is [n, num_points].
 * root_xy_offset (device [n,2] float32, 8-byte aligned, or NULL): where each entry's motion sits on the shared
 * terrain -- added to root x,y of every step after the query and before FK / the observation, i.e.
 * DMEnv._move_to_motion_terrain (envs/ig_parkour/dm_env.py:604-615, applied at :575 and :701). */
int parc_motion_query_steps(const ParcMotionTables* tables, const int64_t* motion_ids, const float* motion_times,
                            int64_t n, const float* time_offsets, int32_t num_steps, const float* root_xy_offset,
                            const ParcCharModel* model, const ParcFrameOut* frame, const ParcFkOut* fk,
                            const ParcHeightfield* hf, const ParcObsSpec* obs, float* obs_out, void* stream);

/* a4: MotionLib.get_motion_frame (anim/motion_lib.py:114-131): integer frame lookup, no blending. */
int parc_get_motion_frame(const ParcMotionTables* tables, const int64_t* motion_ids,
                          const int64_t* frame_idxs, int64_t n, const ParcCharModel* model,
                          const ParcFrameOut* frame, const ParcFkOut* fk, void* stream);

/* The general form of the three queries above: one argument block, plus what a stepping caller needs.
 *   motion_times (blended query, calc_motion_frame) XOR frame_idxs (integer frames, get_motion_frame);
 *   num_steps / time_offsets / root_xy_offset as in parc_motion_query_steps (num_steps 0 or 1 = plain query).
 *   error_flags  device int32[1] (caller-zeroed, 4-byte aligned) or NULL.  The reference raises an IndexError (CPU)
 *                or a device assert (CUDA) on a clip id outside [0, num_clips) / a frame outside the table; the kernel
 *                never reads out of bounds: it ORs PARC_QUERY_ERR_CLIP_ID / PARC_QUERY_ERR_FRAME_IDX into
 *                *error_flags and answers with clip 0 / the clip's nearest valid frame.  The Python mirror turns a
 *                non-zero word into IndexError.
 *   flags        PARC_QUERY_FAST_HEADING   observation heading cos/sin straight from the rotated x axis instead of
 *                                          the reference's atan2 -> cos/sin chain (util/torch_util.py:470-479,
 *                                          :619-631); default is the reference chain.
 *                PARC_QUERY_PDL            launch with programmatic stream serialisation (sm_90+): the kernel's
 *                                          prologue (kinematic tree / template staging) overlaps the tail of the
 *                                          previous kernel of `stream`; it waits for that kernel before reading
 *                                          the inputs or writing anything.
 *                PARC_QUERY_PDL_EARLY_INPUTS  with PARC_QUERY_PDL: the caller guarantees that motion_ids, motion_times
 *                                          / frame_idxs, time_offsets and root_xy_offset were NOT written by the
 *                                          previous kernel of `stream` (e.g. they were uploaded, or produced two
 *                                          kernels earlier); the whole read side -- ids, clip records, frame rows,
 *                                          slerp -- then runs before the wait and only the stores are ordered
 *                                          behind the previous kernel.
 *   variant      0 = chosen by batch size; 1..6 force one instantiation (tuning / tests): 1 = two characters per warp,
 *                whole template sweep in flight (<= 128 registers); 2 = quarter sweep in flight, 64 registers;
 *                3 = half sweep, 80 registers; 4 = one character per warp; 5 = as 1 with every store of a warp's first
 *                work item deferred behind its forward kinematics (for PARC_QUERY_PDL_EARLY_INPUTS: read AND compute
 *                side overlap the previous kernel, only the stores are ordered behind it); 6 = as 5 with the
 *                observation template staged by one TMA bulk copy (cp.async.bulk + mbarrier) instead of per-thread
 *                loads (falls back to 5 when the template is not 16-byte aligned). */
#define PARC_QUERY_FAST_HEADING 1u
#define PARC_QUERY_PDL 2u
#define PARC_QUERY_PDL_EARLY_INPUTS 4u
#define PARC_QUERY_ERR_CLIP_ID 1
#define PARC_QUERY_ERR_FRAME_IDX 2

/* Fused compute_tar_obs (envs/ig_parkour/mgdm_dm_util.py:462-518) for the tracker-step form: with `tar_obs` set in
 * ParcQueryArgs, every query of step k >= 1 (the future targets of fetch_tar_obs_data) also writes its observation row
 *   root_pos_obs 3 | root tan-norm 6 | joint tan-norm 6 (J-1) | key bodies 3 K          (W = 9 + 6 (J-1) + 3 K floats)
 * relative to the SIMULATED character's root (sim_root_pos / sim_root_rot, [n,3] / [n,4], read in place) into
 * obs_out[env * out_env_stride + (k - 1) * W ...] -- the values parc_tar_obs produces from the stored targets, taken
 * from the registers that hold them, without the extra launch and the re-read.  Needs num_steps >= 2 and fk outputs. */
typedef struct ParcTarObsSpec {
  const float* sim_root_pos;       /* [n,3] */
  const float* sim_root_rot;       /* [n,4], 16-byte aligned (unused with global_obs) */
  const int32_t* key_body_ids;     /* [num_keys] body indices, or NULL with num_keys == 0 */
  float* obs_out;                  /* [n, out_env_stride] */
  int64_t out_env_stride;          /* floats between consecutive envs; >= (num_steps - 1) * W */
  int32_t num_keys;
  int32_t global_obs;
  int32_t global_tar_root_h;
  int32_t reserved;
} ParcTarObsSpec;

typedef struct ParcQueryArgs {
  const ParcMotionTables* tables;
  const int64_t* motion_ids;      /* device [n] */
  const float* motion_times;      /* device [n] or NULL */
  const int64_t* frame_idxs;      /* device [n] or NULL */
  int64_t n;                      /* entries */
  const float* time_offsets;      /* device [num_steps] or NULL */
  const float* root_xy_offset;    /* device [n,2] or NULL */
  const ParcCharModel* model;
  const ParcFrameOut* frame;      /* may be NULL */
  const ParcFkOut* fk;            /* may be NULL */
  const ParcHeightfield* hf;      /* required iff obs_out */
  const ParcObsSpec* obs;         /* required iff obs_out */
  float* obs_out;                 /* [n, num_points] or NULL */
  int32_t* error_flags;
  int32_t num_steps;
  uint32_t flags;
  int32_t variant;
  int32_t reserved;
  const ParcTarObsSpec* tar_obs;  /* tracker-step form only, or NULL (appended in ABI version 3) */
} ParcQueryArgs;

int parc_motion_query_ex(const ParcQueryArgs* args, void* stream);

/* a6: KinCharModel.forward_kinematics (anim/kin_char_model.py:509-541) on caller-supplied poses. */
int parc_fk_fwd(const float* root_pos, const float* root_rot, const float* joint_rot, int64_t n,
                const ParcCharModel* model, float* body_pos, float* body_rot, void* stream);

/* a6 backward: vector-Jacobian product of the above (what autograd computes through
 * anim/kin_char_model.py:517-539).  The forward pass is recomputed in registers from (root_rot,
 * joint_rot).  g_body_pos / g_body_rot may be NULL (= zero); outputs are overwritten, any may be NULL. */
int parc_fk_bwd(const float* root_rot, const float* joint_rot, const float* g_body_pos,
                const float* g_body_rot, int64_t n, const ParcCharModel* model, float* g_root_pos,
                float* g_root_rot, float* g_joint_rot, void* stream);

/* a8: KinCharModel.dof_to_rot (anim/kin_char_model.py:478-491, Joint.dof_to_rot :57-77). */
int parc_dof_to_rot_fwd(const float* dof, int64_t n, const ParcCharModel* model, float* joint_rot,
                        void* stream);
int parc_dof_to_rot_bwd(const float* dof, const float* g_joint_rot, int64_t n, const ParcCharModel* model,
                        float* g_dof, void* stream);

/* KinCharModel.rot_to_dof (anim/kin_char_model.py:493-507; Joint.rot_to_dof :79-100): joint quaternions
 * [n,J-1,4] -> DoFs [n,D] (hinge: signed angle about the joint axis; spherical: exp-map).  dof_out must be
 * zero-initialised by the caller for joints without DoFs to read as 0 (all D columns belong to some joint, so
 * in practice every element is written).  Forward only. */
int parc_rot_to_dof(const float* joint_rot, int64_t n, const ParcCharModel* model, float* dof_out, void* stream);

/* util/torch_util.py:414-419 exp_map_to_quat and its VJP (root rotation leaf of the optimiser). */
int parc_exp_map_to_quat_fwd(const float* exp_map, int64_t n, float* quat, void* stream);
int parc_exp_map_to_quat_bwd(const float* exp_map, const float* g_quat, int64_t n, float* g_exp_map,
                             void* stream);

/* a9: SubTerrain.get_hf_val_from_points / get_local_hf_from_terrain (util/terrain_util.py:113-130,
 * :1329-1346).  grid_idx_out (int64 [N,2]) may be NULL. */
int parc_hf_sample(const ParcHeightfield* hf, const float* xy, int64_t n, float* z_out,
                   int64_t* grid_idx_out, void* stream);

/* Self test of the hoisted-reciprocal grid index used in the observation loops: compares it with the
 * reference form clamp(rint((p - min) / cell) ...) using a true IEEE division, for EVERY float bit
 * pattern of p (2^32 inputs), and adds the number of differing indices to *mismatches_dev (device
 * uint64, caller-zeroed).  Expected: 0.  (The packed form used inside the observation sweep is checked
 * on the same inputs except quotients >= 2^63, where the reference wraps through int64 overflow to cell 0
 * and the sweep clamps to the last cell.) */
int parc_selftest_grid_index(float min_coord, float cell_size, int32_t dim, uint64_t* mismatches_dev,
                             void* stream);

/* a10/a11: RefCharEnv._refresh_ray_obs_hfs (envs/ig_parkour/mgdm_dm_util.py:158-179) and
 * sample_hf_z_on_terrain (util/terrain_util.py:2049-2082) with caller-supplied root and heading.
 * root is [N,root_stride] floats (xy at 0,1 and z at 2 when obs->relative).  heading [N], or NULL to take
 * calc_heading(root_rot[N,4]) (util/torch_util.py:470-479) inside the launch, as the caller does at
 * envs/ig_parkour/ig_parkour_env.py:641.  root_offset [N,offset_stride] (or NULL) is added to the root first:
 * the env-local -> terrain shift of _get_global_xyz_pos (:640).  obs_out rows are out_stride floats apart
 * (0 = dense, num_points), so the heightmap can land directly inside a wider policy-observation row. */
int parc_hf_obs(const ParcHeightfield* hf, const ParcObsSpec* obs, const float* root, int32_t root_stride,
                const float* heading, const float* root_rot, const float* root_offset, int32_t offset_stride,
                int64_t n, float* obs_out, int64_t out_stride, void* stream);

/* ---- per-clip terrains, packed (the MDM sampler's terrain gather) ------------------
 * MotionLib keeps one SubTerrain per clip (anim/motion_lib.py:303-321 `_terrains`) and, from the dataset-prep step,
 * per-frame lists of the cells the character's body covers (`_hf_mask_inds`, util/terrain_util.py:1951-1997).  For
 * the device they are concatenated: every clip's heightfield (row-major) and (max, min) band back to back, the
 * per-frame cell lists as bit words (bit ix*Y+iy of frame t of clip c at mask_words[mask_offset + t*W + word],
 * W = ceil(dim_x*dim_y/32) -- the frame_mask_out layout of parc_clip_label). */
typedef struct ParcClipTerrain {
  int64_t cell_offset;       /* first cell of this clip in hf / hf_maxmin */
  int64_t mask_offset;       /* first mask word of this clip, or -1 if the clip carries no masks */
  int32_t dim_x, dim_y;
  int32_t num_frames;        /* frames the mask covers */
  int32_t reserved;
  float min_x, min_y, dx, dy;
} ParcClipTerrain;

typedef struct ParcClipTerrains {
  const ParcClipTerrain* clips;   /* device [num_clips] */
  const float* hf;                /* device [total_cells] */
  const float* hf_maxmin;         /* device [total_cells, 2] (max, min) as SubTerrain.hf_maxmin, or NULL */
  const uint32_t* mask_words;     /* device, or NULL */
  int64_t num_clips;
  int32_t max_mask_words;         /* max over clips of W */
  int32_t reserved;
} ParcClipTerrains;

/* MDMHeightfieldContactMotionSampler.get_hfs_from_data (diffusion/mdm_heightfield_contact_motion_sampler.py:449-474,
 * helper :414-447) without the random augmentation: for sample i of clip motion_ids[i], a grid_x x grid_y template
 * (util/geom_util.py:210-221) is rotated by calc_heading(root_rot[i]) and moved to root_pos[i].xy, the clip's
 * terrain is sampled nearest-cell, and
 *   hf_out[i, gx, gy]        = hf(cell) - ref_z
 *   maxmin_out[i, gx, gy, :] = (band(cell) - ref_z), band = the cell's hf_maxmin where any frame of
 *                              [frame_lo[i], frame_hi[i]] marks the cell, else (free_max, free_min) = (2 max_h, 2 min_h)
 *   centre_h_out[i]          = hf at the template's centre point (centre_x, centre_y)
 * ref_z = canon_root_z[i] (RelativeZStyle.RELATIVE_TO_ROOT) or, with canon_root_z == NULL, the centre height
 * (RELATIVE_TO_ROOT_FLOOR).  maxmin_out / centre_h_out may be NULL. */
typedef struct ParcClipHfQuery {
  const int64_t* motion_ids;   /* [n] */
  const float* root_pos;       /* [n,3] */
  const float* root_rot;       /* [n,4] */
  const float* canon_root_z;   /* [n] or NULL */
  const int32_t* frame_lo;     /* [n] first frame of the sample's window (motion_time_indices[i][0]) */
  const int32_t* frame_hi;     /* [n] last frame, inclusive */
  const float* tmpl_xy;        /* [grid_x*grid_y, 2] */
  int64_t n;
  int32_t grid_x, grid_y, centre_x, centre_y;
  float free_max, free_min;
  float* hf_out;               /* [n, grid_x, grid_y] */
  float* maxmin_out;           /* [n, grid_x, grid_y, 2] or NULL */
  float* centre_h_out;         /* [n] or NULL */
} ParcClipHfQuery;

int parc_clip_hf_gather(const ParcClipTerrains* terrains, const ParcClipHfQuery* query, void* stream);

/* One terrain per sample (hf_batch_stride = X*Y, min_center_stride = 2, base_z_stride = 1) or one
 * terrain shared by the whole batch (strides 0).  x_nodes[X] / y_nodes[Y] are the torch.linspace node
 * offsets the reference adds to min_center (util/terrain_util.py:1855-1860); they are passed in rather
 * than recomputed because linspace's fp32 values are not i*dx.  base_z: device pointer (so that
 * "min(hf) - 10" of tools/procgen/mdm_path.py:77 never has to visit the host) or NULL to use
 * base_z_value. */
typedef struct ParcTerrainBatch {
  const float* hf;            /* device [B or 1, X, Y] */
  int64_t hf_batch_stride;
  const float* min_center;    /* device [B or 1, 2] */
  const float* x_nodes;       /* device [X] */
  const float* y_nodes;       /* device [Y] */
  const float* base_z;        /* device [B or 1] or NULL */
  int32_t dim_x, dim_y;
  int32_t min_center_stride;
  int32_t base_z_stride;
  float half_dx, half_dy;
  float base_z_value;
  int32_t reserved;
} ParcTerrainBatch;

/* a13: terrain_util.points_hf_sdf (util/terrain_util.py:1835-1893): exact min over ALL cells of the
 * box SDF (util/geom_util.py:122-143); solid columns [base_z, hf] or, if inverted, air columns
 * [hf, -base_z] with the result negated.  points [B,N,3] -> sdf_out [B,N]; arg_out (int32 [B,N], flat
 * cell index ix*Y+iy of the minimum, first index on ties) may be NULL.  Any batch size (launches are chunked to the
 * grid limit) and any terrain size: tiles up to ~200 KB are staged in shared memory, larger ones are scanned from
 * global memory with the same exact result. */
int parc_points_hf_sdf(const float* points, int64_t batch, int64_t n_points, const ParcTerrainBatch* terrain,
                       int32_t inverted, float* sdf_out, int32_t* arg_out, void* stream);

/* Gradient of parc_points_hf_sdf with respect to the points -- what autograd yields through
 * util/terrain_util.py:1835-1893 (the MDM's heightfield-collision loss and guidance back-propagate through it:
 * diffusion/mdm.py:729-737, :1484-1496; util/terrain_util.py:1895-1949).  arg = arg_out of the forward call with
 * the same points / terrain / inverted; g_sdf [B,N] upstream gradient; g_points_out [B,N,3] (overwritten).
 * torch.min routes the gradient to the first arg-min cell; sdBox's sub-gradients follow autograd (clamp passes at
 * the bound, norm'(0) = 0, abs'(0) = 0, max -> first index).  No gradient is produced for the heightfield. */
int parc_points_hf_sdf_bwd(const float* points, int64_t batch, int64_t n_points, const ParcTerrainBatch* terrain,
                           int32_t inverted, const int32_t* arg, const float* g_sdf, float* g_points_out,
                           void* stream);

/* Body surface samples: concatenated local points [S,3] (body-major) and per-body offsets
 * point_start[J+1] -- util/geom_util.py:788-870 produces the per-body lists. */
typedef struct ParcBodyPoints {
  const float* points;          /* device [S,3] */
  const int32_t* point_start;   /* device [J+1] */
  int32_t num_points;           /* S */
  int32_t reserved;
} ParcBodyPoints;

/* Body surface points in world space, in the layout the reference's callers build with their per-body loop
 * (util/terrain_util.py:1918-1936 motion_frames_hf_sdf_loss, diffusion/mdm.py:1006-1020 compute_point_hf_sdf):
 * body_pos [B,F,J,3], body_rot [B,F,J,4] -> points_out [B, F*S, 3] where, inside a batch entry, body b's block
 * starts at F * point_start[b] and holds [F, P_b] points frame-major:
 *   points_out[i, F*start_b + f*P_b + k] = quat_rotate(body_rot[i,f,b], points[start_b + k]) + body_pos[i,f,b].
 * _bwd is its VJP: g_points [B, F*S, 3] -> g_body_pos [B,F,J,3], g_body_rot [B,F,J,4] (either may be NULL). */
int parc_body_points_fwd(const float* body_pos, const float* body_rot, int64_t batch, int64_t frames,
                         int32_t num_bodies, const ParcBodyPoints* pts, float* points_out, void* stream);
int parc_body_points_bwd(const float* body_rot, const float* g_points, int64_t batch, int64_t frames,
                         int32_t num_bodies, const ParcBodyPoints* pts, float* g_body_pos, float* g_body_rot,
                         void* stream);

/* a14/a15: the body-point penetration and contact terms, forward AND gradient in one launch.
 * Per (sample b, frame f), with FK of (root_pos, root_rot, joint_rot)[b,f] done in-kernel:
 *   pen_out[b,f]     = sum_points max(0, -sdf_inverted(point))
 *   contact_out[b,f] = sum_body contacts[b,f,body] * min_{p in body} max(0, sdf_solid(p))
 * (tools/procgen/mdm_path.py:79-110; tools/motion_opt/motion_optimization.py:241-272); both are
 * UNWEIGHTED per-frame partial sums (sum over f is the reference's loss term).
 * If any g_* pointer is non-NULL they receive d(w_pen*pen + w_contact*contact)/d(input)[b,f]:
 * g_root_pos [B,F,3], g_root_rot [B,F,4], g_joint_rot [B,F,J-1,4] (what autograd yields through
 * the reference's FK + quat_rotate + points_hf_sdf graph, same sub-gradient/tie conventions). */
int parc_body_loss(const float* root_pos, const float* root_rot, const float* joint_rot, const float* contacts,
                   int64_t batch, int64_t frames, const ParcCharModel* model, const ParcBodyPoints* pts,
                   const ParcTerrainBatch* terrain, float w_pen, float w_contact, float* pen_out,
                   float* contact_out, float* g_root_pos, float* g_root_rot, float* g_joint_rot, void* stream);

/* ---- dataset sweep (BASELINE config 5) ---------------------------------- */

/* Raw motion frames [n, frame_stride] = root_pos(3) | root exp-map(3) | joint DoFs(D) | ... ->
 * exp_map_to_quat + dof_to_rot + forward_kinematics in one launch (the front end of
 * MotionLib.get_frames_for_id anim/motion_lib.py:503-513, terrain_util.compute_hf_mask_inds
 * util/terrain_util.py:1959-1964 and the contact-labelling functions).  Any output may be NULL. */
int parc_frames_fk(const float* frames, int64_t n, int32_t frame_stride, const ParcCharModel* model,
                   float* root_rot_out, float* joint_rot_out, float* body_pos, float* body_rot, void* stream);

#define PARC_MAX_KEY_BODIES 4
/* Bodies the labelling looks at: box-shaped feet (half extents + offset of the box geom, as
 * motion_edit_lib.py:679-686 reads them from char_model._geoms[body][0]) and sphere hands (radius). */
typedef struct ParcKeyBodies {
  int32_t num_feet;
  int32_t num_hands;
  int32_t foot_body[PARC_MAX_KEY_BODIES];
  int32_t hand_body[PARC_MAX_KEY_BODIES];
  float foot_half[PARC_MAX_KEY_BODIES][3];
  float foot_offset[PARC_MAX_KEY_BODIES][3];
  float hand_radius[PARC_MAX_KEY_BODIES];
} ParcKeyBodies;

/* Per-clip labelling over frames [B, F, frame_stride], one terrain per clip (ParcTerrainBatch; its base_z
 * is the solid columns' floor for the hand test, min(hf) - 10 in the reference):
 *   contacts_out [B,F,J]        1.0 where a foot has any box corner below (cell height + eps) or a hand's
 *                               rounded-box SDF to the solid terrain is < eps; 0 elsewhere
 *                               (motion_edit_lib.py:654-747)
 *   pen_correction_out [B,F]    min(0, min over foot corners of z - cell height)      (:683-701)
 *   body_hf_out [B,F,J]         nearest-cell terrain height under every body origin
 *   frame_mask_out [B,F,W]      bit (ix*Y+iy) set iff any body surface point falls in that cell in that
 *                               frame, W = ceil(X*Y/32)           (terrain_util.py:1951-1997, per frame)
 *   min_body_heights [B,X,Y]    running min of surface-point z per cell; CALLER-INITIALISED
 *                               (99999.9999 in the reference, :1969-1970)
 *   body_pos / body_rot         FK outputs [B,F,J,3] / [B,F,J,4]
 * Any output may be NULL. */
int parc_clip_label(const float* frames, int64_t batch, int64_t frames_per_clip, int32_t frame_stride,
                    const ParcCharModel* model, const ParcBodyPoints* pts, const ParcTerrainBatch* terrain,
                    const ParcKeyBodies* keys, float contact_eps, float* contacts_out, float* pen_correction_out,
                    float* body_hf_out, uint32_t* frame_mask_out, float* min_body_heights, float* body_pos,
                    float* body_rot, void* stream);

/* ---- f2: kinematic motion optimisation ------------------------------------------
 * tools/motion_opt/motion_optimization.py:183-395 (motion_terrain_contact_loss) and :404-500
 * (motion_contact_optimization).  One clip; the optimised leaves are the rows of `frames` [F, 6+D] =
 * root position | root exp-map | joint DoFs. */

#define PARC_CONSTRAINT_SPHERE 0
#define PARC_CONSTRAINT_BOX 1
/* One BodyConstraint (motion_optimization.py:29-32) of a body whose first geom is a sphere or a box (:299-327):
 * frames start_frame..end_frame (inclusive) pull the body's surface onto `point`.  Sphere: |sdSphere(point, rotate(
 * body_rot, offset) + body_pos, radius)|.  Box: sum over the body's first 18 surface points (the sole, :320) of
 * clamp(sdSphere(point, p, radius), min=0) with radius = 1.25 |half extents|.  A constrained (body, frame) pays no
 * sliding term (:329-332). */
typedef struct ParcBodyConstraint {
  int32_t body;
  int32_t start_frame;
  int32_t end_frame;
  int32_t shape;            /* PARC_CONSTRAINT_* */
  float point[3];
  float radius;
  float offset[3];
  float reserved;
} ParcBodyConstraint;

typedef struct ParcMotionOptArgs {
  float* frames;                      /* device [F, 6+D]: the leaves; updated in place by the Adam step */
  int64_t num_frames;                 /* F */
  const float* src_root_pos;          /* [F,3]      source clip (constants of the objective) */
  const float* src_root_rot;          /* [F,4]      quaternion */
  const float* src_joint_rot;         /* [F,J-1,4] */
  const float* src_body_vels;         /* [F-1,J,3]  body_pos[1:] - body_pos[:-1] of the source */
  const float* src_body_rot_vels;     /* [F-1,J]    quat_diff_angle(body_rot[1:], body_rot[:-1]) of the source */
  const float* contacts;              /* [F,J] */
  const ParcTerrainBatch* terrain;    /* host pointer: one terrain (strides 0), base_z = -10 in the reference */
  ParcBodyPoints pts;
  const ParcBodyConstraint* constraints;   /* device [num_constraints] or NULL */
  int32_t num_constraints;
  int32_t reserved;
  float w_root_pos, w_root_rot, w_joint_rot, w_smoothness, w_penetration, w_contact, w_sliding,
        w_body_constraints, w_jerk;
  float reserved2;
  double max_jerk_dt3;                /* max_jerk * dt^3 (dt = 1/30, :355-356), formed in double as there */
  double lr, beta1, beta2, eps;       /* torch.optim.Adam: lr = step_size, (0.9, 0.999), 1e-8 */
  float* exp_avg;                     /* [F, 6+D] Adam state, caller-zeroed */
  float* exp_avg_sq;                  /* [F, 6+D] */
  int32_t* step;                      /* device int32[1], caller-zeroed: incremented once per loss_grad launch
                                         (NULL for a loss / gradient evaluation outside an optimisation) */
  /* caller-allocated scratch / outputs */
  float* root_rot;                    /* [F,4]      quaternions of the current leaves */
  float* joint_rot;                   /* [F,J-1,4] */
  float* body_pos;                    /* [F,J,3] */
  float* body_rot;                    /* [F,J,4] */
  float* g_root_pos;                  /* [F,3]      gradient of the weighted penetration + contact terms */
  float* g_root_rot;                  /* [F,4] */
  float* g_joint_rot;                 /* [F,J-1,4] */
  float* pen;                         /* [F] unweighted per-frame penetration term (may be NULL) */
  float* con;                         /* [F] unweighted per-frame contact term (may be NULL) */
  float* grad;                        /* [F, 6+D]   d (weighted objective) / d frames */
  float* terms;                       /* [F, 8] per-frame partial sums, unweighted: root_pos, root_rot, joint_rot,
                                         smoothness, sliding, jerk, body_constraint, 0 (may be NULL) */
} ParcMotionOptArgs;

/* The source-clip constants of the objective (motion_optimization.py:428-436) from the source frames [F, 6+D], with
 * the SAME device code the objective uses on the leaves -- so a target equal to the source has velocity / tracking
 * errors of exactly 0, as in the reference (where both sides are the same torch ops); Adam would otherwise turn
 * last-bit noise into full +-lr steps.  src_root_rot [F,4], src_joint_rot [F,J-1,4], src_body_pos [F,J,3],
 * src_body_rot [F,J,4], src_body_vels [F-1,J,3], src_body_rot_vels [F-1,J].  2 launches. */
int parc_motion_opt_source(const float* src_frames, int64_t num_frames, const ParcCharModel* model,
                           float* src_root_rot, float* src_joint_rot, float* src_body_pos, float* src_body_rot,
                           float* src_body_vels, float* src_body_rot_vels, void* stream);

/* Objective and gradient at the current leaves: 3 launches (frames -> FK; penetration / contact terms + gradient;
 * every other term + the whole backward pass).  Same sub-gradient conventions as autograd on the reference. */
int parc_motion_opt_loss_grad(const ParcMotionOptArgs* args, const ParcCharModel* model, void* stream);
/* torch.optim.Adam's update of `frames` from `grad` (1 launch); bias correction from *step. */
int parc_motion_opt_adam_step(const ParcMotionOptArgs* args, const ParcCharModel* model, void* stream);
/* loss_grad + adam_step: one optimiser iteration, 4 launches, no host synchronisation -- capturable in a CUDA graph. */
int parc_motion_opt_iteration(const ParcMotionOptArgs* args, const ParcCharModel* model, void* stream);

/* ---- f3: tracker step assembly ------------------------------------------------
 * What the tracking environment computes around the motion query every control step.  One launch each. */

/* One character state, struct-of-arrays, all device float32: root_pos [n,3], root_rot [n,4] xyzw,
 * root_vel [n,3], root_ang_vel [n,3], joint_rot [n,J-1,4], dof_vel [n,D], key_pos [n,K,3] (world-space
 * positions of the key bodies; NULL iff K == 0).  root_rot / joint_rot 16-byte aligned.
 * Zero-copy views: env e reads row e * env_stride of every array (1 = dense; S for step 0 of the [n,S,...]
 * outputs of parc_motion_query_steps).  If key_body_ids (device int32 [K]) is non-NULL, key_pos is instead a
 * body-position array [rows, num_bodies, 3] (row e * env_stride) and key k is body key_body_ids[k] of it. */
typedef struct ParcCharState {
  const float* root_pos;
  const float* root_rot;
  const float* root_vel;
  const float* root_ang_vel;
  const float* joint_rot;
  const float* dof_vel;
  const float* key_pos;
  const int32_t* key_body_ids;
  int32_t num_bodies;
  int32_t env_stride;
} ParcCharState;

/* envs/base_env.py:12-16 (DoneFlags) */
#define PARC_DONE_NULL 0
#define PARC_DONE_FAIL 1
#define PARC_DONE_SUCC 2
#define PARC_DONE_TIME 3

/* Scalars of compute_done (envs/ig_parkour/mgdm_dm_util.py:399-460).  They are the reference's python
 * floats, so they travel as doubles: each meets fp32 data as its fp32 rounding, and the squared root
 * distance is formed in double first, as there.  contact_body_mask: bit b set = body b MAY touch the
 * ground (the reference's contact_body_ids); has_contact_bodies = 0 skips the fall test entirely, as an
 * empty id list does there.  pose_termination_dist: device [J-1] float32, per non-root body. */
typedef struct ParcDoneSpec {
  double episode_length;
  double termination_height;
  double root_pos_termination_dist;
  double root_rot_termination_angle;
  const float* pose_termination_dist;
  uint32_t contact_body_mask;
  int32_t has_contact_bodies;
  int32_t pose_termination;
  int32_t enable_early_termination;
  int32_t track_root;
} ParcDoneSpec;

/* compute_char_obs (envs/ig_char_env.py:582-626): obs_out [n, W],
 * W = (root_height_obs ? 1 : 0) + 6 + 3 + 3 + 6*num_joint_rots + dof_size + 3*num_keys, laid out as
 * [root z] | tan-norm of the (heading-local unless global_obs) root rotation | root_vel | root_ang_vel |
 * tan-norm of every joint rotation | dof_vel | key positions relative to the root (heading-local).
 * Rows of obs_out are out_stride floats apart (0 = dense, W): the block can be written straight into a
 * wider policy-observation row. */
int parc_char_obs(const ParcCharState* state, int64_t n, int32_t num_joint_rots, int32_t dof_size,
                  int32_t num_keys, int32_t global_obs, int32_t root_height_obs, float* obs_out,
                  int64_t out_stride, void* stream);

/* compute_tar_obs (envs/ig_parkour/mgdm_dm_util.py:462-518): future targets [n,S,...] expressed against the
 * character (ref_root_pos [n,3], ref_root_rot [n,4]); obs_out [n, S, 3 + 6 + 6*num_joint_rots + 3*num_keys] =
 * root offset | root tan-norm | joint tan-norms | key positions.  tar_key_pos [n,S,K,3] world space.
 * Zero-copy views: the target arrays hold tar_env_stride (>= S) step rows per env, of which the S starting at
 * the given pointers are used (pass S for dense arrays; S+1 and pointers at step 1 for the outputs of
 * parc_motion_query_steps); with key_body_ids (device int32 [K]) non-NULL, tar_key_pos is a body-position
 * array [rows, num_bodies, 3] over the same rows.  Env e's S*W block starts at obs_out + e * out_env_stride
 * (0 = dense, S*W). */
int parc_tar_obs(const float* ref_root_pos, const float* ref_root_rot, const float* tar_root_pos,
                 const float* tar_root_rot, const float* tar_joint_rot, const float* tar_key_pos, int64_t n,
                 int32_t num_steps, int32_t num_joint_rots, int32_t num_keys, int32_t global_obs,
                 int32_t global_tar_root_h_obs, int32_t tar_env_stride, const int32_t* key_body_ids,
                 int32_t num_bodies, float* obs_out, int64_t out_env_stride, void* stream);

/* compute_deepmimic_reward (envs/ig_parkour/mgdm_dm_util.py:328-397): reward_out [n,5] =
 * exp(-0.25 pose), exp(-0.01 vel), exp(-5 (root_pos + 0.1 root_rot)), exp(-(root_vel + 0.1 root_ang_vel)),
 * exp(-10 key_pos).  joint_rot_err_w [J-1], dof_err_w [D] device float32.  num_keys == 0 is PARC_E_SIZE
 * (the reference raises on it too). */
int parc_deepmimic_reward(const ParcCharState* sim, const ParcCharState* tar, int64_t n, int32_t num_joint_rots,
                          int32_t dof_size, int32_t num_keys, const float* joint_rot_err_w,
                          const float* dof_err_w, int32_t track_root_h, int32_t track_root, float* reward_out,
                          void* stream);

/* compute_done with the termination-height lookup of RefCharEnv.update_done fused in
 * (envs/ig_parkour/mgdm_dm_util.py:205-230, :399-460).  time [n]; body_pos / tar_body_pos / contact_force
 * [n,J,3]; root_rot / tar_root_rot [n,4].  Heights: pass term_heights [n,J] (compute_done's own argument),
 * or NULL to sample hf at body xy + env_offsets[:, 0:2] (env_offsets [n, offset_stride], may be NULL) and
 * add termination_height.  tar_root_rot / tar_body_pos read row e * tar_env_stride (1 = dense).
 * done_out [n] int32 (PARC_DONE_*); term_heights_out [n,J] optional. */
int parc_done(const ParcDoneSpec* spec, const float* time, const float* root_rot, const float* body_pos,
              const float* tar_root_rot, const float* tar_body_pos, const float* contact_force,
              const float* term_heights, const ParcHeightfield* hf, const float* env_offsets,
              int32_t offset_stride, int32_t tar_env_stride, int64_t n, int32_t num_bodies, int32_t* done_out,
              float* term_heights_out, void* stream);

/* The simulated character's whole share of a control step in ONE launch: DoF -> joint rotations (never leaving
 * registers), compute_char_obs, compute_deepmimic_reward, compute_done with the termination-height lookup, and
 * the two contact-flag blocks of the policy-observation row.  Same device code as parc_dof_to_rot_fwd +
 * parc_char_obs + parc_deepmimic_reward + parc_done; flags identical, values to fp32 rounding.
 * sim.joint_rot is ignored (dof_pos is converted); sim.key_pos / ref.key_pos follow ParcCharState's rules.
 * ref is the reference frame the character tracks (typically step 0 of parc_motion_query_steps' output:
 * env_stride = number of steps); ref_body_pos [rows, J, 3] over the same rows feeds the pose-termination test.
 * Heights for the fall test are sampled from hf at body xy + env_offsets[:, 0:2].
 * tar_contacts (step 1 of a [n, tar_env_stride, J] array) / char_contacts [n,J] are copied into
 * tar_contacts_out / char_contacts_out when those are non-NULL; all three observation outputs share obs_stride
 * (floats between consecutive envs), i.e. they are column blocks of one policy-observation buffer. */
typedef struct ParcSimStep {
  ParcCharState sim;
  ParcCharState ref;
  const float* dof_pos;            /* [n, D] */
  const float* body_pos;           /* [n, J, 3] simulated body positions */
  const float* ref_body_pos;       /* [rows, J, 3], row e * ref.env_stride */
  const float* contact_force;      /* [n, J, 3] (NULL unless the fall test is enabled) */
  const float* time;               /* [n] */
  const float* env_offsets;        /* [n, offset_stride] or NULL */
  const float* joint_rot_err_w;    /* [J-1] */
  const float* dof_err_w;          /* [D] */
  const float* tar_contacts;
  const float* char_contacts;
  ParcDoneSpec done;
  ParcHeightfield hf;
  int32_t offset_stride;
  int32_t tar_env_stride;
  int32_t num_tar_steps;
  int32_t num_keys;
  int32_t global_obs, root_height_obs, track_root_h, track_root;
  float* joint_rot_out;            /* [n, J-1, 4] or NULL */
  float* char_obs_out;
  float* tar_contacts_out;         /* or NULL */
  float* char_contacts_out;        /* or NULL */
  float* reward_out;               /* [n, 5] */
  int32_t* done_out;               /* [n] */
  int64_t obs_stride;
  int32_t phase;                   /* PARC_SIM_STEP_ALL, or one half of the step (see below); appended in ABI 3 */
  int32_t reserved;
} ParcSimStep;

/* phase: the step splits where it starts to need the reference frame.  _PRE = DoF conversion (joint rotations to
 * joint_rot_out, required), character observation, character contact block: depends on the simulator state only, so it
 * can run BESIDE parc_motion_query_steps.  _POST = target contact block, reward terms, episode flag: reads
 * joint_rot_out back (the same bits the single launch keeps in registers).  _PRE then _POST == _ALL: observations,
 * contact blocks and episode flags bit for bit, reward terms to fp32 rounding. */
#define PARC_SIM_STEP_ALL 0
#define PARC_SIM_STEP_PRE 1
#define PARC_SIM_STEP_POST 2

int parc_sim_step(const ParcSimStep* args, int64_t n, const ParcCharModel* model, void* stream);

/* ---- f4: GPU loader -----------------------------------------------------------------
 * Raw clip frames -> packed frame rows, every frame of every clip in one launch: what MotionLib._load_motions /
 * _load_motion_frames do per clip on the host (anim/motion_lib.py:137-202, :204-380 -- extract_frame_data with
 * quat_pos on the joints :405-423, forward-difference root velocities with the last frame repeated :281-288,
 * KinCharModel.compute_frame_dof_vel anim/kin_char_model.py:543-581) plus parc_pack_frames.
 * frames [total, frame_stride >= 6 + D] (root_pos | root exp-map | DoFs), contacts [total, J] or NULL,
 * frame_clip [total] int32 = clip of each frame, per-clip start / num_frames (int64), fps and the divisor of the
 * DoF velocities (1/fps -- or fps itself to reproduce the motion_frames quirk of :178).  rows_out
 * [total, row_floats], 16-byte aligned.  Clips of a single frame get zero velocities. */
int parc_build_tables(const float* frames, int64_t total_frames, int32_t frame_stride, const float* contacts,
                      const int32_t* frame_clip, const int64_t* clip_start, const int64_t* clip_num_frames,
                      const float* clip_fps, const float* clip_dof_vel_dt, int64_t num_clips,
                      const ParcCharModel* model, float* rows_out, void* stream);

/* ---- multi-GPU: gathering the shards of a sharded query over NVLink / NVSwitch peer memory ------------
 * The reference is single-GPU; this has no counterpart in it.  BASELINE config 4 splits 65 536 envs contiguously over
 * the GPUs of a box; when a consumer wants the full `body_pos` / `obs` batch on every GPU, the shards are exchanged
 * through SYMMETRIC memory (the same allocation on every GPU, peer-mapped into every process, optionally behind an
 * NVSwitch multicast address).  The library does not allocate or map that memory: the caller does (the Python side
 * uses torch.distributed._symmetric_memory) and passes plain addresses.
 *
 * Signals: `num_slots` uint64 counters at the same offset of every rank's symmetric buffer, zero-initialised, plus a
 * rank-private `epoch[num_slots]` (ordinary device memory, zero-initialised).  A hand-shake on slot s adds 1 to slot s
 * on every rank (release) and waits until the local copy reaches world * (epoch[s] + 1) (acquire), then bumps
 * epoch[s]; the launch is therefore replayable from a CUDA graph. */
#define PARC_MAX_PEERS 16
#define PARC_MAX_PUSH_SEGMENTS 4

typedef struct ParcPeerSignals {
  uint64_t* multicast_signal;              /* multicast address of the slots, or NULL (then peer_signal is used) */
  uint64_t* peer_signal[PARC_MAX_PEERS];   /* rank r's slots as mapped into THIS process (r < world; incl. own) */
  uint64_t* local_signal;                  /* this rank's slots */
  uint64_t* epoch;                         /* [num_slots] rank-private launch counters */
  int32_t world;
  int32_t num_slots;
  int64_t timeout_ns;                      /* > 0: a hand-shake that waits longer gives up, ORs 1 into *timeout_flag and
                                              lets the kernel finish (a peer that never launches must not hang the GPU);
                                              0 = wait for ever */
  int32_t* timeout_flag;                   /* device word, or NULL */
  int32_t rank;                            /* this rank: peer-pointer stores start at the next rank and go round, so that
                                              the ranks do not all write into the same peer at the same moment */
  int32_t reserved;
} ParcPeerSignals;

typedef struct ParcPeerSegment {
  const void* src;                         /* this rank's shard (local memory) */
  void* dst_multicast;                     /* multicast address of the shard's place in the gathered tensor, or NULL */
  void* dst_peer[PARC_MAX_PEERS];          /* the same place in rank r's buffer (used when dst_multicast is NULL) */
  int64_t bytes;                           /* multiple of 4; 16-byte vectors are used when everything is 16-aligned */
} ParcPeerSegment;

/* All ranks call it on their stream after the kernels that stored into peer / multicast addresses: returns (on the
 * stream) once every rank has arrived, with the ranks' earlier stores visible.  One 32-thread block. */
int parc_peer_barrier(const ParcPeerSignals* signals, int32_t slot, void* stream);

/* Copies up to PARC_MAX_PUSH_SEGMENTS local shards into every rank's gathered tensors (one multicast store per 16
 * bytes, or one store per peer without multicast) and performs the hand-shake per block (slots 0..num_blocks-1;
 * num_blocks <= 0 selects 64): when the kernel has finished on a rank, every rank's shards have arrived there. */
int parc_peer_push(const ParcPeerSegment* segments, int32_t num_segments, const ParcPeerSignals* signals,
                   int32_t num_blocks, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PARC_B200_H_ */
