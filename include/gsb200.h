/*
 * gsb200.h -- C ABI of libgsb200.so, the B200-native (sm_100a) 3D Gaussian Splatting rasterizer.
 *
 * This is the drop-in boundary for the reference's (zhujinchong/3DGS-native) rasterizer hot path.
 * The reference has no FFI of its own: its boundary is the Python operator surface
 *     forward.render_gaussians (forward.py:629-894), backward.backward (backward.py:955-1196),
 *     and the optimizer / densify kernels train.py launches directly (optimizer.py:6-415).
 * Each entry point below names the reference function / kernel launch it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; memory is caller-owned;
 *     the context only owns scratch (sort double-buffers, histograms, scan state, pinned scalars)
 *   - gsb_stream is a cudaStream_t; all work is enqueued on it, nothing synchronises unless the
 *     function is documented to (those that return a count to the host)
 *   - layouts are the reference's: positions/scales float[N][3], rotations float[N][4] in
 *     (x,y,z,w) order, opacities float[N], SH float[N][16][3]; matrices are 4x4 row-major and
 *     are applied as row-vector * matrix (translation in the last row)
 *   - return value: 0 (GSB_OK) or a negative gsb_status; gsb_last_error_string(ctx) explains
 *   - a context is thread-compatible (one thread at a time); use one context per stream
 *   - there is NO CPU fallback: every call needs a CUDA device
 */
#ifndef GSB200_H_
#define GSB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSB_VERSION 100 /* 0.1.0 */
#define GSB_TILE 16     /* config.py:21-22 TILE_M = TILE_N = 16 */
#define GSB_MAX_RENDERED ((1LL << 30) - 1)
#define GSB_GRAD_FLOATS 59 /* 3 pos + 3 scale + 4 rot + 1 opacity + 48 SH */

typedef enum gsb_status {
  GSB_OK = 0,
  GSB_ERR_INVALID = -1,      /* bad argument */
  GSB_ERR_CUDA = -2,         /* a CUDA call failed, see gsb_last_error_string */
  GSB_ERR_TOO_MANY = -3,     /* num_rendered > 2^30 (forward.py:765-767 raises ValueError) */
  GSB_ERR_CAPACITY = -4,     /* caller-provided point_list / key capacity too small */
  GSB_ERR_NOMEM = -5
} gsb_status;

typedef struct gsb_ctx gsb_ctx;
typedef void* gsb_stream;

/* Per-view constants: the scalar arguments of wp_preprocess / wp_render_gaussians
 * (forward.py:189-223, 384-403) and of the backward kernels. */
typedef struct gsb_frame {
  float view[16];       /* viewmatrix  (world_to_camera, row-major, row-vector convention) */
  float proj[16];       /* projmatrix  (full_proj_matrix) */
  float campos[3];
  float tan_fovx, tan_fovy;
  float scale_modifier; /* forward only: backward ignores it (reference quirk G3) */
  float background[3];
  int32_t width, height;
  int32_t degree;       /* SH degree 0..3; SH rows always have stride 16 */
  int32_t clamped;      /* forward.py:364 */
} gsb_frame;

/* ---- context ---------------------------------------------------------------------------- */
int gsb_version(void);
int gsb_create(gsb_ctx** ctx, int device);
int gsb_destroy(gsb_ctx* ctx);
const char* gsb_last_error_string(gsb_ctx* ctx);
/* pre-size the binning scratch for `num_rendered` duplicates (grow-only; optional) */
int gsb_reserve(gsb_ctx* ctx, gsb_stream s, int64_t num_rendered);
/* number of kernel launches issued through this context so far (bench.py gpu_launches) */
int64_t gsb_launch_count(gsb_ctx* ctx);
/* tuning / A-B knobs outside the reference surface (per context; results never depend on them):
 *   "blend_cull" = 1 (default) / 0: per-block culling masks in the tile kernels
 *   "bwd_reduce" = 2 (default; 1 is a synonym): the backward tile kernel sums the per-pixel terms over a warp's
 *                  pixel block with TF32 tensor-core products; 0: warp-shuffle butterfly (and the contract's
 *                  exponential instead of MUFU: the slow, independent implementation the fast one is tested beside)
 *   "tile_sort"  = 3 (default): one-pass bucket sort of every tile's segment by depth (order-preserving map of the
 *                  depth bits onto >= n buckets, shared-memory atomics, every entry ranks itself inside its bucket;
 *                  bitonic network for degenerate depth distributions and for tiles of more than 4096 entries);
 *                  0: always the bitonic network; 1: always the per-tile LSD radix sort; 2: radix when the frame's
 *                  longest tile list exceeds 2048 entries, bitonic otherwise (round 2's first default)
 *   "speculate"  = 1 (default): gsb_forward queues scatter / per-tile sort / blend behind the tile scan without
 *                  waiting for num_rendered on the host -- the previous frame's capacity and list-length class are
 *                  assumed and checked on the device; a frame that does not fit continues on the waiting path;
 *                  0: always wait first
 *   "pdl"        = 1 (default): the kernels of a frame / training step are launched as programmatic dependents of
 *                  the kernel in front of them (they reach the SMs during its last wave and wait there for its
 *                  completion: launch latency and prologue hidden, never a data dependency); 0: plain launches
 *   "sort_coop"  = 1 (default): gsb_sort_pairs64 sorts inputs of up to num_sms x 12288 pairs with all passes in ONE
 *                  cooperative launch; 0: three kernels per 8-bit pass (any size)
 *   "binning"    = 0 (default): gsb_forward bins by tile with a counting sort and sorts every tile's
 *                  segment in shared memory; 1: duplicate-with-keys + global 64-bit radix sort */
int gsb_set_option(gsb_ctx* ctx, const char* name, int value);

/* ---- forward stages --------------------------------------------------------------------- */

/* replaces wp_preprocess (forward.py:189-382, launched 719-752).  Writes EVERY output element
 * (culled Gaussians get zeros; cov3Ds is written before the later culls like forward.py:260),
 * so outputs need no pre-zeroing. */
int gsb_preprocess(gsb_ctx* ctx, gsb_stream s, const gsb_frame* f, int32_t n, const float* means,
                   const float* scales, const float* rotations, const float* opacities, const float* shs,
                   int32_t* radii, float* points_xy, float* depths, float* cov3Ds, float* rgb,
                   float* conic_opacity, int32_t* tiles_touched, float* clamped_state);

/* replaces wp_prefix_sum (utils/wp_utils.py:46-60, launched forward.py:755-763): inclusive scan.
 * If num_rendered_host != NULL the call synchronises the stream and returns point_offsets[n-1]
 * (the reference's .item() at forward.py:764). */
int gsb_scan_tiles(gsb_ctx* ctx, gsb_stream s, int32_t n, const int32_t* tiles_touched, int32_t* point_offsets,
                   int64_t* num_rendered_host);

/* replaces wp_duplicate_with_keys (forward.py:517-558): key = (tile_id << 32) | bits(depth) */
int gsb_duplicate_with_keys(gsb_ctx* ctx, gsb_stream s, int32_t width, int32_t height, int32_t n,
                            const float* points_xy, const float* depths, const int32_t* point_offsets,
                            const int32_t* radii, int64_t num_rendered, int64_t* keys, int32_t* values);

/* replaces the pad-copy + wp.utils.radix_sort_pairs + copy-back of forward.py:791-824: ascending
 * STABLE sort of (int64 key, int32 value) pairs over key bits [begin_bit, end_bit).  keys/values
 * hold the result on return; tmp_* are scratch of the same size (NULL: use context scratch). */
int gsb_sort_pairs64(gsb_ctx* ctx, gsb_stream s, int64_t* keys, int32_t* values, int64_t* tmp_keys,
                     int32_t* tmp_values, int64_t count, int begin_bit, int end_bit);

/* replaces wp_identify_tile_ranges (forward.py:560-586).  ranges = int32[num_tiles][2]; zeroed here */
int gsb_tile_ranges(gsb_ctx* ctx, gsb_stream s, int64_t num_rendered, const int64_t* sorted_keys,
                    int32_t num_tiles, int32_t* ranges);

/* replaces forward.py:776-840 as a whole (wp_duplicate_with_keys + radix_sort_pairs +
 * wp_identify_tile_ranges): counting sort by tile + per-tile shared-memory sort on (depth, id).
 * Produces the same point_list / ranges as the three stages above.  Needs point_offsets from
 * gsb_scan_tiles.  Synchronises once; *num_rendered_host receives D.  Returns GSB_ERR_CAPACITY
 * if D > point_list_capacity (ranges valid, point_list untouched).  *used_tile_path_host (may be
 * NULL) tells whether the per-tile path ran (1) or a tile was longer than 16384 entries and the
 * global radix sort was used instead (0). */
int gsb_bin_by_tile(gsb_ctx* ctx, gsb_stream s, int32_t width, int32_t height, int32_t n, const float* points_xy,
                    const float* depths, const int32_t* radii, const int32_t* point_offsets, int32_t* point_list,
                    int64_t point_list_capacity, int32_t* ranges, int64_t* num_rendered_host,
                    int32_t* used_tile_path_host);

/* replaces wp_render_gaussians (forward.py:384-515) and the no-op track_pixel_stats (589-627).
 * image float[H][W][3], inv_depth float[H][W], final_T float[H][W], n_contrib int32[H][W].
 * block_masks (int32, one per point_list entry; may be NULL) is an extra output with no counterpart
 * in the reference: the per-entry culling mask the kernel computes anyway (which 8-pixel half rows
 * of the tile the Gaussian can reach with alpha >= 1/255).  Passing it on to gsb_blend_backward
 * saves recomputing it there; entries behind the point where a tile's pixels all terminated are
 * not written (the backward never reads them). */
int gsb_blend_forward(gsb_ctx* ctx, gsb_stream s, const gsb_frame* f, const int32_t* ranges,
                      const int32_t* point_list, const float* points_xy, const float* rgb,
                      const float* conic_opacity, const float* depths, float* image, float* inv_depth,
                      float* final_T, int32_t* n_contrib, int32_t* block_masks);

/* The whole of render_gaussians (forward.py:629-894) in one call.  point_list has room for
 * point_list_capacity entries; *num_rendered_host receives D.  Returns GSB_ERR_CAPACITY (with D
 * set, per-Gaussian outputs valid, image untouched) when D > capacity so the caller can grow
 * and call again, GSB_ERR_TOO_MANY when D > 2^30.  When D == 0 the image outputs are all ZERO
 * (not background), like the reference (forward.py:830).  Synchronises once (to learn D).
 * point_offsets may be NULL when the caller does not need that output (the inclusive scan of tiles_touched is not an
 * input of this binning): the trainer's loop passes NULL and saves the three scan kernels. */
int gsb_forward(gsb_ctx* ctx, gsb_stream s, const gsb_frame* f, int32_t n, const float* means,
                const float* scales, const float* rotations, const float* opacities, const float* shs,
                int32_t* radii, int32_t* point_offsets, float* points_xy, float* depths, float* rgb,
                float* cov3Ds, float* conic_opacity, float* clamped_state, int32_t* point_list,
                int64_t point_list_capacity, int32_t* ranges, float* image, float* inv_depth, float* final_T,
                int32_t* n_contrib, int64_t* num_rendered_host, int32_t* block_masks /* optional, see above */);

/* ---- backward stages -------------------------------------------------------------------- */

/* replaces wp_render_backward_kernel (backward.py:558-706, launched 932-953).  Accumulates into
 * dL_dmean2D float[N][3] (z stays 0), dL_dconic float[N][4] (a, b, 0, c), dL_dopacity float[N],
 * dL_dcolor float[N][3]; the four arrays are zeroed here first (backward.py:1113-1116).
 * block_masks: the optional output of gsb_blend_forward for the SAME point_list / ranges, or NULL
 * (the masks are then recomputed; results are identical). */
int gsb_blend_backward(gsb_ctx* ctx, gsb_stream s, const gsb_frame* f, int32_t n, const int32_t* ranges,
                       const int32_t* point_list, const float* points_xy, const float* conic_opacity,
                       const float* rgb, const float* final_T, const int32_t* n_contrib, const float* dL_dpixels,
                       float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                       const int32_t* block_masks);

/* replaces backward_preprocess (backward.py:770-888): compute_cov2d_backward_kernel,
 * compute_projection_backward_kernel, sh_backward_kernel and compute_cov3d_backward_kernel fused
 * into one pass.  Writes every element of dL_dmean3D float[N][3], dL_dshs float[N][16][3],
 * dL_dscale float[N][3], dL_drot float[N][4].  dL_dcov3D_internal (float[N][6], may be NULL) receives
 * the buffer of backward.py:812 that the reference never returns. */
int gsb_preprocess_backward(gsb_ctx* ctx, gsb_stream s, const gsb_frame* f, int32_t n, const float* means,
                            const int32_t* radii, const float* shs, const float* scales, const float* rotations,
                            const float* cov3Ds, const float* clamped_state, const float* dL_dmean2D,
                            const float* dL_dconic, const float* dL_dcolor, float* dL_dmean3D, float* dL_dshs,
                            float* dL_dscale, float* dL_drot, float* dL_dcov3D_internal);

/* The whole of backward() (backward.py:955-1196): gsb_blend_backward + gsb_preprocess_backward.
 * dL_dcov3D (float[N][6]) is zero-filled: the reference returns a buffer its kernels never touch. */
int gsb_backward(gsb_ctx* ctx, gsb_stream s, const gsb_frame* f, int32_t n, const float* means,
                 const float* opacities, const float* shs, const float* scales, const float* rotations,
                 const int32_t* radii, const float* points_xy, const float* conic_opacity, const float* rgb,
                 const float* clamped_state, const float* cov3Ds, const int32_t* point_list, const int32_t* ranges,
                 const float* final_T, const int32_t* n_contrib, const float* dL_dpixels, float* dL_dmean3D,
                 float* dL_dcolor, float* dL_dshs, float* dL_dopacity, float* dL_dscale, float* dL_drot,
                 float* dL_dmean2D, float* dL_dconic, float* dL_dcov3D, const int32_t* block_masks /* optional */);

/* gsb_preprocess_backward with the SH gradient in COMPACT form (see gsb_backward_compact_sh):
 * dL_dshs_compact is float[N][8]. */
int gsb_preprocess_backward_compact_sh(gsb_ctx* ctx, gsb_stream s, const gsb_frame* f, int32_t n, const float* means,
                                       const int32_t* radii, const float* shs, const float* scales,
                                       const float* rotations, const float* cov3Ds, const float* clamped_state,
                                       const float* dL_dmean2D, const float* dL_dconic, const float* dL_dcolor,
                                       float* dL_dmean3D, float* dL_dshs_compact, float* dL_dscale, float* dL_drot,
                                       float* dL_dcov3D_internal);

/* gsb_backward with the SH gradient in COMPACT form (no reference counterpart; used by the view-sharded
 * trainer).  For one view the reference's dL_dshs[16 i + k] = basis_k(dir_i) * dL_dRGB_i
 * (sh_backward_kernel, backward.py:116-213) is a rank-1 product, so dL_dshs here receives only its two
 * factors: float[N][8] = (dL_dRGB after the clamp mask, unit view direction, 0, 0); zeros for a
 * Gaussian the reference skips.  Everything else is identical to gsb_backward. */
int gsb_backward_compact_sh(gsb_ctx* ctx, gsb_stream s, const gsb_frame* f, int32_t n, const float* means,
                            const float* opacities, const float* shs, const float* scales, const float* rotations,
                            const int32_t* radii, const float* points_xy, const float* conic_opacity, const float* rgb,
                            const float* clamped_state, const float* cov3Ds, const int32_t* point_list,
                            const int32_t* ranges, const float* final_T, const int32_t* n_contrib,
                            const float* dL_dpixels, float* dL_dmean3D, float* dL_dcolor, float* dL_dshs,
                            float* dL_dopacity, float* dL_dscale, float* dL_drot, float* dL_dmean2D, float* dL_dconic,
                            float* dL_dcov3D, const int32_t* block_masks /* optional */);

/* ---- optimizer / densify (optimizer.py, train.py nested kernels) -------------------------- */

/* replaces adam_update (optimizer.py:6-139, launched train.py:750-794); in place. */
int gsb_adam_step(gsb_ctx* ctx, gsb_stream s, int32_t n, const float* g_pos, const float* g_scale,
                  const float* g_rot, const float* g_opac, const float* g_sh, float lr_pos, float lr_scale,
                  float lr_rot, float lr_opac, float lr_sh, float beta1, float beta2, float epsilon,
                  int32_t iteration, float* pos, float* scales, float* rots, float* opac, float* shs, float* m_pos,
                  float* m_scale, float* m_rot, float* m_opac, float* m_sh, float* v_pos, float* v_scale,
                  float* v_rot, float* v_opac, float* v_sh);

/* gsb_adam_step for a subset of the tensors (no reference counterpart; same arithmetic, same bits): phase 1 = positions,
 * scales, rotations, opacities; phase 2 = the SH coefficients (81% of the update's bytes); phase 0 = all.  A trainer may
 * run phase 2 on a side stream beside the next frame's geometry preprocess and binning, which do not read SH, and hand
 * that stream's event to the next gsb_forward with gsb_set_color_dependency (Trainer(overlap_sh=...)). */
int gsb_adam_step_phase(gsb_ctx* ctx, gsb_stream s, int32_t n, const float* g_pos, const float* g_scale,
                  const float* g_rot, const float* g_opac, const float* g_sh, float lr_pos, float lr_scale,
                  float lr_rot, float lr_opac, float lr_sh, float beta1, float beta2, float epsilon,
                  int32_t iteration, float* pos, float* scales, float* rots, float* opac, float* shs, float* m_pos,
                  float* m_scale, float* m_rot, float* m_opac, float* m_sh, float* v_pos, float* v_scale,
                  float* v_rot, float* v_opac, float* v_sh, int32_t phase);

/* Layout of the "flat" per-role buffer used by the data-parallel trainer: positions | scales |
 * rotations | opacities | SH back to back, every segment starting on a 16-byte boundary.
 * offsets5 receives the five offsets in floats, *total the length in floats. */
int gsb_flat_layout(int32_t n, int64_t* offsets5, int64_t* total);

/* Fused gradient exchange + Adam for view-sharded data parallelism (no reference counterpart:
 * the reference trains on one device).  grad_ptrs_host / param_ptrs_host are HOST arrays of `world`
 * device addresses: every rank's flat gradient / parameter buffer as mapped into THIS process
 * (symmetric memory / CUDA IPC); grad_multicast / param_multicast are NVLS multicast addresses of
 * the same buffers or 0.  The kernel sums the gradients of this rank's shard of Gaussians over all
 * ranks (multimem.ld_reduce through the NVSwitch, or one NVLink load per peer), applies
 * adam_update (optimizer.py:6-139) to the shard of m_flat / v_flat, and stores the new parameters
 * into every rank's buffer.  The caller orders it between two cross-rank barriers.
 * publish_position_grad != 0 (densify steps): the SUMMED position gradient of the shard is also stored
 * back into every rank's gradient buffer, so that after the closing barrier every rank's positions
 * segment holds the all-reduced gradient that compute_grad_norms / mark_*_candidates (train.py:398-433)
 * read -- every rank then marks, clones, splits and prunes the same Gaussians. */
int gsb_adam_step_peers(gsb_ctx* ctx, gsb_stream s, int32_t n, int32_t world, int32_t rank,
                        const uint64_t* grad_ptrs_host, const uint64_t* param_ptrs_host, uint64_t grad_multicast,
                        uint64_t param_multicast, float* m_flat, float* v_flat, float lr_pos, float lr_scale,
                        float lr_rot, float lr_opac, float lr_sh, float beta1, float beta2, float epsilon,
                        int32_t iteration, int32_t publish_position_grad);

/* gsb_adam_step_peers for gradient buffers written by gsb_backward_compact_sh: the SH segment of every
 * rank's flat gradient buffer starts with that rank's float[N][8] factors.  The owner of a Gaussian
 * rebuilds each rank's 48 SH gradients from them (same basis code as the backward, same rank order of
 * the sum as gsb_adam_step_peers: identical bits) into sh_local -- device scratch of sh_local_floats >=
 * 48 * (Gaussians of this rank's shard) floats, need not be peer-visible -- and the fused exchange + Adam
 * reads its SH gradient from there.  NVLink carries 76 instead of 236 bytes per Gaussian and peer on the
 * gradient side.  Valid when every rank contributed exactly ONE view (the factors of two views do not add). */
int gsb_adam_step_peers_compact(gsb_ctx* ctx, gsb_stream s, int32_t n, int32_t world, int32_t rank,
                                const uint64_t* grad_ptrs_host, const uint64_t* param_ptrs_host,
                                uint64_t param_multicast, float* m_flat, float* v_flat, float lr_pos, float lr_scale,
                                float lr_rot, float lr_opac, float lr_sh, float beta1, float beta2, float epsilon,
                                int32_t iteration, float* sh_local, int64_t sh_local_floats, int32_t degree,
                                int32_t publish_position_grad);

/* The fused exchange in TWO phases, so that the larger one can run beside the next frame (no reference counterpart:
 * the reference has no multi-GPU path; same arithmetic and the same bits as the one-call forms above):
 *   phase 1: positions, scales, rotations, opacities (11 of the 59 floats per Gaussian; the published position
 *            gradient of densify steps belongs here);   phase 2: the SH coefficients (48 of 59), compact or full;
 *   phase 0: everything (= gsb_adam_step_peers / gsb_adam_step_peers_compact).
 * sh_local == NULL selects the full SH exchange (grad_multicast / param_multicast as in gsb_adam_step_peers), otherwise
 * the compact one (grad_multicast must be 0).  Typical use (Trainer(overlap_sh=True)): barrier, phase 1, barrier on the
 * main stream; phase 2 + barrier on a side stream; the next gsb_forward is given that stream's event with
 * gsb_set_color_dependency and evaluates the colours only behind it. */
int gsb_adam_step_peers_phase(gsb_ctx* ctx, gsb_stream s, int32_t n, int32_t world, int32_t rank,
                              const uint64_t* grad_ptrs_host, const uint64_t* param_ptrs_host, uint64_t grad_multicast,
                              uint64_t param_multicast, float* m_flat, float* v_flat, float lr_pos, float lr_scale,
                              float lr_rot, float lr_opac, float lr_sh, float beta1, float beta2, float epsilon,
                              int32_t iteration, float* sh_local, int64_t sh_local_floats, int32_t degree,
                              int32_t publish_position_grad, int32_t phase);

/* One-shot: the NEXT gsb_forward on this context reads the SH coefficients only behind `cuda_event` (a cudaEvent_t
 * recorded on any stream; NULL clears).  That call runs preprocess without the SH -> RGB evaluation, bins, then makes
 * its stream wait for the event and evaluates rgb / clamped_state with a second kernel right in front of the blend:
 * identical outputs.  Everything else the frame reads (means, scales, rotations, opacities) must be final as usual. */
int gsb_set_color_dependency(gsb_ctx* ctx, void* cuda_event);

/* Diagnostic (no reference counterpart): the reference's per-pixel loops (forward.py:454-501, backward.py:633-706)
 * walked by one thread per pixel with the arithmetic contract, counting work and decisions on a rendered frame.
 * counters7 (device, 7 x uint64, zeroed by the call):
 *   [0] K_fwd: (pixel, Gaussian) pairs the forward loop iterates   [1] pairs that blend   [2] K_bwd: sum of
 *   min(list length, n_contrib)   [3] backward pairs whose alpha gets evaluated   [4] of those, pairs on which the
 *   backward kernel's test (MUFU alpha < 1/255) disagrees with the forward's decision (measured: ~2e-8 of [3])
 *   [5] pairs the conservative exponent threshold skips although the forward blended them: must be 0
 *   [6] evaluated pairs whose MUFU alpha lies within 1e-7 of 1/255. */
int gsb_selftest_work_counters(gsb_ctx* ctx, gsb_stream s, const gsb_frame* f, const int32_t* ranges,
                               const int32_t* point_list, const float* points_xy, const float* conic_opacity,
                               const int32_t* n_contrib, uint64_t* counters7);

/* replaces zero_gradients (train.py:94-115) -- and any other "fill float" need */
int gsb_fill_f32(gsb_ctx* ctx, gsb_stream s, float* dst, int64_t count, float value);

/* Diagnostic (no reference counterpart): the tile kernels skip a (Gaussian, 8-pixel half row) when a
 * conservative ellipse / row test says no pixel of it can reach alpha >= 1/255.  cases: count x 8 floats
 * (gx, gy, conic a, b, c, opacity, tile x0, y0); mask[i] = the 32-bit mask (bit 2r + h: row r, half h);
 * active[i][8] = the 256 pixels (bit 16r + x) that do pass power <= 0 and alpha >= 1/255 under the
 * arithmetic contract.  Every active pixel must be covered by a mask bit. */
int gsb_selftest_block_mask(gsb_ctx* ctx, gsb_stream s, int32_t count, const float* cases, uint32_t* mask,
                            uint32_t* active);

/* Diagnostic (no reference counterpart): the Adam kernels divide with a rescaled fast path instead of
 * the compiler's range-checked `/`; out_fast[i] = that division of a[i] by b[i] (b > 0), out_const[i]
 * = the variant for host-known divisors (reciprocal passed in), out_ref[i] = a[i] / b[i].  The three
 * are bit-identical for every input. */
int gsb_selftest_div(gsb_ctx* ctx, gsb_stream s, int64_t count, const float* a, const float* b, float* out_fast,
                     float* out_const, float* out_ref);
/* out += in (view-batch gradient accumulation; no reference counterpart: batch size is 1 there) */
int gsb_accumulate_f32(gsb_ctx* ctx, gsb_stream s, float* out, const float* in, int64_t count);

/* replaces init_gaussian_params (train.py:36-92) */
int gsb_init_gaussian_params(gsb_ctx* ctx, gsb_stream s, int32_t n, float init_scale, float* pos, float* scales,
                             float* rots, float* opac, float* shs);

/* replaces compute_grad_norms (train.py:398-405) + mark_clone_candidates / mark_split_candidates
 * (optimizer.py:212-242 / 180-210).  grad_norms has n_grads entries; entries i >= n_grads count
 * as 0 (the reference reads a stale shorter array there, quirk G4).  want_split: 0 clone, 1 split */
int gsb_grad_norms(gsb_ctx* ctx, gsb_stream s, int32_t n, const float* pos_grad, float* grad_norms);
int gsb_mark_candidates(gsb_ctx* ctx, gsb_stream s, int32_t n, int32_t n_grads, const float* grad_norms,
                        const float* scales, float grad_threshold, float scene_extent, float percent_dense,
                        int32_t want_split, int32_t* mask);

/* replaces wp.utils.array_scan(inclusive=False) + .numpy()[-1] (train.py:431-433 etc.): exclusive
 * scan; *last_host receives out[n-1] (the reference's "total", which drops the last flag: quirk
 * G5).  Synchronises when last_host != NULL. */
int gsb_scan_mask(gsb_ctx* ctx, gsb_stream s, int32_t n, const int32_t* mask, int32_t* prefix, int32_t* last_host);

/* replaces clone_gaussians (optimizer.py:312-362); out arrays hold new_n Gaussians; writes past
 * new_n (quirk G5) are dropped */
int gsb_clone_gaussians(gsb_ctx* ctx, gsb_stream s, int32_t n, int32_t new_n, const int32_t* mask,
                        const int32_t* prefix, const float* pos, const float* scales, const float* rots,
                        const float* opac, const float* shs, float noise_scale, float* o_pos, float* o_scales,
                        float* o_rots, float* o_opac, float* o_shs);
/* replaces split_gaussians (optimizer.py:244-309) */
int gsb_split_gaussians(gsb_ctx* ctx, gsb_stream s, int32_t n, int32_t new_n, const int32_t* mask,
                        const int32_t* prefix, const float* pos, const float* scales, const float* rots,
                        const float* opac, const float* shs, int32_t n_split, float scale_factor, float* o_pos,
                        float* o_scales, float* o_rots, float* o_opac, float* o_shs);
/* replaces mark_split_originals_for_removal + invert_mask (train.py:547-576) */
int gsb_split_valid_mask(gsb_ctx* ctx, gsb_stream s, int32_t num_points, int32_t offset, const int32_t* split_mask,
                         int32_t* valid);
/* replaces prune_gaussians (optimizer.py:364-382) */
int gsb_prune_mask(gsb_ctx* ctx, gsb_stream s, int32_t n, const float* opac, float threshold, int32_t* valid);
/* replaces compact_gaussians (optimizer.py:384-415); writes past out_n (quirk G5) are dropped */
int gsb_compact_gaussians(gsb_ctx* ctx, gsb_stream s, int32_t n, int32_t out_n, const int32_t* valid,
                          const int32_t* prefix, const float* pos, const float* scales, const float* rots,
                          const float* opac, const float* shs, float* o_pos, float* o_scales, float* o_rots,
                          float* o_opac, float* o_shs);

/* ---- loss ("next" row 8f-1) ------------------------------------------------------------- */

/* replaces l1_loss_kernel + backprop_l1_pixel_gradients (loss.py:11-30, 121-146): one pass that
 * writes pixel_grad = l1_weight * sign(rendered - target) (sign(0) = +1) and reduces
 * sum |rendered - target| into *loss_sum (device double, zeroed here).  count = 3*W*H. */
int gsb_l1_loss_grad(gsb_ctx* ctx, gsb_stream s, int64_t count, const float* rendered, const float* target,
                     float l1_weight, float* pixel_grad, double* loss_sum);

/* replaces gaussian_kernel + ssim_kernel + ssim() (loss.py:33-119, 178-215; "next" row 8f-4 -- dead
 * code in the reference's step loop, train.py:967-974, and without a gradient there, loss.py:243):
 * *ssim_sum (device double, zeroed here) receives the sum over the pixels of the mean-over-channels
 * SSIM in an 11x11 window truncated at the border, with the reference's weights as written;
 * ssim = *ssim_sum / (W*H).  rendered / target: float[H][W][3]. */
int gsb_ssim(gsb_ctx* ctx, gsb_stream s, int32_t width, int32_t height, const float* rendered, const float* target,
             double* ssim_sum);

/* replaces depth_loss_kernel + depth_loss() (loss.py:248-306): *loss_sum (device double, zeroed here)
 * = sum |rendered_depth - target_depth| * depth_mask over count = W*H pixels; loss = sum / (W*H). */
int gsb_depth_loss(gsb_ctx* ctx, gsb_stream s, int64_t count, const float* rendered_depth, const float* target_depth,
                   const float* depth_mask, double* loss_sum);

#ifdef __cplusplus
}
#endif
#endif /* GSB200_H_ */
