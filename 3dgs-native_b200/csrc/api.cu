// api.cu -- context management and the two whole-operator entry points:
//   gsb_forward  = render_gaussians (reference forward.py:629-894)
//   gsb_backward = backward         (reference backward.py:955-1196)
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

int gsb_radix_sort_pingpong(gsb_ctx* ctx, cudaStream_t s, int64_t* k0, int32_t* v0, int64_t* k1, int32_t* v1,
                            int32_t* final_vals, int64_t n, int begin_bit, int end_bit, bool* result_in_second);
int gsb_tile_binning_count(gsb_ctx* ctx, cudaStream_t s, int n, int width, int height, const float* points_xy,
                           const int32_t* radii, int32_t* ranges, int64_t* num_rendered_host, int* max_count_host);
int gsb_tile_binning_prepare(gsb_ctx* ctx, cudaStream_t s, int n, int num_tiles);
int gsb_tile_binning_scan_async(gsb_ctx* ctx, cudaStream_t s, int num_tiles, int32_t* ranges, int spec_cap, int spec_max);
int gsb_blend_forward_impl(gsb_ctx* ctx, gsb_stream s_, const gsb_frame* f, const int32_t* ranges,
                           const int32_t* point_list, const float* points_xy, const float* rgb,
                           const float* conic_opacity, const float* depths, float* image, float* inv_depth,
                           float* final_T, int32_t* n_contrib, int32_t* block_masks, const int* go);
int gsb_tile_binning_wait(gsb_ctx* ctx, int64_t* num_rendered_host, int* max_count_host);
int gsb_tile_binning_sort(gsb_ctx* ctx, cudaStream_t s, int n, int width, int height, const float* points_xy,
                          const float* depths, const int32_t* radii, const int32_t* ranges, int64_t num_rendered,
                          int max_count, int32_t* point_list, const int* go);
int gsb_tile_binning_max();

int gsb_set_error(gsb_ctx* ctx, int code, const char* fmt, ...) {
  if (ctx) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
    va_end(ap);
  }
  return code;
}

int gsb_check_cuda(gsb_ctx* ctx, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return GSB_OK;
  return gsb_set_error(ctx, GSB_ERR_CUDA, "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
}

// grow-only device buffer; contents are NOT preserved
int gsb_grow(gsb_ctx* ctx, void** ptr, int64_t* cap, int64_t need_elems, size_t elem_size, cudaStream_t s) {
  if (need_elems <= *cap) return GSB_OK;
  int64_t new_cap = need_elems + need_elems / 2 + 1024;   // 50% to spare: every growth is a cudaFree (device-wide wait) + cudaMalloc
  if (*ptr) {
    GSB_CUDA(ctx, cudaStreamSynchronize(s));
    GSB_CUDA(ctx, cudaFree(*ptr));
    *ptr = nullptr;
    *cap = 0;
  }
  cudaError_t e = cudaMalloc(ptr, (size_t)new_cap * elem_size);
  if (e != cudaSuccess) {
    *ptr = nullptr;
    cudaGetLastError();
    return gsb_set_error(ctx, GSB_ERR_NOMEM, "cudaMalloc of %lld bytes failed: %s", (long long)(new_cap * (int64_t)elem_size),
                         cudaGetErrorString(e));
  }
  *cap = new_cap;
  return GSB_OK;
}

int gsb_reserve_binning(gsb_ctx* ctx, cudaStream_t s, int64_t num_rendered) {
  if (num_rendered <= ctx->bin_cap) return GSB_OK;
  int64_t c0 = ctx->bin_cap, c1 = ctx->bin_cap, c2 = ctx->bin_cap, c3 = ctx->bin_cap;
  int rc;
  if ((rc = gsb_grow(ctx, (void**)&ctx->keys_a, &c0, num_rendered, sizeof(int64_t), s)) != GSB_OK) return rc;
  if ((rc = gsb_grow(ctx, (void**)&ctx->keys_b, &c1, num_rendered, sizeof(int64_t), s)) != GSB_OK) return rc;
  if ((rc = gsb_grow(ctx, (void**)&ctx->vals_a, &c2, num_rendered, sizeof(int32_t), s)) != GSB_OK) return rc;
  if ((rc = gsb_grow(ctx, (void**)&ctx->vals_b, &c3, num_rendered, sizeof(int32_t), s)) != GSB_OK) return rc;
  ctx->bin_cap = c0 < c1 ? c0 : c1;
  ctx->bin_cap = ctx->bin_cap < c2 ? ctx->bin_cap : c2;
  ctx->bin_cap = ctx->bin_cap < c3 ? ctx->bin_cap : c3;
  return GSB_OK;
}

static int reserve_per_gaussian(gsb_ctx* ctx, cudaStream_t s, int64_t n) {
  if (n <= ctx->n_cap) return GSB_OK;
  int64_t c0 = ctx->n_cap, c1 = ctx->n_cap, c2 = ctx->n_cap, c3 = ctx->n_cap * 12;
  int rc;
  if ((rc = gsb_grow(ctx, (void**)&ctx->bwd_acc, &c3, n * 12, sizeof(float), s)) != GSB_OK) return rc;
  if ((rc = gsb_grow(ctx, (void**)&ctx->tiles_touched, &c0, n, sizeof(int32_t), s)) != GSB_OK) return rc;
  if ((rc = gsb_grow(ctx, (void**)&ctx->dcov3d, &c1, n * 6, sizeof(float), s)) != GSB_OK) return rc;
  if ((rc = gsb_grow(ctx, (void**)&ctx->rank_base, &c2, n, sizeof(int32_t), s)) != GSB_OK) return rc;
  ctx->n_cap = c0 < c1 / 6 ? c0 : c1 / 6;
  ctx->n_cap = ctx->n_cap < c2 ? ctx->n_cap : c2;
  ctx->n_cap = ctx->n_cap < c3 / 12 ? ctx->n_cap : c3 / 12;
  return GSB_OK;
}

GSB_API int gsb_version(void) { return GSB_VERSION; }

GSB_API int gsb_create(gsb_ctx** out, int device) {
  if (!out) return GSB_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
    cudaGetLastError();
    return GSB_ERR_CUDA;  // no CUDA device: there is no CPU fallback
  }
  if (device < 0 || device >= count) return GSB_ERR_INVALID;
  gsb_ctx* ctx = new gsb_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess) {
    delete ctx;
    return GSB_ERR_CUDA;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
    ctx->num_sms = prop.multiProcessorCount;
    ctx->coop_launch = prop.cooperativeLaunch != 0;
  }
  if (cudaMalloc((void**)&ctx->d_scalars, 16 * sizeof(int32_t)) != cudaSuccess ||
      cudaMallocHost((void**)&ctx->h_scalars, 16 * sizeof(int32_t)) != cudaSuccess ||
      cudaMalloc((void**)&ctx->sort_small, (256 + 16) * sizeof(uint32_t)) != cudaSuccess ||
      cudaMalloc((void**)&ctx->d_accum, 2 * sizeof(double)) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_count, cudaEventDisableTiming) != cudaSuccess) {
    delete ctx;
    return GSB_ERR_NOMEM;
  }
  cudaMemset(ctx->d_scalars, 0, 16 * sizeof(int32_t));
  cudaMemset(ctx->sort_small, 0, (256 + 16) * sizeof(uint32_t));
  cudaMemset(ctx->d_accum, 0, 2 * sizeof(double));
  memset(ctx->h_scalars, 0, 16 * sizeof(int32_t));
  *out = ctx;
  return GSB_OK;
}

GSB_API int gsb_destroy(gsb_ctx* ctx) {
  if (!ctx) return GSB_OK;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  void* bufs[] = {ctx->keys_a, ctx->keys_b, ctx->vals_a, ctx->vals_b, ctx->sort_table, ctx->sort_small, ctx->scan_sums,
                  ctx->tiles_touched, ctx->dcov3d, ctx->d_scalars, ctx->tile_count, ctx->d_accum};
  for (void* p : bufs)
    if (p) cudaFree(p);
  if (ctx->h_scalars) cudaFreeHost(ctx->h_scalars);
  if (ctx->rank_base) cudaFree(ctx->rank_base);
  if (ctx->sort_coop_state) cudaFree(ctx->sort_coop_state);
  if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->bwd_acc) cudaFree(ctx->bwd_acc);
  if (ctx->bwd_acc_stage) cudaFree(ctx->bwd_acc_stage);
  if (ctx->ev_count) cudaEventDestroy(ctx->ev_count);
  delete ctx;
  return GSB_OK;
}

GSB_API const char* gsb_last_error_string(gsb_ctx* ctx) { return ctx ? ctx->err : "null context"; }
GSB_API int64_t gsb_launch_count(gsb_ctx* ctx) { return ctx ? ctx->launches : 0; }

GSB_API int gsb_reserve(gsb_ctx* ctx, gsb_stream s, int64_t num_rendered) {
  if (!ctx) return GSB_ERR_INVALID;
  return gsb_reserve_binning(ctx, (cudaStream_t)s, num_rendered);
}

// See gsb200.h.  cuda_event: a cudaEvent_t (torch.cuda.Event.cuda_event), or NULL to clear.
GSB_API int gsb_set_color_dependency(gsb_ctx* ctx, void* cuda_event) {
  if (!ctx) return GSB_ERR_INVALID;
  ctx->color_event = cuda_event;
  return GSB_OK;
}

// A/B knobs (not part of the reference surface; results never depend on them)
GSB_API int gsb_set_option(gsb_ctx* ctx, const char* name, int value) {
  if (!name || !ctx) return GSB_ERR_INVALID;
  if (!strcmp(name, "binning") && (value == 0 || value == 1)) {
    ctx->opt.binning = value;
    return GSB_OK;
  }
  if (!strcmp(name, "blend_cull") && (value == 0 || value == 1)) {
    ctx->opt.blend_cull = value;
    return GSB_OK;
  }
  if (!strcmp(name, "tile_sort") && value >= 0 && value <= 3) {
    ctx->opt.tile_sort = value;
    return GSB_OK;
  }
  if (!strcmp(name, "pdl") && (value == 0 || value == 1)) {
    ctx->opt.pdl = value;
    return GSB_OK;
  }
  if (!strcmp(name, "speculate") && (value == 0 || value == 1)) {
    ctx->opt.speculate = value;
    return GSB_OK;
  }
  if (!strcmp(name, "sort_coop") && (value == 0 || value == 1)) {
    ctx->opt.sort_coop = value;
    return GSB_OK;
  }
  if (!strcmp(name, "bwd_packed") && (value == 0 || value == 1)) {
    ctx->opt.bwd_packed = value;
    return GSB_OK;
  }
  if (!strcmp(name, "bwd_reduce") && value >= 0 && value <= 2) {
    ctx->opt.bwd_reduce = value;
    return GSB_OK;
  }
  return gsb_set_error(ctx, GSB_ERR_INVALID, "unknown option %s=%d", name, value);
}

static int key_bits_for(int num_tiles) {
  int tb = 0;
  while ((1 << tb) < num_tiles) ++tb;
  return 32 + (tb < 1 ? 1 : tb);
}

GSB_API int gsb_bin_by_tile(gsb_ctx* ctx, gsb_stream s_, int32_t width, int32_t height, int32_t n,
                            const float* points_xy, const float* depths, const int32_t* radii,
                            const int32_t* point_offsets, int32_t* point_list, int64_t point_list_capacity,
                            int32_t* ranges, int64_t* num_rendered_host, int32_t* used_tile_path_host) {
  if (!ctx) return GSB_ERR_INVALID;
  GSB_REQUIRE(ctx, n >= 0 && width > 0 && height > 0, "gsb_bin_by_tile: bad size");
  cudaStream_t s = (cudaStream_t)s_;
  const int num_tiles = ((width + kTile - 1) / kTile) * ((height + kTile - 1) / kTile);
  int rc;
  int64_t D = 0;
  int max_count = 0;
  rc = gsb_tile_binning_count(ctx, s, n, width, height, points_xy, radii, ranges, &D, &max_count);
  if (rc != GSB_OK) return rc;
  if (num_rendered_host) *num_rendered_host = D;
  if (used_tile_path_host) *used_tile_path_host = 0;
  if (D > (1LL << 30))  // forward.py:765-767
    return gsb_set_error(ctx, GSB_ERR_TOO_MANY, "Number of rendered points exceeds the maximum supported by Warp.");
  if (D > GSB_MAX_RENDERED)
    return gsb_set_error(ctx, GSB_ERR_TOO_MANY, "num_rendered == 2^30 is not supported by the radix sort (max 2^30-1)");
  if (D > point_list_capacity)
    return gsb_set_error(ctx, GSB_ERR_CAPACITY, "point_list capacity %lld < num_rendered %lld",
                         (long long)point_list_capacity, (long long)D);

  if (D == 0) return GSB_OK;  // ranges are all (0,0) already
  if (ctx->opt.binning == 0 && max_count <= gsb_tile_binning_max()) {
    // duplicate + sort + ranges (forward.py:776-840) as counting sort by tile + per-tile sort
    if ((rc = gsb_reserve_binning(ctx, s, D)) != GSB_OK) return rc;
    rc = gsb_tile_binning_sort(ctx, s, n, width, height, points_xy, depths, radii, ranges, D, max_count, point_list,
                               nullptr);
    if (rc != GSB_OK) return rc;
    if (used_tile_path_host) *used_tile_path_host = 1;
  } else {
    if ((rc = gsb_reserve_binning(ctx, s, D)) != GSB_OK) return rc;
    // forward.py:776-788
    rc = gsb_duplicate_with_keys(ctx, s_, width, height, n, points_xy, depths, point_offsets, radii, D, ctx->keys_a,
                                 ctx->vals_a);
    if (rc != GSB_OK) return rc;
    // forward.py:791-824; the final pass writes the sorted values straight into point_list
    bool in_b = false;
    rc = gsb_radix_sort_pingpong(ctx, s, ctx->keys_a, ctx->vals_a, ctx->keys_b, ctx->vals_b, point_list, D, 0,
                                 key_bits_for(num_tiles), &in_b);
    if (rc != GSB_OK) return rc;
    const int64_t* sorted_keys = in_b ? ctx->keys_b : ctx->keys_a;
    // forward.py:832-840
    if ((rc = gsb_tile_ranges(ctx, s_, D, sorted_keys, num_tiles, ranges)) != GSB_OK) return rc;
  }
  return GSB_OK;
}

GSB_API int gsb_forward(gsb_ctx* ctx, gsb_stream s_, const gsb_frame* f, int32_t n, const float* means,
                        const float* scales, const float* rotations, const float* opacities, const float* shs,
                        int32_t* radii, int32_t* point_offsets, float* points_xy, float* depths, float* rgb,
                        float* cov3Ds, float* conic_opacity, float* clamped_state, int32_t* point_list,
                        int64_t point_list_capacity, int32_t* ranges, float* image, float* inv_depth, float* final_T,
                        int32_t* n_contrib, int64_t* num_rendered_host, int32_t* block_masks) {
  if (!ctx) return GSB_ERR_INVALID;
  GSB_REQUIRE(ctx, f && n >= 0 && f->width > 0 && f->height > 0, "gsb_forward: bad frame or n");
  cudaStream_t s = (cudaStream_t)s_;
  const int gx = (f->width + kTile - 1) / kTile, gy = (f->height + kTile - 1) / kTile;
  const int num_tiles = gx * gy;
  const size_t pixels = (size_t)f->width * f->height;
  int rc;
  if ((rc = reserve_per_gaussian(ctx, s, n)) != GSB_OK) return rc;

  // forward.py:719-788.  Preprocess also runs the counting pass of the tile binning; the per-tile
  // scan and the read-back of D (the frame's one host wait) follow at once, and the inclusive scan of
  // tiles_touched -- an output of the operator, not an input of this binning -- is queued BEHIND the
  // read-back so that it runs while the host wakes up and launches the rest.
  int64_t D = 0;
  int max_count = 0;
  // The rest of the frame -- scatter, per-tile sort, blend -- is queued SPECULATIVELY behind the scan, before the host
  // knows D: launch sizes do not depend on D, only the buffer capacity and the sort kernel's capacity class do, and
  // those are taken from the previous frame (with head-room).  tile_scan_kernel checks them on the device and leaves
  // a go-ahead flag every speculative CTA reads first.  The host then waits for D while the GPU is already blending
  // (the wait used to leave the GPU idle for 9-10 us per frame: scan -> read-back -> wake-up -> launch); when the
  // frame does not fit (first frame, the scene grew, a tile list beyond the class) the speculative kernels have done
  // nothing and the frame continues below as before.
  // One-shot colour dependency (gsb_set_color_dependency): the SH coefficients are still being written on another
  // stream -- the second phase of the multi-GPU exchange, which runs beside this frame's geometry preprocess and
  // binning.  preprocess then leaves rgb / clamped_state to sh_color_kernel, which is queued behind a wait for the event,
  // right in front of the blend.  An error return before that point hands the dependency back to the context.
  struct ColorDep {
    gsb_ctx* ctx;
    void* ev;
    ~ColorDep() {
      if (ev) ctx->color_event = ev;
    }
  } color_dep{ctx, ctx->color_event};
  ctx->color_event = nullptr;
  const bool defer_color = color_dep.ev != nullptr;
  auto finish_color = [&]() -> int {
    if (!color_dep.ev) return GSB_OK;
    GSB_CUDA(ctx, cudaStreamWaitEvent(s, (cudaEvent_t)color_dep.ev, 0));
    color_dep.ev = nullptr;
    return gsb_sh_color_impl(ctx, s, f, n, means, shs, radii, rgb, clamped_state);
  };
  bool spec = false;
  int spec_cap = 0, spec_max = 0;
  if (ctx->opt.speculate && ctx->opt.binning == 0 && ctx->opt.tile_sort == 3 && ctx->last_num_rendered > 0 && n > 0) {
    const int64_t cap = ctx->bin_cap < point_list_capacity ? ctx->bin_cap : point_list_capacity;
    spec_cap = (int)(cap < GSB_MAX_RENDERED ? cap : GSB_MAX_RENDERED);
    const int want = ctx->last_max_count + ctx->last_max_count / 8;
    spec_max = want <= 1024 ? 1024 : want <= 2048 ? 2048 : 4096;
    spec = spec_cap > 0 && ctx->last_max_count <= 4096;
  }
  {
    if ((rc = gsb_tile_binning_prepare(ctx, s, n, num_tiles)) != GSB_OK) return rc;
    PreBin bin{ctx->tile_count, defer_color ? 1 : 0};
    rc = gsb_preprocess_impl(ctx, s, f, n, means, scales, rotations, opacities, shs, radii, points_xy, depths, cov3Ds,
                             rgb, conic_opacity, ctx->tiles_touched, clamped_state, &bin);
    if (rc != GSB_OK) return rc;
    bool offsets_on_side = false;
    if (spec && point_offsets) {
      // forward.py:755-764, the inclusive scan of tiles_touched -- an OUTPUT of the operator, not an input of this
      // binning: three small latency-bound kernels.  They used to run in the shadow of the host's wait; now that the
      // wait no longer idles the GPU they go to a side stream, forked here behind preprocess, and run beside the
      // scan / scatter / sort / blend; the caller's stream joins them at the end of the queue.
      if (!ctx->side_stream) {
        GSB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
        GSB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        GSB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
      }
      GSB_CUDA(ctx, cudaEventRecord(ctx->ev_fork, s));
      GSB_CUDA(ctx, cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
      if ((rc = gsb_scan_tiles(ctx, (gsb_stream)ctx->side_stream, n, ctx->tiles_touched, point_offsets, nullptr)) != GSB_OK)
        return rc;
      GSB_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->side_stream));
      offsets_on_side = true;
    }
    if ((rc = gsb_tile_binning_scan_async(ctx, s, num_tiles, ranges, spec ? spec_cap : 0, spec ? spec_max : 0)) != GSB_OK)
      return rc;
    if (spec) {
      const int* go = ctx->d_scalars + 6;
      // (num_rendered only picks the scatter kernel's lanes per Gaussian, max_count the sort kernel's class)
      rc = gsb_tile_binning_sort(ctx, s, n, f->width, f->height, points_xy, depths, radii, ranges, ctx->last_num_rendered,
                                 spec_max, point_list, go);
      if (rc != GSB_OK) return rc;
      if ((rc = finish_color()) != GSB_OK) return rc;
      rc = gsb_blend_forward_impl(ctx, s_, f, ranges, point_list, points_xy, rgb, conic_opacity, depths, image, inv_depth,
                                  final_T, n_contrib, block_masks, go);
      if (rc != GSB_OK) return rc;
      if (offsets_on_side) GSB_CUDA(ctx, cudaStreamWaitEvent(s, ctx->ev_join, 0));
    }
    // forward.py:755-764: inclusive scan
    // (point_offsets == NULL: the caller does not want this output; the scan is then left out of the frame)
    if (point_offsets && !offsets_on_side &&
        (rc = gsb_scan_tiles(ctx, s_, n, ctx->tiles_touched, point_offsets, nullptr)) != GSB_OK)
      return rc;
    if ((rc = gsb_tile_binning_wait(ctx, &D, &max_count)) != GSB_OK) return rc;
  }
  ctx->last_num_rendered = D <= GSB_MAX_RENDERED ? D : 0;
  ctx->last_max_count = max_count;
  if (spec) {
    if (D > 0 && D <= spec_cap && max_count <= spec_max) {   // the same test tile_scan_kernel made: the frame is queued
      if (num_rendered_host) *num_rendered_host = D;
      return GSB_OK;
    }
    ctx->tile_clean = 0;   // the speculative scatter did not run: the counters still hold this frame's counts
  }
  if (num_rendered_host) *num_rendered_host = D;
  if (D > (1LL << 30))  // forward.py:765-767
    return gsb_set_error(ctx, GSB_ERR_TOO_MANY, "Number of rendered points exceeds the maximum supported by Warp.");
  if (D > GSB_MAX_RENDERED)
    return gsb_set_error(ctx, GSB_ERR_TOO_MANY, "num_rendered == 2^30 is not supported by the radix sort (max 2^30-1)");
  if (D > point_list_capacity)
    return gsb_set_error(ctx, GSB_ERR_CAPACITY, "point_list capacity %lld < num_rendered %lld",
                         (long long)point_list_capacity, (long long)D);
  if (D > 0) {
    if (ctx->opt.binning == 0 && max_count <= gsb_tile_binning_max()) {
      // duplicate + sort + ranges (forward.py:776-840) as counting sort by tile + per-tile sort
      // (first frame / the scene grew: the segment buffer grows here, behind the read-back of D)
      if ((rc = gsb_reserve_binning(ctx, s, D)) != GSB_OK) return rc;
      rc = gsb_tile_binning_sort(ctx, s, n, f->width, f->height, points_xy, depths, radii, ranges, D, max_count,
                                 point_list, nullptr);
      if (rc != GSB_OK) return rc;
    } else {
      // a tile list too long for the shared-memory sort (or the A/B switch): the reference's own
      // sequence on the global radix sort; it rewrites ranges.  It needs the inclusive scan: into the caller's
      // array, or (caller passed NULL) into the rank index of the abandoned counting pass
      int64_t D2 = 0;
      int32_t* offs = point_offsets;
      if (!offs) {
        offs = ctx->rank_base;
        if ((rc = gsb_scan_tiles(ctx, s_, n, ctx->tiles_touched, offs, nullptr)) != GSB_OK) return rc;
      }
      rc = gsb_bin_by_tile(ctx, s_, f->width, f->height, n, points_xy, depths, radii, offs, point_list,
                           point_list_capacity, ranges, &D2, nullptr);
      if (rc != GSB_OK) return rc;
    }
  }
  if ((rc = finish_color()) != GSB_OK) return rc;   // (no-op when the speculative queue above has already done it)
  if (D == 0) {
    // forward.py:830: nothing is launched; every image-shaped output keeps its wp.zeros() state
    // (ranges were already written as all (0,0) by the counting pass)
    GSB_CUDA(ctx, cudaMemsetAsync(image, 0, sizeof(float) * 3 * pixels, s));
    GSB_CUDA(ctx, cudaMemsetAsync(inv_depth, 0, sizeof(float) * pixels, s));
    GSB_CUDA(ctx, cudaMemsetAsync(final_T, 0, sizeof(float) * pixels, s));
    GSB_CUDA(ctx, cudaMemsetAsync(n_contrib, 0, sizeof(int32_t) * pixels, s));
    return GSB_OK;
  }
  // forward.py:844-863 (+ the no-op track_pixel_stats of 867-879)
  return gsb_blend_forward(ctx, s_, f, ranges, point_list, points_xy, rgb, conic_opacity, depths, image, inv_depth,
                           final_T, n_contrib, block_masks);
}

static int backward_impl(gsb_ctx* ctx, gsb_stream s_, const gsb_frame* f, int32_t n, const float* means,
                         const float* opacities, const float* shs, const float* scales, const float* rotations,
                         const int32_t* radii, const float* points_xy, const float* conic_opacity, const float* rgb,
                         const float* clamped_state, const float* cov3Ds, const int32_t* point_list,
                         const int32_t* ranges, const float* final_T, const int32_t* n_contrib,
                         const float* dL_dpixels, float* dL_dmean3D, float* dL_dcolor, float* dL_dshs,
                         float* dL_dopacity, float* dL_dscale, float* dL_drot, float* dL_dmean2D, float* dL_dconic,
                         float* dL_dcov3D, const int32_t* block_masks, int sh_compact) {
  (void)opacities;  // converted and unused by the reference as well (backward.py:1056)
  if (!ctx) return GSB_ERR_INVALID;
  GSB_REQUIRE(ctx, f && n >= 0, "gsb_backward: bad frame or n");
  cudaStream_t s = (cudaStream_t)s_;
  if (n == 0) return GSB_OK;
  int rc;
  if ((rc = reserve_per_gaussian(ctx, s, n)) != GSB_OK) return rc;
  // backward.py:1135-1152.  Default: the tile kernel accumulates into packed per-Gaussian records (vector REDs)
  // and the per-Gaussian pass writes dL_dmean2D / dL_dconic / dL_dcolor / dL_dopacity out in the reference's layouts.
  const bool packed = gsb_blend_backward_uses_packed(ctx);
  if (packed)
    rc = gsb_blend_backward_packed(ctx, s, f, n, ranges, point_list, points_xy, conic_opacity, rgb, final_T, n_contrib,
                                   dL_dpixels, ctx->bwd_acc, block_masks);
  else
    rc = gsb_blend_backward(ctx, s_, f, n, ranges, point_list, points_xy, conic_opacity, rgb, final_T, n_contrib,
                            dL_dpixels, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, block_masks);
  if (rc != GSB_OK) return rc;
  // backward.py:1155-1182
  rc = gsb_preprocess_backward_impl(ctx, s, f, n, means, radii, shs, scales, rotations, cov3Ds, clamped_state,
                                    dL_dmean2D, dL_dconic, dL_dcolor, dL_dmean3D, dL_dshs, dL_dscale, dL_drot, nullptr,
                                    sh_compact, packed ? ctx->bwd_acc : nullptr, dL_dopacity,
                                    dL_dcov3D /* backward.py:1119,1195: returned as zeros; written by the same kernel
                                                 (a memset was a stream operation between this kernel and Adam) */);
  return rc;
}

GSB_API int gsb_backward(gsb_ctx* ctx, gsb_stream s_, const gsb_frame* f, int32_t n, const float* means,
                         const float* opacities, const float* shs, const float* scales, const float* rotations,
                         const int32_t* radii, const float* points_xy, const float* conic_opacity, const float* rgb,
                         const float* clamped_state, const float* cov3Ds, const int32_t* point_list,
                         const int32_t* ranges, const float* final_T, const int32_t* n_contrib,
                         const float* dL_dpixels, float* dL_dmean3D, float* dL_dcolor, float* dL_dshs,
                         float* dL_dopacity, float* dL_dscale, float* dL_drot, float* dL_dmean2D, float* dL_dconic,
                         float* dL_dcov3D, const int32_t* block_masks) {
  return backward_impl(ctx, s_, f, n, means, opacities, shs, scales, rotations, radii, points_xy, conic_opacity, rgb,
                       clamped_state, cov3Ds, point_list, ranges, final_T, n_contrib, dL_dpixels, dL_dmean3D, dL_dcolor,
                       dL_dshs, dL_dopacity, dL_dscale, dL_drot, dL_dmean2D, dL_dconic, dL_dcov3D, block_masks, 0);
}

// gsb_backward whose dL_dshs receives the COMPACT SH gradient: [8 n] floats, per Gaussian (dL_dRGB masked, unit view
// direction, 0, 0) -- the two factors the reference's 48 values are the outer product of (dL_dshs[16 i + k] =
// basis_k(direction) * dL_dRGB, backward.py:127-213).  Consumed by gsb_adam_step_peers_compact.
GSB_API int gsb_backward_compact_sh(gsb_ctx* ctx, gsb_stream s_, const gsb_frame* f, int32_t n, const float* means,
                         const float* opacities, const float* shs, const float* scales, const float* rotations,
                         const int32_t* radii, const float* points_xy, const float* conic_opacity, const float* rgb,
                         const float* clamped_state, const float* cov3Ds, const int32_t* point_list,
                         const int32_t* ranges, const float* final_T, const int32_t* n_contrib,
                         const float* dL_dpixels, float* dL_dmean3D, float* dL_dcolor, float* dL_dshs,
                         float* dL_dopacity, float* dL_dscale, float* dL_drot, float* dL_dmean2D, float* dL_dconic,
                         float* dL_dcov3D, const int32_t* block_masks) {
  return backward_impl(ctx, s_, f, n, means, opacities, shs, scales, rotations, radii, points_xy, conic_opacity, rgb,
                       clamped_state, cov3Ds, point_list, ranges, final_T, n_contrib, dL_dpixels, dL_dmean3D, dL_dcolor,
                       dL_dshs, dL_dopacity, dL_dscale, dL_drot, dL_dmean2D, dL_dconic, dL_dcov3D, block_masks, 1);
}
