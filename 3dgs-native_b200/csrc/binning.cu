// binning.cu -- prefix sum, duplicate-with-keys and identify-tile-ranges.
//
//  * gsb_scan_i32      replaces wp_prefix_sum (utils/wp_utils.py:46-60: ONE thread, serial) and
//                      wp.utils.array_scan (train.py:432,...): a reduce / scan-of-sums / apply
//                      scan, 2048 elements per CTA, no inter-CTA spinning.
//  * duplicate kernel  replaces wp_duplicate_with_keys (forward.py:517-558).
//  * ranges kernel     replaces wp_identify_tile_ranges (forward.py:560-586).
#include "common.cuh"

namespace {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;  // 2048

__device__ __forceinline__ int warp_incl_scan(int v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// block-wide exclusive scan of one value per thread (256 threads); returns exclusive prefix and
// the block total through `total`
__device__ __forceinline__ int block_excl_scan_256(int v, int* s_warp /*[8]*/, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = warp_incl_scan(v);
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  int wbase = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w) {
    int t = s_warp[w];
    if (w < warp) wbase += t;
    tot += t;
  }
  __syncthreads();
  total = tot;
  return wbase + inc - v;
}

__device__ __forceinline__ void load_items(const int* __restrict__ in, int64_t n, int64_t base, int v[kScanItems]) {
  // each thread owns 8 consecutive elements (two 16-byte loads when the tile is full and aligned)
  int64_t p = base + (int64_t)threadIdx.x * kScanItems;
  if (p + kScanItems <= n && ((reinterpret_cast<uintptr_t>(in + p) & 15u) == 0)) {
    int4 a = __ldg(reinterpret_cast<const int4*>(in + p));
    int4 b = __ldg(reinterpret_cast<const int4*>(in + p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) v[k] = (p + k < n) ? in[p + k] : 0;
  }
}

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const int* __restrict__ in, int64_t n,
                                                                   int* __restrict__ sums) {
  __shared__ int s_warp[8];
  int v[kScanItems];
  load_items(in, n, (int64_t)blockIdx.x * kScanTile, v);
  int t = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) t += v[k];
  int total;
  block_excl_scan_256(t, s_warp, total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// single CTA: exclusive scan of the per-tile sums in place (num <= a few thousand)
__global__ void __launch_bounds__(kScanThreads) scan_sums_kernel(int* __restrict__ sums, int num) {
  __shared__ int s_warp[8];
  int carry = 0;
  for (int base = 0; base < num; base += kScanThreads) {
    int i = base + threadIdx.x;
    int v = (i < num) ? sums[i] : 0;
    int total;
    int ex = block_excl_scan_256(v, s_warp, total);
    if (i < num) sums[i] = carry + ex;
    carry += total;
  }
}

__global__ void __launch_bounds__(kScanThreads)
scan_apply_kernel(const int* __restrict__ in, int* __restrict__ out, int64_t n, const int* __restrict__ sums,
                  int exclusive, int* __restrict__ d_last) {
  __shared__ int s_warp[8];
  int v[kScanItems];
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  load_items(in, n, base, v);
  int t = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) t += v[k];
  int total;
  int run = sums[blockIdx.x] + block_excl_scan_256(t, s_warp, total);
  int64_t p = base + (int64_t)threadIdx.x * kScanItems;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int incl = run + v[k];
    if (p + k < n) {
      int o = exclusive ? run : incl;
      out[p + k] = o;
      if (d_last && p + k == n - 1) *d_last = o;
    }
    run = incl;
  }
}

// forward.py:517-558.  One thread per Gaussian; writes its (key, value) run at point_offsets[tid-1].
__global__ void __launch_bounds__(256)
duplicate_kernel(int n, const float2* __restrict__ xy, const float* __restrict__ depths,
                 const int* __restrict__ point_offsets, const int* __restrict__ radii, int grid_x, int grid_y,
                 int64_t capacity, int64_t* __restrict__ keys, int* __restrict__ vals) {
  int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n) return;
  int r = radii[tid];
  if (r <= 0) return;
  int64_t offset = (tid > 0) ? point_offsets[tid - 1] : 0;
  float2 p = xy[tid];
  int rminx, rminy, rmaxx, rmaxy;
  gs_get_rect(p.x, p.y, (float)r, (float)grid_x, (float)grid_y, rminx, rminy, rmaxx, rmaxy);
  const uint64_t depth_bits = (uint64_t)__float_as_uint(depths[tid]);
  for (int y = rminy; y < rmaxy; ++y)
    for (int x = rminx; x < rmaxx; ++x) {
      if (offset >= capacity) return;  // never hit when capacity >= num_rendered
      uint64_t tile_id = (uint64_t)(uint32_t)(y * grid_x + x);
      keys[offset] = (int64_t)((tile_id << 32) | depth_bits);
      vals[offset] = tid;
      offset += 1;
    }
}

// forward.py:560-586
__global__ void __launch_bounds__(256)
tile_ranges_kernel(int64_t num_rendered, const int64_t* __restrict__ keys, int2* __restrict__ ranges) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= num_rendered) return;
  int curr_tile = (int)(keys[idx] >> 32);
  if (idx == 0) {
    ranges[curr_tile].x = 0;
  } else {
    int prev_tile = (int)(keys[idx - 1] >> 32);
    if (curr_tile != prev_tile) {
      ranges[prev_tile].y = (int)idx;
      ranges[curr_tile].x = (int)idx;
    }
  }
  if (idx == num_rendered - 1) ranges[curr_tile].y = (int)num_rendered;
}

}  // namespace

int gsb_scan_i32(gsb_ctx* ctx, cudaStream_t s, int64_t n, const int32_t* in, int32_t* out, bool exclusive,
                 int32_t* d_last) {
  if (n <= 0) return GSB_OK;
  int64_t tiles = gsb_div_up(n, kScanTile);
  int rc = gsb_grow(ctx, (void**)&ctx->scan_sums, &ctx->scan_cap, tiles, sizeof(int32_t), s);
  if (rc != GSB_OK) return rc;
  GSB_LAUNCH(ctx, scan_reduce_kernel, (int)tiles, kScanThreads, 0, s, in, n, ctx->scan_sums);
  GSB_LAUNCH(ctx, scan_sums_kernel, 1, kScanThreads, 0, s, ctx->scan_sums, (int)tiles);
  GSB_LAUNCH(ctx, scan_apply_kernel, (int)tiles, kScanThreads, 0, s, in, out, n, ctx->scan_sums, exclusive ? 1 : 0,
             d_last);
  return GSB_OK;
}

static int read_scalar(gsb_ctx* ctx, cudaStream_t s, int slot, int32_t* host_value) {
  GSB_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars + slot, ctx->d_scalars + slot, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  GSB_CUDA(ctx, cudaStreamSynchronize(s));
  *host_value = ctx->h_scalars[slot];
  return GSB_OK;
}

GSB_API int gsb_scan_tiles(gsb_ctx* ctx, gsb_stream s_, int32_t n, const int32_t* tiles_touched, int32_t* point_offsets,
                           int64_t* num_rendered_host) {
  if (!ctx) return GSB_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)s_;
  if (n <= 0) {
    if (num_rendered_host) *num_rendered_host = 0;
    return GSB_OK;
  }
  int rc = gsb_scan_i32(ctx, s, n, tiles_touched, point_offsets, false, ctx->d_scalars + 0);
  if (rc != GSB_OK) return rc;
  if (num_rendered_host) {
    int32_t v = 0;
    rc = read_scalar(ctx, s, 0, &v);
    if (rc != GSB_OK) return rc;
    // the reference accumulates in int32 too (wp.array(dtype=int)); a wrapped (negative) total
    // means > 2^31 duplicates
    *num_rendered_host = (v < 0) ? (int64_t)(uint32_t)v : (int64_t)v;
  }
  return GSB_OK;
}

GSB_API int gsb_scan_mask(gsb_ctx* ctx, gsb_stream s_, int32_t n, const int32_t* mask, int32_t* prefix,
                          int32_t* last_host) {
  if (!ctx) return GSB_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)s_;
  if (n <= 0) {
    if (last_host) *last_host = 0;
    return GSB_OK;
  }
  int rc = gsb_scan_i32(ctx, s, n, mask, prefix, true, ctx->d_scalars + 1);
  if (rc != GSB_OK) return rc;
  if (last_host) return read_scalar(ctx, s, 1, last_host);
  return GSB_OK;
}

GSB_API int gsb_duplicate_with_keys(gsb_ctx* ctx, gsb_stream s, int32_t width, int32_t height, int32_t n,
                                    const float* points_xy, const float* depths, const int32_t* point_offsets,
                                    const int32_t* radii, int64_t num_rendered, int64_t* keys, int32_t* values) {
  if (!ctx) return GSB_ERR_INVALID;
  if (n <= 0 || num_rendered <= 0) return GSB_OK;
  GSB_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(points_xy) & 7u) == 0, "gsb_duplicate_with_keys: points_xy alignment");
  int gx = (width + kTile - 1) / kTile, gy = (height + kTile - 1) / kTile;
  GSB_LAUNCH(ctx, duplicate_kernel, (int)gsb_div_up(n, 256), 256, 0, (cudaStream_t)s, n,
             reinterpret_cast<const float2*>(points_xy), depths, point_offsets, radii, gx, gy, num_rendered, keys,
             values);
  return GSB_OK;
}

GSB_API int gsb_tile_ranges(gsb_ctx* ctx, gsb_stream s, int64_t num_rendered, const int64_t* sorted_keys,
                            int32_t num_tiles, int32_t* ranges) {
  if (!ctx) return GSB_ERR_INVALID;
  GSB_CUDA(ctx, cudaMemsetAsync(ranges, 0, sizeof(int32_t) * 2 * (size_t)num_tiles, (cudaStream_t)s));
  if (num_rendered <= 0) return GSB_OK;
  GSB_LAUNCH(ctx, tile_ranges_kernel, (int)gsb_div_up(num_rendered, 256), 256, 0, (cudaStream_t)s, num_rendered,
             sorted_keys, reinterpret_cast<int2*>(ranges));
  return GSB_OK;
}
