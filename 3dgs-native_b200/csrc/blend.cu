// blend.cu -- the per-16x16-tile alpha blend and its back-to-front backward.
//   blend_forward_kernel  replaces wp_render_gaussians      (forward.py:384-515)
//   blend_backward_kernel replaces wp_render_backward_kernel (backward.py:558-706)
//
// Design (both kernels): one CTA per tile, 256/PPT threads, PPT pixels per thread.  Lanes run
// along image rows (x = lane & 15), so every per-pixel load/store is a coalesced 64-byte row
// segment; the reference's launch maps adjacent lanes to a pixel COLUMN.  The tile's depth-sorted
// Gaussians are staged through shared memory in batches (one gather per Gaussian per tile instead
// of one per pixel), then broadcast to the threads with three 16-byte LDS per Gaussian.
//
// Culling (both kernels).  The reference bins a Gaussian into every tile of the bounding SQUARE of
// its 3-sigma circle, but a pixel only ever uses it when alpha >= 1/255, i.e. inside the ellipse
// power >= -log(255*opacity).  While staging a Gaussian the loading thread intersects that ellipse
// (with a conservative margin) with each of the tile's 16 pixel rows and keeps a 16-bit row mask:
//   * Gaussians whose mask is empty are dropped from the batch by a stable compaction (on the
//     headline scene 31% of all list entries), their list position is kept for n_contrib;
//   * a warp skips a staged Gaussian with one LDS + one test unless the mask touches its rows
//     (only 38% of the (Gaussian, 2-row strip) combinations can contribute at all);
//   * lanes that remain are filtered by the same threshold on the exponent before the exponential.
// Pairs that survive run the exact arithmetic of the contract, so n_contrib / final_T are
// bit-identical to the oracle: culling only ever removes pairs the reference would `continue` on.
// Forward: a CTA stops as soon as every one of its pixels has terminated (__syncthreads_and).
//
// Backward: the reference issues 11 scalar global atomics per (pixel, Gaussian) pair.  Here each
// thread first sums its PPT pixels in registers, the warp then reduces the nine live gradient
// scalars with a transposed butterfly (9 + 5 shuffles instead of 45), and nine lanes issue one
// RED each per (warp, Gaussian) -- and only for Gaussians that touched the warp at all.
#include "blend_common.cuh"

namespace {

template <int PPT>
__global__ void __launch_bounds__(256 / PPT)
blend_forward_kernel(const BlendParams P, const int2* __restrict__ ranges, const int* __restrict__ point_list,
                     const float2* __restrict__ xy, const float* __restrict__ rgb,
                     const float4* __restrict__ conic_opacity, const float* __restrict__ depths,
                     float* __restrict__ image, float* __restrict__ inv_depth, float* __restrict__ final_T,
                     int* __restrict__ n_contrib) {
  constexpr int NT = 256 / PPT;
  constexpr int NW = NT / 32;
  __shared__ float4 s_a[NT];  // x, y, conic.a, conic.b
  __shared__ float4 s_b[NT];  // conic.c, opacity, power threshold, 1/depth
  __shared__ float4 s_c[NT];  // r, g, b, -
  __shared__ int2 s_meta[NT]; // 1-based position in the tile's list, row mask
  __shared__ int s_wcnt[NW];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile_x = blockIdx.x, tile_y = blockIdx.y;
  const int tile_id = tile_y * P.grid_x + tile_x;
  const int px = tile_x * kTile + (lane & 15);
  const int row0 = tile_y * kTile + warp * (2 * PPT) + (lane >> 4);
  const float pxf = (float)px;
  const unsigned my_rows = ((1u << (2 * PPT)) - 1u) << (warp * 2 * PPT);
  const float tile_x0 = (float)(tile_x * kTile), tile_y0 = (float)(tile_y * kTile);

  float pyf[PPT], T[PPT], C0[PPT], C1[PPT], C2[PPT], Dp[PPT];
  int last[PPT];
  bool done[PPT];
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    int py = row0 + 2 * k;
    pyf[k] = (float)py;
    done[k] = !(px < P.W && py < P.H);
    T[k] = 1.0f;
    C0[k] = C1[k] = C2[k] = Dp[k] = 0.0f;
    last[k] = 0;
  }

  const int2 range = ranges[tile_id];
  const int todo = range.y - range.x;
  for (int base = 0; base < todo; base += NT) {
    bool all_done = true;
#pragma unroll
    for (int k = 0; k < PPT; ++k) all_done = all_done && done[k];
    if (__syncthreads_and(all_done)) break;
    float4 ea, eb, ec;
    unsigned rowmask = 0u;
    if (base + tid < todo) {
      const int gid = point_list[range.x + base + tid];
      const float2 p = xy[gid];
      const float4 co = conic_opacity[gid];
      const float thr = gs_power_threshold(co.w);
      rowmask = P.cull ? gs_row_mask(p.x, p.y, co.x, co.y, co.z, thr, tile_x0, tile_y0) : 0xffffu;
      ea = make_float4(p.x, p.y, co.x, co.y);
      eb = make_float4(co.z, co.w, thr, 1.0f / depths[gid]);
      ec = make_float4(rgb[3 * gid + 0], rgb[3 * gid + 1], rgb[3 * gid + 2], 0.0f);
    }
    int cnt;
    const int slot = compact_slot<NW>(rowmask != 0u, lane, warp, s_wcnt, cnt);
    if (rowmask != 0u) {
      s_a[slot] = ea;
      s_b[slot] = eb;
      s_c[slot] = ec;
      s_meta[slot] = make_int2(base + tid + 1, (int)rowmask);
    }
    __syncthreads();
    for (int j = 0; j < cnt; ++j) {
      const int2 meta = s_meta[j];
      if (!((unsigned)meta.y & my_rows)) continue;  // warp-uniform: this Gaussian cannot touch our rows
      const float4 a = s_a[j];
      const float4 b = s_b[j];
      const float4 c = s_c[j];
      const float dx = a.x - pxf;
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        if (done[k]) continue;
        const float dy = a.y - pyf[k];
        const float power = gs_power(a.z, a.w, b.x, dx, dy);
        if (power > 0.0f) continue;        // forward.py:474
        if (power < b.z) continue;         // provably alpha < 1/255 (see gs_power_threshold)
        const float alpha = f_min(0.99f, b.y * gs_expf(power));
        if (alpha < (1.0f / 255.0f)) continue;
        const float test_T = T[k] * (1.0f - alpha);
        if (test_T < 0.0001f) {            // forward.py:487: the breaking Gaussian is not counted
          done[k] = true;
          continue;
        }
        C0[k] += c.x * alpha * T[k];
        C1[k] += c.y * alpha * T[k];
        C2[k] += c.z * alpha * T[k];
        Dp[k] += b.w * alpha * T[k];
        T[k] = test_T;
        last[k] = meta.x;
      }
    }
  }

#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    const int py = row0 + 2 * k;
    if (px < P.W && py < P.H) {
      const size_t pix = (size_t)py * P.W + px;
      final_T[pix] = T[k];
      n_contrib[pix] = last[k];
      image[3 * pix + 0] = C0[k] + T[k] * P.bg0;
      image[3 * pix + 1] = C1[k] + T[k] * P.bg1;
      image[3 * pix + 2] = C2[k] + T[k] * P.bg2;
      inv_depth[pix] = Dp[k];
    }
  }
}

template <int PPT>
int launch_fwd(gsb_ctx* ctx, cudaStream_t s, const BlendParams& P, dim3 grid, const int32_t* ranges,
               const int32_t* point_list, const float* xy, const float* rgb, const float* conic_opacity,
               const float* depths, float* image, float* inv_depth, float* final_T, int32_t* n_contrib) {
  GSB_LAUNCH(ctx, blend_forward_kernel<PPT>, grid, 256 / PPT, 0, s, P, reinterpret_cast<const int2*>(ranges), point_list,
             reinterpret_cast<const float2*>(xy), rgb, reinterpret_cast<const float4*>(conic_opacity), depths, image,
             inv_depth, final_T, n_contrib);
  return GSB_OK;
}

}  // namespace

int g_blend_fwd_ppt = 1;

GSB_API int gsb_blend_forward(gsb_ctx* ctx, gsb_stream s_, const gsb_frame* f, const int32_t* ranges,
                              const int32_t* point_list, const float* points_xy, const float* rgb,
                              const float* conic_opacity, const float* depths, float* image, float* inv_depth,
                              float* final_T, int32_t* n_contrib) {
  if (!ctx) return GSB_ERR_INVALID;
  GSB_REQUIRE(ctx, f && f->width > 0 && f->height > 0, "gsb_blend_forward: bad frame");
  GSB_REQUIRE(ctx, gsb_aligned16(conic_opacity), "gsb_blend_forward: conic_opacity must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)s_;
  BlendParams P = make_blend_params(f);
  dim3 grid(P.grid_x, (f->height + kTile - 1) / kTile);
  switch (g_blend_fwd_ppt) {
    case 2: return launch_fwd<2>(ctx, s, P, grid, ranges, point_list, points_xy, rgb, conic_opacity, depths, image, inv_depth, final_T, n_contrib);
    case 4: return launch_fwd<4>(ctx, s, P, grid, ranges, point_list, points_xy, rgb, conic_opacity, depths, image, inv_depth, final_T, n_contrib);
    case 8: return launch_fwd<8>(ctx, s, P, grid, ranges, point_list, points_xy, rgb, conic_opacity, depths, image, inv_depth, final_T, n_contrib);
    default: return launch_fwd<1>(ctx, s, P, grid, ranges, point_list, points_xy, rgb, conic_opacity, depths, image, inv_depth, final_T, n_contrib);
  }
}

