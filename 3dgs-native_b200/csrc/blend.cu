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
// Forward: a per-Gaussian conservative threshold on the exponent (computed once at staging time)
// rejects most (pixel, Gaussian) pairs before the exponential; pairs that survive run the exact
// arithmetic of the contract, so n_contrib / final_T are bit-identical to the oracle.  A CTA stops
// as soon as every one of its pixels has terminated (__syncthreads_and).
//
// Backward: the reference issues 11 scalar global atomics per (pixel, Gaussian) pair.  Here each
// thread first sums its PPT pixels in registers, the warp then reduces the nine live gradient
// scalars with a transposed butterfly (9 + 5 shuffles instead of 45), and nine lanes issue one
// RED each per (warp, Gaussian) -- and only for Gaussians that touched the warp at all.
#include "common.cuh"

namespace {

struct BlendParams {
  int W, H, grid_x;
  float bg0, bg1, bg2;
};

template <int PPT>
__global__ void __launch_bounds__(256 / PPT)
blend_forward_kernel(const BlendParams P, const int2* __restrict__ ranges, const int* __restrict__ point_list,
                     const float2* __restrict__ xy, const float* __restrict__ rgb,
                     const float4* __restrict__ conic_opacity, const float* __restrict__ depths,
                     float* __restrict__ image, float* __restrict__ inv_depth, float* __restrict__ final_T,
                     int* __restrict__ n_contrib) {
  constexpr int NT = 256 / PPT;
  __shared__ float4 s_a[NT];  // x, y, conic.a, conic.b
  __shared__ float4 s_b[NT];  // conic.c, opacity, power threshold, 1/depth
  __shared__ float4 s_c[NT];  // r, g, b, -

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile_x = blockIdx.x, tile_y = blockIdx.y;
  const int tile_id = tile_y * P.grid_x + tile_x;
  const int px = tile_x * kTile + (lane & 15);
  const int row0 = tile_y * kTile + warp * (2 * PPT) + (lane >> 4);
  const float pxf = (float)px;

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
    if (base + tid < todo) {
      const int gid = point_list[range.x + base + tid];
      const float2 p = xy[gid];
      const float4 co = conic_opacity[gid];
      s_a[tid] = make_float4(p.x, p.y, co.x, co.y);
      s_b[tid] = make_float4(co.z, co.w, gs_power_threshold(co.w), 1.0f / depths[gid]);
      s_c[tid] = make_float4(rgb[3 * gid + 0], rgb[3 * gid + 1], rgb[3 * gid + 2], 0.0f);
    }
    __syncthreads();
    const int cnt = min(NT, todo - base);
    for (int j = 0; j < cnt; ++j) {
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
        last[k] = base + j + 1;
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

// Sum v[0..8] over the 32 lanes and add the totals into the four gradient arrays of Gaussian gid.
__device__ __forceinline__ void warp_reduce9_red(float v[9], int lane, int gid, float* __restrict__ dL_dmean2D,
                                                 float* __restrict__ dL_dconic, float* __restrict__ dL_dopacity,
                                                 float* __restrict__ dL_dcolor) {
  const unsigned full = 0xffffffffu;
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float send = h16 ? v[i] : v[i + 4];
    float keep = h16 ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(full, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float send = h8 ? v[i] : v[i + 2];
    float keep = h8 ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(full, send, 8);
  }
  float send = h4 ? v[0] : v[1];
  float keep = h4 ? v[1] : v[0];
  float s = keep + __shfl_xor_sync(full, send, 4);
  s += __shfl_xor_sync(full, s, 2);
  s += __shfl_xor_sync(full, s, 1);
  float o = v[8];
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) o += __shfl_xor_sync(full, o, m);

  // lanes 0,4,...,28 hold value index 4*h16 + 2*h8 + h4; lane 1 takes the opacity
  const int idx = (h16 ? 4 : 0) + (h8 ? 2 : 0) + (h4 ? 1 : 0);
  if ((lane & 3) == 0) {
    float* dst;
    if (idx < 3) dst = dL_dcolor + 3 * (size_t)gid + idx;              // r, g, b
    else if (idx < 5) dst = dL_dmean2D + 3 * (size_t)gid + (idx - 3);  // x, y
    else dst = dL_dconic + 4 * (size_t)gid + (idx == 7 ? 3 : idx - 5); // a, b, (0), c
    atomicAdd(dst, s);
  } else if (lane == 1) {
    atomicAdd(dL_dopacity + gid, o);
  }
}

template <int PPT>
__global__ void __launch_bounds__(256 / PPT)
blend_backward_kernel(const BlendParams P, const int2* __restrict__ ranges, const int* __restrict__ point_list,
                      const float2* __restrict__ xy, const float4* __restrict__ conic_opacity,
                      const float* __restrict__ rgb, const float* __restrict__ final_T,
                      const int* __restrict__ n_contrib, const float* __restrict__ dL_dpixels,
                      float* __restrict__ dL_dmean2D, float* __restrict__ dL_dconic, float* __restrict__ dL_dopacity,
                      float* __restrict__ dL_dcolor) {
  constexpr int NT = 256 / PPT;
  constexpr int NW = NT / 32;
  __shared__ float4 s_a[NT];  // x, y, conic.a, conic.b
  __shared__ float4 s_b[NT];  // conic.c, opacity, power threshold, -
  __shared__ float4 s_c[NT];  // r, g, b, gid (bits)
  __shared__ int s_max[NW];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile_x = blockIdx.x, tile_y = blockIdx.y;
  const int tile_id = tile_y * P.grid_x + tile_x;
  const int px = tile_x * kTile + (lane & 15);
  const int row0 = tile_y * kTile + warp * (2 * PPT) + (lane >> 4);
  const float pxf = (float)px;
  const int2 range = ranges[tile_id];

  float pyf[PPT], T[PPT], T_final[PPT], acc0[PPT], acc1[PPT], acc2[PPT], last_alpha[PPT], lc0[PPT], lc1[PPT], lc2[PPT];
  float dp0[PPT], dp1[PPT], dp2[PPT], bgdot[PPT];
  int kept[PPT];
  int my_max = 0;
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    const int py = row0 + 2 * k;
    pyf[k] = (float)py;
    const bool inside = (px < P.W && py < P.H);
    const size_t pix = inside ? ((size_t)py * P.W + px) : 0;
    T_final[k] = inside ? final_T[pix] : 0.0f;
    T[k] = T_final[k];
    // backward.py:619 last_kept = min(range_end, range_start + n_contrib), relative to range_start
    kept[k] = inside ? min(range.y - range.x, n_contrib[pix]) : 0;
    my_max = max(my_max, kept[k]);
    dp0[k] = inside ? dL_dpixels[3 * pix + 0] : 0.0f;
    dp1[k] = inside ? dL_dpixels[3 * pix + 1] : 0.0f;
    dp2[k] = inside ? dL_dpixels[3 * pix + 2] : 0.0f;
    bgdot[k] = gs_dot3(P.bg0, P.bg1, P.bg2, dp0[k], dp1[k], dp2[k]);  // backward.py:679
    acc0[k] = acc1[k] = acc2[k] = 0.0f;
    last_alpha[k] = 0.0f;
    lc0[k] = lc1[k] = lc2[k] = 0.0f;
  }
  my_max = __reduce_max_sync(0xffffffffu, my_max);
  if (lane == 0) s_max[warp] = my_max;
  __syncthreads();
  int tile_max = 0;
#pragma unroll
  for (int w = 0; w < NW; ++w) tile_max = max(tile_max, s_max[w]);

  const float ddelx_dx = 0.5f * (float)P.W;
  const float ddely_dy = 0.5f * (float)P.H;

  for (int hi = tile_max; hi > 0; hi -= NT) {
    const int cnt = min(NT, hi);
    __syncthreads();
    if (tid < cnt) {
      const int gid = point_list[range.x + hi - 1 - tid];
      const float2 p = xy[gid];
      const float4 co = conic_opacity[gid];
      s_a[tid] = make_float4(p.x, p.y, co.x, co.y);
      s_b[tid] = make_float4(co.z, co.w, gs_power_threshold(co.w), 0.0f);
      s_c[tid] = make_float4(rgb[3 * gid + 0], rgb[3 * gid + 1], rgb[3 * gid + 2], __int_as_float(gid));
    }
    __syncthreads();
    // a warp whose pixels all stopped before this batch has nothing to do in it
    if (my_max <= hi - cnt) continue;
    for (int j = 0; j < cnt; ++j) {
      const int pos = hi - 1 - j;  // position in the tile's list; pixel k replays it iff pos < kept[k]
      if (pos >= my_max) continue; // warp-uniform
      const float4 a = s_a[j];
      const float4 b = s_b[j];
      const float4 c = s_c[j];
      const float dx = a.x - pxf;
      float g[9];
#pragma unroll
      for (int q = 0; q < 9; ++q) g[q] = 0.0f;
      bool any = false;
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        if (pos >= kept[k]) continue;
        const float dy = a.y - pyf[k];
        const float power = gs_power(a.z, a.w, b.x, dx, dy);
        if (power > 0.0f) continue;  // backward.py:647
        if (power < b.z) continue;   // provably alpha < 1/255
        const float G = gs_expf(power);
        const float alpha = f_min(0.99f, b.y * G);
        if (alpha < (1.0f / 255.0f)) continue;  // backward.py:655
        T[k] = T[k] / (1.0f - alpha);
        const float dchannel_dcolor = alpha * T[k];
        acc0[k] = last_alpha[k] * lc0[k] + (1.0f - last_alpha[k]) * acc0[k];
        acc1[k] = last_alpha[k] * lc1[k] + (1.0f - last_alpha[k]) * acc1[k];
        acc2[k] = last_alpha[k] * lc2[k] + (1.0f - last_alpha[k]) * acc2[k];
        lc0[k] = c.x;
        lc1[k] = c.y;
        lc2[k] = c.z;
        float dL_dalpha = gs_dot3(c.x - acc0[k], c.y - acc1[k], c.z - acc2[k], dp0[k], dp1[k], dp2[k]);
        g[0] += dchannel_dcolor * dp0[k];
        g[1] += dchannel_dcolor * dp1[k];
        g[2] += dchannel_dcolor * dp2[k];
        dL_dalpha *= T[k];
        last_alpha[k] = alpha;
        dL_dalpha += (-T_final[k] / (1.0f - alpha)) * bgdot[k];
        const float dL_dG = b.y * dL_dalpha;
        const float gdx = G * dx;
        const float gdy = G * dy;
        const float dG_ddelx = -gdx * a.z - gdy * a.w;
        const float dG_ddely = -gdy * b.x - gdx * a.w;
        g[3] += dL_dG * dG_ddelx * ddelx_dx;
        g[4] += dL_dG * dG_ddely * ddely_dy;
        g[5] += -0.5f * gdx * dx * dL_dG;
        g[6] += -0.5f * gdx * dy * dL_dG;
        g[7] += -0.5f * gdy * dy * dL_dG;
        g[8] += G * dL_dalpha;
        any = true;
      }
      if (__any_sync(0xffffffffu, any))
        warp_reduce9_red(g, lane, __float_as_int(c.w), dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor);
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

template <int PPT>
int launch_bwd(gsb_ctx* ctx, cudaStream_t s, const BlendParams& P, dim3 grid, const int32_t* ranges,
               const int32_t* point_list, const float* xy, const float* conic_opacity, const float* rgb,
               const float* final_T, const int32_t* n_contrib, const float* dL_dpixels, float* dL_dmean2D,
               float* dL_dconic, float* dL_dopacity, float* dL_dcolor) {
  GSB_LAUNCH(ctx, blend_backward_kernel<PPT>, grid, 256 / PPT, 0, s, P, reinterpret_cast<const int2*>(ranges),
             point_list, reinterpret_cast<const float2*>(xy), reinterpret_cast<const float4*>(conic_opacity), rgb,
             final_T, n_contrib, dL_dpixels, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor);
  return GSB_OK;
}

BlendParams make_params(const gsb_frame* f) {
  BlendParams P;
  P.W = f->width;
  P.H = f->height;
  P.grid_x = (f->width + kTile - 1) / kTile;
  P.bg0 = f->background[0];
  P.bg1 = f->background[1];
  P.bg2 = f->background[2];
  return P;
}

}  // namespace

int g_blend_fwd_ppt = 1;
int g_blend_bwd_ppt = 1;

GSB_API int gsb_blend_forward(gsb_ctx* ctx, gsb_stream s_, const gsb_frame* f, const int32_t* ranges,
                              const int32_t* point_list, const float* points_xy, const float* rgb,
                              const float* conic_opacity, const float* depths, float* image, float* inv_depth,
                              float* final_T, int32_t* n_contrib) {
  if (!ctx) return GSB_ERR_INVALID;
  GSB_REQUIRE(ctx, f && f->width > 0 && f->height > 0, "gsb_blend_forward: bad frame");
  GSB_REQUIRE(ctx, gsb_aligned16(conic_opacity), "gsb_blend_forward: conic_opacity must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)s_;
  BlendParams P = make_params(f);
  dim3 grid(P.grid_x, (f->height + kTile - 1) / kTile);
  switch (g_blend_fwd_ppt) {
    case 2: return launch_fwd<2>(ctx, s, P, grid, ranges, point_list, points_xy, rgb, conic_opacity, depths, image, inv_depth, final_T, n_contrib);
    case 4: return launch_fwd<4>(ctx, s, P, grid, ranges, point_list, points_xy, rgb, conic_opacity, depths, image, inv_depth, final_T, n_contrib);
    case 8: return launch_fwd<8>(ctx, s, P, grid, ranges, point_list, points_xy, rgb, conic_opacity, depths, image, inv_depth, final_T, n_contrib);
    default: return launch_fwd<1>(ctx, s, P, grid, ranges, point_list, points_xy, rgb, conic_opacity, depths, image, inv_depth, final_T, n_contrib);
  }
}

GSB_API int gsb_blend_backward(gsb_ctx* ctx, gsb_stream s_, const gsb_frame* f, int32_t n, const int32_t* ranges,
                               const int32_t* point_list, const float* points_xy, const float* conic_opacity,
                               const float* rgb, const float* final_T, const int32_t* n_contrib,
                               const float* dL_dpixels, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                               float* dL_dcolor) {
  if (!ctx) return GSB_ERR_INVALID;
  GSB_REQUIRE(ctx, f && f->width > 0 && f->height > 0 && n >= 0, "gsb_blend_backward: bad frame");
  GSB_REQUIRE(ctx, gsb_aligned16(conic_opacity), "gsb_blend_backward: conic_opacity must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)s_;
  if (n == 0) return GSB_OK;
  GSB_CUDA(ctx, cudaMemsetAsync(dL_dmean2D, 0, sizeof(float) * 3 * (size_t)n, s));
  GSB_CUDA(ctx, cudaMemsetAsync(dL_dconic, 0, sizeof(float) * 4 * (size_t)n, s));
  GSB_CUDA(ctx, cudaMemsetAsync(dL_dopacity, 0, sizeof(float) * (size_t)n, s));
  GSB_CUDA(ctx, cudaMemsetAsync(dL_dcolor, 0, sizeof(float) * 3 * (size_t)n, s));
  BlendParams P = make_params(f);
  dim3 grid(P.grid_x, (f->height + kTile - 1) / kTile);
  switch (g_blend_bwd_ppt) {
    case 2: return launch_bwd<2>(ctx, s, P, grid, ranges, point_list, points_xy, conic_opacity, rgb, final_T, n_contrib, dL_dpixels, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor);
    case 4: return launch_bwd<4>(ctx, s, P, grid, ranges, point_list, points_xy, conic_opacity, rgb, final_T, n_contrib, dL_dpixels, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor);
    case 8: return launch_bwd<8>(ctx, s, P, grid, ranges, point_list, points_xy, conic_opacity, rgb, final_T, n_contrib, dL_dpixels, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor);
    default: return launch_bwd<1>(ctx, s, P, grid, ranges, point_list, points_xy, conic_opacity, rgb, final_T, n_contrib, dL_dpixels, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor);
  }
}
