// blend_bwd.cu -- back-to-front replay of the tile blend: replaces wp_render_backward_kernel
// (reference backward.py:558-706).  See blend.cu for the design notes shared with the forward.
//
// This translation unit is compiled with -fmad=true: gradients are compared with a tolerance
// (rel 1e-3), so the gradient arithmetic may contract to FMAs and use one reciprocal instead of two
// IEEE divisions.  Everything that feeds a DECISION (the Gaussian exponent, the exponential, the
// alpha tests) goes through gs_power / gs_expf, which round every operation explicitly and are
// therefore identical to the forward's and the oracle's no matter how this file is compiled.
#include "blend_common.cuh"

namespace {

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

__global__ void __launch_bounds__(256, 4)
blend_backward_kernel(const BlendParams P, const int2* __restrict__ ranges, const int* __restrict__ point_list,
                      const float2* __restrict__ xy, const float4* __restrict__ conic_opacity,
                      const float* __restrict__ rgb, const float* __restrict__ final_T,
                      const int* __restrict__ n_contrib, const float* __restrict__ dL_dpixels,
                      float* __restrict__ dL_dmean2D, float* __restrict__ dL_dconic, float* __restrict__ dL_dopacity,
                      float* __restrict__ dL_dcolor) {
  constexpr int NT = 256, NW = 8;
  __shared__ float4 s_a[NT];   // x, y, conic.a, conic.b
  __shared__ float4 s_b[NT];   // conic.c, opacity, power threshold, -
  __shared__ float4 s_c[NT];   // r, g, b, gid (bits)
  __shared__ int2 s_meta[NT];  // position in the tile's list, block mask
  __shared__ int s_max[NW];
  __shared__ int s_wcnt[NW];
  __shared__ unsigned char s_widx[NW][NT];  // per warp: the staged entries it has to replay

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile_x = blockIdx.x, tile_y = blockIdx.y;
  const int tile_id = tile_y * P.grid_x + tile_x;
  const int px = tile_x * kTile + (warp & 1) * 8 + (lane & 7);
  const int py = tile_y * kTile + (warp >> 1) * 4 + (lane >> 3);
  const float pxf = (float)px, pyf = (float)py;
  const int2 range = ranges[tile_id];
  const unsigned my_mask = gs_warp_mask(warp);
  const float tile_x0 = (float)(tile_x * kTile), tile_y0 = (float)(tile_y * kTile);

  const bool inside = (px < P.W && py < P.H);
  const size_t pix = inside ? ((size_t)py * P.W + px) : 0;
  const float T_final = inside ? final_T[pix] : 0.0f;
  float T = T_final;
  // backward.py:619 last_kept = min(range_end, range_start + n_contrib), relative to range_start
  const int kept = inside ? min(range.y - range.x, n_contrib[pix]) : 0;
  const float dp0 = inside ? dL_dpixels[3 * pix + 0] : 0.0f;
  const float dp1 = inside ? dL_dpixels[3 * pix + 1] : 0.0f;
  const float dp2 = inside ? dL_dpixels[3 * pix + 2] : 0.0f;
  const float bgdot = gs_dot3(P.bg0, P.bg1, P.bg2, dp0, dp1, dp2);  // backward.py:679
  float acc0 = 0.0f, acc1 = 0.0f, acc2 = 0.0f, last_alpha = 0.0f, lc0 = 0.0f, lc1 = 0.0f, lc2 = 0.0f;

  const int my_max = __reduce_max_sync(0xffffffffu, kept);
  if (lane == 0) s_max[warp] = my_max;
  __syncthreads();
  int tile_max = 0;
#pragma unroll
  for (int w = 0; w < NW; ++w) tile_max = max(tile_max, s_max[w]);

  const float ddelx_dx = 0.5f * (float)P.W;
  const float ddely_dy = 0.5f * (float)P.H;

  for (int hi = tile_max; hi > 0; hi -= NT) {
    const int n_in = min(NT, hi);
    __syncthreads();
    float4 ea, eb, ec;
    unsigned bmask = 0u;
    if (tid < n_in) {
      const int gid = point_list[range.x + hi - 1 - tid];
      const float2 p = xy[gid];
      const float4 co = conic_opacity[gid];
      const float thr = gs_power_threshold(co.w);
      bmask = P.cull ? gs_block_mask(p.x, p.y, co.x, co.y, co.z, thr, tile_x0, tile_y0) : 0xffffffffu;
      ea = make_float4(p.x, p.y, co.x, co.y);
      eb = make_float4(co.z, co.w, thr, 0.0f);
      ec = make_float4(rgb[3 * gid + 0], rgb[3 * gid + 1], rgb[3 * gid + 2], __int_as_float(gid));
    }
    int cnt;
    const int slot = compact_slot<NW>(bmask != 0u, lane, warp, s_wcnt, cnt);
    if (bmask != 0u) {
      s_a[slot] = ea;
      s_b[slot] = eb;
      s_c[slot] = ec;
      s_meta[slot] = make_int2(hi - 1 - tid, (int)bmask);
    }
    __syncthreads();
    // a warp whose pixels all stopped before this batch has nothing to do in it
    if (my_max <= hi - n_in) continue;
    // entries that touch this warp's block and lie inside its replay range (position < my_max)
    const int wn = warp_compact_hits(s_meta, cnt, my_mask, my_max, lane, s_widx[warp]);
    for (int q = 0; q < wn; ++q) {
      const int j = s_widx[warp][q];
      const int pos = s_meta[j].x;  // position in the tile's list; the pixel replays it iff pos < kept
      const float4 a = s_a[j];
      const float4 b = s_b[j];
      const float dx = a.x - pxf;
      const float dy = a.y - pyf;
      float g[9];
#pragma unroll
      for (int q = 0; q < 9; ++q) g[q] = 0.0f;
      bool any = false;
      const float power = gs_power(a.z, a.w, b.x, dx, dy);
      // backward.py:647 (power > 0), the conservative exponent threshold, and the replay limit
      if (pos < kept && !(power > 0.0f) && !(power < b.z)) {
        const float G = gs_expf(power);
        const float alpha = f_min(0.99f, b.y * G);
        if (!(alpha < (1.0f / 255.0f))) {  // backward.py:655
          const float4 c = s_c[j];
          const float inv_1ma = 1.0f / (1.0f - alpha);   // backward.py:658,680 divide twice by (1 - alpha)
          T = T * inv_1ma;
          const float dchannel_dcolor = alpha * T;
          acc0 = last_alpha * lc0 + (1.0f - last_alpha) * acc0;
          acc1 = last_alpha * lc1 + (1.0f - last_alpha) * acc1;
          acc2 = last_alpha * lc2 + (1.0f - last_alpha) * acc2;
          lc0 = c.x;
          lc1 = c.y;
          lc2 = c.z;
          float dL_dalpha = gs_dot3(c.x - acc0, c.y - acc1, c.z - acc2, dp0, dp1, dp2);
          g[0] = dchannel_dcolor * dp0;
          g[1] = dchannel_dcolor * dp1;
          g[2] = dchannel_dcolor * dp2;
          dL_dalpha *= T;
          last_alpha = alpha;
          dL_dalpha += (-T_final * inv_1ma) * bgdot;
          const float dL_dG = b.y * dL_dalpha;
          const float gdx = G * dx;
          const float gdy = G * dy;
          const float dG_ddelx = -gdx * a.z - gdy * a.w;
          const float dG_ddely = -gdy * b.x - gdx * a.w;
          g[3] = dL_dG * dG_ddelx * ddelx_dx;
          g[4] = dL_dG * dG_ddely * ddely_dy;
          g[5] = -0.5f * gdx * dx * dL_dG;
          g[6] = -0.5f * gdx * dy * dL_dG;
          g[7] = -0.5f * gdy * dy * dL_dG;
          g[8] = G * dL_dalpha;
          any = true;
        }
      }
      if (__any_sync(0xffffffffu, any))
        warp_reduce9_red(g, lane, __float_as_int(s_c[j].w), dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor);
    }
  }
}

}  // namespace

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
  BlendParams P = make_blend_params(f);
  dim3 grid(P.grid_x, (f->height + kTile - 1) / kTile);
  GSB_LAUNCH(ctx, blend_backward_kernel, grid, 256, 0, s, P, reinterpret_cast<const int2*>(ranges), point_list,
             reinterpret_cast<const float2*>(points_xy), reinterpret_cast<const float4*>(conic_opacity), rgb, final_T,
             n_contrib, dL_dpixels, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor);
  return GSB_OK;
}
