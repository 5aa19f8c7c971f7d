// Shared by blend.cu (forward) and blend_bwd.cu (backward): launch parameters, the conservative
// per-row culling mask and the stable batch compaction.
#pragma once
#include "common.cuh"

extern int g_blend_cull;

namespace {

struct BlendParams {
  int W, H, grid_x;
  float bg0, bg1, bg2;
  int cull;  // 1: per-row culling masks (default); 0: keep every list entry (A/B switch, same results)
};

// 16-bit mask of the tile rows y0..y0+15 on which the Gaussian can reach alpha >= 1/255 for some
// pixel x in [x0, x0+15].  Conservative: every comparison is written so that NaN / degenerate
// conics / rounding fall on the "keep" side; `thr` is gs_power_threshold(opacity).
__device__ __forceinline__ unsigned gs_row_mask(float gx, float gy, float ca, float cb, float cc, float thr, float x0,
                                                float y0) {
  if (thr == __int_as_float(0x7f800000)) return 0u;  // opacity < 1/255: alpha can never reach 1/255
  const float t2 = -2.0f * thr * 1.0001f + 1e-3f;    // q = a dx^2 + 2b dx dy + c dy^2 <= t2  <=>  power >= thr
  const float inv_a = 1.0f / ca;
  const float xlo = x0 - 0.05f, xhi = x0 + 15.05f;
  unsigned mask = 0u;
#pragma unroll 2
  for (int r = 0; r < 16; ++r) {
    const float dy = gy - (y0 + (float)r);
    const float bd = cb * dy;
    const float bd2 = bd * bd;
    const float ac = ca * (cc * dy * dy - t2);
    // discriminant of a dx^2 + 2 bd dx + (c dy^2 - t2) <= 0, inflated by 100x the worst rounding error
    const float disc = (bd2 - ac) + 1e-5f * (bd2 + fabsf(ac)) + 1e-6f;
    const float sq = sqrtf(fmaxf(disc, 0.0f)) + 1e-3f;
    // dx in [(-bd - sq)/a, (-bd + sq)/a]  =>  pixel x = gx - dx
    const float px_lo = gx - (-bd + sq) * inv_a;
    const float px_hi = gx - (-bd - sq) * inv_a;
    const bool miss = (ca > 0.0f) && ((disc < 0.0f) || (px_lo > xhi) || (px_hi < xlo));
    if (!miss) mask |= (1u << r);
  }
  return mask;
}

// Stable block-wide compaction slot for `keep`; returns the slot (valid when keep) and the total.
template <int NW>
__device__ __forceinline__ int compact_slot(bool keep, int lane, int warp, int* s_wcnt, int& total) {
  const unsigned bal = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) s_wcnt[warp] = __popc(bal);
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    const int c = s_wcnt[w];
    if (w < warp) base += c;
    tot += c;
  }
  total = tot;
  return base + __popc(bal & ((1u << lane) - 1u));
}

inline BlendParams make_blend_params(const gsb_frame* f) {
  BlendParams P;
  P.W = f->width;
  P.H = f->height;
  P.grid_x = (f->width + kTile - 1) / kTile;
  P.bg0 = f->background[0];
  P.bg1 = f->background[1];
  P.bg2 = f->background[2];
  P.cull = g_blend_cull;
  return P;
}

}  // namespace
