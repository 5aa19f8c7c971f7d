// Shared by blend.cu (forward) and blend_bwd.cu (backward): launch parameters, the pixel mapping,
// the conservative per-block culling mask and the stable batch compaction.
//
// Pixel mapping: a CTA is one 16x16 tile, 256 threads.  Warp w owns the 4-row x 8-column block
// (bx = w & 1, by = w >> 1) and lane l the pixel (8*bx + (l & 7), 4*by + (l >> 3)) of it: a compact
// footprint means fewer warps are touched by a given Gaussian and more of their lanes are live
// when they are (19.5% of the (Gaussian, warp) combinations with 17 live lanes, against 23.7% with
// 14 for 2x16 strips on the headline scene).  Loads/stores are 32-byte row segments.
#pragma once
#include "common.cuh"


namespace {

struct BlendParams {
  int W, H, grid_x;
  float bg0, bg1, bg2;
  int cull;  // 1: per-block culling masks (default); 0: keep every list entry (A/B switch, same results)
};

__device__ __forceinline__ float gs_sqrt_approx(float x) {  // MUFU.SQRT, ~1 ulp
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// 32-bit mask, bit (2*r + h): on tile row r the Gaussian can reach alpha >= 1/255 for some pixel of
// the 8-column half h.  `thr` is gs_power_threshold(opacity):  power >= thr  <=>
//   q(dx, dy) = a dx^2 + 2 b dx dy + c dy^2 <= t2 = -2 thr        (dx = gx - px, dy = gy - py).
// On a row (dy fixed) that is  px in [xc - hw, xc + hw]  with  xc = gx + (b/a) dy  and
// hw^2 = t2/a - (c/a - (b/a)^2) dy^2.  Everything is conservative: t2 carries the threshold's own
// margin, hw^2 is inflated by 1e-5 of the magnitude of each of its terms (two orders above their
// rounding errors), the interval by 0.05 px + 1e-3/a, and Gaussians whose coordinates are large
// enough for fp32 rounding to approach that margin, or whose conic is degenerate / NaN, keep every
// bit.  Every comparison is written so that a NaN falls on the "keep" side.  Explicit FMAs: the
// function gives the same value in translation units built with and without -fmad.
__device__ __forceinline__ unsigned gs_block_mask(float gx, float gy, float ca, float cb, float cc, float thr,
                                                  float x0, float y0) {
  // a non-finite centre or conic makes NaN exponents, which pass both of the reference's tests (and
  // poison the pixel there): never culled, whatever the opacity
  if (!(fabsf((gx + gy) + (ca + cb) + cc) < 3.0e38f)) return 0xffffffffu;
  if (thr == __int_as_float(0x7f800000)) return 0u;  // opacity < 1/255: alpha can never reach 1/255
  if (!(ca > 0.0f)) return 0xffffffffu;
  const float t2 = __fmaf_rn(-2.0f * thr, 1.0001f, 1e-3f);
  const float inv_a = __fdiv_rn(1.0f, ca);
  const float kb = __fmul_rn(cb, inv_a);
  const float ck = __fmul_rn(cc, inv_a);
  const float kb2 = __fmul_rn(kb, kb);
  const float K0 = __fmul_rn(t2, inv_a);
  const float K0i = __fmaf_rn(1e-5f, fabsf(K0), K0) + __fmul_rn(1e-6f * inv_a, inv_a);
  const float K3 = __fmaf_rn(1e-5f, kb2 + fabsf(ck), kb2 - ck);   // -(c/a - (b/a)^2), inflated
  const float m = __fmaf_rn(1e-3f, inv_a, 0.05f);
  // magnitude guard: with every coordinate below 3e4 the fp32 / MUFU errors stay under 0.03 px
  const float dy_max = fabsf(gy - y0) + 16.0f;
  const float reach = fabsf(gx) + fabsf(kb) * dy_max + gs_sqrt_approx(fabsf(K0i) + fabsf(K3) * dy_max * dy_max);
  if (!(reach < 3e4f) || !(dy_max < 3e4f)) return 0xffffffffu;
  const float gx_lo = gx - m, gx_hi = gx + m;
  const float xa = x0 + 7.0f, xb = x0 + 8.0f, xe = x0 + 15.0f;
  unsigned mask = 0u;
#pragma unroll 4
  for (int r = 0; r < 16; ++r) {
    const float dy = gy - (y0 + (float)r);
    const float h2 = __fmaf_rn(K3, __fmul_rn(dy, dy), K0i);
    const float hw = gs_sqrt_approx(fmaxf(h2, 0.0f));
    const float lo = __fmaf_rn(kb, dy, gx_lo) - hw;
    const float hi = __fmaf_rn(kb, dy, gx_hi) + hw;
    const bool row_miss = h2 < 0.0f;
    const bool miss0 = row_miss || (lo > xa) || (hi < x0);
    const bool miss1 = row_miss || (lo > xe) || (hi < xb);
    if (!miss0) mask |= (1u << (2 * r));
    if (!miss1) mask |= (1u << (2 * r + 1));
  }
  return mask;
}

// the mask bits of warp w's 4x8 block
__device__ __forceinline__ unsigned gs_warp_mask(int warp) {
  const int bx = warp & 1, by = warp >> 1;
  return (0x55u << bx) << (8 * by);  // rows 4*by .. 4*by+3, half bx
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

// Per-warp compaction of the staged batch: the indices (0..cnt-1, in order) of the entries whose
// block mask touches this warp's block -- and, for the backward, that lie inside the warp's replay
// range (position < pos_limit) -- are written to widx[0..n).  One ballot per 32 entries; the main
// loop then never sees an entry it would skip.
__device__ __forceinline__ int warp_compact_hits(const int2* s_meta, int cnt, unsigned my_mask, int pos_limit, int lane,
                                                 unsigned char* widx) {
  int n = 0;
  for (int base = 0; base < cnt; base += 32) {
    const int e = base + lane;
    bool hit = false;
    if (e < cnt) {
      const int2 m = s_meta[e];
      hit = (((unsigned)m.y & my_mask) != 0u) && (m.x < pos_limit);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    GSB_DCHECK(n + __popc(bal) <= 256 && e < 256 + 32);
    if (hit) widx[n + __popc(bal & ((1u << lane) - 1u))] = (unsigned char)e;
    n += __popc(bal);
  }
  __syncwarp();
  return n;
}

inline BlendParams make_blend_params(const gsb_ctx* ctx, const gsb_frame* f) {
  BlendParams P;
  P.W = f->width;
  P.H = f->height;
  P.grid_x = (f->width + kTile - 1) / kTile;
  P.bg0 = f->background[0];
  P.bg1 = f->background[1];
  P.bg2 = f->background[2];
  P.cull = ctx->opt.blend_cull;
  return P;
}

}  // namespace
