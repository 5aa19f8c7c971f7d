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
                      float* __restrict__ dL_dcolor, const unsigned* __restrict__ block_masks) {
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
      const int e = range.x + hi - 1 - tid;
      // the forward's culling mask when it was handed on, else recomputed below (same value)
      const bool have = P.cull && block_masks != nullptr;
      bmask = have ? block_masks[e] : 0xffffffffu;
      if (bmask != 0u) {
        const int gid = point_list[e];
        const float2 p = xy[gid];
        const float4 co = conic_opacity[gid];
        const float thr = gs_power_threshold(co.w);
        if (P.cull && !have) bmask = gs_block_mask(p.x, p.y, co.x, co.y, co.z, thr, tile_x0, tile_y0);
        ea = make_float4(p.x, p.y, co.x, co.y);
        eb = make_float4(co.z, co.w, thr, 0.0f);
        ec = make_float4(rgb[3 * gid + 0], rgb[3 * gid + 1], rgb[3 * gid + 2], __int_as_float(gid));
      }
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


// ------------------------------------------------------------------------------------------
// Tensor-core pixel reduction (default).  The per-Gaussian gradients are sums over the pixels of a
// warp's 4x8 block:  with  s = G * dL_dalpha  and  w = alpha * T  per (pixel, Gaussian),
//   dL_dcolor[c]           = sum_p w * dL_dpixel[p][c]
//   dL_dopacity            = sum_p s
//   dL_dmean2D, dL_dconic  = linear combinations (per-Gaussian coefficients) of the six moments
//                            sum_p s * {1, i, r, i^2, i*r, r^2},  (i, r) = block-local pixel coords
// i.e. two small matrix products  [16 Gaussians x 32 pixels] x [32 pixels x 6]  and  x [32 x 3].
// Each lane stores its s and w for 16 consecutive hits into a per-warp shared-memory tile; the warp
// then multiplies the tile with the constant pixel matrices on the tensor cores
// (mma.sync.m16n8k8 TF32, fp32 accumulate; operands split hi + lo so the products carry ~21
// mantissa bits; the monomials are small integers, exact in TF32) and 16 lanes turn the moments
// into the nine gradient scalars and issue the REDs.  This replaces the 14-shuffle / 14-select /
// 14-add butterfly per (warp, Gaussian) -- a third of the old kernel's instructions -- by
// 2 STS per hit plus ~13 instructions per hit amortised over the group.
constexpr int kGrp = 16;   // hits per MMA group (the M dimension)
constexpr int kSRow = 36;  // padded row length (floats) of the S / W tiles: conflict-free 16-byte reads

struct BwdSmem {
  float4 a[256];   // x, y, conic.a, conic.c   (two packed pairs: see gs_power_packed)
  float4 b[256];   // conic.b, opacity, power threshold, position in the tile's list (int bits)
  float4 c[256];   // r, g, b, gid (int bits)
  int2 meta[256];  // position in the tile's list, block mask
  int smax[8];
  int wcnt[8];
  unsigned char widx[8][256];     // per warp: the staged entries it has to replay
  float sw[8][2][kGrp][kSRow];    // per warp: S tile (s per hit x pixel), W tile; the S tile is reused for results
};
static_assert(offsetof(BwdSmem, sw) % 16 == 0, "the S / W tiles are read with 16-byte loads");

// TF32 head of x (mantissa truncated to 10 bits: one LOP3; cvt.rna.tf32 costs four instructions on
// sm_100a).  x - tf32_hi(x) is exact, so hi + lo carries >= 20 mantissa bits through the MMA.
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// exp(x) for x <= 0 on the MUFU unit (2 ulp; with the rounding of x * log2(e), < 6e-7 relative for
// the exponents that pass the threshold test).  The backward's alpha VALUE only feeds tolerance-compared
// gradients.
__device__ __forceinline__ float exp_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
  return r;
}

// The backward's one DECISION per pair, alpha >= 1/255 (backward.py:655), is taken on the MUFU value of alpha; the
// forward took it on the contract's gs_expf (forward.py:479).  The two can only disagree for a pair whose alpha lies
// within the MUFU error (< 6e-7 relative) of 1/255; such a pair shifts the T that pixel replays by 0.39%.
// MEASURED, not assumed (gsb_selftest_work_counters walks every pair of a frame with both rules; bench.py prints
// the counters, tests/test_gpu_large.py bounds them): 1 disagreement in 42.1 M evaluated pairs at the headline
// size, 2 in 143.7 M at 1M Gaussians / 1080p.  Two exact variants were built and measured in round 2 -- re-deciding
// with gs_expf inside a band around 1/255 (on alpha: 324-328 us; on the exponent, off the MUFU dependency chain:
// 314 us) against 300 us for this test: the replay loop pays 5-9% for any extra compare + branch per hit, so the
// measured 2e-8 disagreement rate is accepted and reported instead.
__device__ __forceinline__ bool bwd_alpha_hit(const float power, const float opacity, float& G, float& alpha) {
  G = exp_approx(power);
  alpha = f_min(0.99f, opacity * G);
  return !(alpha < (1.0f / 255.0f));
}

__device__ __forceinline__ float rcp_approx(float x) {  // MUFU.RCP, 1 ulp; x must be a normal number
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// D[16x8] += A[16x4] * B[4x8] (mma.sync m16n8k4, TF32 inputs, fp32 accumulate).  Fragment layout (PTX ISA):
//   A: a0 (g, t)  a1 (g+8, t);  B: b0 (k=t, n=g);  D: d0 (g, 2t) d1 (g, 2t+1) d2 (g+8, 2t) d3 (g+8, 2t+1)
//   with g = lane >> 2, t = lane & 3.  (The m16n8k8 form needs four A operands live together: experiment 19.)
__device__ __forceinline__ void mma_tf32_k4(float d[4], float a0, float a1, float b0) {
  asm volatile(
      "mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(__float_as_uint(a0)), "r"(__float_as_uint(a1)), "r"(__float_as_uint(b0)));
}

__device__ __forceinline__ float f4_get(const float4& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w)); }

// The tensor-core part of a group flush, shared by the two tensor-core kernels: the warp's S / W tiles (kGrp hits
// x 32 pixels) times the constant pixel matrices.  On return lane l (and l + 16) holds, for the group's hit
// l & 15, the six moments  sum_p s * {1, i, r, i^2, i r, r^2}  (m03, m45) and the three colour sums
// sum_p w * dL_dpixel[p][c]  (col); the S tile is used as scratch for the results.
__device__ __forceinline__ void bwd_group_moments(float* tS, const float* tW, const int lane, const int fg, const int ft,
                                                  const float m0, const float m1, const float m2, const float* dp_row,
                                                  const int dp_cols, float4& m03, float2& m45, float4& col) {
  __syncwarp();
  float dm[4] = {0.0f, 0.0f, 0.0f, 0.0f}, dc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  // k-step ks (m16n8k4) covers the pixels (i = ks, r = 0..3); lane (fg, ft) supplies rows fg, fg + 8
  // at pixel 8 ft + ks: two float4 per half row serve four k-steps.
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const float4 va = *reinterpret_cast<const float4*>(tS + fg * kSRow + 8 * ft + 4 * half);
    const float4 vb = *reinterpret_cast<const float4*>(tS + (fg + 8) * kSRow + 8 * ft + 4 * half);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float ks = (float)(4 * half + e);
      const float a0 = f4_get(va, e), a1 = f4_get(vb, e);
      const float h0 = tf32_hi(a0), h1 = tf32_hi(a1);
      const float b0 = m0 + ks * (m1 + ks * m2);
      mma_tf32_k4(dm, h0, h1, b0);
      mma_tf32_k4(dm, a0 - h0, a1 - h1, b0);
    }
  }
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const float4 va = *reinterpret_cast<const float4*>(tW + fg * kSRow + 8 * ft + 4 * half);
    const float4 vb = *reinterpret_cast<const float4*>(tW + (fg + 8) * kSRow + 8 * ft + 4 * half);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = 4 * half + e;
      const float a0 = f4_get(va, e), a1 = f4_get(vb, e);
      const float h0 = tf32_hi(a0), h1 = tf32_hi(a1);
      const float dpv = (i < dp_cols) ? __ldg(dp_row + 3 * i) : 0.0f;
      const float p0 = tf32_hi(dpv);
      mma_tf32_k4(dc, h0, h1, p0);
      mma_tf32_k4(dc, a0 - h0, a1 - h1, p0);
      mma_tf32_k4(dc, h0, h1, dpv - p0);
    }
  }
  __syncwarp();  // every lane has read its operands: the S tile may now take the results
  // (12 floats at the head of each of its 16 rows: the padding columns 32..35 of the rows stay untouched -- they
  // carry the metadata of an open group across batches, see blend_backward_mma_kernel)
  *reinterpret_cast<float2*>(tS + fg * kSRow + 2 * ft) = make_float2(dm[0], dm[1]);
  *reinterpret_cast<float2*>(tS + (fg + 8) * kSRow + 2 * ft) = make_float2(dm[2], dm[3]);
  if (ft < 2) {
    *reinterpret_cast<float2*>(tS + fg * kSRow + 8 + 2 * ft) = make_float2(dc[0], dc[1]);
    *reinterpret_cast<float2*>(tS + (fg + 8) * kSRow + 8 + 2 * ft) = make_float2(dc[2], dc[3]);
  }
  __syncwarp();
  m03 = *reinterpret_cast<const float4*>(tS + (lane & 15) * kSRow);
  m45 = *reinterpret_cast<const float2*>(tS + (lane & 15) * kSRow + 4);
  col = *reinterpret_cast<const float4*>(tS + (lane & 15) * kSRow + 8);
}

// a hit none of the 32 pixels used sums to nine exact zeros: adding them would be a no-op
__device__ __forceinline__ bool bwd_group_nonzero(const float4& m03, const float2& m45, const float4& col) {
  return ((__float_as_uint(m03.x) | __float_as_uint(m03.y) | __float_as_uint(m03.z) | __float_as_uint(m03.w) |
           __float_as_uint(m45.x) | __float_as_uint(m45.y) | __float_as_uint(col.x) | __float_as_uint(col.y) |
           __float_as_uint(col.z)) << 1) != 0u;
}

// The moments of one hit (ga = x, y, conic.a, conic.c; gb = conic.b, opacity, ..; block origin bx0, by0) turned
// into its gradients (backward.py:683-706): g[0..2] dL_dconic a, b, c; g[3] dL_dmean2D x; g[4] dL_dmean2D y.
// dL_dopacity is m03.x, dL_dcolor is col.
__device__ __forceinline__ void bwd_group_gradients(const float4& ga, const float4& gb, const float bx0f, const float by0f,
                                                    const float4& m03, const float2& m45, const float ddelx_dx,
                                                    const float ddely_dy, float g[5]) {
  const float ux = ga.x - bx0f, uy = ga.y - by0f;   // dx = ux - i, dy = uy - r
  const float S0 = m03.x, Si = m03.y, Sr = m03.z, Sii = m03.w, Sir = m45.x, Srr = m45.y;
  const float Sdx = ux * S0 - Si;
  const float Sdy = uy * S0 - Sr;
  const float Sdxx = ux * (Sdx - Si) + Sii;
  const float Sdxy = ux * Sdy - uy * Si + Sir;
  const float Sdyy = uy * (Sdy - Sr) + Srr;
  const float o = gb.y;  // dL_dG = opacity * dL_dalpha (backward.py:683)
  g[3] = -o * (ga.z * Sdx + gb.x * Sdy) * ddelx_dx;   // backward.py:691-695
  g[4] = -o * (ga.w * Sdy + gb.x * Sdx) * ddely_dy;
  g[0] = -0.5f * o * Sdxx, g[1] = -0.5f * o * Sdxy, g[2] = -0.5f * o * Sdyy;   // backward.py:698-703
}

// 16-byte vector reduction (sm_90+): four float adds at one 16-byte-aligned address in ONE instruction and one
// L2 sector operation (REDG.E.ADD.F32x4, same FTZ.RN flavour as the scalar RED atomicAdd compiles to).
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// PACKED: the nine sums of a flushed Gaussian go to ONE 48-byte record
//   acc[12 gid + 0..3] = dL_dconic a, b, c, dL_dopacity      acc[+4..7] = dL_dcolor r, g, b, dL_dmean2D x
//   acc[+8] = dL_dmean2D y                                    (+9..11 unused)
// with two vector REDs and one scalar RED instead of nine scalar REDs into four arrays: per flush the 16 lanes
// touch 48 instead of 144 sectors (the replay is co-limited by the LSU data pipe, profiles/r01_ncu_bwd_warp_
// autonomous.md).  preprocess_backward_kernel<.., PACKED> unpacks the record into the reference's arrays.
template <bool PACKED>   // four resident CTAs per SM: 64 registers (3 CTAs with 80 registers measured slower in round 1)
__global__ void __launch_bounds__(256, 4)
blend_backward_mma_kernel(const BlendParams P, const int2* __restrict__ ranges, const int* __restrict__ point_list,
                          const float2* __restrict__ xy, const float4* __restrict__ conic_opacity,
                          const float* __restrict__ rgb, const float* __restrict__ final_T,
                          const int* __restrict__ n_contrib, const float* __restrict__ dL_dpixels,
                          float* __restrict__ dL_dmean2D, float* __restrict__ dL_dconic,
                          float* __restrict__ dL_dopacity, float* __restrict__ dL_dcolor,
                          const unsigned* __restrict__ block_masks, float* __restrict__ acc_packed) {
  constexpr int NT = 256, NW = 8;
  gsb_pdl_wait();
  gsb_pdl_launch_dependents();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile_x = blockIdx.x, tile_y = blockIdx.y;
  const int tile_id = tile_y * P.grid_x + tile_x;
  const int bx0 = tile_x * kTile + (warp & 1) * 8, by0 = tile_y * kTile + (warp >> 1) * 4;  // block origin
  const int px = bx0 + (lane & 7);
  const int py = by0 + (lane >> 3);
  const gs_f2 npxy = gs_pack2(-(float)px, -(float)py);
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
  // Behind-sum in scalar form.  backward.py:667-680 carries accum_rec (a colour) and evaluates
  //   dL_dalpha_i = T_i (c_i - accum_rec_i) . dL_dpix - T_final/(1 - alpha_i) (bg . dL_dpix).
  // With accum_rec_i T_i = (sum_{k behind i} c_k alpha_k T_k) / (1 - alpha_i) this is
  //   dL_dalpha_i = T_i (c_i . dL_dpix) - gamma_i / (1 - alpha_i),
  //   gamma_i = T_final (bg . dL_dpix) + sum_{k behind i} (c_k . dL_dpix) alpha_k T_k :
  // one scalar of state instead of seven and a third of the arithmetic (same value up to rounding).
  float gamma = T_final * gs_dot3(P.bg0, P.bg1, P.bg2, dp0, dp1, dp2);

  // Constant B fragments.  Logical column (kk, t) / (kk, t+4) of k-step kk is the pixel owned by
  // lane 8t + kk / 8t + 4 + kk (block coords i = kk or 4 + kk, r = t), so that a lane's A operands
  // for the four k-steps are two contiguous float4 of a tile row.
  const int fg = lane >> 2, ft = lane & 3;
  // monomial column fg of pixel (i, r = ft) as  m0 + i * (m1 + i * m2):  1, i, r, i^2, i*r, r^2, 0, 0
  const float fr = (float)ft;
  const float m0 = fg == 0 ? 1.0f : fg == 2 ? fr : fg == 5 ? fr * fr : 0.0f;
  const float m1 = fg == 1 ? 1.0f : fg == 4 ? fr : 0.0f;
  const float m2 = fg == 3 ? 1.0f : 0.0f;
  // dL_dpixels as a B operand: column fg < 3 of pixel (i, r = ft); re-read (L1) at every group
  // rather than held in 16 registers.  Pixels outside the image read pixel 0 and are zeroed.
  const bool dp_row_ok = fg < 3 && (by0 + ft) < P.H;
  const float* const dp_row = dL_dpixels + (dp_row_ok ? 3 * ((size_t)(by0 + ft) * P.W + bx0) + fg : 0);
  const int dp_cols = dp_row_ok ? min(8, P.W - bx0) : 0;  // valid i are 0 .. dp_cols-1

  const int my_max = __reduce_max_sync(0xffffffffu, kept);
  if (lane == 0) sm.smax[warp] = my_max;
  __syncthreads();
  int tile_max = 0;
#pragma unroll
  for (int w = 0; w < NW; ++w) tile_max = max(tile_max, sm.smax[w]);

  const float ddelx_dx = 0.5f * (float)P.W;
  const float ddely_dy = 0.5f * (float)P.H;
  float* const tS = &sm.sw[warp][0][0][0];
  float* const tW = &sm.sw[warp][1][0][0];
  const unsigned char* const wlist = sm.widx[warp];

  // The open tensor-core group survives the end of a batch: its S / W rows stay in the per-warp tiles, and the
  // per-row metadata the flush needs (centre, conic, opacity, Gaussian id: 7 words) moves from the staging arrays --
  // which the next batch overwrites -- into the padding columns of those rows.  (Flushing the partial group at every
  // batch end cost 5.6% of the kernel's instructions: 2.5 batches per tile, half-empty groups.)
  int gslot = 0;              // slot of the next hit in the current group
  int carried = 0;            // rows [0, carried) of the group come from earlier batches (metadata in the padding)
  float* pS = tS + lane;      // this lane's column of the S tile, row gslot (the W tile is 16 rows further)
  for (int hi = tile_max; hi > 0; hi -= NT) {
    const int n_in = min(NT, hi);
    __syncthreads();
    float4 ea, eb, ec;
    unsigned bmask = 0u;
    if (tid < n_in) {
      const int e = range.x + hi - 1 - tid;
      // the forward's culling mask when it was handed on, else recomputed below (same value)
      const bool have = P.cull && block_masks != nullptr;
      bmask = have ? block_masks[e] : 0xffffffffu;
      if (bmask != 0u) {
        const int gid = point_list[e];
        const float2 p = xy[gid];
        const float4 co = conic_opacity[gid];
        const float thr = gs_power_threshold(co.w);
        if (P.cull && !have) bmask = gs_block_mask(p.x, p.y, co.x, co.y, co.z, thr, tile_x0, tile_y0);
        ea = make_float4(p.x, p.y, co.x, co.z);
        eb = make_float4(co.y, co.w, thr, __int_as_float(hi - 1 - tid));
        ec = make_float4(rgb[3 * gid + 0], rgb[3 * gid + 1], rgb[3 * gid + 2], __int_as_float(gid));
      }
    }
    int cnt;
    const int slot = compact_slot<NW>(bmask != 0u, lane, warp, sm.wcnt, cnt);
    GSB_DCHECK(bmask == 0u || (slot >= 0 && slot < NT && slot <= tid));
    GSB_DCHECK(cnt >= 0 && cnt <= NT);
    if (bmask != 0u) {
      sm.a[slot] = ea;
      sm.b[slot] = eb;
      sm.c[slot] = ec;
      sm.meta[slot] = make_int2(hi - 1 - tid, (int)bmask);
    }
    __syncthreads();
    // a warp whose pixels all stopped before this batch has nothing to do in it
    if (my_max <= hi - n_in) continue;
    // entries that touch this warp's block and lie inside its replay range (position < my_max)
    const int wn = warp_compact_hits(sm.meta, cnt, my_mask, my_max, lane, sm.widx[warp]);

    // One replayed hit whose alpha is known and >= 1/255 (backward.py:655-683): returns s, w.
    auto replay_hit = [&](const float G, const float alpha, const int j, float& sv, float& wv) {
      const float4 c = sm.c[j];
      const float inv_1ma = rcp_approx(1.0f - alpha);     // backward.py:658,680; 1 - alpha is in [0.01, 1]
      T = T * inv_1ma;                                   // T_i
      wv = alpha * T;                                    // d(channel)/d(colour), backward.py:672
      const float dc = gs_dot3(c.x, c.y, c.z, dp0, dp1, dp2);
      const float dL_dalpha = T * dc - gamma * inv_1ma;
      gamma = fmaf(dc, wv, gamma);
      sv = G * dL_dalpha;      // dL_dG = opacity * dL_dalpha is applied after the reduction
    };

    // Group of `count` hits complete (rows 0 .. count-1 of the tiles, list entries qbase ..): reduce
    // them over the 32 pixels on the tensor cores and add the results to the gradient arrays.
    // qbase: list entry of the group's first row that was recorded in THIS batch (row `carried`)
    auto flush_group = [&](const int qbase, const int count) {
      float4 m03, col;
      float2 m45;
      bwd_group_moments(tS, tW, lane, fg, ft, m0, m1, m2, dp_row, dp_cols, m03, m45, col);
      // (+ 1: a group closed by the all-zero padding row has one row without a list entry; its sums are zero)
      GSB_DCHECK(count >= 1 && count <= kGrp && carried <= count && qbase >= 0 && qbase + count - carried <= wn + 1);
      if (lane < count && bwd_group_nonzero(m03, m45, col)) {
        float4 ga, gb;
        int gid;
        if (lane < carried) {
          ga = *reinterpret_cast<const float4*>(tS + lane * kSRow + 32);
          gb = *reinterpret_cast<const float4*>(tW + lane * kSRow + 32);   // conic.b, opacity, gid bits, -
          gid = __float_as_int(gb.z);
        } else {
          const int je = wlist[qbase + lane - carried];
          ga = sm.a[je];
          gb = sm.b[je];
          gid = __float_as_int(sm.c[je].w);
        }
        float g[5];
        bwd_group_gradients(ga, gb, (float)bx0, (float)by0, m03, m45, ddelx_dx, ddely_dy, g);
        if (PACKED) {
          float* const rec = acc_packed + 12 * (size_t)gid;
          red_add_v4(rec, g[0], g[1], g[2], m03.x);
          red_add_v4(rec + 4, col.x, col.y, col.z, g[3]);
          atomicAdd(rec + 8, g[4]);
        } else {
          atomicAdd(dL_dcolor + 3 * (size_t)gid + 0, col.x);
          atomicAdd(dL_dcolor + 3 * (size_t)gid + 1, col.y);
          atomicAdd(dL_dcolor + 3 * (size_t)gid + 2, col.z);
          atomicAdd(dL_dmean2D + 3 * (size_t)gid + 0, g[3]);
          atomicAdd(dL_dmean2D + 3 * (size_t)gid + 1, g[4]);
          atomicAdd(dL_dconic + 4 * (size_t)gid + 0, g[0]);
          atomicAdd(dL_dconic + 4 * (size_t)gid + 1, g[1]);
          atomicAdd(dL_dconic + 4 * (size_t)gid + 3, g[2]);
          atomicAdd(dL_dopacity + gid, m03.x);   // backward.py:706
        }
      }
      __syncwarp();  // results consumed before the next group overwrites the tile
      gslot = 0;
      carried = 0;
      pS = tS + lane;
    };

    // Two hits per iteration (the list is already in replay order): exponents share the packed
    // FFMA2 sequence, the independent parts of the two hits overlap, the T / gamma recurrences stay
    // sequential.  kGrp is even, so a pair never straddles a group.
    int q = 0;
    for (; q + 1 < wn; q += 2) {
      const int jA = wlist[q], jB = wlist[q + 1];
      GSB_DCHECK(jA < cnt && jB < cnt && jA < jB && gslot + 2 <= kGrp);
      const float4 aA = sm.a[jA], bA = sm.b[jA];
      const float4 aB = sm.a[jB], bB = sm.b[jB];
      const float pwA = gs_power_packed(gs_pack2(aA.x, aA.y), npxy, gs_pack2(aA.z, aA.w), bA.x);
      const float pwB = gs_power_packed(gs_pack2(aB.x, aB.y), npxy, gs_pack2(aB.z, aB.w), bB.x);
      // the replay limit (pixel replays the entry iff position < kept), backward.py:647 (power > 0)
      // and the conservative exponent threshold
      const bool actA = __float_as_int(bA.w) < kept && !(pwA > 0.0f) && !(pwA < bA.z);
      const bool actB = __float_as_int(bB.w) < kept && !(pwB > 0.0f) && !(pwB < bB.z);
      float svA = 0.0f, wvA = 0.0f, svB = 0.0f, wvB = 0.0f;
      if (actA || actB) {
        float GA, GB, alphaA, alphaB;
        const bool hitA = bwd_alpha_hit(pwA, bA.y, GA, alphaA), hitB = bwd_alpha_hit(pwB, bB.y, GB, alphaB);
        if (actA && hitA) replay_hit(GA, alphaA, jA, svA, wvA);  // backward.py:655
        if (actB && hitB) replay_hit(GB, alphaB, jB, svB, wvB);
      }
      pS[0] = svA;
      pS[kGrp * kSRow] = wvA;
      pS[kSRow] = svB;
      pS[kGrp * kSRow + kSRow] = wvB;
      gslot += 2;
      pS += 2 * kSRow;
      if (gslot == kGrp) flush_group(q + 2 - (kGrp - carried), kGrp);
    }
    if (q < wn) {  // odd tail
      const int j = wlist[q];
      const float4 a = sm.a[j], b4 = sm.b[j];
      const float power = gs_power_packed(gs_pack2(a.x, a.y), npxy, gs_pack2(a.z, a.w), b4.x);
      float sv = 0.0f, wv = 0.0f;
      if (__float_as_int(b4.w) < kept && !(power > 0.0f) && !(power < b4.z)) {
        float G, alpha;
        if (bwd_alpha_hit(power, b4.y, G, alpha)) replay_hit(G, alpha, j, sv, wv);
      }
      pS[0] = sv;
      pS[kGrp * kSRow] = wv;
      ++gslot;
      pS += kSRow;
      ++q;
    }
    // q == wn here: the rows [carried, gslot) were recorded in this batch, from the list entries [wn - (gslot - carried), wn)
    if (hi <= NT) {                       // the tile's last batch: what is open is flushed
      if (gslot > 0) flush_group(wn - (gslot - carried), gslot);
    } else if (gslot > carried) {         // rows recorded in this batch stay open
      const int fresh = gslot - carried;  // their number (before the padding row)
      if (gslot & 1) {                    // pairs must not straddle a group: an all-zero row (the flush skips it)
        pS[0] = 0.0f;
        pS[kGrp * kSRow] = 0.0f;
        ++gslot;
        pS += kSRow;
      }
      if (gslot == kGrp) {
        flush_group(wn - fresh, kGrp);    // (the padding row filled the group)
      } else {
        if (lane >= carried && lane < carried + fresh) {
          const int je = wlist[wn - fresh + lane - carried];
          const float4 b4 = sm.b[je];
          *reinterpret_cast<float4*>(tS + lane * kSRow + 32) = sm.a[je];
          *reinterpret_cast<float4*>(tW + lane * kSRow + 32) = make_float4(b4.x, b4.y, sm.c[je].w, 0.0f);
        }
        __syncwarp();
        carried = gslot;
      }
    }
  }
}

// Stage-level entry point only: the packed records written out in the reference's four layouts (inside
// gsb_backward the per-Gaussian pass does this on the fly).
__global__ void __launch_bounds__(256)
unpack_records_kernel(int n, const float4* __restrict__ acc, float* __restrict__ dL_dmean2D, float4* __restrict__ dL_dconic,
                      float* __restrict__ dL_dopacity, float* __restrict__ dL_dcolor) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 q0 = acc[3 * (size_t)i], q1 = acc[3 * (size_t)i + 1];
  const float my = reinterpret_cast<const float*>(acc + 3 * (size_t)i + 2)[0];
  dL_dconic[i] = make_float4(q0.x, q0.y, 0.0f, q0.z);
  dL_dopacity[i] = q0.w;
  dL_dcolor[3 * (size_t)i + 0] = q1.x;
  dL_dcolor[3 * (size_t)i + 1] = q1.y;
  dL_dcolor[3 * (size_t)i + 2] = q1.z;
  dL_dmean2D[3 * (size_t)i + 0] = q1.w;
  dL_dmean2D[3 * (size_t)i + 1] = my;
  dL_dmean2D[3 * (size_t)i + 2] = 0.0f;
}

// The four accumulation targets of the tile kernel zeroed by ONE launch (four memsets were four stream
// operations of ~2 us each in front of a 0.3 ms kernel).  16-byte stores over each array's aligned body.
struct ZeroJob {
  float* p[4];
  long long n[4];  // floats
};
__global__ void __launch_bounds__(256) zero_arrays_kernel(const ZeroJob job) {
  gsb_pdl_wait();
  gsb_pdl_launch_dependents();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    float* const p = job.p[a];
    const long long n = job.n[a];
    if ((reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
      float4* const p4 = reinterpret_cast<float4*>(p);
      const long long n4 = n >> 2;
      for (long long i = t; i < n4; i += stride) p4[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (t < (n & 3)) p[4 * n4 + t] = 0.0f;
    } else {
      for (long long i = t; i < n; i += stride) p[i] = 0.0f;
    }
  }
}

// Diagnostic: the reference's per-pixel loops (forward.py:454-501, backward.py:633-706) walked by one thread per
// pixel with the contract's arithmetic, counting work and decisions:
//   [0] K_fwd   (pixel, Gaussian) pairs the forward loop iterates (until the pixel breaks or its list ends)
//   [1] pairs that blend (alpha >= 1/255, before the break)      [2] K_bwd = sum of min(list, n_contrib)
//   [3] backward pairs whose exponent passes power <= 0 and the conservative threshold (alpha gets evaluated)
//   [4] of those: the backward kernel's test (MUFU alpha < 1/255) disagrees with the forward's decision
//   [5] pairs the conservative exponent threshold skips although the forward blended them: must be 0
//   [6] evaluated pairs whose MUFU alpha lies within 1e-7 of 1/255
__global__ void __launch_bounds__(256)
work_counters_kernel(const BlendParams P, const int2* __restrict__ ranges, const int* __restrict__ point_list,
                     const float2* __restrict__ xy, const float4* __restrict__ conic_opacity,
                     const int* __restrict__ n_contrib, unsigned long long* __restrict__ counters) {
  const int px = blockIdx.x * 16 + (threadIdx.x & 15), py = blockIdx.y * 16 + (threadIdx.x >> 4);
  unsigned long long c[7] = {0, 0, 0, 0, 0, 0, 0};
  if (px < P.W && py < P.H) {
    const int2 range = ranges[blockIdx.y * P.grid_x + blockIdx.x];
    const gs_f2 npxy = gs_pack2(-(float)px, -(float)py);
    const int kept = min(range.y - range.x, n_contrib[(size_t)py * P.W + px]);
    c[2] = (unsigned long long)kept;
    float T = 1.0f;
    for (int e = range.x; e < range.y; ++e) {
      const int gid = point_list[e];
      const float2 p = xy[gid];
      const float4 co = conic_opacity[gid];
      c[0] += 1;
      const float power = gs_power_packed(gs_pack2(p.x, p.y), npxy, gs_pack2(co.x, co.z), co.y);
      const bool fwd_hit = !(power > 0.0f) && !(f_min(0.99f, __fmul_rn(co.w, gs_expf(power))) < (1.0f / 255.0f));
      if (e - range.x < kept && !(power > 0.0f)) {
        if (!(power < gs_power_threshold(co.w))) {
          c[3] += 1;
          float G, alpha;
          c[4] += bwd_alpha_hit(power, co.w, G, alpha) != fwd_hit;
          c[6] += fabsf(alpha - (1.0f / 255.0f)) < 1e-7f;
        } else {
          c[5] += fwd_hit;     // the conservative exponent threshold said "skip": the forward must not have blended it
        }
      }
      if (!fwd_hit) continue;
      const float alpha = f_min(0.99f, __fmul_rn(co.w, gs_expf(power)));
      const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
      if (test_T < 0.0001f) break;
      c[1] += 1;
      T = test_T;
    }
  }
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    unsigned long long v = c[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(counters + k, v);
  }
}

}  // namespace

GSB_API int gsb_selftest_work_counters(gsb_ctx* ctx, gsb_stream s_, const gsb_frame* f, const int32_t* ranges,
                                       const int32_t* point_list, const float* points_xy, const float* conic_opacity,
                                       const int32_t* n_contrib, uint64_t* counters7) {
  if (!ctx) return GSB_ERR_INVALID;
  GSB_REQUIRE(ctx, f && f->width > 0 && f->height > 0 && counters7, "gsb_selftest_work_counters: bad arguments");
  cudaStream_t s = (cudaStream_t)s_;
  GSB_CUDA(ctx, cudaMemsetAsync(counters7, 0, 7 * sizeof(uint64_t), s));
  BlendParams P = make_blend_params(ctx, f);
  dim3 grid(P.grid_x, (f->height + kTile - 1) / kTile);
  GSB_LAUNCH(ctx, work_counters_kernel, grid, 256, 0, s, P, reinterpret_cast<const int2*>(ranges), point_list,
             reinterpret_cast<const float2*>(points_xy), reinterpret_cast<const float4*>(conic_opacity), n_contrib,
             reinterpret_cast<unsigned long long*>(counters7));
  return GSB_OK;
}

bool gsb_blend_backward_uses_packed(const gsb_ctx* ctx) {
  return ctx->opt.bwd_packed != 0 && ctx->opt.bwd_reduce != 0;
}

// The tensor-core kernel, either accumulating into the caller's four arrays (packed == nullptr; they must be
// zero) or into packed records (which must be zero).
static int launch_backward_mma(gsb_ctx* ctx, cudaStream_t s, const BlendParams& P, dim3 grid, const int32_t* ranges,
                               const int32_t* point_list, const float* points_xy, const float* conic_opacity,
                               const float* rgb, const float* final_T, const int32_t* n_contrib,
                               const float* dL_dpixels, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                               float* dL_dcolor, const unsigned* masks, float* packed) {
  bool& attr_set = ctx->smem_optin_blend_bwd;  // > 48 KB of dynamic shared memory needs the opt-in (per device)
  if (!attr_set) {
    GSB_CUDA(ctx, cudaFuncSetAttribute(blend_backward_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdSmem)));
    GSB_CUDA(ctx, cudaFuncSetAttribute(blend_backward_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdSmem)));
    attr_set = true;
  }
#define GSB_BWD_MMA(PK)                                                                                                    \
  GSB_LAUNCH_PDL(ctx, (blend_backward_mma_kernel<PK>), grid, 256, sizeof(BwdSmem), s, P,                                        \
             reinterpret_cast<const int2*>(ranges), point_list, reinterpret_cast<const float2*>(points_xy),               \
             reinterpret_cast<const float4*>(conic_opacity), rgb, final_T, n_contrib, dL_dpixels, dL_dmean2D, dL_dconic,  \
             dL_dopacity, dL_dcolor, masks, packed)
  if (packed) GSB_BWD_MMA(true);
  else GSB_BWD_MMA(false);
#undef GSB_BWD_MMA
  return GSB_OK;
}

int gsb_blend_backward_packed(gsb_ctx* ctx, cudaStream_t s, const gsb_frame* f, int32_t n, const int32_t* ranges,
                              const int32_t* point_list, const float* points_xy, const float* conic_opacity,
                              const float* rgb, const float* final_T, const int32_t* n_contrib,
                              const float* dL_dpixels, float* packed, const int32_t* block_masks) {
  GSB_REQUIRE(ctx, f && f->width > 0 && f->height > 0 && n >= 0, "gsb_blend_backward: bad frame");
  GSB_REQUIRE(ctx, gsb_aligned16(conic_opacity) && gsb_aligned16(packed), "gsb_blend_backward: 16-byte alignment");
  if (n == 0) return GSB_OK;
  ZeroJob job;
  for (int a = 0; a < 4; ++a) job.p[a] = packed, job.n[a] = 0;
  job.n[0] = 12LL * n;
  const long long blocks = gsb_div_up(3LL * n, 256);  // one 16-byte store per thread
  GSB_LAUNCH_PDL(ctx, zero_arrays_kernel, (unsigned)(blocks < 8192 ? blocks : 8192), 256, 0, s, job);
  BlendParams P = make_blend_params(ctx, f);
  dim3 grid(P.grid_x, (f->height + kTile - 1) / kTile);
  return launch_backward_mma(ctx, s, P, grid, ranges, point_list, points_xy, conic_opacity, rgb, final_T, n_contrib,
                             dL_dpixels, nullptr, nullptr, nullptr, nullptr, reinterpret_cast<const unsigned*>(block_masks),
                             packed);
}

GSB_API int gsb_blend_backward(gsb_ctx* ctx, gsb_stream s_, const gsb_frame* f, int32_t n, const int32_t* ranges,
                               const int32_t* point_list, const float* points_xy, const float* conic_opacity,
                               const float* rgb, const float* final_T, const int32_t* n_contrib,
                               const float* dL_dpixels, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                               float* dL_dcolor, const int32_t* block_masks) {
  if (!ctx) return GSB_ERR_INVALID;
  GSB_REQUIRE(ctx, f && f->width > 0 && f->height > 0 && n >= 0, "gsb_blend_backward: bad frame");
  GSB_REQUIRE(ctx, gsb_aligned16(conic_opacity), "gsb_blend_backward: conic_opacity must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)s_;
  if (n == 0) return GSB_OK;
  if (gsb_blend_backward_uses_packed(ctx) && gsb_aligned16(dL_dconic)) {
    // stage-level call of the default path: packed records in context scratch, then written out
    if (ctx->bwd_acc_cap < 12LL * n) {
      int rc = gsb_grow(ctx, (void**)&ctx->bwd_acc_stage, &ctx->bwd_acc_cap, 12LL * n, sizeof(float), s);
      if (rc != GSB_OK) return rc;
    }
    int rc = gsb_blend_backward_packed(ctx, s, f, n, ranges, point_list, points_xy, conic_opacity, rgb, final_T, n_contrib,
                                       dL_dpixels, ctx->bwd_acc_stage, block_masks);
    if (rc != GSB_OK) return rc;
    GSB_LAUNCH(ctx, unpack_records_kernel, (int)gsb_div_up(n, 256), 256, 0, s, n,
               reinterpret_cast<const float4*>(ctx->bwd_acc_stage), dL_dmean2D, reinterpret_cast<float4*>(dL_dconic),
               dL_dopacity, dL_dcolor);
    return GSB_OK;
  }
  {
    ZeroJob job;
    job.p[0] = dL_dmean2D, job.n[0] = 3LL * n;
    job.p[1] = dL_dconic, job.n[1] = 4LL * n;
    job.p[2] = dL_dopacity, job.n[2] = (long long)n;
    job.p[3] = dL_dcolor, job.n[3] = 3LL * n;
    const long long blocks = gsb_div_up(n, 256);  // one 16-byte store per thread and array
    GSB_LAUNCH_PDL(ctx, zero_arrays_kernel, (unsigned)(blocks < 4096 ? blocks : 4096), 256, 0, s, job);
  }
  BlendParams P = make_blend_params(ctx, f);
  dim3 grid(P.grid_x, (f->height + kTile - 1) / kTile);
  const unsigned* masks = reinterpret_cast<const unsigned*>(block_masks);
  if (ctx->opt.bwd_reduce == 0) {  // A/B: warp-shuffle butterfly reduction
    GSB_LAUNCH(ctx, blend_backward_kernel, grid, 256, 0, s, P, reinterpret_cast<const int2*>(ranges), point_list,
               reinterpret_cast<const float2*>(points_xy), reinterpret_cast<const float4*>(conic_opacity), rgb, final_T,
               n_contrib, dL_dpixels, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, masks);
    return GSB_OK;
  }
  return launch_backward_mma(ctx, s, P, grid, ranges, point_list, points_xy, conic_opacity, rgb, final_T, n_contrib,
                             dL_dpixels, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, masks, nullptr);
}
