// blend.cu -- the per-16x16-tile alpha blend.  Its back-to-front backward lives in blend_bwd.cu.
//   blend_forward_kernel   replaces wp_render_gaussians       (forward.py:384-515)
//   blend_backward_* (blend_bwd.cu) replace wp_render_backward_kernel (backward.py:558-706)
//
// Design (forward and backward): one CTA per tile, 256 threads, one pixel per thread; a warp owns a 4x8
// pixel block (blend_common.cuh), so per-pixel loads/stores are 32-byte row segments (the
// reference's launch maps adjacent lanes to a pixel COLUMN).  The tile's depth-sorted Gaussians
// are staged through shared memory in batches of 256 (one gather per Gaussian per tile instead of
// one per pixel), then broadcast to the threads with three 16-byte LDS per Gaussian.
//
// Culling (both kernels).  The reference bins a Gaussian into every tile of the bounding SQUARE of
// its 3-sigma circle, but a pixel only ever uses it when alpha >= 1/255, i.e. inside the ellipse
// power >= -log(255*opacity).  While staging a Gaussian the loading thread intersects that ellipse
// (with a conservative margin) with each of the tile's 16 pixel rows, split in two 8-pixel halves,
// and keeps a 32-bit mask:
//   * Gaussians whose mask is empty are dropped from the batch by a stable compaction (on the
//     headline scene 31% of all list entries), their list position is kept for n_contrib;
//   * a warp skips a staged Gaussian with one LDS + one test unless the mask touches its 4x8 block
//     (only about a fifth of the (Gaussian, block) combinations can contribute at all);
//   * lanes that remain are filtered by the same threshold on the exponent before the exponential.
// Pairs that survive run the exact arithmetic of the contract, so n_contrib / final_T are
// bit-identical to the oracle: culling only ever removes pairs the reference would `continue` on.
// Forward: a CTA stops as soon as every one of its pixels has terminated (__syncthreads_and).
//
// Backward (blend_bwd.cu): the reference issues 11 scalar global atomics per (pixel, Gaussian) pair.
// There a warp sums its pixels' terms for 16 Gaussians at a time on the tensor cores and adds the nine
// results per Gaussian to one packed record with two vector REDs and a scalar one -- and only for
// Gaussians that touched the warp at all.
#include "blend_common.cuh"


namespace {

#ifndef GSB_FWD_MINB
#define GSB_FWD_MINB 5
#endif

// Staging is dynamic and double-buffered.  Round 1 staged every batch of 256 list entries with all eight warps
// between CTA barriers (three per batch), so the warp with the most hits in its block decided when the CTA moved on:
// at the barrier on top of the batch loop alone the warps spent 13% of their time (ncu source page,
// profiles/r02_ncu_step.md), and the slowest warp also staged its 32 entries like everybody else.  Here
//   * the staging arrays are double-buffered; a warp that has finished its hits of batch k at once takes CHUNKS
//     (32 entries) of batch k+1 from a shared counter and stages them into the other buffer -- so the warps with few
//     hits stage for the ones with many, and the slowest warp of a batch does no staging at all;
//   * a chunk is compacted inside the staging warp (ballot) to the slots [32 c, 32 c + count_c): no CTA-wide prefix,
//     no barrier inside the staging;
//   * ONE barrier per batch remains (__syncthreads_and(done): batch k staged by everybody, batch k-1 blended by
//     everybody, and the early exit when all 256 pixels are finished).
// The entries reach a pixel in list order (chunks in order, slots in order).  Measured against the round-1 kernel on
// the headline scene: 217.1 against 221.2 us, every output bit-identical (profiles/r02_experiments.md).
struct FwdSmem {
  float4 a[2][256];
  float4 b[2][256];
  float4 c[2][256];
  int2 meta[2][256];
  int ccnt[2][8];   // entries kept per chunk
  int ctr[2];       // next chunk of the batch being staged into this buffer
  unsigned char widx[8][256];
};

__global__ void __launch_bounds__(256, GSB_FWD_MINB)
blend_forward_kernel(const BlendParams P, const int2* __restrict__ ranges, const int* __restrict__ point_list,
                         const float2* __restrict__ xy, const float* __restrict__ rgb,
                         const float4* __restrict__ conic_opacity, const float* __restrict__ depths,
                         float* __restrict__ image, float* __restrict__ inv_depth, float* __restrict__ final_T,
                         int* __restrict__ n_contrib, unsigned* __restrict__ block_masks, const int* __restrict__ go) {
  constexpr int NT = 256;
  gsb_pdl_wait();
  gsb_pdl_launch_dependents();
  if (go && *go == 0) return;   // queued speculatively and the frame does not fit (tilesort.cu, tile_scan_kernel)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FwdSmem& sm = *reinterpret_cast<FwdSmem*>(smem_raw);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile_x = blockIdx.x, tile_y = blockIdx.y;
  const int tile_id = tile_y * P.grid_x + tile_x;
  const int px = tile_x * kTile + (warp & 1) * 8 + (lane & 7);
  const int py = tile_y * kTile + (warp >> 1) * 4 + (lane >> 3);
  const float pxf = (float)px;
  const float pyf = (float)py;
  const unsigned my_mask = gs_warp_mask(warp);
  const float tile_x0 = (float)(tile_x * kTile), tile_y0 = (float)(tile_y * kTile);
  const bool inside = (px < P.W && py < P.H);

  bool done = !inside;
  gs_f2 npxy = gs_pack2(done ? __int_as_float(0x7fc00000) : -pxf, -pyf);  // (-px, -py), NaN once finished
  float T = 1.0f, C0 = 0.0f, C1 = 0.0f, C2 = 0.0f, Dp = 0.0f;
  int last = 0;

  const int2 range = ranges[tile_id];
  const int todo = range.y - range.x;
  const int nbatch = (todo + NT - 1) / NT;

  // chunk `chunk` (32 entries) of batch `batch`, staged by this warp into buffer batch & 1
  auto stage_chunk = [&](const int batch, const int chunk) {
    const int buf = batch & 1;
    const int ei = batch * NT + chunk * 32 + lane;   // position in the tile's list
    float4 ea, eb, ec;
    unsigned bmask = 0u;
    if (ei < todo) {
      const int gid = point_list[range.x + ei];
      const float2 p = xy[gid];
      const float4 co = conic_opacity[gid];
      const float thr = gs_power_threshold(co.w);
      bmask = P.cull ? gs_block_mask(p.x, p.y, co.x, co.y, co.z, thr, tile_x0, tile_y0) : 0xffffffffu;
      if (block_masks) block_masks[range.x + ei] = bmask;  // handed on to the backward
      ea = make_float4(p.x, p.y, co.x, co.z);
      eb = make_float4(co.y, co.w, thr, 1.0f / depths[gid]);
      ec = make_float4(rgb[3 * gid + 0], rgb[3 * gid + 1], rgb[3 * gid + 2], __int_as_float(ei + 1));
    }
    const unsigned keep = __ballot_sync(0xffffffffu, bmask != 0u);
    if (bmask != 0u) {
      const int slot = chunk * 32 + __popc(keep & ((1u << lane) - 1u));
      sm.a[buf][slot] = ea;
      sm.b[buf][slot] = eb;
      sm.c[buf][slot] = ec;
      sm.meta[buf][slot] = make_int2(ei + 1, (int)bmask);
    }
    if (lane == 0) sm.ccnt[buf][chunk] = __popc(keep);
  };

  if (tid < 2) sm.ctr[tid] = 0;
  if (nbatch > 0 && warp * 32 < min(NT, todo)) stage_chunk(0, warp);   // the first batch: one chunk per warp

  const unsigned char* const wlist = sm.widx[warp];
  for (int k = 0; k < nbatch; ++k) {
    // batch k is staged (every warp staged its chunks before it arrived here) and nobody reads buffer (k+1) & 1 any
    // more (every warp has finished batch k-1); all pixels finished => the tile is done
    if (__syncthreads_and(done)) break;
    const int buf = k & 1;
    if (tid == 0) sm.ctr[buf] = 0;   // this buffer's counter is used again for batch k+2, after the next barrier
    const int nchunk = (min(NT, todo - k * NT) + 31) >> 5;
    if (!__all_sync(0xffffffffu, done)) {
      // the staged entries of the batch that touch this warp's block, in list order
      int wn = 0;
      for (int c = 0; c < nchunk; ++c) {
        const int cc = sm.ccnt[buf][c];
        bool hit = false;
        if (lane < cc) hit = (((unsigned)sm.meta[buf][c * 32 + lane].y & my_mask) != 0u);
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (hit) sm.widx[warp][wn + __popc(bal & ((1u << lane) - 1u))] = (unsigned char)(c * 32 + lane);
        wn += __popc(bal);
      }
      __syncwarp();
      const float4* const sa = sm.a[buf];
      const float4* const sb = sm.b[buf];
      const float4* const sc = sm.c[buf];
      // One hit: alpha is known and >= 1/255.  forward.py:483-499.
      auto blend_hit = [&](const float alpha, const int j, const float inv_z) {
        const float test_T = T * (1.0f - alpha);
        if (test_T < 0.0001f) {            // forward.py:487: the breaking Gaussian is not counted
          done = true;
          npxy = gs_pack2(__int_as_float(0x7fc00000), -pyf);
          return;
        }
        const float4 c = sc[j];
        const float w = alpha * T;
        C0 = __fmaf_rn(c.x, w, C0);
        C1 = __fmaf_rn(c.y, w, C1);
        C2 = __fmaf_rn(c.z, w, C2);
        Dp = __fmaf_rn(inv_z, w, Dp);
        T = test_T;
        last = __float_as_int(c.w);
      };
      int q = 0;
      for (; q + 1 < wn; q += 2) {
        if ((q & 3) == 0 && __all_sync(0xffffffffu, done)) break;  // every pixel of the block is finished
        const int jA = wlist[q], jB = wlist[q + 1];
        GSB_DCHECK(jA < jB && jB < NT);
        const float4 aA = sa[jA], bA = sb[jA];
        const float4 aB = sa[jB], bB = sb[jB];
        const float pwA = gs_power_packed(gs_pack2(aA.x, aA.y), npxy, gs_pack2(aA.z, aA.w), bA.x);
        const float pwB = gs_power_packed(gs_pack2(aB.x, aB.y), npxy, gs_pack2(aB.z, aB.w), bB.x);
        const bool okA = (pwA <= 0.0f) && (pwA >= bA.z);
        const bool okB = (pwB <= 0.0f) && (pwB >= bB.z);
        if (!(okA || okB)) continue;
        float GA, GB;
        gs_expf2(pwA, pwB, GA, GB);
        const float alphaA = f_min(0.99f, bA.y * GA);
        const float alphaB = f_min(0.99f, bB.y * GB);
        if (okA && !(alphaA < (1.0f / 255.0f))) blend_hit(alphaA, jA, bA.w);
        if (okB && !done && !(alphaB < (1.0f / 255.0f))) blend_hit(alphaB, jB, bB.w);
      }
      if (q < wn && q + 1 >= wn) {  // odd tail (not reached after an early exit: then q + 1 < wn)
        const int j = wlist[q];
        const float4 a = sa[j], b4 = sb[j];
        const float power = gs_power_packed(gs_pack2(a.x, a.y), npxy, gs_pack2(a.z, a.w), b4.x);
        if ((power <= 0.0f) && (power >= b4.z)) {
          const float alpha = f_min(0.99f, b4.y * gs_expf(power));
          if (!(alpha < (1.0f / 255.0f))) blend_hit(alpha, j, b4.w);
        }
      }
    }
    // help staging the next batch: take chunks until none is left
    if (k + 1 < nbatch) {
      const int nnext = (min(NT, todo - (k + 1) * NT) + 31) >> 5;
      for (;;) {
        int c = 0;
        if (lane == 0) c = atomicAdd(&sm.ctr[buf ^ 1], 1);
        c = __shfl_sync(0xffffffffu, c, 0);
        if (c >= nnext) break;
        stage_chunk(k + 1, c);
      }
    }
  }

  if (inside) {
    const size_t pix = (size_t)py * P.W + px;
    final_T[pix] = T;
    n_contrib[pix] = last;
    image[3 * pix + 0] = C0 + T * P.bg0;
    image[3 * pix + 1] = C1 + T * P.bg1;
    image[3 * pix + 2] = C2 + T * P.bg2;
    inv_depth[pix] = Dp;
  }
}

}  // namespace

// Diagnostic: for each test case (one Gaussian against one tile) the culling mask of gs_block_mask and
// the exact set of pixels the forward would blend it into if nothing else had been drawn (power <= 0 and
// alpha >= 1/255 with the contract's arithmetic).  The mask must cover that set: tests/test_gpu_parity.py.
namespace {
__global__ void selftest_block_mask_kernel(int n, const float* __restrict__ g /* n x 8: gx gy a b c opacity x0 y0 */,
                                           unsigned* __restrict__ mask, unsigned* __restrict__ active /* n x 8 */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gx = g[8 * i + 0], gy = g[8 * i + 1], ca = g[8 * i + 2], cb = g[8 * i + 3], cc = g[8 * i + 4];
  const float o = g[8 * i + 5], x0 = g[8 * i + 6], y0 = g[8 * i + 7];
  mask[i] = gs_block_mask(gx, gy, ca, cb, cc, gs_power_threshold(o), x0, y0);
  for (int w = 0; w < 8; ++w) {
    unsigned bits = 0u;
    for (int b = 0; b < 32; ++b) {
      const int p = 32 * w + b, r = p >> 4, c = p & 15;
      const float dx = gx - (x0 + (float)c), dy = gy - (y0 + (float)r);
      const float power = gs_power(ca, cb, cc, dx, dy);
      if (power > 0.0f) continue;
      const float alpha = f_min(0.99f, o * gs_expf(power));
      if (alpha < (1.0f / 255.0f)) continue;
      bits |= 1u << b;   // NaN power / alpha fall through to here, like in the tile kernels
    }
    active[8 * i + w] = bits;
  }
}
}  // namespace

GSB_API int gsb_selftest_block_mask(gsb_ctx* ctx, gsb_stream s, int32_t count, const float* cases, uint32_t* mask,
                                    uint32_t* active) {
  if (!ctx) return GSB_ERR_INVALID;
  if (count <= 0) return GSB_OK;
  GSB_LAUNCH(ctx, selftest_block_mask_kernel, (count + 127) / 128, 128, 0, (cudaStream_t)s, count, cases, mask, active);
  return GSB_OK;
}

int gsb_blend_forward_impl(gsb_ctx* ctx, gsb_stream s_, const gsb_frame* f, const int32_t* ranges,
                           const int32_t* point_list, const float* points_xy, const float* rgb,
                           const float* conic_opacity, const float* depths, float* image, float* inv_depth,
                           float* final_T, int32_t* n_contrib, int32_t* block_masks, const int* go);

GSB_API int gsb_blend_forward(gsb_ctx* ctx, gsb_stream s_, const gsb_frame* f, const int32_t* ranges,
                              const int32_t* point_list, const float* points_xy, const float* rgb,
                              const float* conic_opacity, const float* depths, float* image, float* inv_depth,
                              float* final_T, int32_t* n_contrib, int32_t* block_masks) {
  return gsb_blend_forward_impl(ctx, s_, f, ranges, point_list, points_xy, rgb, conic_opacity, depths, image, inv_depth,
                                final_T, n_contrib, block_masks, nullptr);
}

// go: device flag checked by every CTA (gsb_forward's speculative launch), or nullptr
int gsb_blend_forward_impl(gsb_ctx* ctx, gsb_stream s_, const gsb_frame* f, const int32_t* ranges,
                           const int32_t* point_list, const float* points_xy, const float* rgb,
                           const float* conic_opacity, const float* depths, float* image, float* inv_depth,
                           float* final_T, int32_t* n_contrib, int32_t* block_masks, const int* go) {
  if (!ctx) return GSB_ERR_INVALID;
  GSB_REQUIRE(ctx, f && f->width > 0 && f->height > 0, "gsb_blend_forward: bad frame");
  GSB_REQUIRE(ctx, gsb_aligned16(conic_opacity), "gsb_blend_forward: conic_opacity must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)s_;
  BlendParams P = make_blend_params(ctx, f);
  dim3 grid(P.grid_x, (f->height + kTile - 1) / kTile);
  GSB_LAUNCH_PDL(ctx, blend_forward_kernel, grid, 256, sizeof(FwdSmem), s, P, reinterpret_cast<const int2*>(ranges),
             point_list, reinterpret_cast<const float2*>(points_xy), rgb, reinterpret_cast<const float4*>(conic_opacity),
             depths, image, inv_depth, final_T, n_contrib, reinterpret_cast<unsigned*>(block_masks), go);
  return GSB_OK;
}
