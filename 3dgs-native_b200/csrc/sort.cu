// sort.cu -- stable LSD radix sort of (int64 key, int32 value) pairs, 8 bits per pass.
// Replaces the pad-copy + wp.utils.radix_sort_pairs + copy-back of forward.py:791-824 (which on
// the reference's CUDA device is cub::DeviceRadixSort inside Warp's prebuilt library [Warp]).
//
// Only the live key bits are sorted: bits [0, 32 + ceil(log2(num_tiles))) for the tile|depth
// keys (44 bits = 6 passes at 800x800 instead of 8).
//
// Per pass three kernels, no inter-CTA spinning, nothing serial:
//   radix_count_kernel    every CTA counts the digits of its slice (warp-level peer masks from ballots, so a
//                         digit all lanes share costs one add) into table[digit][cta] -- digit-major, so that
//   radix_scan_kernel     one WARP per digit turns its row into exclusive prefixes with every load of the
//                         row in flight at once (the round-1 version let the last CTA of the counting
//                         kernel walk the table row by row: 34 of that pass's 81 us), and leaves the digit's
//                         total;
//   radix_scatter_kernel  every CTA scans the 256 totals itself (no fourth kernel), re-reads its slice in
//                         sub-tiles of 4096 pairs, ranks them stably (warp peer-mask ranking + per-warp
//                         digit counters), reorders the sub-tile through shared memory so that each digit's
//                         run leaves as contiguous stores, and advances its 256 running offsets.  512
//                         threads x 8 pairs, the values landing in shared memory by cp.async while the keys
//                         are ranked: 64 registers (round 1: 128 registers, 47 us).
// Algorithmic traffic: 32 B per pair per pass (key read twice, value once, pair written once).
#include "common.cuh"

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kItems = 8;
constexpr int kSub = kThreads * kItems;  // 4096 pairs per sub-tile
constexpr int kMaxBlocks = 2048;
constexpr int kCountThreads = 1024;  // a slice of ~5.6 K keys (headline size) is one batch of loads per thread

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// The lanes of the warp that hold the same 8-bit digit as this one (valid lanes only): eight ballots, one per
// digit bit.  __match_any_sync computes the same mask in one instruction, but MATCH goes through the MIO pipe at
// about one warp instruction per 64 cycles per SM: with it both kernels of a pass sat behind mio_throttle /
// short_scoreboard (31 and 18 stall cycles per issued instruction, ncu) at under 20% issue utilisation.
__device__ __forceinline__ unsigned warp_peers8(const unsigned d, const bool valid) {
  unsigned peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const bool bit = (d >> b) & 1u;
    const unsigned m = __ballot_sync(0xffffffffu, bit);
    peers &= bit ? m : ~m;
  }
  return peers;
}

__global__ void __launch_bounds__(kCountThreads, 2)
radix_count_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift, unsigned digit_mask, int64_t chunk,
                   uint32_t* __restrict__ table /*[256][nb]*/) {
  constexpr int NW = kCountThreads / 32;
  __shared__ uint32_t s_hist[NW][256];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < NW * 256; i += kCountThreads) (&s_hist[0][0])[i] = 0;
  __syncthreads();
  const int64_t begin = (int64_t)blockIdx.x * chunk;
  const int64_t end = min(n, begin + chunk);
  for (int64_t base = begin; base < end; base += kCountThreads * 8) {
    uint64_t k[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {   // eight independent loads in flight
      const int64_t idx = base + u * kCountThreads + tid;
      k[u] = idx < end ? __ldg(keys + idx) : 0ull;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const bool valid = base + u * kCountThreads + tid < end;
      const unsigned d = (unsigned)((k[u] >> shift) & digit_mask);
      const unsigned m = warp_peers8(d, valid);
      if (valid && lane == (__ffs(m) - 1)) s_hist[warp][d] += __popc(m);
      __syncwarp();
    }
  }
  __syncthreads();
  if (tid < 256) {
    uint32_t c = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) c += s_hist[w][tid];
    table[(size_t)tid * gridDim.x + blockIdx.x] = c;
  }
}

// One warp per digit: exclusive prefix of the digit's counts over the CTAs of the pass (a row of `nb` <= 2048
// entries, read with up to eight independent 128-byte loads in flight per lane), written to a second table, and
// the digit's total.
__global__ void __launch_bounds__(256)
radix_scan_kernel(const uint32_t* __restrict__ table, uint32_t* __restrict__ prefix, uint32_t* __restrict__ digit_total,
                  int nb) {
  const int lane = threadIdx.x & 31;
  const int d = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (d >= 256) return;
  const uint32_t* const row = table + (size_t)d * nb;
  uint32_t* const out = prefix + (size_t)d * nb;
  uint32_t carry = 0;
  for (int base = 0; base < nb; base += 256) {
    uint32_t v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base + 32 * u + lane;
      v[u] = i < nb ? __ldcg(row + i) : 0u;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      uint32_t inc = v[u];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      const int i = base + 32 * u + lane;
      if (i < nb) out[i] = carry + inc - v[u];
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
  }
  if (lane == 0) digit_total[d] = carry;
}

struct ScatterSmem {
  uint64_t keys[kSub];
  int32_t vals[kSub];
  int32_t vals_in[kSub];  // the sub-tile's values in input order, landed by cp.async while the keys are ranked
  uint32_t whist[kWarps][256];
  uint32_t off[256];     // running global offset of each digit for this CTA
  uint32_t dstart[256];  // start of each digit's run inside the reordered sub-tile
  uint32_t wtot[kWarps];
};

// exclusive scan over the 256 per-digit values held by threads 0..255 (one per thread); all kThreads threads call
__device__ __forceinline__ uint32_t block_exclusive_256(uint32_t v, int tid, int lane, int warp, uint32_t* s_wtot,
                                                        uint32_t& inclusive) {
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31 && warp < 8) s_wtot[warp] = inc;
  __syncthreads();
  uint32_t wb = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w)
    if (w < warp) wb += s_wtot[w];
  inclusive = wb + inc;
  return wb + inc - v;
}

__global__ void __launch_bounds__(kThreads, 2)
radix_scatter_kernel(const uint64_t* __restrict__ in_keys, const int32_t* __restrict__ in_vals,
                     uint64_t* __restrict__ out_keys, int32_t* __restrict__ out_vals, int64_t n, int shift,
                     unsigned digit_mask, int64_t chunk, const uint32_t* __restrict__ prefix /*[256][nb]*/,
                     const uint32_t* __restrict__ digit_total /*[256]*/) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ScatterSmem& sm = *reinterpret_cast<ScatterSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned lt = lanemask_lt();

  {  // this CTA's first output position of every digit: digits below + the same digit in the CTAs before
    uint32_t incl;
    const uint32_t tot = tid < 256 ? digit_total[tid] : 0u;
    const uint32_t before = tid < 256 ? prefix[(size_t)tid * gridDim.x + blockIdx.x] : 0u;   // both loads in flight
    const uint32_t base = block_exclusive_256(tot, tid, lane, warp, sm.wtot, incl);
    if (tid < 256) sm.off[tid] = base + before;
  }
  const int64_t begin = (int64_t)blockIdx.x * chunk;
  const int64_t end = min(n, begin + chunk);

  for (int64_t sub = begin; sub < end; sub += kSub) {
    const int sub_count = (int)min((int64_t)kSub, end - sub);
    for (int i = tid; i < kWarps * 256; i += kThreads) (&sm.whist[0][0])[i] = 0;
    __syncthreads();

    uint64_t key[kItems];
    uint32_t rank[kItems];
    const int64_t wbase = sub + warp * (32 * kItems) + lane;
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
      const int64_t idx = wbase + i * 32;
      key[i] = idx < end ? in_keys[idx] : ~0ull;
    }
    // the values go straight to shared memory (no registers, no wait until after the ranking)
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
      const int64_t idx = wbase + i * 32;
      if (idx < end) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(&sm.vals_in[warp * (32 * kItems) + i * 32 + lane]);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(in_vals + idx) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
      const bool valid = (wbase + i * 32) < end;
      const unsigned d = (unsigned)((key[i] >> shift) & digit_mask);
      const unsigned m = warp_peers8(d, valid);
      uint32_t old = 0;
      if (valid) {
        const int leader = __ffs(m) - 1;
        if (lane == leader) {
          old = sm.whist[warp][d];
          sm.whist[warp][d] = old + __popc(m);
        }
        old = __shfl_sync(m, old, leader);
        rank[i] = old + __popc(m & lt);
      }
      __syncwarp();
    }
    __syncthreads();

    // thread d: exclusive scan of digit d's counts across the warps; then scan over digits
    uint32_t cnt = 0;
    if (tid < 256) {
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        const uint32_t c = sm.whist[w][tid];
        sm.whist[w][tid] = cnt;
        cnt += c;
      }
    }
    {
      uint32_t incl;
      const uint32_t ex = block_exclusive_256(cnt, tid, lane, warp, sm.wtot, incl);
      if (tid < 256) sm.dstart[tid] = ex;
    }
    __syncthreads();

    asm volatile("cp.async.wait_group 0;" ::: "memory");   // this thread's own values have landed (it reads only those)
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
      const int64_t idx = wbase + i * 32;
      if (idx < end) {
        const unsigned d = (unsigned)((key[i] >> shift) & digit_mask);
        const uint32_t p = sm.dstart[d] + sm.whist[warp][d] + rank[i];
        GSB_DCHECK(p < (uint32_t)sub_count);
        sm.keys[p] = key[i];
        sm.vals[p] = sm.vals_in[warp * (32 * kItems) + i * 32 + lane];
      }
    }
    __syncthreads();

#pragma unroll 4
    for (int p = tid; p < sub_count; p += kThreads) {
      const uint64_t k = sm.keys[p];
      const unsigned d = (unsigned)((k >> shift) & digit_mask);
      const uint32_t g = sm.off[d] + ((uint32_t)p - sm.dstart[d]);
      GSB_DCHECK((int64_t)g < n && (uint32_t)p >= sm.dstart[d]);
      out_keys[g] = k;
      out_vals[g] = sm.vals[p];
    }
    __syncthreads();
    if (tid < 256) sm.off[tid] += cnt;
  }
}

// The slice of the input one CTA handles in both kernels of a pass.  Up to `slots` CTAs of the scatter kernel are
// resident at once (two per SM): when the input is more than one sub-tile per slot the slices are made equal
// (n / slots, rounded up to 256 pairs) so that the pass is ONE wave of evenly loaded CTAs -- with one 4096-pair
// sub-tile per CTA the headline size gave 394 CTAs on 296 slots, i.e. two rounds for most SMs.
int plan(const gsb_ctx* ctx, int64_t n, int64_t* chunk) {
  const int64_t slots = 2LL * (ctx->num_sms > 0 ? ctx->num_sms : 148);
  const int64_t total_subs = gsb_div_up(n, kSub);
  int64_t c = kSub;
  if (total_subs > slots) c = gsb_div_up(gsb_div_up(n, slots), 256) * 256;
  if (gsb_div_up(n, c) > kMaxBlocks) c = gsb_div_up(gsb_div_up(n, kMaxBlocks), 256) * 256;
  *chunk = c;
  return (int)gsb_div_up(n, c);
}

}  // namespace

// Ping-pong sort.  Input in (k0,v0); pass p reads buffer p%2 and writes buffer (p+1)%2, except that
// the final pass writes its values to final_vals when that is non-null.  Returns through
// *result_in_second whether the sorted keys ended up in k1 (odd number of passes).
int gsb_radix_sort_pingpong(gsb_ctx* ctx, cudaStream_t s, int64_t* k0, int32_t* v0, int64_t* k1, int32_t* v1,
                            int32_t* final_vals, int64_t n, int begin_bit, int end_bit, bool* result_in_second) {
  *result_in_second = false;
  if (n <= 0) return GSB_OK;
  if (n > GSB_MAX_RENDERED) return gsb_set_error(ctx, GSB_ERR_TOO_MANY, "radix sort: %lld pairs exceed 2^30-1", (long long)n);
  if (begin_bit < 0 || end_bit > 64 || end_bit <= begin_bit)
    return gsb_set_error(ctx, GSB_ERR_INVALID, "radix sort: bad bit range [%d,%d)", begin_bit, end_bit);
  bool& attr_set = ctx->smem_optin_radix;
  if (!attr_set) {
    GSB_CUDA(ctx, cudaFuncSetAttribute(radix_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(ScatterSmem)));
    attr_set = true;
  }
  int64_t chunk;
  const int nb = plan(ctx, n, &chunk);
  // [256][nb] counts + [256][nb] prefixes
  int rc = gsb_grow(ctx, (void**)&ctx->sort_table, &ctx->sort_table_cap, (int64_t)nb * 512, sizeof(uint32_t), s);
  if (rc != GSB_OK) return rc;
  uint32_t* const counts = ctx->sort_table;
  uint32_t* const prefix = ctx->sort_table + (size_t)nb * 256;
  uint32_t* const digit_total = ctx->sort_small;     // [256]
  const int passes = (end_bit - begin_bit + 7) / 8;
  uint64_t* kb[2] = {reinterpret_cast<uint64_t*>(k0), reinterpret_cast<uint64_t*>(k1)};
  int32_t* vb[2] = {v0, v1};
  for (int p = 0; p < passes; ++p) {
    int shift = begin_bit + 8 * p;
    int bits = (end_bit - shift) < 8 ? (end_bit - shift) : 8;
    unsigned mask = (1u << bits) - 1u;
    const uint64_t* ik = kb[p & 1];
    const int32_t* iv = vb[p & 1];
    uint64_t* ok = kb[(p + 1) & 1];
    int32_t* ov = (p == passes - 1 && final_vals) ? final_vals : vb[(p + 1) & 1];
    GSB_LAUNCH(ctx, radix_count_kernel, nb, kCountThreads, 0, s, ik, n, shift, mask, chunk, counts);
    GSB_LAUNCH(ctx, radix_scan_kernel, 32, 256, 0, s, counts, prefix, digit_total, nb);
    GSB_LAUNCH(ctx, radix_scatter_kernel, nb, kThreads, sizeof(ScatterSmem), s, ik, iv, ok, ov, n, shift, mask, chunk,
               prefix, digit_total);
  }
  *result_in_second = (passes & 1) != 0;
  return GSB_OK;
}

GSB_API int gsb_sort_pairs64(gsb_ctx* ctx, gsb_stream s_, int64_t* keys, int32_t* values, int64_t* tmp_keys,
                             int32_t* tmp_values, int64_t count, int begin_bit, int end_bit) {
  if (!ctx) return GSB_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)s_;
  if (count <= 0) return GSB_OK;
  if (!tmp_keys || !tmp_values) {
    int rc = gsb_reserve_binning(ctx, s, count);
    if (rc != GSB_OK) return rc;
    if (!tmp_keys) tmp_keys = ctx->keys_b;
    if (!tmp_values) tmp_values = ctx->vals_b;
  }
  bool in_tmp = false;
  int rc = gsb_radix_sort_pingpong(ctx, s, keys, values, tmp_keys, tmp_values, nullptr, count, begin_bit, end_bit, &in_tmp);
  if (rc != GSB_OK) return rc;
  if (in_tmp) {
    GSB_CUDA(ctx, cudaMemcpyAsync(keys, tmp_keys, sizeof(int64_t) * (size_t)count, cudaMemcpyDeviceToDevice, s));
    GSB_CUDA(ctx, cudaMemcpyAsync(values, tmp_values, sizeof(int32_t) * (size_t)count, cudaMemcpyDeviceToDevice, s));
  }
  return GSB_OK;
}
