// sort.cu -- stable LSD radix sort of (int64 key, int32 value) pairs, 8 bits per pass.
// Replaces the pad-copy + wp.utils.radix_sort_pairs + copy-back of forward.py:791-824 (which on
// the reference's CUDA device is cub::DeviceRadixSort inside Warp's prebuilt library [Warp]).
//
// Only the live key bits are sorted: bits [0, 32 + ceil(log2(num_tiles))) for the tile|depth
// keys (44 bits = 6 passes at 800x800 instead of 8).
//
// Per pass, two kernels and no inter-CTA spinning:
//   radix_hist_kernel    every CTA counts the digits of its slice into table[cta][256]; the LAST
//                        CTA to finish (atomic ticket) turns the table into exclusive prefixes
//                        per digit column and writes the 256 global digit bases.
//   radix_scatter_kernel every CTA re-reads its slice in sub-tiles of 4096 pairs, ranks them
//                        stably (warp match_any ranking + per-warp digit counters), reorders the
//                        sub-tile through shared memory so that each digit's run leaves as
//                        contiguous, coalesced stores, and advances its 256 running offsets.
// Algorithmic traffic: 32 B per pair per pass (key read twice, value once, pair written once).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kItems = 16;
constexpr int kSub = kThreads * kItems;  // 4096 pairs per sub-tile
constexpr int kMaxBlocks = 2048;

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

__global__ void __launch_bounds__(kThreads)
radix_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift, unsigned digit_mask, int subs_per_block,
                  uint32_t* __restrict__ table /*[nb][256]*/, uint32_t* __restrict__ digit_base /*[256]*/,
                  unsigned int* __restrict__ ticket) {
  __shared__ uint32_t s_hist[kWarps][256];
  __shared__ bool s_last;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < kWarps * 256; i += kThreads) (&s_hist[0][0])[i] = 0;
  __syncthreads();
  const int64_t begin = (int64_t)blockIdx.x * subs_per_block * kSub;
  const int64_t end = min(n, begin + (int64_t)subs_per_block * kSub);
  for (int64_t base = begin; base < end; base += kThreads * 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int64_t idx = base + u * kThreads + tid;
      bool valid = idx < end;
      unsigned d = valid ? (unsigned)((keys[idx] >> shift) & digit_mask) : 256u;
      unsigned m = __match_any_sync(0xffffffffu, d);
      if (valid && lane == (__ffs(m) - 1)) s_hist[warp][d] += __popc(m);
      __syncwarp();
    }
  }
  __syncthreads();
  {
    uint32_t c = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) c += s_hist[w][tid];
    table[(size_t)blockIdx.x * 256 + tid] = c;
  }
  // ---- last CTA: column-wise exclusive scan of the table + global digit bases ----
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    unsigned t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  uint32_t run = 0;
  const int nb = gridDim.x;
  for (int b = 0; b < nb; ++b) {
    uint32_t v = __ldcg(table + (size_t)b * 256 + tid);
    table[(size_t)b * 256 + tid] = run;
    run += v;
  }
  // exclusive scan of the 256 digit totals
  __shared__ uint32_t s_tot[256];
  s_tot[tid] = run;
  __syncthreads();
  if (tid < 32) {
    uint32_t loc[8], sum = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      loc[k] = s_tot[tid * 8 + k];
      sum += loc[k];
    }
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    uint32_t ex = inc - sum;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      digit_base[tid * 8 + k] = ex;
      ex += loc[k];
    }
  }
  if (tid == 0) *ticket = 0;  // ready for the next pass
}

struct ScatterSmem {
  uint64_t keys[kSub];
  int32_t vals[kSub];
  uint32_t whist[kWarps][256];
  uint32_t off[256];     // running global offset of each digit for this CTA
  uint32_t dstart[256];  // start of each digit's run inside the reordered sub-tile
  uint32_t wtot[kWarps];
};

__global__ void __launch_bounds__(kThreads)
radix_scatter_kernel(const uint64_t* __restrict__ in_keys, const int32_t* __restrict__ in_vals,
                     uint64_t* __restrict__ out_keys, int32_t* __restrict__ out_vals, int64_t n, int shift,
                     unsigned digit_mask, int subs_per_block, const uint32_t* __restrict__ table,
                     const uint32_t* __restrict__ digit_base) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ScatterSmem& sm = *reinterpret_cast<ScatterSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned lt = lanemask_lt();

  sm.off[tid] = digit_base[tid] + table[(size_t)blockIdx.x * 256 + tid];
  const int64_t begin = (int64_t)blockIdx.x * subs_per_block * kSub;
  const int64_t end = min(n, begin + (int64_t)subs_per_block * kSub);

  for (int64_t sub = begin; sub < end; sub += kSub) {
    const int sub_count = (int)min((int64_t)kSub, end - sub);
    for (int i = tid; i < kWarps * 256; i += kThreads) (&sm.whist[0][0])[i] = 0;
    __syncthreads();

    uint64_t key[kItems];
    int32_t val[kItems];
    uint32_t rank[kItems];
    const int64_t wbase = sub + warp * (32 * kItems) + lane;
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
      int64_t idx = wbase + i * 32;
      bool valid = idx < end;
      key[i] = valid ? in_keys[idx] : ~0ull;
      val[i] = valid ? in_vals[idx] : 0;
    }
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
      bool valid = (wbase + i * 32) < end;
      unsigned d = valid ? (unsigned)((key[i] >> shift) & digit_mask) : 256u;
      unsigned m = __match_any_sync(0xffffffffu, d);
      uint32_t old = 0;
      if (valid) {
        int leader = __ffs(m) - 1;
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
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      uint32_t c = sm.whist[w][tid];
      sm.whist[w][tid] = cnt;
      cnt += c;
    }
    {
      uint32_t inc = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (lane == 31) sm.wtot[warp] = inc;
      __syncthreads();
      uint32_t wb = 0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w)
        if (w < warp) wb += sm.wtot[w];
      sm.dstart[tid] = wb + inc - cnt;
    }
    __syncthreads();

#pragma unroll
    for (int i = 0; i < kItems; ++i) {
      if ((wbase + i * 32) < end) {
        unsigned d = (unsigned)((key[i] >> shift) & digit_mask);
        uint32_t p = sm.dstart[d] + sm.whist[warp][d] + rank[i];
        sm.keys[p] = key[i];
        sm.vals[p] = val[i];
      }
    }
    __syncthreads();

#pragma unroll 4
    for (int p = tid; p < sub_count; p += kThreads) {
      uint64_t k = sm.keys[p];
      unsigned d = (unsigned)((k >> shift) & digit_mask);
      uint32_t g = sm.off[d] + ((uint32_t)p - sm.dstart[d]);
      out_keys[g] = k;
      out_vals[g] = sm.vals[p];
    }
    __syncthreads();
    sm.off[tid] += cnt;
  }
}

int plan(int64_t n, int* subs_per_block) {
  int64_t total_subs = gsb_div_up(n, kSub);
  int spb = (int)gsb_div_up(total_subs, kMaxBlocks);
  if (spb < 1) spb = 1;
  *subs_per_block = spb;
  return (int)gsb_div_up(total_subs, spb);
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
  int spb;
  int nb = plan(n, &spb);
  int rc = gsb_grow(ctx, (void**)&ctx->sort_table, &ctx->sort_table_cap, (int64_t)nb * 256, sizeof(uint32_t), s);
  if (rc != GSB_OK) return rc;
  uint32_t* digit_base = ctx->sort_small;            // [256]
  unsigned int* ticket = ctx->sort_small + 256;      // zero-initialised at context creation
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
    GSB_LAUNCH(ctx, radix_hist_kernel, nb, kThreads, 0, s, ik, n, shift, mask, spb, ctx->sort_table, digit_base, ticket);
    GSB_LAUNCH(ctx, radix_scatter_kernel, nb, kThreads, sizeof(ScatterSmem), s, ik, iv, ok, ov, n, shift, mask, spb,
               ctx->sort_table, digit_base);
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
