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
template <int BITS = 8>
__device__ __forceinline__ unsigned warp_peers8(const unsigned d, const bool valid) {
  unsigned peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
  for (int b = 0; b < BITS; ++b) {
    const bool bit = (d >> b) & 1u;
    const unsigned m = __ballot_sync(0xffffffffu, bit);
    peers &= bit ? m : ~m;
  }
  return peers;
}

// Stable rank of this lane's key among the keys of digit d the warp has seen so far (wh: the warp's 256 running
// counters in shared memory), for one round of 32 keys.  Every lane executes the shuffle with the FULL mask: with the
// peer mask as the shuffle's member mask (round 2's first version) nvcc verifies the mask with a MATCH.ANY per
// shuffle -- the MIO-pipe instruction the ballots were introduced to avoid -- and the ranking of 12 rounds took
// 12.4 us per pass in the cooperative kernel instead of 2.
template <int BITS = 8, typename Counter>
__device__ __forceinline__ uint32_t warp_rank_round(Counter* wh, const unsigned d, const bool valid, const unsigned lt,
                                                    const int lane) {
  const unsigned m = warp_peers8<BITS>(d, valid);
  const int leader = __ffs(m) - 1;            // -1 on a lane without a key: it reads lane 31 and ignores the result
  uint32_t old = 0;
  if (valid && lane == leader) {
    old = wh[d];
    wh[d] = (Counter)(old + __popc(m));
  }
  old = __shfl_sync(0xffffffffu, old, leader & 31);
  __syncwarp();                               // the next round's leader may be another lane
  return old + __popc(m & lt);
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
      rank[i] = warp_rank_round(sm.whist[warp], d, valid, lt, lane);
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

// ------------------------------------------------------------------------------------------
// All passes in ONE cooperative launch (inputs up to num_sms x 12288 pairs: the headline frame's 1.6 M duplicates).
// At that size a pass of the three-kernel scheme above is bound by launch gaps and dependent latencies (41.6 us for
// 38 MB that sit in L2), not by bytes.  Here one CTA of 1024 threads per SM owns a contiguous slice of <= 12288 pairs
// for the whole sort, the digit is 8 or 9 bits wide (9 when that saves a pass: 44 live bits = 5 passes), and a pass is
//   load   the slice's keys into REGISTERS (12 per thread; they stay there until the scatter: a key is read once
//          per pass, not twice), its values into shared memory by cp.async;
//   rank   ballot peer masks + per-warp digit counters (16 bits).  tools/ubench/warp_ops.cu measured what the
//          instructions cost per scheduler with 8 warps each: VOTE 8 cycles, POPC 8, FLO 16, MATCH.ANY 256 (!),
//          SHFL 5, LDS 4.4 -- so a round is one VOTE per digit bit, ONE POPC (the highest lane of a peer group
//          writes the counter: its rank + 1 is the group's size), no FLO, no shuffle, no ballot for the slice's tail
//          (lanes behind the end rank as the all-ones digit, behind every real key, and are subtracted afterwards);
//   count  transposed: lane w of a warp reads warp w's counter of a digit, a shuffle scan gives every warp's first
//          slot of the digit in the digit-sorted slice and the CTA's count, which goes to table[digit][cta] and,
//          by RED, to the digit's partial sum of the CTA's GROUP of 16;
//   ---- grid barrier ----
//   base   the CTA's first output position of every digit = exclusive scan of the digit totals + the digit's counts
//          in the CTAs before this one, both from <= 10 group sums and <= 15 counts of the own group (reading the
//          whole 256 x G table in every CTA, the first version, was 22 MB per pass: L2-bandwidth bound at 3.6 us);
//   move   keys to their digit-sorted place in shared memory, then out in contiguous runs (a digit's run is
//          24-48 pairs on average), values through a 16-bit source index;
//   ---- grid barrier ----  (the next pass reads what this one wrote)
// i.e. 24 B per pair and pass and two grid barriers (one counter in L2, ~2 us each) instead of three launches.
// Loads of data written earlier in the same kernel by other SMs go through L2 (ld.global.cg / cp.async.cg).
// Measured (tools/sort_phases.py, -DGSB_SORT_TIMING): profiles/r02_experiments.md section 3.
constexpr int kCoopThreads = 1024;
constexpr int kCoopWarps = kCoopThreads / 32;
constexpr int kCoopItems = 12;
constexpr int kCoopCap = kCoopThreads * kCoopItems;   // pairs per CTA
constexpr int kCoopMaxGrid = 160;                     // CTAs = row length of the table
constexpr int kCoopGroup = 16;                        // CTAs per group (<= 10 groups)
constexpr int kCoopStateWords = 2 + 2 * 512 * 16;     // grid-barrier counter, exit counter, 2 x [digit][16 group sums]

template <int BITS>
struct CoopSmem {
  static constexpr int NB = 1 << BITS;
  static constexpr int ROW = NB;
  uint64_t keys[kCoopCap];       // the slice's keys, stably sorted by the pass's digit
  int32_t vals_in[kCoopCap];     // the slice's values in input order
  uint16_t src[kCoopCap];        // input position (within the slice) of each sorted key
  uint16_t whist[kCoopWarps][ROW];  // per-warp digit counters -> the warp's first slot of the digit in keys[]
  uint32_t dstart[NB];           // start of the digit's run in keys[]
  uint32_t base[NB];             // global output position of the run's first key minus dstart
  uint32_t part[kCoopThreads];   // partial sums of the base computation
  uint32_t wtot[2][16];
  int last;
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Every CTA of the (co-resident: cooperative launch) grid arrives; returns when `target` arrivals have been counted.
__device__ __forceinline__ void coop_grid_barrier(unsigned* counter, const unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();              // the CTA's writes (ordered before by the barrier above) are visible device-wide
    atomicAdd(counter, 1u);
    while (ld_acquire_u32(counter) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

// exclusive scan of the NB values held by threads 0 .. NB-1; all threads call (one __syncthreads inside)
template <int NB>
__device__ __forceinline__ uint32_t coop_scan(const uint32_t v, const int lane, const int warp, uint32_t* s_wtot) {
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31 && warp < NB / 32) s_wtot[warp] = inc;
  __syncthreads();
  uint32_t wb = 0;
#pragma unroll
  for (int w = 0; w < NB / 32; ++w)
    if (w < warp) wb += s_wtot[w];
  return wb + inc - v;
}

#ifdef GSB_SORT_TIMING
__device__ long long g_coop_clk[160 * 16];
#define COOP_TICK(slot) do { if (tid == 0 && p == 1) g_coop_clk[cta * 16 + (slot)] = clock64(); } while (0)
#else
#define COOP_TICK(slot) ((void)0)
#endif

template <int BITS>
__global__ void __launch_bounds__(kCoopThreads, 1)
radix_coop_kernel(uint64_t* __restrict__ k0, int32_t* __restrict__ v0, uint64_t* __restrict__ k1, int32_t* __restrict__ v1,
                  int32_t* __restrict__ final_vals, const int n, const int begin_bit, const int end_bit, const int chunk,
                  uint32_t* __restrict__ table /*[gridDim.x][NB]*/, unsigned* __restrict__ state /*[kCoopStateWords], zero on entry and exit*/) {
  constexpr int NB = 1 << BITS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CoopSmem<BITS>& sm = *reinterpret_cast<CoopSmem<BITS>*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x, cta = blockIdx.x;
  const int mygroup = cta / kCoopGroup, ngroups = (G + kCoopGroup - 1) / kCoopGroup;
  const unsigned lt = lanemask_lt();
  const int begin = cta * chunk;
  const int end = min(n, begin + chunk);
  const int count = max(0, end - begin);
  const int passes = (end_bit - begin_bit + BITS - 1) / BITS;
  const int lbase = warp * (32 * kCoopItems) + lane;   // slice-local index of this thread's item 0
  // items of this warp behind the slice's end (they rank as the all-ones digit)
  const int warp_invalid = min(32 * kCoopItems, max(0, (warp + 1) * (32 * kCoopItems) - count));
  unsigned* const bar = state;
  unsigned arrivals = 0;

  for (int p = 0; p < passes; ++p) {
    const int shift = begin_bit + BITS * p;
    const unsigned digit_mask = (1u << min(BITS, end_bit - shift)) - 1u;
    const uint64_t* const in_k = (p & 1) ? k1 : k0;
    const int32_t* const in_v = (p & 1) ? v1 : v0;
    uint64_t* const out_k = (p & 1) ? k0 : k1;
    int32_t* const out_v = (p == passes - 1 && final_vals) ? final_vals : ((p & 1) ? v0 : v1);
    unsigned* const gpart = state + 2 + (p & 1) * (512 * 16);          // [16][NB] group sums of this pass
    unsigned* const gpart_next = state + 2 + ((p + 1) & 1) * (512 * 16);

    COOP_TICK(0);
    uint64_t key[kCoopItems];
#pragma unroll
    for (int i = 0; i < kCoopItems; ++i) {
      const int li = lbase + 32 * i;
      key[i] = li < count ? __ldcg(reinterpret_cast<const unsigned long long*>(in_k) + begin + li) : ~0ull;
    }
    // the slice's values go straight to shared memory, 16 bytes at a time through L2 (cp.async.cg), zero-filled
    // behind the slice's end; they are needed after the second block barrier from here at the earliest
#pragma unroll
    for (int j = 0; j < kCoopCap / 4 / kCoopThreads; ++j) {
      const int c4 = 4 * (tid + j * kCoopThreads);
      if (c4 < count) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(&sm.vals_in[c4]);
        const int bytes = min(16, 4 * (count - c4));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(in_v + begin + c4), "r"(bytes) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int i = tid; i < kCoopWarps * CoopSmem<BITS>::ROW / 2; i += kCoopThreads)
      reinterpret_cast<uint32_t*>(&sm.whist[0][0])[i] = 0;
    __syncthreads();
    COOP_TICK(1);

    uint32_t rank2[kCoopItems / 2];   // two 16-bit ranks (< 384) per register
    {
      uint16_t* const wh = sm.whist[warp];
#pragma unroll
      for (int i = 0; i < kCoopItems; ++i) {
        const unsigned d = (unsigned)(key[i] >> shift) & digit_mask;
        unsigned m = 0xffffffffu;
#pragma unroll
        for (int b = 0; b < BITS; ++b) {
          const bool bit = (d >> b) & 1u;
          const unsigned v = __ballot_sync(0xffffffffu, bit);
          m &= bit ? v : ~v;
        }
        const uint32_t below = __popc(m & lt);
        const uint32_t r = wh[d] + below;
        __syncwarp();
        if ((m >> lane) == 1u) wh[d] = (uint16_t)(r + 1u);   // the highest lane of the peer group
        __syncwarp();
        rank2[i >> 1] = (i & 1) ? (rank2[i >> 1] | (r << 16)) : r;
      }
      if (lane == 0 && warp_invalid) wh[digit_mask] = (uint16_t)(wh[digit_mask] - warp_invalid);
    }
    __syncthreads();
    COOP_TICK(2);

    // thread d: digit d's counters over the warps -> exclusive prefixes in place; the CTA's count goes to the table
    // (CTA-major: coalesced) and, by RED, to the sum of the CTA's group
    uint32_t cnt = 0;
    if (tid < NB) {
#pragma unroll 8
      for (int w = 0; w < kCoopWarps; ++w) {
        const uint32_t c = sm.whist[w][tid];
        sm.whist[w][tid] = (uint16_t)cnt;
        cnt += c;
      }
      table[(size_t)cta * NB + tid] = cnt;
      if (cnt) atomicAdd(gpart + mygroup * NB + tid, cnt);
    }
    const uint32_t ds = coop_scan<NB>(cnt, lane, warp, sm.wtot[0]);   // start of the digit's run in the sorted slice
    if (tid < NB) sm.dstart[tid] = ds;
    asm volatile("cp.async.wait_group 0;" ::: "memory");   // own copies landed; the grid barrier publishes them to the CTA

    COOP_TICK(3);
    arrivals += (unsigned)G;
    coop_grid_barrier(bar, arrivals);
    COOP_TICK(4);

    // digit totals and the counts of the CTAs before this one: thread (d, part 0) adds the <= 10 group sums of
    // digit d, the threads (d, part > 0) the counts of the CTAs of the own group before this one; every load is a
    // coalesced row of 32 consecutive digits
    {
      constexpr int P = kCoopThreads / NB;                                   // threads per digit
      constexpr int PER = (kCoopGroup - 1 + (P - 2)) / (P - 1);              // in-group CTAs per thread of part > 0
      const int d = tid & (NB - 1), part = tid >> BITS;
      uint32_t tot = 0, mine = 0;
      if (part == 0) {
#pragma unroll
        for (int g = 0; g < kCoopMaxGrid / kCoopGroup; ++g) {
          const uint32_t v = g < ngroups ? __ldcg(gpart + g * NB + d) : 0u;
          tot += v;
          mine += g < mygroup ? v : 0u;
        }
        sm.part[tid] = mine;
      } else {
#pragma unroll
        for (int j = 0; j < PER; ++j) {
          const int c = mygroup * kCoopGroup + (part - 1) * PER + j;
          mine += c < cta ? __ldcg(table + (size_t)c * NB + d) : 0u;
        }
        sm.part[tid] = mine;
      }
      __syncthreads();
      if (tid < NB) {
#pragma unroll
        for (int q = 1; q < P; ++q) mine += sm.part[q * NB + tid];
      }
    // the other buffer of group sums is zeroed for the next pass (its last readers passed the previous barrier)
      for (int i = cta * kCoopThreads + tid; i < NB * 16; i += G * kCoopThreads) gpart_next[i] = 0u;
      COOP_TICK(5);
      const uint32_t gex = coop_scan<NB>(tid < NB ? tot : 0u, lane, warp, sm.wtot[1]);
      if (tid < NB) sm.base[tid] = gex + mine - ds;
    }
    __syncthreads();
    COOP_TICK(6);

#pragma unroll
    for (int i = 0; i < kCoopItems; ++i) {
      const int li = lbase + 32 * i;
      if (li < count) {
        const unsigned d = (unsigned)(key[i] >> shift) & digit_mask;
        const uint32_t q = sm.dstart[d] + sm.whist[warp][d] + ((i & 1) ? (rank2[i >> 1] >> 16) : (rank2[i >> 1] & 0xffffu));
        GSB_DCHECK(q < (uint32_t)count);
        sm.keys[q] = key[i];
        sm.src[q] = (uint16_t)li;
      }
    }
    __syncthreads();
    COOP_TICK(7);
#pragma unroll 4
    for (int q = tid; q < count; q += kCoopThreads) {
      const uint64_t k = sm.keys[q];
      const unsigned d = (unsigned)(k >> shift) & digit_mask;
      const uint32_t g = sm.base[d] + (uint32_t)q;
      GSB_DCHECK(g < (uint32_t)n);
      out_k[g] = k;
      out_v[g] = sm.vals_in[sm.src[q]];
    }
    COOP_TICK(8);
    if (p + 1 < passes) {
      arrivals += (unsigned)G;
      coop_grid_barrier(bar, arrivals);
    }
    COOP_TICK(9);
  }
  // leave the state zero for the next launch: the CTA that arrives here last knows nobody waits or adds any more
  if (tid == 0) sm.last = (atomicAdd(bar + 1, 1u) == (unsigned)G - 1u);
  __syncthreads();
  if (sm.last) {
    unsigned* const used = state + 2 + ((passes - 1) & 1) * (512 * 16);
    for (int i = tid; i < NB * 16; i += kCoopThreads) used[i] = 0u;
    if (tid < 2) state[tid] = 0u;
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
  const int passes = (end_bit - begin_bit + 7) / 8;
  uint64_t* kb[2] = {reinterpret_cast<uint64_t*>(k0), reinterpret_cast<uint64_t*>(k1)};
  int32_t* vb[2] = {v0, v1};
  const int sms = min(ctx->num_sms > 0 ? ctx->num_sms : 148, kCoopMaxGrid);
  if (ctx->opt.sort_coop && ctx->coop_launch && n <= (int64_t)sms * kCoopCap && gsb_aligned16(v0) && gsb_aligned16(v1)) {
    // one cooperative launch for all passes: a slice per CTA, at most one CTA per SM (co-resident by construction);
    // 9-bit digits when they save a pass
    const int key_bits = end_bit - begin_bit;
    const bool wide = (key_bits + 8) / 9 < (key_bits + 7) / 8;
    if (!ctx->smem_optin_radix_coop) {
      GSB_CUDA(ctx, cudaFuncSetAttribute(radix_coop_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(CoopSmem<8>)));
      GSB_CUDA(ctx, cudaFuncSetAttribute(radix_coop_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(CoopSmem<9>)));
      ctx->smem_optin_radix_coop = true;
    }
    if (!ctx->sort_coop_state) {
      GSB_CUDA(ctx, cudaMalloc((void**)&ctx->sort_coop_state, sizeof(uint32_t) * kCoopStateWords));
      // (once per context, blocking: the first sort on ANOTHER stream must not overtake an asynchronous memset)
      GSB_CUDA(ctx, cudaMemset(ctx->sort_coop_state, 0, sizeof(uint32_t) * kCoopStateWords));
    }
    int chunk = (int)(gsb_div_up(gsb_div_up(n, (int64_t)sms), 32) * 32);
    if (chunk < 2048) chunk = 2048;
    if (chunk > kCoopCap) chunk = kCoopCap;
    const int G = (int)gsb_div_up(n, (int64_t)chunk);
    // [512][kCoopMaxGrid] counts
    int rc = gsb_grow(ctx, (void**)&ctx->sort_table, &ctx->sort_table_cap, 512 * kCoopMaxGrid, sizeof(uint32_t), s);
    if (rc != GSB_OK) return rc;
    uint32_t* table = ctx->sort_table;
    unsigned* state = ctx->sort_coop_state;
    int n32 = (int)n;
    void* args[] = {&kb[0], &vb[0], &kb[1], &vb[1], &final_vals, &n32, &begin_bit, &end_bit, &chunk, &table, &state};
    const void* fn = wide ? (const void*)radix_coop_kernel<9> : (const void*)radix_coop_kernel<8>;
#ifdef GSB_COOP_PLAIN_LAUNCH   // measurement only: what the cooperative launch itself costs
    if (wide)
      radix_coop_kernel<9><<<G, kCoopThreads, sizeof(CoopSmem<9>), s>>>(kb[0], vb[0], kb[1], vb[1], final_vals, n32, begin_bit, end_bit, chunk, table, state);
    else
      radix_coop_kernel<8><<<G, kCoopThreads, sizeof(CoopSmem<8>), s>>>(kb[0], vb[0], kb[1], vb[1], final_vals, n32, begin_bit, end_bit, chunk, table, state);
    (void)fn; (void)args;
    GSB_CUDA(ctx, cudaGetLastError());
#else
    const cudaError_t ce = cudaLaunchCooperativeKernel(fn, dim3(G), dim3(kCoopThreads), args,
                                                       wide ? sizeof(CoopSmem<9>) : sizeof(CoopSmem<8>), s);
    if (ce == cudaErrorCooperativeLaunchTooLarge) {
      // fewer SMs than the device reports can hold CTAs of this kernel (a partitioned device): nothing was launched;
      // this context sorts with the three-kernel path from now on
      cudaGetLastError();
      ctx->coop_launch = false;
    } else {
      GSB_CUDA(ctx, ce);
#endif
      ctx->launches += 1;
      const int coop_passes = wide ? (key_bits + 8) / 9 : passes;
      *result_in_second = (coop_passes & 1) != 0;
      return GSB_OK;
#ifndef GSB_COOP_PLAIN_LAUNCH
    }
#endif
  }
  int64_t chunk;
  const int nb = plan(ctx, n, &chunk);
  // [256][nb] counts + [256][nb] prefixes
  int rc = gsb_grow(ctx, (void**)&ctx->sort_table, &ctx->sort_table_cap, (int64_t)nb * 512, sizeof(uint32_t), s);
  if (rc != GSB_OK) return rc;
  uint32_t* const counts = ctx->sort_table;
  uint32_t* const prefix = ctx->sort_table + (size_t)nb * 256;
  uint32_t* const digit_total = ctx->sort_small;     // [256]
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

#ifdef GSB_SORT_TIMING
// measurement builds only: the cooperative kernel's phase clocks of pass 1, [cta][16]
GSB_API int gsb_debug_sort_clocks(long long* host_out, int count) {
  return cudaMemcpyFromSymbol(host_out, g_coop_clk, sizeof(long long) * (size_t)count) == cudaSuccess ? 0 : 1;
}
#endif
