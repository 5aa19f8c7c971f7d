// tilesort.cu -- the binning fast path of gsb_forward: an MSD/LSD hybrid of the reference's
// "duplicate with keys -> radix sort of 64-bit tile|depth keys -> identify tile ranges"
// (forward.py:517-586, 791-840) that produces the SAME point_list and ranges.
//
// The 64-bit key is (tile_id << 32) | depth_bits and the unsorted list is in ascending Gaussian
// order, so the reference's stable sort orders the duplicates by (tile, depth_bits, gaussian_id).
// Here the most significant digit -- the tile id -- is resolved by a counting sort:
//   tile_count_kernel   one RED per (Gaussian, tile) on a per-tile counter (one 128-byte line per counter:
//                       neighbouring tiles must not serialise on one L2 line); inside gsb_forward this pass is
//                       part of preprocess_kernel
//   tile_scan_kernel    exclusive scan of the counts = the tile ranges (and D, and the max count)
//   tile_scatter_kernel every duplicate takes a slot of its tile's segment by counting the tile's counter back
//                       DOWN (one returning atomic; the order inside the segment is arbitrary, the sort fixes it)
//                       and goes there as (depth_bits<<32 | id); the counters end the frame at zero
//   tile_sort_kernel    one CTA per tile sorts its segment (bitonic network on the 64-bit
//                       composites -- they are unique, so any correct sort gives the stable order;
//                       strides below 64 run in registers with shuffles, only the wide strides go
//                       through shared memory) and writes the Gaussian ids to point_list
// Traffic: 8 B written + 8 B read + 4 B written per duplicate, instead of 32 B per duplicate per
// radix pass (6 passes at 800x800).  Tiles longer than kMaxTileSort fall back to the global radix
// sort (sort.cu), which is also what the stage-level entry point gsb_sort_pairs64 runs.
#include "tilesort.cuh"

constexpr int kMaxTileSort = 16384;

namespace {

constexpr int kCntStride = 32;  // ints between two tiles' counters = one 128-byte line each

__global__ void __launch_bounds__(256)
tile_count_kernel(int n, const float2* __restrict__ xy, const int* __restrict__ radii, int grid_x, int grid_y,
                  int* __restrict__ tile_count) {
  int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n) return;
  int r = radii[tid];
  if (r <= 0) return;
  float2 p = xy[tid];
  int rminx, rminy, rmaxx, rmaxy;
  gs_get_rect(p.x, p.y, (float)r, (float)grid_x, (float)grid_y, rminx, rminy, rmaxx, rmaxy);
  for (int y = rminy; y < rmaxy; ++y)
    for (int x = rminx; x < rmaxx; ++x) atomicAdd(tile_count + (size_t)(y * grid_x + x) * kCntStride, 1);
}

// Single CTA: exclusive scan over the tiles.  Writes ranges (start,end) -- (0,0) for empty tiles,
// like the zero-initialised reference buffer --, the per-tile write cursors, the total and the max.
// The counters keep their counts: the scatter pass takes its slots from them and leaves them at zero for the next
// frame (gsb_tile_binning_prepare then skips its memset).
__global__ void __launch_bounds__(1024)
tile_scan_kernel(int num_tiles, const int* __restrict__ tile_count, int2* __restrict__ ranges,
                 int* __restrict__ out_total_max, volatile int* __restrict__ host_total_max, const int spec_cap,
                 const int spec_max) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  __shared__ int s_max[32];
  gsb_pdl_wait();
  gsb_pdl_launch_dependents();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  int my_max = 0;
  __syncthreads();
  // four consecutive tiles per thread, all four counters (one 128-byte line each) requested before the scan:
  // 4096 tiles per round -- one round at 800x800, where the chain of dependent loads used to cost three
  constexpr int kPer = 4;
  for (int base = 0; base < num_tiles; base += 1024 * kPer) {
    const int i0 = base + tid * kPer;
    int c[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) c[k] = (i0 + k < num_tiles) ? tile_count[(size_t)(i0 + k) * kCntStride] : 0;
    int local = 0;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      my_max = max(my_max, c[k]);
      local += c[k];
    }
    int inc = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int wb = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 32; ++w) {
      const int t = s_warp[w];
      if (w < warp) wb += t;
      tot += t;
    }
    int start = s_carry + wb + inc - local;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      if (i0 + k < num_tiles) ranges[i0 + k] = (c[k] > 0) ? make_int2(start, start + c[k]) : make_int2(0, 0);
      start += c[k];
    }
    __syncthreads();
    if (tid == 0) s_carry += tot;
    __syncthreads();
  }
  my_max = __reduce_max_sync(0xffffffffu, my_max);
  if (lane == 0) s_max[warp] = my_max;
  __syncthreads();
  if (warp == 0) {
    const int m = __reduce_max_sync(0xffffffffu, s_max[lane]);   // 32 warps: one value per lane
    if (lane != 0) return;
    out_total_max[0] = s_carry;
    out_total_max[1] = m;
    // the go-ahead of the kernels gsb_forward queued behind this one WITHOUT waiting for D on the host: they run only
    // if the frame fits what the host assumed (buffer capacity, the sort kernel's capacity class)
    out_total_max[2] = (s_carry > 0 && s_carry <= spec_cap && m <= spec_max) ? 1 : 0;
    // D and the longest list go straight to the host's (mapped, pinned) scalars: no copy operation sits in the
    // stream between this kernel and the scatter pass (copy + broken chain of dependent launches cost the step 12 us)
    host_total_max[0] = s_carry;   // (visible to the host when the event behind this kernel has completed)
    host_total_max[1] = m;
  }
}

// LANES threads per Gaussian: lane t of the group writes the tiles t, t + LANES, ... of the Gaussian's rectangle.
// With eight lanes a Gaussian that covers a hundred tiles (6M Gaussians at 4K: 18 per visible Gaussian, with a long
// tail) no longer makes one thread walk them all while its warp waits.  At the headline size (5.7 tiles per visible
// Gaussian) the seven extra lanes mostly idle and one thread per Gaussian is faster: the host picks by D / n.
// The slot inside the tile's segment is the tile's counter, counted down: up to four returning atomics in flight.
template <int kScatterLanes>
__global__ void __launch_bounds__(256)
tile_scatter_kernel(int n, const float2* __restrict__ xy, const float* __restrict__ depths,
                    const int* __restrict__ radii, int grid_x, int grid_y, const int2* __restrict__ ranges,
                    int* __restrict__ tile_count, unsigned long long* __restrict__ binned, const int* __restrict__ go) {
  gsb_pdl_wait();
  gsb_pdl_launch_dependents();
  if (go && *go == 0) return;   // queued speculatively and the frame does not fit (see tile_scan_kernel)
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int gid = (int)(t / kScatterLanes), sub = (int)(t % kScatterLanes);
  if (gid >= n) return;
  const int r = radii[gid];
  if (r <= 0) return;
  const float2 p = xy[gid];
  int rminx, rminy, rmaxx, rmaxy;
  gs_get_rect(p.x, p.y, (float)r, (float)grid_x, (float)grid_y, rminx, rminy, rmaxx, rmaxy);
  const unsigned long long v = ((unsigned long long)__float_as_uint(depths[gid]) << 32) | (unsigned)gid;
  const int w = rmaxx - rminx, cnt = w * (rmaxy - rminy);
  if (w <= 0) return;
  int y = rminy + sub / w, x = rminx + sub % w;   // tile `sub` of the rectangle, row-major
  for (int k = sub; k < cnt; k += 4 * kScatterLanes) {
    int tile[4], slot[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      tile[u] = -1;
      if (k + u * kScatterLanes < cnt) {
        GSB_DCHECK(y < rmaxy && x >= rminx && x < rmaxx);
        tile[u] = y * grid_x + x;
        slot[u] = atomicSub(tile_count + (size_t)tile[u] * kCntStride, 1) - 1;
        x += kScatterLanes;
        while (x >= rmaxx) {
          x -= w;
          ++y;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (tile[u] >= 0) {
        const int2 rg = ranges[tile[u]];
        GSB_DCHECK(slot[u] >= 0 && slot[u] < rg.y - rg.x);
        binned[rg.x + slot[u]] = v;
      }
  }
}

// One CTA per tile: see tile_sort_segment (tilesort.cuh).
template <int CAP>
__global__ void __launch_bounds__(256)
tile_sort_kernel(const int2* __restrict__ ranges, const unsigned long long* __restrict__ binned,
                 int* __restrict__ point_list, int lo, int hi) {
  extern __shared__ __align__(16) unsigned long long s_key[];
  const int2 rg = ranges[blockIdx.x];
  const int count = rg.y - rg.x;
  if (count <= lo || count > hi) return;  // empty, or a tile another launch sorts
  tile_sort_segment(rg, binned, point_list, s_key);
}

// One CTA per tile: stable LSD radix sort of the segment by the 32 depth bits (8 bits per pass, passes
// whose digit is the same for every key are skipped -- the top byte of a positive float almost always
// is), ids as payload, then ties (equal depth bits) put in ascending-id order: the order of the
// reference's stable sort of (tile | depth) keys over an id-ordered list.  O(n) instead of the bitonic
// network's O(n log^2 n); NOT the default -- see gsb_tile_binning_sort for the measurement.
//   count phase   one shared-memory atomic per element on hist[digit][warp]
//   scan          thread d scans digit d's eight warp counters, block-wide exclusive scan of the digits
//   scatter       a warp walks its contiguous block in order; lanes with the same digit are ranked
//                 by lane (peer masks from ballots), the first of them advances hist[digit][warp]
// Tiles whose count is outside (lo, hi] return at once (they belong to the bitonic kernel).
template <int CAP>
__global__ void __launch_bounds__(256)
tile_radix_kernel(const int2* __restrict__ ranges, const unsigned long long* __restrict__ binned,
                  int* __restrict__ point_list, int lo, int hi) {
  // 12 bytes per element (two key buffers, two 16-bit payload buffers: the payload is the element's
  // index in the unsorted segment, < CAP <= 4096) + the 8 KB histogram: 32 KB at CAP = 2048, so that
  // seven CTAs share an SM -- the kernel is a chain of dependent shared-memory operations and lives
  // on occupancy.
  extern __shared__ __align__(16) unsigned s_u[];
  unsigned* ks = s_u;                                                // keys, source
  unsigned* kd = ks + CAP;                                           // keys, destination
  unsigned* hist = kd + CAP;                                         // [256 digits][8 warps]
  unsigned short* vs = reinterpret_cast<unsigned short*>(hist + 2048);  // payload, source
  unsigned short* vd = vs + CAP;                                     // payload, destination
  __shared__ unsigned s_red[8];
  __shared__ unsigned s_wtot[8];
  const int2 rg = ranges[blockIdx.x];
  const int count = rg.y - rg.x;
  if (count <= lo || count > hi) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (count == 1) {
    if (tid == 0) point_list[rg.x] = (int)(unsigned)binned[rg.x];
    return;
  }
  unsigned diff = 0u;
  {
    const unsigned k0 = (unsigned)(binned[rg.x] >> 32);
    for (int i = tid; i < count; i += 256) {
      const unsigned k = (unsigned)(binned[rg.x + i] >> 32);
      ks[i] = k;
      vs[i] = (unsigned short)i;
      diff |= k ^ k0;
    }
  }
  diff = __reduce_or_sync(0xffffffffu, diff);
  if (lane == 0) s_red[warp] = diff;
  __syncthreads();
  diff = 0u;
#pragma unroll
  for (int w = 0; w < 8; ++w) diff |= s_red[w];

  const int per = (((count + 7) >> 3) + 31) & ~31;  // a warp's contiguous block, whole rounds of 32
  const int wbeg = warp * per, wend = min(count, wbeg + per);
  for (int shift = 0; shift < 32; shift += 8) {
    if (((diff >> shift) & 255u) == 0u) continue;  // every key has the same digit here
    // (the __syncthreads at the end of the previous pass / after the load orders the buffers)
#pragma unroll
    for (int k = 0; k < 8; ++k) hist[tid + 256 * k] = 0u;
    __syncthreads();
    for (int i = wbeg + lane; i < wend; i += 32) atomicAdd(&hist[((ks[i] >> shift) & 255u) * 8 + warp], 1u);
    __syncthreads();
    {
      // thread d: exclusive scan of digit d's eight counters, then of the digit totals over the block
      uint4 c0 = *reinterpret_cast<uint4*>(hist + tid * 8), c1 = *reinterpret_cast<uint4*>(hist + tid * 8 + 4);
      const unsigned tot = c0.x + c0.y + c0.z + c0.w + c1.x + c1.y + c1.z + c1.w;
      unsigned inc = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (lane == 31) s_wtot[warp] = inc;
      __syncthreads();
      unsigned base = inc - tot;
#pragma unroll
      for (int w = 0; w < 8; ++w)
        if (w < warp) base += s_wtot[w];
      uint4 e0, e1;
      e0.x = base;
      e0.y = e0.x + c0.x;
      e0.z = e0.y + c0.y;
      e0.w = e0.z + c0.z;
      e1.x = e0.w + c0.w;
      e1.y = e1.x + c1.x;
      e1.z = e1.y + c1.y;
      e1.w = e1.z + c1.z;
      *reinterpret_cast<uint4*>(hist + tid * 8) = e0;
      *reinterpret_cast<uint4*>(hist + tid * 8 + 4) = e1;
    }
    __syncthreads();
    for (int base_i = wbeg; base_i < wend; base_i += 32) {
      const int i = base_i + lane;
      const bool live = i < wend;
      const unsigned key = live ? ks[i] : 0u;
      const unsigned short val = live ? vs[i] : (unsigned short)0;
      const unsigned digit = live ? ((key >> shift) & 255u) : 256u + lane;  // dead lanes are nobody's peers
      // peer mask from one ballot per digit bit: MATCH.ANY goes through the MIO pipe at about one warp
      // instruction per 64 cycles per SM (sort.cu); measured: forward at 6M Gaussians / 4K 5.53 ms with the
      // ballots against 6.1 with match_any (and 6.51 with the bitonic network)
      unsigned peers = __ballot_sync(0xffffffffu, live);
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const bool bit = (digit >> b) & 1u;
        const unsigned bm = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? bm : ~bm;
      }
      if (!live) peers = 1u << lane;
      const int leader = __ffs(peers) - 1;
      const unsigned rank = __popc(peers & ((1u << lane) - 1u));
      unsigned start = 0u;
      if (live && lane == leader) {
        start = hist[digit * 8 + warp];
        hist[digit * 8 + warp] = start + __popc(peers);
      }
      start = __shfl_sync(0xffffffffu, start, leader);
      if (live) {
        kd[start + rank] = key;
        vd[start + rank] = val;
      }
      __syncwarp();
    }
    __syncthreads();
    unsigned* t = ks;
    ks = kd;
    kd = t;
    unsigned short* u = vs;
    vs = vd;
    vd = u;
  }
  // The payloads become Gaussian ids (kd is free: every pass ended with a swap).  Ties -- equal depth
  // bits -- must come out in ascending id order: odd-even rounds, only if any tie exists.
  unsigned* ids = kd;
  bool tie = false;
  for (int i = tid; i < count; i += 256) {
    ids[i] = (unsigned)binned[rg.x + vs[i]];
    if (i + 1 < count) tie |= (ks[i] == ks[i + 1]);
  }
  if (__syncthreads_or(tie)) {
    bool swapped;
    do {
      swapped = false;
#pragma unroll
      for (int parity = 0; parity < 2; ++parity) {
        for (int i = 2 * tid + parity; i + 1 < count; i += 512) {
          if (ks[i] == ks[i + 1] && ids[i] > ids[i + 1]) {
            const unsigned t = ids[i];
            ids[i] = ids[i + 1];
            ids[i + 1] = t;
            swapped = true;
          }
        }
        __syncthreads();
      }
    } while (__syncthreads_or(swapped));
  }
  for (int i = tid; i < count; i += 256) point_list[rg.x + i] = (int)ids[i];
}


// One CTA per tile: ONE-pass bucket sort of the segment's (depth_bits << 32 | id) composites (round 2, default).
// The per-tile LSD radix sort above pays one VOTE per key bit and round of 32 keys (8 cycles per scheduler each,
// tools/ubench/warp_ops.cu), the bitonic network log^2 n compare-exchanges; but the keys are depths -- floats spread
// over the view's range -- so an order-preserving map  bucket = (depth_bits - min) >> shift  onto CAP >= n buckets
// leaves about one entry per bucket:
//   load      the entries into registers (CAP / 256 per thread), block minimum / maximum of the depth bits;
//   count     one shared-memory atomic per entry;   scan: exclusive prefix over the buckets;
//   scatter   one returning shared-memory atomic per entry (the order inside a bucket is arbitrary);
//   finish    an entry's final position is the start of its bucket + the number of smaller 64-bit composites in the
//             bucket (counted by the entry's own thread) -- the order of the reference's stable sort (tile | depth
//             keys over an id-ordered list): ties of the depth bits come out by ascending id, whatever order the
//             atomics arrived in.
// Degenerate segments (one bucket with more than kBucketLimit entries: many equal or nearly equal depths) fall back
// to the bitonic network inside the same CTA.  Tiles whose count is outside (lo, hi] return at once.
constexpr int kBucketLimit = 32;

// The sort of one segment of 2 <= count <= CAP entries with CAP buckets (CAP / 256 entries and buckets per thread).
template <int CAP>
__device__ __forceinline__ void tile_bucket_segment(const int2 rg, const unsigned long long* __restrict__ binned,
                                                    int* __restrict__ point_list, unsigned long long* s_dst,
                                                    unsigned* s_min, unsigned* s_max, unsigned* s_wtot) {
  constexpr int E = CAP / 256;   // entries per thread, buckets per thread
  constexpr int LOG_CAP = CAP == 1024 ? 10 : CAP == 2048 ? 11 : 12;
  static_assert(CAP == (1 << LOG_CAP), "CAP must be 1024, 2048 or 4096");
  unsigned* const cursor = reinterpret_cast<unsigned*>(s_dst + CAP);  // [CAP] bucket counts -> starts -> ends
  const int count = rg.y - rg.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned long long e[E];
  unsigned kmin = 0xffffffffu, kmax = 0u;
#pragma unroll
  for (int j = 0; j < E; ++j) {
    const int i = tid + 256 * j;
    e[j] = i < count ? binned[rg.x + i] : ~0ull;
    if (i < count) {
      const unsigned k = (unsigned)(e[j] >> 32);
      kmin = min(kmin, k);
      kmax = max(kmax, k);
    }
  }
#pragma unroll
  for (int j = 0; j < E / 4; ++j) reinterpret_cast<uint4*>(cursor)[tid + 256 * j] = make_uint4(0u, 0u, 0u, 0u);
  kmin = __reduce_min_sync(0xffffffffu, kmin);
  kmax = __reduce_max_sync(0xffffffffu, kmax);
  if (lane == 0) s_min[warp] = kmin, s_max[warp] = kmax;
  __syncthreads();
#pragma unroll
  for (int w = 0; w < 8; ++w) kmin = min(kmin, s_min[w]), kmax = max(kmax, s_max[w]);
  const unsigned range = kmax - kmin;
  const int sh = max(0, (32 - __clz(range)) - LOG_CAP);   // (range >> sh) < CAP
#pragma unroll
  for (int j = 0; j < E; ++j)
    if (tid + 256 * j < count) atomicAdd(&cursor[((unsigned)(e[j] >> 32) - kmin) >> sh], 1u);
  __syncthreads();
  // exclusive scan over the buckets; thread t owns buckets t E .. t E + E - 1
  unsigned c[E], sum = 0u, cmax = 0u;
#pragma unroll
  for (int j = 0; j < E / 4; ++j) {
    const uint4 v = reinterpret_cast<const uint4*>(cursor)[tid * (E / 4) + j];
    c[4 * j + 0] = v.x, c[4 * j + 1] = v.y, c[4 * j + 2] = v.z, c[4 * j + 3] = v.w;
  }
#pragma unroll
  for (int j = 0; j < E; ++j) {
    cmax = max(cmax, c[j]);
    const unsigned t = c[j];
    c[j] = sum;
    sum += t;
  }
  unsigned inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_wtot[warp] = inc;
  if (__syncthreads_or(cmax > (unsigned)kBucketLimit)) {   // (also orders s_wtot and the reads of cursor)
    tile_sort_segment(rg, binned, point_list, s_dst);      // degenerate depths: the bitonic network (n_pad <= CAP)
    return;                                                // (uniform: the whole CTA takes this path)
  }
  unsigned base = inc - sum;
#pragma unroll
  for (int w = 0; w < 8; ++w)
    if (w < warp) base += s_wtot[w];
#pragma unroll
  for (int j = 0; j < E / 4; ++j)
    reinterpret_cast<uint4*>(cursor)[tid * (E / 4) + j] =
        make_uint4(base + c[4 * j], base + c[4 * j + 1], base + c[4 * j + 2], base + c[4 * j + 3]);
  __syncthreads();
#pragma unroll
  for (int j = 0; j < E; ++j)
    if (tid + 256 * j < count) {
      const unsigned pos = atomicAdd(&cursor[((unsigned)(e[j] >> 32) - kmin) >> sh], 1u);
      GSB_DCHECK(pos < (unsigned)count);
      s_dst[pos] = e[j];
    }
  __syncthreads();
  // cursor[b] is now the END of bucket b.  Every entry finds its final place by itself: the start of its bucket plus
  // the number of entries of the bucket below it (composites are unique: the id is part of them).  One thread per
  // ENTRY, no dependent chain -- the first version let thread t insertion-sort the slice of its buckets, and the
  // CTA waited a third of the kernel's time for the thread with the longest slice (ncu: 31% of the samples at the
  // barrier behind it).
#pragma unroll
  for (int j = 0; j < E; ++j)
    if (tid + 256 * j < count) {
      const unsigned b = ((unsigned)(e[j] >> 32) - kmin) >> sh;
      const int s0 = b == 0u ? 0 : (int)cursor[b - 1], s1 = (int)cursor[b];
      int r = 0;
      for (int i = s0; i < s1; ++i) r += (s_dst[i] < e[j]) ? 1 : 0;
      GSB_DCHECK(s0 + r < count);
      point_list[rg.x + s0 + r] = (int)(unsigned)e[j];
    }
}

// MAXCAP sizes the launch's shared memory (12 bytes per entry); each tile runs the instance that fits its count, so
// a frame whose longest list is 1100 entries still sorts its 600-entry tiles with 1024 buckets and 4 entries per thread.
template <int MAXCAP>
__global__ void __launch_bounds__(256)
tile_bucket_kernel(const int2* __restrict__ ranges, const unsigned long long* __restrict__ binned,
                   int* __restrict__ point_list, int lo, int hi, const int* __restrict__ go) {
  extern __shared__ __align__(16) unsigned long long s_dst[];      // [CAP] entries in bucket order + [CAP] cursors
  __shared__ unsigned s_min[8], s_max[8], s_wtot[8];
  gsb_pdl_wait();
  gsb_pdl_launch_dependents();
  if (go && *go == 0) return;   // queued speculatively and the frame does not fit (see tile_scan_kernel)
  const int2 rg = ranges[blockIdx.x];
  const int count = rg.y - rg.x;
  if (count <= lo || count > hi) return;
  if (count == 1) {
    if (threadIdx.x == 0) point_list[rg.x] = (int)(unsigned)binned[rg.x];
    return;
  }
  if (MAXCAP == 1024 || count <= 1024)
    tile_bucket_segment<1024>(rg, binned, point_list, s_dst, s_min, s_max, s_wtot);
  else if (MAXCAP == 2048 || count <= 2048)
    tile_bucket_segment<(MAXCAP >= 2048 ? 2048 : 1024)>(rg, binned, point_list, s_dst, s_min, s_max, s_wtot);
  else
    tile_bucket_segment<MAXCAP>(rg, binned, point_list, s_dst, s_min, s_max, s_wtot);
}

}  // namespace


// ---- host side -------------------------------------------------------------------------------
// prepare: size and zero the per-tile counters
int gsb_tile_binning_prepare(gsb_ctx* ctx, cudaStream_t s, int n, int num_tiles) {
  (void)n;
  const int64_t need = (int64_t)num_tiles * kCntStride, old_cap = ctx->tile_cap;
  int rc = gsb_grow(ctx, (void**)&ctx->tile_count, &ctx->tile_cap, need, sizeof(int32_t), s);
  if (rc != GSB_OK) return rc;
  if (ctx->tile_cap != old_cap) ctx->tile_clean = 0;  // reallocated: contents unknown
  // The previous frame's scatter pass counted every counter back to zero (tile_clean); only the first frame, a
  // larger tile grid, or a frame that did not reach its scatter pass (an error, the global-sort path) pays the memset.
  if (ctx->tile_clean < need) GSB_CUDA(ctx, cudaMemsetAsync(ctx->tile_count, 0, sizeof(int32_t) * (size_t)need, s));
  ctx->tile_clean = 0;  // dirty from here until the scatter pass has been queued
  return GSB_OK;
}

// scan the counters into ranges, start the read-back of (D, max count) and mark it with an event
int gsb_tile_binning_scan_async(gsb_ctx* ctx, cudaStream_t s, int num_tiles, int32_t* ranges, int spec_cap, int spec_max) {
  GSB_LAUNCH_PDL(ctx, tile_scan_kernel, 1, 1024, 0, s, num_tiles, ctx->tile_count, reinterpret_cast<int2*>(ranges),
             ctx->d_scalars + 4, ctx->h_scalars + 4, spec_cap, spec_max);
  // speculative frames record the event behind the scatter pass (the host is in no hurry while the GPU has the
  // blend ahead; measured: an event record between two dependent launches costs nothing either way)
  if (spec_cap <= 0) GSB_CUDA(ctx, cudaEventRecord(ctx->ev_count, s));
  return GSB_OK;
}


// the one host wait of a frame: on the event, so work queued behind the read-back keeps running
int gsb_tile_binning_wait(gsb_ctx* ctx, int64_t* num_rendered_host, int* max_count_host) {
  GSB_CUDA(ctx, cudaEventSynchronize(ctx->ev_count));
  const int32_t total = ctx->h_scalars[4];
  *num_rendered_host = (total < 0) ? (int64_t)(uint32_t)total : (int64_t)total;
  *max_count_host = ctx->h_scalars[5];
  return GSB_OK;
}

// Stand-alone counting pass (gsb_bin_by_tile): on return *num_rendered_host and *max_count_host are set.
int gsb_tile_binning_count(gsb_ctx* ctx, cudaStream_t s, int n, int width, int height, const float* points_xy,
                           const int32_t* radii, int32_t* ranges, int64_t* num_rendered_host, int* max_count_host) {
  const int gx = (width + kTile - 1) / kTile, gy = (height + kTile - 1) / kTile;
  const int num_tiles = gx * gy;
  int rc = gsb_tile_binning_prepare(ctx, s, n, num_tiles);
  if (rc != GSB_OK) return rc;
  if (n > 0)
    GSB_LAUNCH(ctx, tile_count_kernel, (int)gsb_div_up(n, 256), 256, 0, s, n, reinterpret_cast<const float2*>(points_xy),
               radii, gx, gy, ctx->tile_count);
  if ((rc = gsb_tile_binning_scan_async(ctx, s, num_tiles, ranges, 0, 0)) != GSB_OK) return rc;
  return gsb_tile_binning_wait(ctx, num_rendered_host, max_count_host);
}

// precondition (checked by the caller): ctx->bin_cap >= num_rendered and the counters hold this frame's counts
// go != nullptr: queued speculatively (num_rendered / max_count are the host's assumptions; the kernels check *go);
// only the one-pass bucket sort with max_count <= 4096 may be queued that way.
int gsb_tile_binning_sort(gsb_ctx* ctx, cudaStream_t s, int n, int width, int height, const float* points_xy,
                          const float* depths, const int32_t* radii, const int32_t* ranges, int64_t num_rendered,
                          int max_count, int32_t* point_list, const int* go) {
  const int gx = (width + kTile - 1) / kTile, gy = (height + kTile - 1) / kTile;
  const int num_tiles = gx * gy;
  unsigned long long* binned = reinterpret_cast<unsigned long long*>(ctx->keys_a);
  const int2* rg = reinterpret_cast<const int2*>(ranges);
  if (num_rendered > 8 * (int64_t)n) {
    GSB_LAUNCH_PDL(ctx, tile_scatter_kernel<8>, (unsigned)gsb_div_up((int64_t)n * 8, 256), 256, 0, s, n,
               reinterpret_cast<const float2*>(points_xy), depths, radii, gx, gy, rg, ctx->tile_count, binned, go);
  } else {
    GSB_LAUNCH_PDL(ctx, tile_scatter_kernel<1>, (unsigned)gsb_div_up(n, 256), 256, 0, s, n,
               reinterpret_cast<const float2*>(points_xy), depths, radii, gx, gy, rg, ctx->tile_count, binned, go);
  }
  ctx->tile_clean = (int64_t)num_tiles * kCntStride;   // every counter is back at zero when the scatter pass has run
  if (go) GSB_CUDA(ctx, cudaEventRecord(ctx->ev_count, s));   // speculative frame: the host's wait for D ends here
  // Per-tile sort: the bitonic network (O(n log^2 n), pure register / shuffle / shared-memory compare-exchange) or
  // the O(n) shared-memory LSD radix sort for tiles of up to 4096 entries (longer ones always go to the bitonic
  // kernel; each kernel skips the other's tiles).  Measured on a B200, whole forward, L2 flushed:
  //   ~650 entries per tile (300k Gaussians, 800^2):    bitonic 351 us, radix 353 us
  //   ~920 (1M Gaussians, 1080p):                        both 1032 us
  //   ~2100 (6M Gaussians, 4K):                          bitonic 6.51 ms, radix 5.53 ms
  // (round 1's radix kernel ranked with match_any and lost everywhere: MATCH is an MIO-pipe instruction).
  // tile_sort = 2 (default): radix when the frame's longest list exceeds 2048 entries; 0 / 1 force one of them.
  constexpr int kRadixCap = 4096;
  auto radix_smem = [](int cap) { return (size_t)(2 * cap + 2048) * sizeof(unsigned) + (size_t)2 * cap * sizeof(unsigned short); };
  bool& attr_set = ctx->smem_optin_tilesort;
  if (!attr_set) {
    GSB_CUDA(ctx, cudaFuncSetAttribute(tile_sort_kernel<kMaxTileSort>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kMaxTileSort * 8));
    GSB_CUDA(ctx, cudaFuncSetAttribute(tile_radix_kernel<kRadixCap>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)radix_smem(kRadixCap)));
    GSB_CUDA(ctx, cudaFuncSetAttribute(tile_bucket_kernel<kRadixCap>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kRadixCap * 12));
    attr_set = true;
  }
  int bitonic_lo = 0;  // the bitonic kernel sorts tiles with more than this many entries
  if (ctx->opt.tile_sort == 3) {
    // one-pass bucket sort for tiles of up to 4096 entries (12 bytes of shared memory per entry of the capacity)
    if (max_count <= 1024) {
      GSB_LAUNCH_PDL(ctx, tile_bucket_kernel<1024>, num_tiles, 256, 1024 * 12, s, rg, binned, point_list, 0, 1024, go);
    } else if (max_count <= 2048) {
      GSB_LAUNCH_PDL(ctx, tile_bucket_kernel<2048>, num_tiles, 256, 2048 * 12, s, rg, binned, point_list, 0, 2048, go);
    } else {
      GSB_LAUNCH_PDL(ctx, tile_bucket_kernel<kRadixCap>, num_tiles, 256, kRadixCap * 12, s, rg, binned, point_list, 0, kRadixCap, go);
    }
    if (max_count <= kRadixCap) return GSB_OK;
    bitonic_lo = kRadixCap;
  }
  if (ctx->opt.tile_sort == 1 || (ctx->opt.tile_sort == 2 && max_count > 2048)) {
    if (max_count <= 1024) {
      GSB_LAUNCH(ctx, tile_radix_kernel<1024>, num_tiles, 256, radix_smem(1024), s, rg, binned, point_list, 0, 1024);
    } else if (max_count <= 2048) {
      GSB_LAUNCH(ctx, tile_radix_kernel<2048>, num_tiles, 256, radix_smem(2048), s, rg, binned, point_list, 0, 2048);
    } else {
      GSB_LAUNCH(ctx, tile_radix_kernel<kRadixCap>, num_tiles, 256, radix_smem(kRadixCap), s, rg, binned, point_list, 0,
                 kRadixCap);
    }
    if (max_count <= kRadixCap) return GSB_OK;
    bitonic_lo = kRadixCap;
  }
  // The shared-memory footprint is the launch's capacity (8 B per entry), so a frame with a few very
  // long lists sorts the short ones with the small-footprint instance (several CTAs per SM) and only
  // the long ones with the 128 KB instance.
  const int kBig = 1 << 30;
  if (bitonic_lo == 0 && max_count <= 1024) {
    GSB_LAUNCH(ctx, tile_sort_kernel<1024>, num_tiles, 256, 1024 * 8, s, rg, binned, point_list, 0, kBig);
  } else if (max_count <= 4096) {
    GSB_LAUNCH(ctx, tile_sort_kernel<4096>, num_tiles, 256, 4096 * 8, s, rg, binned, point_list, bitonic_lo, kBig);
  } else {
    if (bitonic_lo < 4096)
      GSB_LAUNCH(ctx, tile_sort_kernel<4096>, num_tiles, 256, 4096 * 8, s, rg, binned, point_list, bitonic_lo, 4096);
    GSB_LAUNCH(ctx, tile_sort_kernel<kMaxTileSort>, num_tiles, 256, kMaxTileSort * 8, s, rg, binned, point_list, 4096,
               kBig);
  }
  return GSB_OK;
}

int gsb_tile_binning_max() { return kMaxTileSort; }

