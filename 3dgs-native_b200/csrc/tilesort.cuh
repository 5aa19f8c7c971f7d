// tilesort.cuh -- the per-tile bitonic sort shared by tile_sort_kernel (tilesort.cu) and the forward tile
// kernel's fused prologue (blend.cu).
#pragma once
#include "common.cuh"

namespace {

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
  unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)v, m);
  unsigned hi = __shfl_xor_sync(0xffffffffu, (unsigned)(v >> 32), m);
  return ((unsigned long long)hi << 32) | lo;
}

// One CTA per tile: bitonic sort of the segment's (depth_bits<<32 | id) composites.
// A warp owns whole 64-element chunks (dealt round-robin); in chunk c lane l holds elements
// e0 = 64c + l and e1 = e0 + 32.  Compare-exchange strides j < 32 are shuffles, j == 32 is
// in-thread, j >= 64 goes through shared memory.  Direction of element i in merge size k:
// ascending iff (i & k) == 0 (automatically true for the final merge k == n_pad).
// Called by all 256 threads of a CTA with the tile's range rg (count >= 1) and s_key = shared memory for the
// padded segment (8 bytes per entry of the next power of two >= max(count, 64)).  No barrier at the end:
// the caller synchronises before it reuses s_key or reads point_list back.
__device__ __forceinline__ void tile_sort_segment(const int2 rg, const unsigned long long* __restrict__ binned,
                                                  int* point_list, unsigned long long* s_key) {
  const int count = rg.y - rg.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (count == 1) {
    if (tid == 0) point_list[rg.x] = (int)(unsigned)binned[rg.x];
    return;
  }
  int n_pad = 64;
  while (n_pad < count) n_pad <<= 1;  // <= CAP by construction (host checked the max count)
  const int chunks = n_pad >> 6;

  // ---- phase 1: sort every 64-element chunk in registers (k = 2 .. 64)
  for (int c = warp; c < chunks; c += 8) {
    const int e0 = (c << 6) + lane, e1 = e0 + 32;
    unsigned long long a = (e0 < count) ? binned[rg.x + e0] : ~0ull;
    unsigned long long b = (e1 < count) ? binned[rg.x + e1] : ~0ull;
#pragma unroll
    for (int k = 2; k <= 64; k <<= 1) {
      const bool asc0 = (e0 & k) == 0, asc1 = (e1 & k) == 0;
      if (k == 64 && ((a > b) == asc0)) {  // stride 32: the partner is this thread's other element
        const unsigned long long t = a;
        a = b;
        b = t;
      }
#pragma unroll
      for (int j = (k == 64 ? 16 : k >> 1); j > 0; j >>= 1) {
        const unsigned long long pa = shfl_xor_u64(a, j), pb = shfl_xor_u64(b, j);
        const bool lower = (lane & j) == 0;
        a = ((a < pa) == (lower == asc0)) ? a : pa;  // lower half keeps the min when ascending
        b = ((b < pb) == (lower == asc1)) ? b : pb;
      }
    }
    s_key[e0] = a;
    s_key[e1] = b;
  }
  // ---- phase 2: merges k = 128 .. n_pad; strides >= 64 through shared memory, the rest in registers
  for (int k = 128; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j >= 64; j >>= 1) {
      __syncthreads();
      for (int i = tid; i < (n_pad >> 1); i += 256) {
        const int l = ((i & ~(j - 1)) << 1) | (i & (j - 1));
        const int r = l + j;
        const unsigned long long a = s_key[l], b = s_key[r];
        if ((a > b) == ((l & k) == 0)) {
          s_key[l] = b;
          s_key[r] = a;
        }
      }
    }
    __syncthreads();
    for (int c = warp; c < chunks; c += 8) {
      const int e0 = (c << 6) + lane, e1 = e0 + 32;
      unsigned long long a = s_key[e0], b = s_key[e1];
      const bool asc = (e0 & k) == 0;
      if ((a > b) == asc) {
        const unsigned long long t = a;
        a = b;
        b = t;
      }
#pragma unroll
      for (int j = 16; j > 0; j >>= 1) {
        const unsigned long long pa = shfl_xor_u64(a, j), pb = shfl_xor_u64(b, j);
        const bool lower = (lane & j) == 0;
        a = ((a < pa) == (lower == asc)) ? a : pa;
        b = ((b < pb) == (lower == asc)) ? b : pb;
      }
      s_key[e0] = a;
      s_key[e1] = b;
    }
  }
  __syncthreads();
  for (int i = tid; i < count; i += 256) point_list[rg.x + i] = (int)(unsigned)s_key[i];
}

}  // namespace
