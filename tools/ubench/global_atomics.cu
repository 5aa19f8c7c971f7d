// Micro-benchmark: throughput of global atomics on per-tile counters as the binning passes issue them -- 1.6 M
// operations from 300k threads (5.4 each) on K distinct counters, each counter on its own 128-byte line (or its own
// 32-byte sector), returning (atomicSub) and not (RED).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o global_atomics global_atomics.cu && ./global_atomics
#include <cstdio>
#include <cuda_runtime.h>

template <bool RETURNING>
__global__ void __launch_bounds__(256) k_atomics(int n, int per, int K, int stride, int* counters, int* sink) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  unsigned h = t * 2654435761u;
  int acc = 0;
  for (int j = 0; j < per; ++j) {
    h = h * 1664525u + 1013904223u;
    int* c = counters + (size_t)((h >> 8) % (unsigned)K) * stride;
    if (RETURNING) acc += atomicSub(c, 1);
    else atomicAdd(c, 1);
  }
  if (RETURNING && acc == 0x7fffffff) sink[0] = acc;
}

int main() {
  const int n = 300000, per = 5;
  int *counters, *sink;
  cudaMalloc(&counters, (size_t)80000 * 32 * sizeof(int));
  cudaMalloc(&sink, 4);
  cudaMemset(counters, 0, (size_t)80000 * 32 * sizeof(int));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int Ks[] = {2500, 5000, 10000, 20000, 80000};
  const int strides[] = {32, 8, 1};
  for (int stride : strides)
    for (int K : Ks)
      for (int ret = 0; ret < 2; ++ret) {
        float best = 1e9f;
        for (int it = 0; it < 5; ++it) {
          cudaEventRecord(e0);
          if (ret) k_atomics<true><<<(n + 255) / 256, 256>>>(n, per, K, stride, counters, sink);
          else k_atomics<false><<<(n + 255) / 256, 256>>>(n, per, K, stride, counters, sink);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          if (ms < best) best = ms;
        }
        printf("stride %2d ints, %6d counters, %s: %7.1f us for %d atomics (%5.1f G/s, %6.0f per counter)\n", stride, K,
               ret ? "returning" : "RED      ", best * 1e3, n * per, n * per / (best * 1e-3) / 1e9, (double)n * per / K);
      }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
