// Micro-benchmark: issue cost of warp-collective instructions on sm_100a with 32 warps per SM (8 per scheduler),
// the configuration of the cooperative radix sort.  Prints cycles per warp-instruction per scheduler.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o warp_ops warp_ops.cu && ./warp_ops
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096;

template <int OP>
__global__ void __launch_bounds__(1024, 1) bench(unsigned* out, long long* cycles, unsigned seed) {
  const int lane = threadIdx.x & 31;
  unsigned x = seed * (threadIdx.x + 1) + blockIdx.x, acc = 0;
  __shared__ unsigned sh[1024];
  sh[threadIdx.x] = x;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < kIters; ++i) {
    if (OP == 0) {          // VOTE (ballot) on a data-dependent predicate
      acc += __ballot_sync(0xffffffffu, (x >> (i & 31)) & 1u);
    } else if (OP == 1) {   // MATCH.ANY
      acc += __match_any_sync(0xffffffffu, (x + i) & 0xffu);
    } else if (OP == 2) {   // SHFL.IDX
      acc += __shfl_sync(0xffffffffu, x + i, (lane + i) & 31);
    } else if (OP == 3) {   // POPC
      acc += __popc(x + i * 0x9e3779b9u);
    } else if (OP == 4) {   // FLO (ffs)
      acc += __ffs(x + i * 0x9e3779b9u);
    } else if (OP == 5) {   // REDUX.SUM
      acc += __reduce_add_sync(0xffffffffu, x + i);
    } else if (OP == 6) {   // LOP3 baseline (ALU pipe)
      acc = (acc ^ (x + i)) & (acc | 0x55555555u + i);
    } else if (OP == 7) {   // LDS, conflict-free
      acc += sh[(threadIdx.x + i) & 1023];
    } else if (OP == 8) {   // ATOMS.OR on distinct addresses
      atomicOr(&sh[(threadIdx.x + i) & 1023], 1u << (i & 31));
    } else if (OP == 9) {   // ATOMS.OR, all lanes of a warp on the same address
      atomicOr(&sh[(threadIdx.x >> 5) + (i & 7)], 1u << lane);
    } else if (OP == 10) {  // ATOMS.ADD returning, distinct addresses
      acc += atomicAdd(&sh[(threadIdx.x + i) & 1023], 1u);
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + sh[lane];
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, unsigned* out, long long* cyc) {
  bench<OP><<<148, 1024>>>(out, cyc, 12345u);
  bench<OP><<<148, 1024>>>(out, cyc, 12345u);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < 148; ++i) s += (double)h[i];
  s /= 148.0;
  // 8 warps per scheduler each issue kIters instructions of the op (plus a few ALU instructions per iteration)
  printf("%-34s %8.2f cycles per warp-instruction per scheduler (loop incl. ~2-3 ALU ops)\n", name, s / (kIters * 8.0));
}

int main() {
  unsigned* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(unsigned));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  run<6>("LOP3 pair (baseline)", out, cyc);
  run<0>("VOTE.ANY (ballot)", out, cyc);
  run<1>("MATCH.ANY", out, cyc);
  run<2>("SHFL.IDX", out, cyc);
  run<3>("POPC", out, cyc);
  run<4>("FLO (ffs)", out, cyc);
  run<5>("REDUX.SUM", out, cyc);
  run<7>("LDS", out, cyc);
  run<8>("ATOMS.OR distinct", out, cyc);
  run<9>("ATOMS.OR same address per warp", out, cyc);
  run<10>("ATOMS.ADD returning, distinct", out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
