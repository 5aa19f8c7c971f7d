// loss.cu -- the parts of loss.py the reference's step loop leaves commented out (train.py:967-974)
// but that belong to its operator surface:
//   ssim_kernel        replaces gaussian_kernel + ssim_kernel + ssim()   (loss.py:33-119, 178-215)
//   depth_loss_kernel  replaces depth_loss_kernel + depth_loss()        (loss.py:248-306)
// (The L1 loss and its sign gradient -- the loss the step loop does use -- live in optimizer.cu.)
//
// SSIM: one CTA per 16x16 pixel tile; the tile and its 5-pixel halo of both images are staged once
// in shared memory (26 x 26 x 6 floats), every thread then walks its 11x11 window out of it.  The
// reference launches one thread per pixel that reads 2 x 121 vec3 from global memory and adds its
// result into ONE address atomically; here a CTA adds one double.  The window is truncated at the
// image border and the weights are the reference's as written (loss.py:84 indexes the 11-tap kernel
// by |offset|, so the centre tap gets exp(-25/4.5) and the corner taps 1.0).
#include <math.h>

#include "common.cuh"

namespace {

constexpr int kWin = 11, kHalf = 5, kTs = 16, kHalo = kTs + 2 * kHalf;  // 26

struct SsimWeights {
  float g[kHalf + 1];  // gaussian_weights[0..5] of loss.py:45: exp(-(k - 5)^2 / (2 sigma^2))
};

__global__ void __launch_bounds__(256)
ssim_kernel(int W, int H, const float* __restrict__ rendered, const float* __restrict__ target, const SsimWeights wt,
            double* __restrict__ ssim_sum) {
  __shared__ float s_r[kHalo * kHalo * 3];
  __shared__ float s_t[kHalo * kHalo * 3];
  __shared__ double s_part[8];
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * kTs - kHalf, y0 = blockIdx.y * kTs - kHalf;
  for (int e = tid; e < kHalo * kHalo; e += 256) {
    const int hy = e / kHalo, hx = e - hy * kHalo;
    const int x = x0 + hx, y = y0 + hy;
    const bool in = x >= 0 && x < W && y >= 0 && y < H;
    const size_t p = in ? ((size_t)y * W + x) * 3 : 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      s_r[e * 3 + c] = in ? rendered[p + c] : 0.0f;
      s_t[e * 3 + c] = in ? target[p + c] : 0.0f;
    }
  }
  __syncthreads();
  const int lx = tid & 15, ly = tid >> 4;
  const int px = blockIdx.x * kTs + lx, py = blockIdx.y * kTs + ly;
  float ssim_val = 0.0f;
  if (px < W && py < H) {
    // per-axis weights of this pixel's window: 0 outside the image (the reference's loops skip those taps)
    float wx[kWin], wy[kWin];
#pragma unroll
    for (int k = 0; k < kWin; ++k) {
      const int d = k - kHalf, a = d < 0 ? -d : d;
      wx[k] = (px + d >= 0 && px + d < W) ? wt.g[a] : 0.0f;
      wy[k] = (py + d >= 0 && py + d < H) ? wt.g[a] : 0.0f;
    }
    float mu1[3] = {0.f, 0.f, 0.f}, mu2[3] = {0.f, 0.f, 0.f}, s1[3] = {0.f, 0.f, 0.f}, s2[3] = {0.f, 0.f, 0.f},
          s12[3] = {0.f, 0.f, 0.f};
    float weight_sum = 0.0f;
#pragma unroll
    for (int ky = 0; ky < kWin; ++ky) {
      if (wy[ky] == 0.0f) continue;
      const float* rr = s_r + ((ly + ky) * kHalo + lx) * 3;
      const float* tr = s_t + ((ly + ky) * kHalo + lx) * 3;
#pragma unroll
      for (int kx = 0; kx < kWin; ++kx) {
        if (wx[kx] == 0.0f) continue;
        const float w = wx[kx] * wy[ky];  // loss.py:84
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float p1 = rr[kx * 3 + c], p2 = tr[kx * 3 + c];
          mu1[c] = mu1[c] + p1 * w;
          mu2[c] = mu2[c] + p2 * w;
          s1[c] = s1[c] + (p1 * p1) * w;
          s2[c] = s2[c] + (p2 * p2) * w;
          s12[c] = s12[c] + (p1 * p2) * w;
        }
        weight_sum = weight_sum + w;
      }
    }
    const float c1 = 0.01f * 0.01f, c2 = 0.03f * 0.03f;
    float acc = 0.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float m1 = mu1[c], m2 = mu2[c], v1 = s1[c], v2 = s2[c], v12 = s12[c];
      if (weight_sum > 0.0f) {  // loss.py:97-102
        m1 = m1 / weight_sum;
        m2 = m2 / weight_sum;
        v1 = v1 / weight_sum;
        v2 = v2 / weight_sum;
        v12 = v12 / weight_sum;
      }
      v1 = v1 - m1 * m1;
      v2 = v2 - m2 * m2;
      v12 = v12 - m1 * m2;
      const float s = ((2.0f * m1 * m2 + c1) * (2.0f * v12 + c2)) / ((m1 * m1 + m2 * m2 + c1) * (v1 + v2 + c2));
      acc = acc + s;  // (ssim_r + ssim_g + ssim_b), loss.py:116
    }
    ssim_val = acc / 3.0f;
  }
  double v = (double)ssim_val;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((tid & 31) == 0) s_part[tid >> 5] = v;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_part[w];
    atomicAdd(ssim_sum, t);
  }
}

__global__ void __launch_bounds__(256)
depth_loss_kernel(long long count, const float* __restrict__ rendered, const float* __restrict__ target,
                  const float* __restrict__ mask, double* __restrict__ loss_sum) {
  __shared__ double s_part[8];
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    acc += (double)(fabsf(rendered[i] - target[i]) * mask[i]);  // loss.py:264
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_part[w];
    atomicAdd(loss_sum, t);
  }
}

}  // namespace

GSB_API int gsb_ssim(gsb_ctx* ctx, gsb_stream s_, int32_t width, int32_t height, const float* rendered,
                     const float* target, double* ssim_sum) {
  if (!ctx) return GSB_ERR_INVALID;
  GSB_REQUIRE(ctx, width > 0 && height > 0 && rendered && target && ssim_sum, "gsb_ssim: bad arguments");
  cudaStream_t s = (cudaStream_t)s_;
  GSB_CUDA(ctx, cudaMemsetAsync(ssim_sum, 0, sizeof(double), s));
  SsimWeights wt;
  const float sigma = 1.5f;
  for (int k = 0; k <= kHalf; ++k) {  // loss.py:33-45, binary32 like the kernel that fills the array
    const int x = k - kWin / 2;
    wt.g[k] = expf(-1.0f * (float)(x * x) / (2.0f * sigma * sigma));
  }
  dim3 grid((width + kTs - 1) / kTs, (height + kTs - 1) / kTs);
  GSB_LAUNCH(ctx, ssim_kernel, grid, 256, 0, s, width, height, rendered, target, wt, ssim_sum);
  return GSB_OK;
}

GSB_API int gsb_depth_loss(gsb_ctx* ctx, gsb_stream s_, int64_t count, const float* rendered_depth,
                           const float* target_depth, const float* depth_mask, double* loss_sum) {
  if (!ctx) return GSB_ERR_INVALID;
  GSB_REQUIRE(ctx, count >= 0 && loss_sum, "gsb_depth_loss: bad arguments");
  cudaStream_t s = (cudaStream_t)s_;
  GSB_CUDA(ctx, cudaMemsetAsync(loss_sum, 0, sizeof(double), s));
  if (count == 0) return GSB_OK;
  long long blocks = gsb_div_up(count, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  GSB_LAUNCH(ctx, depth_loss_kernel, (unsigned)blocks, 256, 0, s, (long long)count, rendered_depth, target_depth,
             depth_mask, loss_sum);
  return GSB_OK;
}
