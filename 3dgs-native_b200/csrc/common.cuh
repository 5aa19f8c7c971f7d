// Internal header of libgsb200: context, launch bookkeeping and the device-side math that
// implements the arithmetic contract:
//   * every reference expression is evaluated in binary32, in Python's parse order, each
//     operation individually rounded.  All translation units are compiled with -fmad=false (no
//     FMA contraction), default -prec-div/-prec-sqrt (IEEE division and sqrt), no fast-math,
//     -ftz=false.
//   * float->int conversions truncate and saturate (cvt.rzi.s32.f32), NaN -> 0
//   * exp() is gs_expf below: a fixed sequence of IEEE operations, so that decisions that hang
//     on it (alpha < 1/255, T < 1e-4 -> n_contrib) are reproducible bit for bit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gsb200.h"

#define GSB_API extern "C" __attribute__((visibility("default")))

constexpr int kTile = GSB_TILE;

struct gsb_ctx {
  int device = 0;
  int num_sms = 148;
  char err[512] = {0};
  int64_t launches = 0;

  // binning scratch (grow-only): sort double buffers
  int64_t* keys_a = nullptr;
  int64_t* keys_b = nullptr;
  int32_t* vals_a = nullptr;
  int32_t* vals_b = nullptr;
  int64_t bin_cap = 0;
  // radix-sort block histogram table + counters
  uint32_t* sort_table = nullptr;
  int64_t sort_table_cap = 0;  // in uint32 entries
  uint32_t* sort_small = nullptr;  // [8*256] global digit bases + counters
  // generic scan scratch
  int32_t* scan_sums = nullptr;
  int64_t scan_cap = 0;
  // tile-binning scratch: [num_tiles] counts + [num_tiles] write cursors
  int32_t* tile_count = nullptr;
  int64_t tile_cap = 0;
  // per-Gaussian internal buffers for gsb_forward / gsb_backward
  int32_t* tiles_touched = nullptr;
  float* dcov3d = nullptr;
  int64_t n_cap = 0;
  // device + pinned host scalars
  int32_t* d_scalars = nullptr;  // [16]
  int32_t* h_scalars = nullptr;  // pinned [16]
};

int gsb_set_error(gsb_ctx* ctx, int code, const char* fmt, ...);
int gsb_check_cuda(gsb_ctx* ctx, cudaError_t e, const char* what);
int gsb_grow(gsb_ctx* ctx, void** ptr, int64_t* cap, int64_t need_elems, size_t elem_size, cudaStream_t s);
int gsb_reserve_binning(gsb_ctx* ctx, cudaStream_t s, int64_t num_rendered);

#define GSB_CUDA(ctx, call)                                   \
  do {                                                        \
    int _rc = gsb_check_cuda((ctx), (call), #call);           \
    if (_rc != GSB_OK) return _rc;                            \
  } while (0)

#define GSB_LAUNCH(ctx, kernel, grid, block, smem, stream, ...)                       \
  do {                                                                                \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                       \
    (ctx)->launches += 1;                                                             \
    int _rc = gsb_check_cuda((ctx), cudaGetLastError(), #kernel);                     \
    if (_rc != GSB_OK) return _rc;                                                    \
  } while (0)

#define GSB_REQUIRE(ctx, cond, msg)                                        \
  do {                                                                     \
    if (!(cond)) return gsb_set_error((ctx), GSB_ERR_INVALID, "%s", msg);  \
  } while (0)

static inline bool gsb_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int64_t gsb_div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------------------------------
// device math
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ float f_min(float a, float b) { return (a < b) ? a : b; }  // wp.min
__device__ __forceinline__ float f_max(float a, float b) { return (a > b) ? a : b; }  // wp.max
__device__ __forceinline__ int f2i(float x) { return __float2int_rz(x); }             // int(x)

// Deterministic exp (arithmetic contract): a fixed sequence of IEEE operations.
__device__ __forceinline__ float gs_expf(float x) {
  if (x < -87.0f) return 0.0f;
  float t = __fmul_rn(x, 1.44269504088896341f);
  float n = __fsub_rn(__fadd_rn(t, 12582912.0f), 12582912.0f);
  float r = __fmaf_rn(n, -0.693145751953125f, x);
  r = __fmaf_rn(n, -1.42860682030941723212e-6f, r);
  float p = 1.9875691500e-4f;
  p = __fmaf_rn(p, r, 1.3981999507e-3f);
  p = __fmaf_rn(p, r, 8.3334519073e-3f);
  p = __fmaf_rn(p, r, 4.1665795894e-2f);
  p = __fmaf_rn(p, r, 1.6666665459e-1f);
  p = __fmaf_rn(p, r, 5.0000001201e-1f);
  float r2 = __fmul_rn(r, r);
  float y = __fmaf_rn(p, r2, r);
  y = __fadd_rn(y, 1.0f);
  int ni = __float2int_rz(n);
  float scale = __int_as_float((ni + 127) << 23);
  return __fmul_rn(y, scale);
}

// Gaussian exponent of forward.py:471 / backward.py:644, explicit rounding of every operation:
//   power = -0.5*(a*dx*dx + c*dy*dy) - b*dx*dy
__device__ __forceinline__ float gs_power(float ca, float cb, float cc, float dx, float dy) {
  float t0 = __fmul_rn(__fmul_rn(ca, dx), dx);
  float t1 = __fmul_rn(__fmul_rn(cc, dy), dy);
  float q = __fmul_rn(-0.5f, __fadd_rn(t0, t1));
  float t2 = __fmul_rn(__fmul_rn(cb, dx), dy);
  return __fsub_rn(q, t2);
}

// Conservative skip threshold on `power` for one Gaussian: power < thr  ==>  alpha < 1/255
// (margin 1e-3 in the exponent dwarfs every rounding error of gs_expf and logf).  Anything not
// provably skippable (incl. NaN opacity) takes the exact path, so results never depend on it.
__device__ __forceinline__ float gs_power_threshold(float opacity) {
  if (opacity < (1.0f / 255.0f)) return __int_as_float(0x7f800000);  // +inf: can never reach 1/255
  return -logf(255.0f * opacity) - 1e-3f;
}

// forward.py:63-76 get_rect (grid as floats, truncating casts)
__device__ __forceinline__ void gs_get_rect(float px, float py, float max_radius, float grid_x, float grid_y,
                                            int& min_x, int& min_y, int& max_x, int& max_y) {
  min_x = min(f2i(grid_x), max(0, f2i((px - max_radius) / 16.0f)));
  min_y = min(f2i(grid_y), max(0, f2i((py - max_radius) / 16.0f)));
  max_x = min(f2i(grid_x), max(0, f2i((px + max_radius + 16.0f - 1.0f) / 16.0f)));
  max_y = min(f2i(grid_y), max(0, f2i((py + max_radius + 16.0f - 1.0f) / 16.0f)));
}

// [Warp] row-vector * mat44 (row-major m), accumulation order i = 0..3
__device__ __forceinline__ void gs_vec4_mul_mat44(float v0, float v1, float v2, float v3, const float* m, float r[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float s = m[0 * 4 + j] * v0;
    s = s + m[1 * 4 + j] * v1;
    s = s + m[2 * 4 + j] * v2;
    s = s + m[3 * 4 + j] * v3;
    r[j] = s;
  }
}

// [Warp] mat33 * mat33: t[i][j] = 0; for k: t[i][j] += a[i][k]*b[k][j]
__device__ __forceinline__ void gs_mat33_mul(const float a[9], const float b[9], float t[9]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float s = 0.0f;
#pragma unroll
      for (int k = 0; k < 3; ++k) s = s + a[i * 3 + k] * b[k * 3 + j];
      t[i * 3 + j] = s;
    }
}
__device__ __forceinline__ void gs_mat33_transpose(const float a[9], float t[9]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) t[i * 3 + j] = a[j * 3 + i];
}
__device__ __forceinline__ float gs_dot3(float a0, float a1, float a2, float b0, float b1, float b2) {
  float s = a0 * b0;
  s = s + a1 * b1;
  s = s + a2 * b2;
  return s;
}

// SH constants (forward.py:44-45,330-344; backward.py:158-162,190-196)
#define GS_SH_C0 0.28209479177387814f
#define GS_SH_C1 0.4886025119029199f
#define GS_C2_0 1.0925484305920792f
#define GS_C2_1 (-1.0925484305920792f)
#define GS_C2_2 0.31539156525252005f
#define GS_C2_3 (-1.0925484305920792f)
#define GS_C2_4 0.5462742152960396f
#define GS_C3_0 (-0.5900435899266435f)
#define GS_C3_1 2.890611442640554f
#define GS_C3_2 (-0.4570457994644658f)
#define GS_C3_3 0.3731763325901154f
#define GS_C3_4 (-0.4570457994644658f)
#define GS_C3_5 1.445305721320277f
#define GS_C3_6 (-0.5900435899266435f)

// Per-view constants in kernel-parameter space
struct FrameK {
  float view[16];
  float proj[16];
  float campos[3];
  float tan_fovx, tan_fovy, scale_modifier;
  float bg[3];
  int W, H, degree, clamped;
  float focal_x, focal_y;  // W/(2 tan_fovx), H/(2 tan_fovy) evaluated in double on the host, then cast
  int grid_x, grid_y;
};

#endif  // __CUDACC__

struct FrameK;
void gsb_make_framek(const gsb_frame* f, FrameK* k);

// ---- stage launchers implemented across the .cu files (host side) --------------------------
int gsb_scan_i32(gsb_ctx* ctx, cudaStream_t s, int64_t n, const int32_t* in, int32_t* out, bool exclusive,
                 int32_t* d_last /*device, may be null: receives out[n-1] (+in[n-1] if inclusive ==> total)*/);
