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
  // cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: done once per context, not per process
  bool smem_optin_blend_bwd = false, smem_optin_radix = false, smem_optin_tilesort = false, smem_optin_radix_coop = false;
  bool coop_launch = false;   // cudaDevAttrCooperativeLaunch
  // A/B knobs of gsb_set_option (per context; results never depend on them)
  struct Options {
    int binning = 0;     // 0: per-tile counting sort + shared-memory sort (default); 1: global 64-bit radix sort
    int blend_cull = 1;  // per-block culling masks in the tile kernels
    int tile_sort = 3;   // 3 (default): one-pass bucket sort by depth (bitonic for tiles > 4096); 0: bitonic network for every tile; 1: per-tile LSD radix sort (bitonic for tiles > 4096); 2: radix when the longest list > 2048
    int bwd_reduce = 2;  // 2 (1 is accepted as a synonym): tensor-core pixel sums; 0: warp-shuffle butterfly with the exact exponential
    int pdl = 1;         // 1: the kernels of a frame / training step are launched as programmatic dependents of each other (GSB_LAUNCH_PDL); 0: plain launches
    int speculate = 1;   // 1: gsb_forward queues scatter / sort / blend behind the scan without waiting for D (checked on the device); 0: waits first
    int sort_coop = 1;   // 1: gsb_sort_pairs64 sorts inputs of up to num_sms x 12288 pairs in one cooperative launch; 0: three kernels per pass
    int bwd_packed = 1;  // 1: the tensor-core backward accumulates into packed records with vector REDs; 0: nine scalar REDs
  } opt;

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
  uint32_t* sort_coop_state = nullptr;  // cooperative sort: grid-barrier counters + digit totals (zero between launches)
  // generic scan scratch
  int32_t* scan_sums = nullptr;
  int64_t scan_cap = 0;
  // tile-binning scratch: [num_tiles] counts + [num_tiles] write cursors
  int32_t* tile_count = nullptr;
  int64_t tile_cap = 0;
  int64_t tile_clean = 0;  // leading ints of tile_count known to be zero: the scatter pass counts every counter
                           // back down, so a steady-state frame needs no memset
  // per-Gaussian internal buffers for gsb_forward / gsb_backward
  int32_t* tiles_touched = nullptr;
  float* dcov3d = nullptr;
  float* bwd_acc = nullptr;      // [12 n] packed accumulation records of the backward tile kernel
  float* bwd_acc_stage = nullptr;  // the same for the stage-level gsb_blend_backward (grown on demand)
  int64_t bwd_acc_cap = 0;
  int32_t* rank_base = nullptr;  // [n] scratch: the inclusive scan of tiles_touched when the caller passed no point_offsets (global-sort path)
  int64_t n_cap = 0;
  int64_t last_num_rendered = 0;   // D and the longest tile list of the previous gsb_forward: what the next frame's
  int last_max_count = 0;          // speculative launch assumes
  void* color_event = nullptr;     // gsb_set_color_dependency: one-shot, consumed by the next gsb_forward
  cudaStream_t side_stream = nullptr;   // gsb_forward: the point_offsets scan beside the speculative kernels
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_count = nullptr;  // recorded behind the read-back of D: the host waits on it, not on the stream
  // device + pinned host scalars
  int32_t* d_scalars = nullptr;  // [16]
  int32_t* h_scalars = nullptr;  // pinned [16]
  double* d_accum = nullptr;     // [2] loss accumulator + CTA ticket (8 bytes); self-resetting, see l1_loss_grad_kernel
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

// Programmatic dependent launch (sm_90+): a kernel launched with GSB_LAUNCH_PDL may be brought onto the SMs while the
// last CTAs of the kernel in front of it in the stream are still running; its threads call gsb_pdl_wait() FIRST -- it
// returns when that kernel has completed and its writes are visible -- so the launch latency and the prologue are
// hidden, never a data dependency.  gsb_pdl_launch_dependents() (called right behind the wait) lets the kernel behind
// this one do the same.  Launched without the attribute both are no-ops.  Only kernels that begin with gsb_pdl_wait()
// may be launched with GSB_LAUNCH_PDL; everything else keeps the stream's full serialisation.
// Measured (kbench, step loop, B200): 705.5 us without, 694.7 us with every kernel of the step launched this way except
// preprocess_backward_kernel -- that one, brought in behind the backward tile kernel, costs 7 us instead of saving any
// (705.3 against 698.8 with only the forward's kernels dependent; with ALL kernels dependent the step was 722 us), so
// it keeps the plain launch; its own hooks stay, for the Adam kernel behind it.
#ifdef __CUDACC__
__device__ __forceinline__ void gsb_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void gsb_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

#define GSB_LAUNCH_PDL(ctx, kernel, grid, block, smem, stream_, ...)                                         \
  do {                                                                                                \
    if ((ctx)->opt.pdl) {                                                                                         \
      cudaLaunchConfig_t _cfg = {};                                                                   \
      _cfg.gridDim = dim3(grid);                                                                      \
      _cfg.blockDim = dim3(block);                                                                    \
      _cfg.dynamicSmemBytes = (smem);                                                                 \
      _cfg.stream = (stream_);                                                                        \
      cudaLaunchAttribute _at[1];                                                                     \
      _at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                 \
      _at[0].val.programmaticStreamSerializationAllowed = 1;                                          \
      _cfg.attrs = _at;                                                                               \
      _cfg.numAttrs = 1;                                                                              \
      (ctx)->launches += 1;                                                                           \
      int _rc = gsb_check_cuda((ctx), cudaLaunchKernelEx(&_cfg, kernel, __VA_ARGS__), #kernel);       \
      if (_rc != GSB_OK) return _rc;                                                                  \
    } else {                                                                                          \
      GSB_LAUNCH(ctx, kernel, grid, block, smem, stream_, __VA_ARGS__);                               \
    }                                                                                                 \
  } while (0)

// Device-side bounds checks of the index arithmetic the kernels rely on (shared-memory slots, list positions, scatter
// targets).  Compiled in only by `make variant TAG=chk EXTRA=-DGSB_DEBUG_CHECKS=1` (tools/sanitize_case.py runs the small
// end-to-end cases on that build: compute-sanitizer is closed on this pool); a failed check traps the kernel and the
// next CUDA call returns cudaErrorAssert.
#ifdef GSB_DEBUG_CHECKS
#include <assert.h>
#define GSB_DCHECK(cond) assert(cond)
#else
#define GSB_DCHECK(cond) ((void)0)
#endif

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
  // 2^n without a conversion (F2I is an XU-pipe instruction: 8 cycles per scheduler, tools/ubench/warp_ops.cu): n
  // sits in the low mantissa bits of t + 1.5 * 2^23 (two's complement) and the constant's own bits vanish in the
  // shift -- (n + 127) << 23 for every n in [-126, 127].  Same bits as (__float2int_rz(n) + 127) << 23: forward tile
  // kernel 217.1 -> 214.5 us with identical n_contrib / final_T / image.
  float scale = __int_as_float((__float_as_int(__fadd_rn(t, 12582912.0f)) << 23) + 0x3f800000);
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

// ---- packed binary32 pairs (FFMA2, sm_100) -----------------------------------------------------
// Blackwell issues one FFMA2 for two independent fused multiply-adds on 64-bit register pairs.  The
// tile kernels are instruction-issue bound, so arithmetic that comes in pairs is packed.  ptxas
// contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even at -fmad=false, so products and sums are
// written as FMAs themselves: a*b + (-0) and a*1 + b are exactly RN(a*b) and RN(a+b) (signed zeros
// included), and an FMA cannot be contracted further.  Every lane rounds as the scalar code does.
typedef unsigned long long gs_f2;
__device__ __forceinline__ gs_f2 gs_pack2(float lo, float hi) {
  gs_f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void gs_unpack2(gs_f2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ gs_f2 gs_fma2(gs_f2 a, gs_f2 b, gs_f2 c) {
  gs_f2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ gs_f2 gs_mul2(gs_f2 a, gs_f2 b) { return gs_fma2(a, b, gs_pack2(-0.0f, -0.0f)); }
__device__ __forceinline__ gs_f2 gs_add2(gs_f2 a, gs_f2 b) { return gs_fma2(a, gs_pack2(1.0f, 1.0f), b); }
__device__ __forceinline__ gs_f2 gs_splat2(float c) { return gs_pack2(c, c); }

// gs_power with the x / y halves packed: gxy = (gx, gy), npxy = (-px, -py), cac = (conic.a, conic.c).
// Same operations in the same order as gs_power => the same bits.
__device__ __forceinline__ float gs_power_packed(gs_f2 gxy, gs_f2 npxy, gs_f2 cac, float cb) {
  const gs_f2 d = gs_add2(gxy, npxy);            // (dx, dy)
  const gs_f2 t = gs_mul2(gs_mul2(cac, d), d);   // ((a*dx)*dx, (c*dy)*dy)
  float dx, dy, t0, t1;
  gs_unpack2(d, dx, dy);
  gs_unpack2(t, t0, t1);
  const float q = __fmul_rn(-0.5f, __fadd_rn(t0, t1));
  const float t2 = __fmul_rn(__fmul_rn(cb, dx), dy);
  return __fsub_rn(q, t2);
}

// gs_expf of two arguments at once (lo, hi): the same operation sequence per half => the same bits.
__device__ __forceinline__ void gs_expf2(float x0, float x1, float& y0, float& y1) {
  const gs_f2 x = gs_pack2(x0, x1);
  const gs_f2 t = gs_mul2(x, gs_splat2(1.44269504088896341f));
  const gs_f2 tm = gs_add2(t, gs_splat2(12582912.0f));
  const gs_f2 n = gs_add2(tm, gs_splat2(-12582912.0f));
  gs_f2 r = gs_fma2(n, gs_splat2(-0.693145751953125f), x);
  r = gs_fma2(n, gs_splat2(-1.42860682030941723212e-6f), r);
  gs_f2 p = gs_splat2(1.9875691500e-4f);
  p = gs_fma2(p, r, gs_splat2(1.3981999507e-3f));
  p = gs_fma2(p, r, gs_splat2(8.3334519073e-3f));
  p = gs_fma2(p, r, gs_splat2(4.1665795894e-2f));
  p = gs_fma2(p, r, gs_splat2(1.6666665459e-1f));
  p = gs_fma2(p, r, gs_splat2(5.0000001201e-1f));
  const gs_f2 r2 = gs_mul2(r, r);
  gs_f2 y = gs_fma2(p, r2, r);
  y = gs_add2(y, gs_splat2(1.0f));
  float m0, m1;   // see gs_expf: 2^n from the bits of t + 1.5 * 2^23, no F2I
  gs_unpack2(tm, m0, m1);
  const float s0 = __int_as_float((__float_as_int(m0) << 23) + 0x3f800000);
  const float s1 = __int_as_float((__float_as_int(m1) << 23) + 0x3f800000);
  y = gs_mul2(y, gs_pack2(s0, s1));
  gs_unpack2(y, y0, y1);
  if (x0 < -87.0f) y0 = 0.0f;
  if (x1 < -87.0f) y1 = 0.0f;
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

// The 16 real SH basis values of sh_backward_kernel (backward.py:127-213) at the unit direction
// (x, y, z), zero above `deg`: dL_dshs[16 i + k] = basis[k] * dL_dRGB.  One definition for the two places
// that evaluate it (preprocess_bwd.cu and the compact SH-gradient exchange in optimizer.cu), so both
// produce the same bits.
__device__ __forceinline__ void gs_sh_basis(const int deg, const float x, const float y, const float z, float basis[16]) {
#pragma unroll
  for (int k = 0; k < 16; ++k) basis[k] = 0.0f;
  const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
  basis[0] = GS_SH_C0;
  if (deg > 0) {
    basis[1] = -GS_SH_C1 * y;
    basis[2] = GS_SH_C1 * z;
    basis[3] = -GS_SH_C1 * x;
    if (deg > 1) {
      basis[4] = GS_C2_0 * xy;
      basis[5] = GS_C2_1 * yz;
      basis[6] = GS_C2_2 * (2.0f * zz - xx - yy);
      basis[7] = GS_C2_3 * xz;
      basis[8] = GS_C2_4 * (xx - yy);
      if (deg > 2) {
        basis[9] = GS_C3_0 * y * (3.0f * xx - yy);
        basis[10] = GS_C3_1 * xy * z;
        basis[11] = GS_C3_2 * y * (4.0f * zz - xx - yy);
        basis[12] = GS_C3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
        basis[13] = GS_C3_4 * x * (4.0f * zz - xx - yy);
        basis[14] = GS_C3_5 * z * (xx - yy);
        basis[15] = GS_C3_6 * x * (xx - 3.0f * yy);
      }
    }
  }
}

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

// Counting pass of the tile binning fused into preprocess (gsb_forward only): the per-tile counters.
struct PreBin {
  int32_t* tile_count;
  int defer_color;   // 1: rgb / clamped_state are left to gsb_sh_color_impl (the SH rows are not read)
};
int gsb_sh_color_impl(gsb_ctx* ctx, cudaStream_t s, const gsb_frame* f, int32_t n, const float* means, const float* shs,
                      const int32_t* radii, float* rgb, float* clamped_state);
int gsb_preprocess_impl(gsb_ctx* ctx, cudaStream_t s, const gsb_frame* f, int32_t n, const float* means,
                        const float* scales, const float* rotations, const float* opacities, const float* shs,
                        int32_t* radii, float* points_xy, float* depths, float* cov3Ds, float* rgb,
                        float* conic_opacity, int32_t* tiles_touched, float* clamped_state, const PreBin* bin);

// gsb_preprocess_backward with the choice of SH-gradient output: the full [48 n] array, or (sh_compact
// = 1) the rank-1 factors it is the outer product of, [8 n]: (dL_dRGB masked, unit direction, 0, 0).
int gsb_preprocess_backward_impl(gsb_ctx* ctx, cudaStream_t s, const gsb_frame* f, int32_t n, const float* means,
                                 const int32_t* radii, const float* shs, const float* scales, const float* rotations,
                                 const float* cov3Ds, const float* clamped_state, const float* dL_dmean2D,
                                 const float* dL_dconic, const float* dL_dcolor, float* dL_dmean3D, float* dL_dshs,
                                 float* dL_dscale, float* dL_drot, float* dL_dcov3D_internal, int sh_compact,
                                 const float* packed /* may be null: see preprocess_backward_kernel<.., PACKED> */,
                                 float* dL_dopacity_out, float* zero_cov3D /* may be null: [6 n] written as zeros */);
// the backward tile kernel accumulating into the packed 12-float records of `packed` (zeroed here)
int gsb_blend_backward_packed(gsb_ctx* ctx, cudaStream_t s, const gsb_frame* f, int32_t n, const int32_t* ranges,
                              const int32_t* point_list, const float* points_xy, const float* conic_opacity,
                              const float* rgb, const float* final_T, const int32_t* n_contrib,
                              const float* dL_dpixels, float* packed, const int32_t* block_masks);
bool gsb_blend_backward_uses_packed(const gsb_ctx* ctx);

// ---- stage launchers implemented across the .cu files (host side) --------------------------
int gsb_scan_i32(gsb_ctx* ctx, cudaStream_t s, int64_t n, const int32_t* in, int32_t* out, bool exclusive,
                 int32_t* d_last /*device, may be null: receives out[n-1] (+in[n-1] if inclusive ==> total)*/);
