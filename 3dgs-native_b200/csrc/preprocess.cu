// preprocess.cu -- replaces wp_preprocess (reference forward.py:189-382).
//
// One thread per Gaussian, 128 Gaussians per CTA.  The 192-byte SH rows of the CTA's Gaussians
// are one contiguous 24 KB span: the CTA streams it in with coalesced 16-byte loads into a
// padded shared-memory tile (row stride 49 words => conflict-free row reads), instead of every
// thread walking its own 192-byte row with 32 wavefronts per load.  All other per-Gaussian
// inputs are contiguous spans read with stride-12/16 accesses whose sectors are fully used.
// Every output element is written (zeros for culled Gaussians), so no memset precedes it.
//
// HBM roofline: reads 236 B + writes 84 B per Gaussian (320*N).
#include "common.cuh"

namespace {

constexpr int kPreThreads = 128;
constexpr int kShStride = 49;

constexpr int kCntStride = 32;  // ints between two tiles' counters (tilesort.cu: one 128-byte line each)

// SH -> RGB of one Gaussian (forward.py:304-362): sh = its 16 x 3 coefficients (row stride 3), o_rgb the colour
// (clamped at 0 when f.clamped), o_cl the clamp flags.  Shared by preprocess_kernel and sh_color_kernel: one
// expression order, one translation unit (-fmad=false), the same bits.
__device__ __forceinline__ void gs_sh_to_rgb(const FrameK& f, const float px, const float py, const float pz,
                                             const float* sh, float o_rgb[3], float o_cl[3]) {
    float dx = px - f.campos[0], dy = py - f.campos[1], dz = pz - f.campos[2];
    float len = sqrtf(gs_dot3(dx, dy, dz, dx, dy, dz));
    float x = 0.f, y = 0.f, z = 0.f;  // [Warp] normalize: v/len if len > 0 else 0
    if (len > 0.0f) {
      x = dx / len;
      y = dy / len;
      z = dz / len;
    }
    float result[3];
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#define SHK(k) sh[(k) * 3 + c]
      float r = GS_SH_C0 * SHK(0);
      if (f.degree > 0) {
        r = r - GS_SH_C1 * y * SHK(1) + GS_SH_C1 * z * SHK(2) - GS_SH_C1 * x * SHK(3);
        if (f.degree > 1) {
          r = r + GS_C2_0 * xy * SHK(4);
          r = r + GS_C2_1 * yz * SHK(5);
          r = r + GS_C2_2 * (2.0f * zz - xx - yy) * SHK(6);
          r = r + GS_C2_3 * xz * SHK(7);
          r = r + GS_C2_4 * (xx - yy) * SHK(8);
          if (f.degree > 2) {
            r = r + GS_C3_0 * y * (3.0f * xx - yy) * SHK(9);
            r = r + GS_C3_1 * xy * z * SHK(10);
            r = r + GS_C3_2 * y * (4.0f * zz - xx - yy) * SHK(11);
            r = r + GS_C3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * SHK(12);
            r = r + GS_C3_4 * x * (4.0f * zz - xx - yy) * SHK(13);
            r = r + GS_C3_5 * z * (xx - yy) * SHK(14);
            r = r + GS_C3_6 * x * (xx - 3.0f * yy) * SHK(15);
          }
        }
      }
#undef SHK
      r = r + 0.5f;
      result[c] = r;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      o_cl[c] = (result[c] < 0.0f) ? 1.0f : 0.0f;  // forward.py:351-362
      o_rgb[c] = f.clamped ? f_max(result[c], 0.0f) : result[c];
    }
}

// BIN = true (gsb_forward): the kernel also runs the counting pass of the tile binning -- one RED per
// (Gaussian, tile) on the tile's counter -- instead of a second kernel that re-derives every rectangle.
// The pass needs no prefix sum of tiles_touched; that scan (an output of the operator) runs off the
// critical path.
// COLOR = false (gsb_forward with a colour dependency, see gsb_set_color_dependency): the SH rows are neither loaded
// nor evaluated and rgb / clamped_state are left to sh_color_kernel.
template <bool BIN, bool COLOR>
__global__ void __launch_bounds__(kPreThreads)
preprocess_kernel(const FrameK f, const int n, const float* __restrict__ means, const float* __restrict__ scales,
                  const float* __restrict__ rots, const float* __restrict__ opac, const float* __restrict__ shs,
                  int* __restrict__ radii, float2* __restrict__ xy_out, float* __restrict__ depths,
                  float* __restrict__ cov3Ds, float* __restrict__ rgb, float4* __restrict__ conic_opacity,
                  int* __restrict__ tiles_touched, float* __restrict__ clamped_state, const PreBin bin) {
  __shared__ float s_sh[kPreThreads * kShStride];
  gsb_pdl_wait();
  gsb_pdl_launch_dependents();
  const int base = blockIdx.x * kPreThreads;
  const int tid = threadIdx.x;
  const int rows = min(kPreThreads, n - base);

  // This thread's own inputs first, so that they are in flight together with the SH tile (the loads
  // below are independent; issued after the barrier they would cost a second DRAM round trip).
  const int i = base + tid;
  const bool live = i < n;
  const int il = live ? i : base;  // clamped: dead threads of the last CTA re-read row 0 and store nothing
  const float px = means[3 * il + 0], py = means[3 * il + 1], pz = means[3 * il + 2];
  const float4 q = __ldg(reinterpret_cast<const float4*>(rots) + il);  // (x,y,z,w)
  const float sc0 = scales[3 * il + 0], sc1 = scales[3 * il + 1], sc2 = scales[3 * il + 2];
  const float opacity = __ldg(opac + il);

  // cooperative, coalesced SH load (12 float4 per Gaussian), all twelve in flight
  if (COLOR) {
    const float4* src = reinterpret_cast<const float4*>(shs + (size_t)base * 48);
    const int chunks = rows * 12;
    float4 v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const int c = tid + k * kPreThreads;
      v[k] = (c < chunks) ? __ldg(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const int c = tid + k * kPreThreads;
      if (c < chunks) {
        const int g = c / 12, qd = c - g * 12;
        float* d = s_sh + g * kShStride + qd * 4;
        d[0] = v[k].x;
        d[1] = v[k].y;
        d[2] = v[k].z;
        d[3] = v[k].w;
      }
    }
  }
  __syncthreads();
  if (!live) return;

  // outputs default to zero: forward.py:703-710 allocates them with wp.zeros
  int o_radius = 0, o_tiles = 0;
  float o_x = 0.f, o_y = 0.f, o_depth = 0.f;
  float o_cov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float o_rgb[3] = {0.f, 0.f, 0.f};
  float o_cl[3] = {0.f, 0.f, 0.f};
  float4 o_con = make_float4(0.f, 0.f, 0.f, 0.f);
  int rminx = 0, rminy = 0, rmaxx = 0, rmaxy = 0;

  do {
    if (BIN && !live) break;
    float p_view[4];
    gs_vec4_mul_mat44(px, py, pz, 1.0f, f.view, p_view);
    if (p_view[2] < 0.2f) break;  // forward.py:250

    float p_hom[4];
    gs_vec4_mul_mat44(px, py, pz, 1.0f, f.proj, p_hom);
    float p_w = 1.0f / (p_hom[3] + 0.0000001f);
    float p_proj_x = p_hom[0] * p_w, p_proj_y = p_hom[1] * p_w;

    // ---- compute_cov3d, forward.py:146-186 ([Warp] quat_to_matrix via quat_rotate) ----
    float cov3d[6];
    {
      const float s0 = f.scale_modifier * sc0;
      const float s1 = f.scale_modifier * sc1;
      const float s2 = f.scale_modifier * sc2;
      float R[9];
      float c = 2.0f * q.w * q.w - 1.0f;
      float d;
      d = 2.0f * q.x;
      R[0] = c + q.x * d;
      R[3] = q.y * d + q.z * q.w * 2.0f;
      R[6] = q.z * d + (-q.y) * q.w * 2.0f;
      d = 2.0f * q.y;
      R[1] = q.x * d + (-q.z) * q.w * 2.0f;
      R[4] = c + q.y * d;
      R[7] = q.z * d + q.x * q.w * 2.0f;
      d = 2.0f * q.z;
      R[2] = q.x * d + q.y * q.w * 2.0f;
      R[5] = q.y * d + (-q.x) * q.w * 2.0f;
      R[8] = c + q.z * d;
      float S[9] = {s0, 0.f, 0.f, 0.f, s1, 0.f, 0.f, 0.f, s2};
      float M[9], Mt[9], sigma[9];
      gs_mat33_mul(R, S, M);
      gs_mat33_transpose(M, Mt);
      gs_mat33_mul(M, Mt, sigma);
      cov3d[0] = sigma[0];
      cov3d[1] = sigma[1];
      cov3d[2] = sigma[2];
      cov3d[3] = sigma[4];
      cov3d[4] = sigma[5];
      cov3d[5] = sigma[8];
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) o_cov[k] = cov3d[k];  // forward.py:260 (stored before later culls)

    // ---- compute_cov2d, forward.py:79-144 (quirk G1: W = view[:3,:3] as stored) ----
    float cov2d0, cov2d1, cov2d2;
    {
      float t0 = p_view[0], t1 = p_view[1], t2 = p_view[2];
      float limx = 1.3f * f.tan_fovx;
      float limy = 1.3f * f.tan_fovy;
      float txtz = t0 / t2;
      float tytz = t1 / t2;
      t0 = f_min(limx, f_max(-limx, txtz)) * t2;
      t1 = f_min(limy, f_max(-limy, tytz)) * t2;
      float focal_x = (float)f.W / (2.0f * f.tan_fovx);
      float focal_y = (float)f.H / (2.0f * f.tan_fovy);
      float J[9] = {focal_x / t2, 0.f, -(focal_x * t0) / (t2 * t2), 0.f, focal_y / t2, -(focal_y * t1) / (t2 * t2),
                    0.f, 0.f, 0.f};
      float Wm[9] = {f.view[0], f.view[1], f.view[2], f.view[4], f.view[5], f.view[6], f.view[8], f.view[9], f.view[10]};
      float T[9], Tt[9], A[9], cov[9];
      gs_mat33_mul(J, Wm, T);
      float Vrkt[9] = {cov3d[0], cov3d[1], cov3d[2], cov3d[1], cov3d[3], cov3d[4], cov3d[2], cov3d[4], cov3d[5]};
      gs_mat33_transpose(T, Tt);
      gs_mat33_mul(T, Vrkt, A);  // T * transpose(Vrk) * transpose(T); Vrk is symmetric by construction
      gs_mat33_mul(A, Tt, cov);
      cov2d0 = cov[0];
      cov2d1 = cov[1];
      cov2d2 = cov[4];
    }

    const float h_var = 0.3f;
    float cb0 = cov2d0 + h_var, cb1 = cov2d1, cb2 = cov2d2 + h_var;
    float det = cb0 * cb2 - cb1 * cb1;
    if (det == 0.0f) break;  // forward.py:278
    float det_inv = 1.0f / det;
    float conic0 = cb2 * det_inv, conic1 = -cb1 * det_inv, conic2 = cb0 * det_inv;
    float mid = 0.5f * (cb0 + cb2);
    float lambda1 = mid + sqrtf(f_max(0.1f, mid * mid - det));
    float lambda2 = mid - sqrtf(f_max(0.1f, mid * mid - det));
    float my_radius = ceilf(3.0f * sqrtf(f_max(lambda1, lambda2)));
    float pix = ((p_proj_x + 1.0f) * (float)f.W - 1.0f) * 0.5f;  // ndc2pix, forward.py:59-61
    float piy = ((p_proj_y + 1.0f) * (float)f.H - 1.0f) * 0.5f;

    gs_get_rect(pix, piy, my_radius, (float)f.grid_x, (float)f.grid_y, rminx, rminy, rmaxx, rmaxy);
    if ((rmaxx - rminx) * (rmaxy - rminy) == 0) break;  // forward.py:301

    // ---- SH -> RGB, forward.py:304-346 ----
    if (COLOR) gs_sh_to_rgb(f, px, py, pz, s_sh + tid * kShStride, o_rgb, o_cl);
    o_depth = p_view[2];
    o_radius = f2i(my_radius);
    o_x = pix;
    o_y = piy;
    o_con = make_float4(conic0, conic1, conic2, opacity);
    o_tiles = (rmaxy - rminy) * (rmaxx - rminx);
  } while (false);

  if (BIN && o_tiles > 0) {
    // counting pass of the tile binning: one RED per (Gaussian, tile) on the tile's counter.  (Round 1 drew the
    // duplicate's arrival rank from a RETURNING atomic here and stored it for the scatter pass -- 16 us on top of
    // the 23 us of the kernel, a block-wide scan for the rank slots and 4 B written + read per duplicate; now the
    // scatter pass takes its slot from the counter itself, counting it back down to zero.)
    for (int ty = rminy; ty < rmaxy; ++ty)
      for (int tx = rminx; tx < rmaxx; ++tx) atomicAdd(bin.tile_count + (size_t)(ty * f.grid_x + tx) * kCntStride, 1);
  }

  radii[i] = o_radius;
  tiles_touched[i] = o_tiles;
  depths[i] = o_depth;
  xy_out[i] = make_float2(o_x, o_y);
  conic_opacity[i] = o_con;
  float2* c3 = reinterpret_cast<float2*>(cov3Ds + (size_t)i * 6);
  c3[0] = make_float2(o_cov[0], o_cov[1]);
  c3[1] = make_float2(o_cov[2], o_cov[3]);
  c3[2] = make_float2(o_cov[4], o_cov[5]);
  if (COLOR) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      rgb[3 * i + c] = o_rgb[c];
      clamped_state[3 * i + c] = o_cl[c];
    }
  }
}

// The colour half of preprocess_kernel<.., COLOR = false>: rgb and clamped_state of every Gaussian (zeros for the
// culled ones, radii == 0, like the wp.zeros outputs of forward.py:703-710).  Same CTA shape and SH staging.
__global__ void __launch_bounds__(kPreThreads)
sh_color_kernel(const FrameK f, const int n, const float* __restrict__ means, const float* __restrict__ shs,
                const int* __restrict__ radii, float* __restrict__ rgb, float* __restrict__ clamped_state) {
  __shared__ float s_sh[kPreThreads * kShStride];
  gsb_pdl_wait();
  gsb_pdl_launch_dependents();
  const int base = blockIdx.x * kPreThreads;
  const int tid = threadIdx.x;
  const int rows = min(kPreThreads, n - base);
  const int i = base + tid;
  const bool live = i < n;
  const int il = live ? i : base;
  const float px = means[3 * il + 0], py = means[3 * il + 1], pz = means[3 * il + 2];
  const int radius = radii[il];
  {
    const float4* src = reinterpret_cast<const float4*>(shs + (size_t)base * 48);
    const int chunks = rows * 12;
    float4 v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const int c = tid + k * kPreThreads;
      v[k] = (c < chunks) ? __ldg(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const int c = tid + k * kPreThreads;
      if (c < chunks) {
        const int g = c / 12, qd = c - g * 12;
        float* d = s_sh + g * kShStride + qd * 4;
        d[0] = v[k].x;
        d[1] = v[k].y;
        d[2] = v[k].z;
        d[3] = v[k].w;
      }
    }
  }
  __syncthreads();
  if (!live) return;
  float o_rgb[3] = {0.f, 0.f, 0.f}, o_cl[3] = {0.f, 0.f, 0.f};
  if (radius > 0) gs_sh_to_rgb(f, px, py, pz, s_sh + tid * kShStride, o_rgb, o_cl);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    rgb[3 * i + c] = o_rgb[c];
    clamped_state[3 * i + c] = o_cl[c];
  }
}

}  // namespace

void gsb_make_framek(const gsb_frame* f, FrameK* k) {
  for (int i = 0; i < 16; ++i) {
    k->view[i] = f->view[i];
    k->proj[i] = f->proj[i];
  }
  for (int i = 0; i < 3; ++i) {
    k->campos[i] = f->campos[i];
    k->bg[i] = f->background[i];
  }
  k->tan_fovx = f->tan_fovx;
  k->tan_fovy = f->tan_fovy;
  k->scale_modifier = f->scale_modifier;
  k->W = f->width;
  k->H = f->height;
  k->degree = f->degree;
  k->clamped = f->clamped;
  // backward.py:1044-1045 evaluates these in Python doubles and passes them as float arguments
  k->focal_x = (float)((double)f->width / (2.0 * (double)f->tan_fovx));
  k->focal_y = (float)((double)f->height / (2.0 * (double)f->tan_fovy));
  k->grid_x = (f->width + kTile - 1) / kTile;
  k->grid_y = (f->height + kTile - 1) / kTile;
}

int gsb_preprocess_impl(gsb_ctx* ctx, cudaStream_t s, const gsb_frame* f, int32_t n, const float* means,
                        const float* scales, const float* rotations, const float* opacities, const float* shs,
                        int32_t* radii, float* points_xy, float* depths, float* cov3Ds, float* rgb,
                        float* conic_opacity, int32_t* tiles_touched, float* clamped_state, const PreBin* bin) {
  if (!ctx) return GSB_ERR_INVALID;
  GSB_REQUIRE(ctx, f && n >= 0, "gsb_preprocess: bad frame or n");
  if (n == 0) return GSB_OK;
  GSB_REQUIRE(ctx, gsb_aligned16(shs) && gsb_aligned16(rotations) && gsb_aligned16(conic_opacity),
              "gsb_preprocess: shs, rotations and conic_opacity must be 16-byte aligned");
  GSB_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(points_xy) & 7u) == 0 && (reinterpret_cast<uintptr_t>(cov3Ds) & 7u) == 0,
              "gsb_preprocess: points_xy and cov3Ds must be 8-byte aligned");
  FrameK k;
  gsb_make_framek(f, &k);
  int grid = (int)gsb_div_up(n, kPreThreads);
  if (bin && bin->defer_color) {
    GSB_LAUNCH_PDL(ctx, (preprocess_kernel<true, false>), grid, kPreThreads, 0, s, k, n, means, scales, rotations, opacities,
               shs, radii, reinterpret_cast<float2*>(points_xy), depths, cov3Ds, rgb,
               reinterpret_cast<float4*>(conic_opacity), tiles_touched, clamped_state, *bin);
  } else if (bin) {
    GSB_LAUNCH_PDL(ctx, (preprocess_kernel<true, true>), grid, kPreThreads, 0, s, k, n, means, scales, rotations, opacities,
               shs, radii, reinterpret_cast<float2*>(points_xy), depths, cov3Ds, rgb,
               reinterpret_cast<float4*>(conic_opacity), tiles_touched, clamped_state, *bin);
  } else {
    GSB_LAUNCH_PDL(ctx, (preprocess_kernel<false, true>), grid, kPreThreads, 0, s, k, n, means, scales, rotations, opacities,
               shs, radii, reinterpret_cast<float2*>(points_xy), depths, cov3Ds, rgb,
               reinterpret_cast<float4*>(conic_opacity), tiles_touched, clamped_state, PreBin{});
  }
  return GSB_OK;
}

// rgb / clamped_state of a frame whose preprocess ran with bin->defer_color (gsb_forward with a colour dependency)
int gsb_sh_color_impl(gsb_ctx* ctx, cudaStream_t s, const gsb_frame* f, int32_t n, const float* means, const float* shs,
                      const int32_t* radii, float* rgb, float* clamped_state) {
  if (n == 0) return GSB_OK;
  FrameK k;
  gsb_make_framek(f, &k);
  GSB_LAUNCH_PDL(ctx, sh_color_kernel, (int)gsb_div_up(n, kPreThreads), kPreThreads, 0, s, k, n, means, shs, radii, rgb,
             clamped_state);
  return GSB_OK;
}

GSB_API int gsb_preprocess(gsb_ctx* ctx, gsb_stream s, const gsb_frame* f, int32_t n, const float* means,
                           const float* scales, const float* rotations, const float* opacities, const float* shs,
                           int32_t* radii, float* points_xy, float* depths, float* cov3Ds, float* rgb,
                           float* conic_opacity, int32_t* tiles_touched, float* clamped_state) {
  return gsb_preprocess_impl(ctx, (cudaStream_t)s, f, n, means, scales, rotations, opacities, shs, radii, points_xy,
                             depths, cov3Ds, rgb, conic_opacity, tiles_touched, clamped_state, nullptr);
}
