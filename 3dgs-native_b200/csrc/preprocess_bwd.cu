// preprocess_bwd.cu -- replaces backward_preprocess (reference backward.py:770-888), i.e. the four
// kernels compute_cov2d_backward_kernel (258-435), compute_projection_backward_kernel (708-768),
// sh_backward_kernel (68-255) and compute_cov3d_backward_kernel (438-556) plus the dL_dcov3D
// round trip and the two host syncs between them, as ONE pass over the Gaussians.
//
// One thread per Gaussian, 128 per CTA.  SH rows come in through a padded shared-memory tile
// (coalesced 16-byte loads), and the 48 SH gradients of each Gaussian leave through the same tile
// as coalesced 16-byte stores.  Every output element is written (zeros for skipped Gaussians).
//
// The reference's defects on this path are reproduced on purpose (SURVEY.md 8a note G):
//   G2 cov3D backward applies a column-major formula to row-major matrices; G3 it always runs
//   with scale_modifier = 1; the stray w = 1.0 in the view-space -> world-space product; the SH
//   clamp mask is applied whether or not the forward clamped.
//
// HBM roofline: reads 312 B + writes 232 B per Gaussian (544*N).
#include "common.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kShStride = 49;

// COMPACT: instead of the 48 SH gradients the kernel stores the two factors of that rank-1 product --
// (dL_dRGB masked, unit view direction) as two float4 per Gaussian at dL_dshs[8 i] -- for the
// multi-GPU exchange, which ships 32 bytes per Gaussian and view instead of 192 and expands them on
// the receiving side with the same gs_sh_basis.
// PACKED: the tile-stage gradients come as the packed per-Gaussian record the backward tile kernel
// accumulates into (12 floats: conic a, b, c, opacity | r, g, b, mean2D.x | mean2D.y, -, -, -; see
// blend_bwd.cu), and this kernel also writes them out in the reference's layouts (dL_dmean2D float[N][3],
// dL_dconic float[N][4], dL_dcolor float[N][3], dL_dopacity float[N]): those four arrays are then outputs.
template <bool COMPACT, bool PACKED>
__global__ void __launch_bounds__(kThreads, 6)
preprocess_backward_kernel(const FrameK f, const int n, const float* __restrict__ means, const int* __restrict__ radii,
                           const float* __restrict__ shs, const float* __restrict__ scales,
                           const float* __restrict__ rots, const float* __restrict__ cov3Ds,
                           const float* __restrict__ clamped_state, const float* __restrict__ dL_dmean2D,
                           const float* __restrict__ dL_dconic, const float* __restrict__ dL_dcolor,
                           float* __restrict__ dL_dmean3D, float* __restrict__ dL_dshs, float* __restrict__ dL_dscale,
                           float4* __restrict__ dL_drot, float* __restrict__ dL_dcov3D_out,
                           const float4* __restrict__ packed, float* __restrict__ dL_dopacity_out,
                           float* __restrict__ zero_cov3D) {
  __shared__ float s_sh[kThreads * kShStride];
  gsb_pdl_wait();
  gsb_pdl_launch_dependents();
  const int base = blockIdx.x * kThreads;
  const int tid = threadIdx.x;
  const int rows = min(kThreads, n - base);
  // This thread's own inputs first: they are independent of the SH tile and would otherwise cost a
  // second DRAM round trip after the barrier.
  const int i = base + tid;
  const int il = (i < n) ? i : base;  // clamped: the dead threads of the last CTA re-read row 0
  const int in_radius = radii[il];
  const float mx = means[3 * il + 0], my = means[3 * il + 1], mz = means[3 * il + 2];
  const float2 c3a = *reinterpret_cast<const float2*>(cov3Ds + (size_t)il * 6);
  const float2 c3b = *reinterpret_cast<const float2*>(cov3Ds + (size_t)il * 6 + 2);
  const float2 c3c = *reinterpret_cast<const float2*>(cov3Ds + (size_t)il * 6 + 4);
  float4 in_dcon;
  float in_g0, in_g1, in_dcol[3], in_dopac = 0.0f;
  if (PACKED) {
    const float4 q0 = __ldcs(packed + 3 * (size_t)il), q1 = __ldcs(packed + 3 * (size_t)il + 1);
    in_g1 = __ldcs(reinterpret_cast<const float*>(packed + 3 * (size_t)il + 2));
    in_dcon = make_float4(q0.x, q0.y, 0.0f, q0.z);
    in_dopac = q0.w;
    in_dcol[0] = q1.x, in_dcol[1] = q1.y, in_dcol[2] = q1.z;
    in_g0 = q1.w;
  } else {
    in_dcon = __ldg(reinterpret_cast<const float4*>(dL_dconic) + il);
    in_g0 = dL_dmean2D[3 * il + 0], in_g1 = dL_dmean2D[3 * il + 1];
    in_dcol[0] = dL_dcolor[3 * il + 0], in_dcol[1] = dL_dcolor[3 * il + 1], in_dcol[2] = dL_dcolor[3 * il + 2];
  }
  const float in_cl[3] = {clamped_state[3 * il + 0], clamped_state[3 * il + 1], clamped_state[3 * il + 2]};
  const float4 in_q = __ldg(reinterpret_cast<const float4*>(rots) + il);
  const float in_s0 = scales[3 * il + 0], in_s1 = scales[3 * il + 1], in_s2 = scales[3 * il + 2];
  {
    const float4* src = reinterpret_cast<const float4*>(shs + (size_t)base * 48);
    const int chunks = rows * 12;
    float4 v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const int c = tid + k * kThreads;
      v[k] = (c < chunks) ? __ldg(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const int c = tid + k * kThreads;
      if (c < chunks) {
        const int g = c / 12, q = c - g * 12;
        float* d = s_sh + g * kShStride + q * 4;
        d[0] = v[k].x;
        d[1] = v[k].y;
        d[2] = v[k].z;
        d[3] = v[k].w;
      }
    }
  }
  __syncthreads();

  float* sh = s_sh + tid * kShStride;  // this thread's row: SH in, dL_dSH out
  if (i < n) {
    float o_mean[3] = {0.f, 0.f, 0.f};
    float o_scale[3] = {0.f, 0.f, 0.f};
    float4 o_rot = make_float4(0.f, 0.f, 0.f, 0.f);
    float o_dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    bool sh_written = false;
    float4 shc0 = make_float4(0.f, 0.f, 0.f, 0.f), shc1 = shc0;  // COMPACT: zeros for a skipped Gaussian

    if (in_radius > 0) {
      // ================= compute_cov2d_backward_kernel, backward.py:258-435 =================
      {
        const float c3[6] = {c3a.x, c3a.y, c3b.x, c3b.y, c3c.x, c3c.y};
        const float dcon0 = in_dcon.x, dcon1 = in_dcon.y, dcon2 = in_dcon.w;
        float t[4];
        gs_vec4_mul_mat44(mx, my, mz, 1.0f, f.view, t);
        const float limx = 1.3f * f.tan_fovx;
        const float limy = 1.3f * f.tan_fovy;
        const float tz = t[2];
        const float inv_tz = 1.0f / tz;
        const float txtz = t[0] * inv_tz;
        const float tytz = t[1] * inv_tz;
        const bool x_cl = (txtz < -limx) || (txtz > limx);
        const bool y_cl = (tytz < -limy) || (tytz > limy);
        const float x_grad_mul = 1.0f - (x_cl ? 1.0f : 0.0f);
        const float y_grad_mul = 1.0f - (y_cl ? 1.0f : 0.0f);
        const float tx = f_min(limx, f_max(-limx, txtz)) * tz;
        const float ty = f_min(limy, f_max(-limy, tytz)) * tz;
        const float inv_tz2 = inv_tz * inv_tz;
        const float inv_tz3 = inv_tz2 * inv_tz;
        const float h_x = f.focal_x, h_y = f.focal_y;
        const float J00 = h_x * inv_tz;
        const float J11 = h_y * inv_tz;
        const float J02 = -h_x * tx * inv_tz2;
        const float J12 = -h_y * ty * inv_tz2;
        const float J[9] = {J00, 0.f, 0.f, 0.f, J11, 0.f, J02, J12, 0.f};  // transpose(mat33(J00,0,J02, 0,J11,J12, 0,0,0))
        const float Wm[9] = {f.view[0], f.view[1], f.view[2], f.view[4], f.view[5], f.view[6], f.view[8], f.view[9], f.view[10]};
        float T[9];
        gs_mat33_mul(Wm, J, T);
        const float Vrk[9] = {c3[0], c3[1], c3[2], c3[1], c3[3], c3[4], c3[2], c3[4], c3[5]};
        float Tt[9], A[9], cov2D[9];
        gs_mat33_transpose(T, Tt);
        gs_mat33_mul(Tt, Vrk, A);  // transpose(T) * transpose(Vrk) * T, Vrk symmetric
        gs_mat33_mul(A, T, cov2D);
        const float a = cov2D[0] + 0.3f;
        const float b = cov2D[1];
        const float c = cov2D[4] + 0.3f;
        const float denom = a * c - b * b;
        float dL_da = 0.0f, dL_db = 0.0f, dL_dc = 0.0f;
        if (denom != 0.0f) {
          const float denom2inv = 1.0f / (denom * denom + 1e-7f);
          dL_da = denom2inv * (-c * c * dcon0 + 2.0f * b * c * dcon1 + (denom - a * c) * dcon2);
          dL_dc = denom2inv * (-a * a * dcon2 + 2.0f * a * b * dcon1 + (denom - a * c) * dcon0);
          dL_db = denom2inv * 2.0f * (b * c * dcon0 - (denom + 2.0f * b * b) * dcon1 + a * b * dcon2);
        }
#define T_(r, q) T[(r) * 3 + (q)]
#define V_(r, q) Vrk[(r) * 3 + (q)]
        o_dcov[0] = T_(0, 0) * T_(0, 0) * dL_da + T_(0, 0) * T_(0, 1) * dL_db + T_(0, 1) * T_(0, 1) * dL_dc;
        o_dcov[1] = 2.0f * T_(0, 0) * T_(1, 0) * dL_da + (T_(0, 0) * T_(1, 1) + T_(1, 0) * T_(0, 1)) * dL_db +
                    2.0f * T_(0, 1) * T_(1, 1) * dL_dc;
        o_dcov[2] = 2.0f * T_(0, 0) * T_(2, 0) * dL_da + (T_(0, 0) * T_(2, 1) + T_(2, 0) * T_(0, 1)) * dL_db +
                    2.0f * T_(0, 1) * T_(2, 1) * dL_dc;
        o_dcov[3] = T_(1, 0) * T_(1, 0) * dL_da + T_(1, 0) * T_(1, 1) * dL_db + T_(1, 1) * T_(1, 1) * dL_dc;
        o_dcov[4] = 2.0f * T_(2, 0) * T_(1, 0) * dL_da + (T_(1, 0) * T_(2, 1) + T_(2, 0) * T_(1, 1)) * dL_db +
                    2.0f * T_(1, 1) * T_(2, 1) * dL_dc;
        o_dcov[5] = T_(2, 0) * T_(2, 0) * dL_da + T_(2, 0) * T_(2, 1) * dL_db + T_(2, 1) * T_(2, 1) * dL_dc;

        const float dL_dT00 = 2.0f * (T_(0, 0) * V_(0, 0) + T_(1, 0) * V_(1, 0) + T_(2, 0) * V_(2, 0)) * dL_da +
                              (T_(0, 1) * V_(0, 0) + T_(1, 1) * V_(1, 0) + T_(2, 1) * V_(2, 0)) * dL_db;
        const float dL_dT01 = 2.0f * (T_(0, 0) * V_(0, 1) + T_(1, 0) * V_(1, 1) + T_(2, 0) * V_(2, 1)) * dL_da +
                              (T_(0, 1) * V_(0, 1) + T_(1, 1) * V_(1, 1) + T_(2, 1) * V_(2, 1)) * dL_db;
        const float dL_dT02 = 2.0f * (T_(0, 0) * V_(0, 2) + T_(1, 0) * V_(1, 2) + T_(2, 0) * V_(2, 2)) * dL_da +
                              (T_(0, 1) * V_(0, 2) + T_(1, 1) * V_(1, 2) + T_(2, 1) * V_(2, 2)) * dL_db;
        const float dL_dT10 = 2.0f * (T_(0, 1) * V_(0, 0) + T_(1, 1) * V_(1, 0) + T_(2, 1) * V_(2, 0)) * dL_dc +
                              (T_(0, 0) * V_(0, 0) + T_(1, 0) * V_(1, 0) + T_(2, 0) * V_(2, 0)) * dL_db;
        const float dL_dT11 = 2.0f * (T_(0, 1) * V_(0, 1) + T_(1, 1) * V_(1, 1) + T_(2, 1) * V_(2, 1)) * dL_dc +
                              (T_(0, 0) * V_(0, 1) + T_(1, 0) * V_(1, 1) + T_(2, 0) * V_(2, 1)) * dL_db;
        const float dL_dT12 = 2.0f * (T_(0, 1) * V_(0, 2) + T_(1, 1) * V_(1, 2) + T_(2, 1) * V_(2, 2)) * dL_dc +
                              (T_(0, 0) * V_(0, 2) + T_(1, 0) * V_(1, 2) + T_(2, 0) * V_(2, 2)) * dL_db;
#undef T_
#undef V_
#define W_(r, q) Wm[(r) * 3 + (q)]
        const float dL_dJ00 = W_(0, 0) * dL_dT00 + W_(1, 0) * dL_dT01 + W_(2, 0) * dL_dT02;
        const float dL_dJ02 = W_(0, 2) * dL_dT00 + W_(1, 2) * dL_dT01 + W_(2, 2) * dL_dT02;
        const float dL_dJ11 = W_(0, 1) * dL_dT10 + W_(1, 1) * dL_dT11 + W_(2, 1) * dL_dT12;
        const float dL_dJ12 = W_(0, 2) * dL_dT10 + W_(1, 2) * dL_dT11 + W_(2, 2) * dL_dT12;
#undef W_
        const float dL_dtx = -h_x * inv_tz2 * dL_dJ02;
        const float dL_dty = -h_y * inv_tz2 * dL_dJ12;
        const float dL_dtz = -h_x * inv_tz2 * dL_dJ00 - h_y * inv_tz2 * dL_dJ11 + 2.0f * h_x * tx * inv_tz3 * dL_dJ02 +
                             2.0f * h_y * ty * inv_tz3 * dL_dJ12;
        const float d0 = dL_dtx * x_grad_mul, d1 = dL_dty * y_grad_mul, d2 = dL_dtz, d3 = 1.0f;  // stray w = 1.0
#pragma unroll
        for (int j = 0; j < 3; ++j) {  // vec4 * transpose(view): r[j] = sum_k v[k] * view[j][k]
          float s = f.view[j * 4 + 0] * d0;
          s = s + f.view[j * 4 + 1] * d1;
          s = s + f.view[j * 4 + 2] * d2;
          s = s + f.view[j * 4 + 3] * d3;
          o_mean[j] = o_mean[j] + s;
        }
      }
      // ================= compute_projection_backward_kernel, backward.py:708-768 =================
      {
        const float g0 = in_g0, g1 = in_g1;
        float m_hom[4];
        gs_vec4_mul_mat44(mx, my, mz, 1.0f, f.proj, m_hom);
        const float m_w = 1.0f / (m_hom[3] + 0.0000001f);
#define P_(r, q) f.proj[(r) * 4 + (q)]
        const float mul1 = (P_(0, 0) * mx + P_(1, 0) * my + P_(2, 0) * mz + P_(3, 0)) * m_w * m_w;
        const float mul2 = (P_(0, 1) * mx + P_(1, 1) * my + P_(2, 1) * mz + P_(3, 1)) * m_w * m_w;
        const float e0 = (P_(0, 0) * m_w - P_(0, 3) * mul1) * g0 + (P_(0, 1) * m_w - P_(0, 3) * mul2) * g1;
        const float e1 = (P_(1, 0) * m_w - P_(1, 3) * mul1) * g0 + (P_(1, 1) * m_w - P_(1, 3) * mul2) * g1;
        const float e2 = (P_(2, 0) * m_w - P_(2, 3) * mul1) * g0 + (P_(2, 1) * m_w - P_(2, 3) * mul2) * g1;
#undef P_
        o_mean[0] = o_mean[0] + e0;
        o_mean[1] = o_mean[1] + e1;
        o_mean[2] = o_mean[2] + e2;
      }
      // ================= sh_backward_kernel, backward.py:68-255 =================
      {
        const float dox = mx - f.campos[0], doy = my - f.campos[1], doz = mz - f.campos[2];
        const float dir_len = sqrtf(gs_dot3(dox, doy, doz, dox, doy, doz));
        if (!(dir_len < 1e-8f)) {
          const float x = dox / dir_len, y = doy / dir_len, z = doz / dir_len;
          float dRGB[3];
#pragma unroll
          for (int c = 0; c < 3; ++c)  // mask applied regardless of the forward `clamped` flag
            dRGB[c] = in_dcol[c] * (1.0f + (-1.0f * in_cl[c]));
          float ddx[3] = {0.f, 0.f, 0.f}, ddy[3] = {0.f, 0.f, 0.f}, ddz[3] = {0.f, 0.f, 0.f};
          float basis[16];
          const int deg = f.degree;
          const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
          gs_sh_basis(deg, x, y, z, basis);
          if (deg > 0) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              ddx[c] = -GS_SH_C1 * sh[3 * 3 + c];
              ddy[c] = -GS_SH_C1 * sh[1 * 3 + c];
              ddz[c] = GS_SH_C1 * sh[2 * 3 + c];
            }
            if (deg > 1) {
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                const float sh4 = sh[4 * 3 + c], sh5 = sh[5 * 3 + c], sh6 = sh[6 * 3 + c], sh7 = sh[7 * 3 + c],
                            sh8 = sh[8 * 3 + c];
                ddx[c] += GS_C2_0 * y * sh4 + GS_C2_2 * 2.0f * -x * sh6 + GS_C2_3 * z * sh7 + GS_C2_4 * 2.0f * x * sh8;
                ddy[c] += GS_C2_0 * x * sh4 + GS_C2_1 * z * sh5 + GS_C2_2 * 2.0f * -y * sh6 + GS_C2_4 * 2.0f * -y * sh8;
                ddz[c] += GS_C2_1 * y * sh5 + GS_C2_2 * 2.0f * 2.0f * z * sh6 + GS_C2_3 * x * sh7;
              }
              if (deg > 2) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                  const float sh9 = sh[9 * 3 + c], sh10 = sh[10 * 3 + c], sh11 = sh[11 * 3 + c], sh12 = sh[12 * 3 + c],
                              sh13 = sh[13 * 3 + c], sh14 = sh[14 * 3 + c], sh15 = sh[15 * 3 + c];
                  ddx[c] += (GS_C3_0 * sh9 * 3.0f * 2.0f * xy + GS_C3_1 * sh10 * yz + GS_C3_2 * sh11 * -2.0f * xy +
                             GS_C3_3 * sh12 * -3.0f * 2.0f * xz + GS_C3_4 * sh13 * (-3.0f * xx + 4.0f * zz - yy) +
                             GS_C3_5 * sh14 * 2.0f * xz + GS_C3_6 * sh15 * 3.0f * (xx - yy));
                  ddy[c] += (GS_C3_0 * sh9 * 3.0f * (xx - yy) + GS_C3_1 * sh10 * xz +
                             GS_C3_2 * sh11 * (-3.0f * yy + 4.0f * zz - xx) + GS_C3_3 * sh12 * -3.0f * 2.0f * yz +
                             GS_C3_4 * sh13 * -2.0f * xy + GS_C3_5 * sh14 * -2.0f * yz + GS_C3_6 * sh15 * -3.0f * 2.0f * xy);
                  ddz[c] += (GS_C3_1 * sh10 * xy + GS_C3_2 * sh11 * 4.0f * 2.0f * yz +
                             GS_C3_3 * sh12 * 3.0f * (2.0f * zz - xx - yy) + GS_C3_4 * sh13 * 4.0f * 2.0f * xz +
                             GS_C3_5 * sh14 * (xx - yy));
                }
              }
            }
          }
          // all SH reads are done: the row now becomes dL_dSH (assignment, backward.py:127-213;
          // coefficients above the degree stay zero as in the zero-initialised reference buffer)
          if (COMPACT) {
            shc0 = make_float4(dRGB[0], dRGB[1], dRGB[2], x);
            shc1 = make_float4(y, z, 0.0f, 0.0f);
          } else {
#pragma unroll
            for (int k = 0; k < 16; ++k)
#pragma unroll
              for (int c = 0; c < 3; ++c) sh[k * 3 + c] = basis[k] * dRGB[c];
          }
          sh_written = true;
          const float dd0 = gs_dot3(ddx[0], ddx[1], ddx[2], dRGB[0], dRGB[1], dRGB[2]);
          const float dd1 = gs_dot3(ddy[0], ddy[1], ddy[2], dRGB[0], dRGB[1], dRGB[2]);
          const float dd2 = gs_dot3(ddz[0], ddz[1], ddz[2], dRGB[0], dRGB[1], dRGB[2]);
          // dnormvdv, backward.py:42-64
          const float sum2 = dox * dox + doy * doy + doz * doz;
          if (!(sum2 < 1e-10f)) {
            const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
            const float r0 = ((sum2 - dox * dox) * dd0 - doy * dox * dd1 - doz * dox * dd2) * invsum32;
            const float r1 = (-dox * doy * dd0 + (sum2 - doy * doy) * dd1 - doz * doy * dd2) * invsum32;
            const float r2 = (-dox * doz * dd0 - doy * doz * dd1 + (sum2 - doz * doz) * dd2) * invsum32;
            o_mean[0] = o_mean[0] + r0;
            o_mean[1] = o_mean[1] + r1;
            o_mean[2] = o_mean[2] + r2;
          } else {
            o_mean[0] = o_mean[0] + 0.0f;
            o_mean[1] = o_mean[1] + 0.0f;
            o_mean[2] = o_mean[2] + 0.0f;
          }
        }
      }
      // ================= compute_cov3d_backward_kernel, backward.py:438-556 (G2, G3) =================
      {
        const float4 q = in_q;
        const float r = q.w, x = q.x, y = q.y, z = q.z;
        const float R[9] = {1.0f - 2.0f * (y * y + z * z), 2.0f * (x * y - r * z),        2.0f * (x * z + r * y),
                            2.0f * (x * y + r * z),        1.0f - 2.0f * (x * x + z * z), 2.0f * (y * z - r * x),
                            2.0f * (x * z - r * y),        2.0f * (y * z + r * x),        1.0f - 2.0f * (x * x + y * y)};
        const float scale_modifier = 1.0f;  // quirk G3: backward() never forwards the caller's value
        const float sv0 = scale_modifier * in_s0, sv1 = scale_modifier * in_s1, sv2 = scale_modifier * in_s2;
        const float S[9] = {sv0, 0.f, 0.f, 0.f, sv1, 0.f, 0.f, 0.f, sv2};
        float M[9];
        gs_mat33_mul(S, R, M);
        const float dSig[9] = {o_dcov[0],        0.5f * o_dcov[1], 0.5f * o_dcov[2], 0.5f * o_dcov[1], o_dcov[3],
                               0.5f * o_dcov[4], 0.5f * o_dcov[2], 0.5f * o_dcov[4], o_dcov[5]};
        float M2[9], dL_dM[9], Rt[9], dMt[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) M2[k] = 2.0f * M[k];
        gs_mat33_mul(M2, dSig, dL_dM);
        gs_mat33_transpose(R, Rt);
        gs_mat33_transpose(dL_dM, dMt);
        o_scale[0] = gs_dot3(Rt[0], Rt[1], Rt[2], dMt[0], dMt[1], dMt[2]) * scale_modifier;
        o_scale[1] = gs_dot3(Rt[3], Rt[4], Rt[5], dMt[3], dMt[4], dMt[5]) * scale_modifier;
        o_scale[2] = gs_dot3(Rt[6], Rt[7], Rt[8], dMt[6], dMt[7], dMt[8]) * scale_modifier;
        const float sv[3] = {sv0, sv1, sv2};
        float D[9];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) D[a * 3 + b] = dMt[a * 3 + b] * sv[a];
#define D_(a, b) D[(a) * 3 + (b)]
        const float dL_dr = 2.0f * (z * (D_(0, 1) - D_(1, 0)) + y * (D_(2, 0) - D_(0, 2)) + x * (D_(1, 2) - D_(2, 1)));
        const float dL_dx = 2.0f * (y * (D_(1, 0) + D_(0, 1)) + z * (D_(2, 0) + D_(0, 2)) + r * (D_(1, 2) - D_(2, 1))) -
                            4.0f * x * (D_(2, 2) + D_(1, 1));
        const float dL_dy = 2.0f * (x * (D_(1, 0) + D_(0, 1)) + r * (D_(2, 0) - D_(0, 2)) + z * (D_(1, 2) + D_(2, 1))) -
                            4.0f * y * (D_(2, 2) + D_(0, 0));
        const float dL_dz = 2.0f * (r * (D_(0, 1) - D_(1, 0)) + x * (D_(2, 0) + D_(0, 2)) + y * (D_(1, 2) + D_(2, 1))) -
                            4.0f * z * (D_(1, 1) + D_(0, 0));
#undef D_
        o_rot = make_float4(dL_dx, dL_dy, dL_dz, dL_dr);
      }
    }
    if (COMPACT) {
      float4* o = reinterpret_cast<float4*>(dL_dshs) + 2 * (size_t)i;
      o[0] = shc0;
      o[1] = shc1;
    } else if (!sh_written) {
#pragma unroll
      for (int k = 0; k < 48; ++k) sh[k] = 0.0f;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      dL_dmean3D[3 * i + c] = o_mean[c];
      dL_dscale[3 * i + c] = o_scale[c];
    }
    dL_drot[i] = o_rot;
    if (PACKED) {   // the reference's layouts: dL_dmean2D (x, y, 0), dL_dconic (a, b, 0, c) -- backward.py:691-706
      float* m2 = const_cast<float*>(dL_dmean2D) + 3 * (size_t)i;
      m2[0] = in_g0, m2[1] = in_g1, m2[2] = 0.0f;
      reinterpret_cast<float4*>(const_cast<float*>(dL_dconic))[i] = in_dcon;
      float* dc = const_cast<float*>(dL_dcolor) + 3 * (size_t)i;
      dc[0] = in_dcol[0], dc[1] = in_dcol[1], dc[2] = in_dcol[2];
      dL_dopacity_out[i] = in_dopac;
    }
    if (zero_cov3D) {   // backward.py:1119,1195: the operator returns dL_dcov3D as a zero buffer (8-byte aligned here)
      float2* z = reinterpret_cast<float2*>(zero_cov3D + (size_t)i * 6);
      z[0] = z[1] = z[2] = make_float2(0.0f, 0.0f);
    }
    if (dL_dcov3D_out) {
      float2* o = reinterpret_cast<float2*>(dL_dcov3D_out + (size_t)i * 6);
      o[0] = make_float2(o_dcov[0], o_dcov[1]);
      o[1] = make_float2(o_dcov[2], o_dcov[3]);
      o[2] = make_float2(o_dcov[4], o_dcov[5]);
    }
  }
  if (COMPACT) return;
  __syncthreads();
  {  // coalesced store of the CTA's SH gradients
    float4* dst = reinterpret_cast<float4*>(dL_dshs + (size_t)base * 48);
    const int chunks = rows * 12;
#pragma unroll 4
    for (int c = tid; c < chunks; c += kThreads) {
      int g = c / 12, q = c - g * 12;
      const float* d = s_sh + g * kShStride + q * 4;
      dst[c] = make_float4(d[0], d[1], d[2], d[3]);
    }
  }
}

}  // namespace

GSB_API int gsb_preprocess_backward(gsb_ctx* ctx, gsb_stream s, const gsb_frame* f, int32_t n, const float* means,
                                    const int32_t* radii, const float* shs, const float* scales,
                                    const float* rotations, const float* cov3Ds, const float* clamped_state,
                                    const float* dL_dmean2D, const float* dL_dconic, const float* dL_dcolor,
                                    float* dL_dmean3D, float* dL_dshs, float* dL_dscale, float* dL_drot,
                                    float* dL_dcov3D_internal) {
  return gsb_preprocess_backward_impl(ctx, (cudaStream_t)s, f, n, means, radii, shs, scales, rotations, cov3Ds,
                                      clamped_state, dL_dmean2D, dL_dconic, dL_dcolor, dL_dmean3D, dL_dshs, dL_dscale,
                                      dL_drot, dL_dcov3D_internal, 0, nullptr, nullptr, nullptr);
}

GSB_API int gsb_preprocess_backward_compact_sh(gsb_ctx* ctx, gsb_stream s, const gsb_frame* f, int32_t n,
                                               const float* means, const int32_t* radii, const float* shs,
                                               const float* scales, const float* rotations, const float* cov3Ds,
                                               const float* clamped_state, const float* dL_dmean2D,
                                               const float* dL_dconic, const float* dL_dcolor, float* dL_dmean3D,
                                               float* dL_dshs_compact, float* dL_dscale, float* dL_drot,
                                               float* dL_dcov3D_internal) {
  return gsb_preprocess_backward_impl(ctx, (cudaStream_t)s, f, n, means, radii, shs, scales, rotations, cov3Ds,
                                      clamped_state, dL_dmean2D, dL_dconic, dL_dcolor, dL_dmean3D, dL_dshs_compact,
                                      dL_dscale, dL_drot, dL_dcov3D_internal, 1, nullptr, nullptr, nullptr);
}

int gsb_preprocess_backward_impl(gsb_ctx* ctx, cudaStream_t s, const gsb_frame* f, int32_t n, const float* means,
                                 const int32_t* radii, const float* shs, const float* scales, const float* rotations,
                                 const float* cov3Ds, const float* clamped_state, const float* dL_dmean2D,
                                 const float* dL_dconic, const float* dL_dcolor, float* dL_dmean3D, float* dL_dshs,
                                 float* dL_dscale, float* dL_drot, float* dL_dcov3D_internal, int sh_compact,
                                 const float* packed, float* dL_dopacity_out, float* zero_cov3D) {
  if (!ctx) return GSB_ERR_INVALID;
  if (zero_cov3D && (reinterpret_cast<uintptr_t>(zero_cov3D) & 7u) != 0) {
    // an unaligned dL_dcov3D result buffer: zeroed by a stream operation instead of by the kernel
    GSB_CUDA(ctx, cudaMemsetAsync(zero_cov3D, 0, sizeof(float) * 6 * (size_t)n, s));
    zero_cov3D = nullptr;
  }
  GSB_REQUIRE(ctx, f && n >= 0, "gsb_preprocess_backward: bad frame or n");
  if (n == 0) return GSB_OK;
  GSB_REQUIRE(ctx, gsb_aligned16(shs) && gsb_aligned16(dL_dshs) && gsb_aligned16(rotations) && gsb_aligned16(dL_drot),
              "gsb_preprocess_backward: shs, dL_dshs, rotations, dL_drot must be 16-byte aligned");
  GSB_REQUIRE(ctx, gsb_aligned16(dL_dconic) && (reinterpret_cast<uintptr_t>(cov3Ds) & 7u) == 0,
              "gsb_preprocess_backward: dL_dconic must be 16-byte and cov3Ds 8-byte aligned");
  GSB_REQUIRE(ctx, !dL_dcov3D_internal || (reinterpret_cast<uintptr_t>(dL_dcov3D_internal) & 7u) == 0,
              "gsb_preprocess_backward: dL_dcov3D_internal must be 8-byte aligned");
  FrameK k;
  gsb_make_framek(f, &k);
  GSB_REQUIRE(ctx, !packed || (gsb_aligned16(packed) && dL_dopacity_out), "gsb_preprocess_backward: bad packed record");
  const float4* pk = reinterpret_cast<const float4*>(packed);
#define GSB_PPB_LAUNCH(C, K)                                                                                              \
  GSB_LAUNCH(ctx, (preprocess_backward_kernel<C, K>), (int)gsb_div_up(n, kThreads), kThreads, 0, s, k, n, means, radii, shs, \
             scales, rotations, cov3Ds, clamped_state, dL_dmean2D, dL_dconic, dL_dcolor, dL_dmean3D, dL_dshs, dL_dscale, \
             reinterpret_cast<float4*>(dL_drot), dL_dcov3D_internal, pk, dL_dopacity_out, zero_cov3D)
  if (sh_compact && packed) GSB_PPB_LAUNCH(true, true);
  else if (sh_compact) GSB_PPB_LAUNCH(true, false);
  else if (packed) GSB_PPB_LAUNCH(false, true);
  else GSB_PPB_LAUNCH(false, false);
#undef GSB_PPB_LAUNCH
  return GSB_OK;
}
