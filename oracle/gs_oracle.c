/*
 * gs_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A literal plain-C restatement of the reference's (zhujinchong/3DGS-native) rasterizer hot
 * path: forward.py, backward.py, optimizer.py, utils/wp_utils.py, loss.py and the three nested
 * kernels of train.py.  Every function cites the reference file:line it follows.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library.  The product (3dgs-native_b200/) never links, imports or calls it.
 *
 * PARITY STATUS: the reference cannot be executed here (every module imports warp-lang==1.7.0,
 * which is not installed and not installable).  The restatement is pinned by
 *   (1) tests/golden/<name>.npz -- outputs of the reference's OWN kernel source executed under the
 *       pure-Python Warp shim in tests/warp_shim/ (generator: tests/golden/make_golden.py),
 *   (2) the known-answer colours recovered from the reference's assets/example_render.png.
 * Semantics of Warp built-ins (mat/vec products, quat_to_matrix, normalize, randf, sign) are
 * restated from upstream knowledge and marked [Warp]; they are "parity unpinned" at the
 * bit level.
 *
 * ARITHMETIC CONTRACT (shared with the CUDA kernels, see DESIGN.md "Arithmetic contract"):
 *   every reference expression is evaluated in IEEE binary32, in the order Python parses it,
 *   each operation individually rounded (no FMA contraction: build with -ffp-contract=off);
 *   division and sqrt are correctly rounded; float->int conversion truncates and SATURATES
 *   (CUDA semantics, NaN -> 0); exp() is either libm expf (GSO_EXP_LIBM, faithful to Warp's
 *   CPU device) or the deterministic gs_expf below (GSO_EXP_DET, default; the CUDA kernels
 *   implement the same operation sequence so integer outputs can be compared bit-exactly).
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -fno-fast-math -pthread -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#define GSO_API __attribute__((visibility("default")))

#define TILE_M 16 /* config.py:21 */
#define TILE_N 16 /* config.py:22 */

/* ----------------------------------------------------------------------------------------- */
/* global switches                                                                            */
/* ----------------------------------------------------------------------------------------- */
#define GSO_EXP_DET 0
#define GSO_EXP_LIBM 1
static int g_exp_mode = GSO_EXP_DET;
static int g_threads = 1; /* 1 = serial, thread order of Warp's CPU device [Warp] */

GSO_API void gso_set_exp_mode(int mode) { g_exp_mode = mode; }
GSO_API int gso_get_exp_mode(void) { return g_exp_mode; }
GSO_API void gso_set_threads(int n) { g_threads = n < 1 ? 1 : (n > 256 ? 256 : n); }
GSO_API int gso_get_threads(void) { return g_threads; }
GSO_API int gso_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n < 1 ? 1 : (int)n;
}

/* Minimal pthread parallel-for (this image has no libgomp).  With g_threads == 1 the body runs
 * inline in index order = the serial thread order of Warp's CPU device [Warp]. */
typedef void (*pf_body)(int begin, int end, void* ctx, long* acc);
typedef struct {
  pf_body body;
  void* ctx;
  int n, chunk;
  atomic_int* next;
  long acc;
} pf_arg;
static void* pf_worker(void* p) {
  pf_arg* a = (pf_arg*)p;
  for (;;) {
    int b = atomic_fetch_add(a->next, a->chunk);
    if (b >= a->n) break;
    int e = b + a->chunk < a->n ? b + a->chunk : a->n;
    a->body(b, e, a->ctx, &a->acc);
  }
  return NULL;
}
static long parallel_for(int n, int chunk, pf_body body, void* ctx) {
  long acc = 0;
  if (g_threads <= 1 || n <= chunk) {
    body(0, n, ctx, &acc);
    return acc;
  }
  int nt = g_threads;
  pthread_t th[256];
  pf_arg args[256];
  atomic_int next = 0;
  for (int t = 0; t < nt; ++t) {
    args[t] = (pf_arg){body, ctx, n, chunk, &next, 0};
    pthread_create(&th[t], NULL, pf_worker, &args[t]);
  }
  for (int t = 0; t < nt; ++t) {
    pthread_join(th[t], NULL);
    acc += args[t].acc;
  }
  return acc;
}

/* ----------------------------------------------------------------------------------------- */
/* scalar helpers                                                                             */
/* ----------------------------------------------------------------------------------------- */
static inline float f_min(float a, float b) { return (a < b) ? a : b; } /* [Warp] wp.min */
static inline float f_max(float a, float b) { return (a > b) ? a : b; } /* [Warp] wp.max */
static inline int i_min(int a, int b) { return (a < b) ? a : b; }
static inline int i_max(int a, int b) { return (a > b) ? a : b; }

/* float -> int32, truncating, saturating, NaN -> 0 (what int(x)/wp.int32(x) compiles to on the
 * reference's CUDA device: cvt.rzi.s32.f32) */
static inline int f2i(float x) {
  if (x != x) return 0;
  if (x >= 2147483648.0f) return 2147483647;
  if (x <= -2147483648.0f) return (-2147483647 - 1);
  return (int)x;
}

static inline uint32_t float_bits(float x) { /* forward.py:51-57 */
  uint32_t u;
  memcpy(&u, &x, 4);
  return u;
}
static inline float bits_float(uint32_t u) {
  float x;
  memcpy(&x, &u, 4);
  return x;
}

/* Deterministic exp for x <= ~88 (the blend only calls it with x <= 0).  Operation sequence is
 * part of the arithmetic contract: Cody-Waite reduction with round-to-nearest-even via the
 * 1.5*2^23 trick, degree-5 polynomial (Cephes expf coefficients) in Horner form with fused
 * multiply-adds, exact scaling by 2^n.  Max observed error vs correctly-rounded exp: < 1 ulp
 * (tests/test_oracle_math.py). */
GSO_API float gso_expf_det(float x) {
  if (x < -87.0f) return 0.0f;
  float t = x * 1.44269504088896341f;
  float n = (t + 12582912.0f) - 12582912.0f;
  float r = fmaf(n, -0.693145751953125f, x);
  r = fmaf(n, -1.42860682030941723212e-6f, r);
  float p = 1.9875691500e-4f;
  p = fmaf(p, r, 1.3981999507e-3f);
  p = fmaf(p, r, 8.3334519073e-3f);
  p = fmaf(p, r, 4.1665795894e-2f);
  p = fmaf(p, r, 1.6666665459e-1f);
  p = fmaf(p, r, 5.0000001201e-1f);
  float r2 = r * r;
  float y = fmaf(p, r2, r);
  y = y + 1.0f;
  int ni = (int)n;
  float scale = bits_float((uint32_t)(ni + 127) << 23);
  return y * scale;
}

static inline float gs_exp(float x) { return g_exp_mode == GSO_EXP_DET ? gso_expf_det(x) : expf(x); }

/* [Warp] row-vector * mat44 :  r[j] = sum_i v[i]*m[i][j], accumulated i = 0..3 in order
 * (native/mat.h mul(vec,mat): r = row(0)*v[0]; r += row(i)*v[i]).  m is row-major 4x4. */
static inline void vec4_mul_mat44(const float v[4], const float* m, float r[4]) {
  for (int j = 0; j < 4; ++j) {
    float s = m[0 * 4 + j] * v[0];
    s = s + m[1 * 4 + j] * v[1];
    s = s + m[2 * 4 + j] * v[2];
    s = s + m[3 * 4 + j] * v[3];
    r[j] = s;
  }
}

/* [Warp] mat33 * mat33: t[i][j] = 0; for k: t[i][j] += a[i][k]*b[k][j] */
static inline void mat33_mul(const float a[9], const float b[9], float t[9]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      float s = 0.0f;
      for (int k = 0; k < 3; ++k) s = s + a[i * 3 + k] * b[k * 3 + j];
      t[i * 3 + j] = s;
    }
}
static inline void mat33_transpose(const float a[9], float t[9]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t[i * 3 + j] = a[j * 3 + i];
}
/* [Warp] dot: r = a0*b0; r += ai*bi */
static inline float dot3(const float a[3], const float b[3]) {
  float s = a[0] * b[0];
  s = s + a[1] * b[1];
  s = s + a[2] * b[2];
  return s;
}
static inline float length3(const float a[3]) { return sqrtf(dot3(a, a)); } /* [Warp] wp.length */

/* ----------------------------------------------------------------------------------------- */
/* forward device functions                                                                   */
/* ----------------------------------------------------------------------------------------- */

/* forward.py:59-61 */
static inline float ndc2pix(float x, float size) { return ((x + 1.0f) * size - 1.0f) * 0.5f; }

/* forward.py:63-76.  tile grid arrives as FLOATS (wp.vec3) and is cast with wp.int32(). */
static inline void get_rect(float px, float py, float max_radius, float grid_x, float grid_y, int* rect_min_x,
                            int* rect_min_y, int* rect_max_x, int* rect_max_y) {
  *rect_min_x = i_min(f2i(grid_x), i_max(0, f2i((px - max_radius) / (float)TILE_M)));
  *rect_min_y = i_min(f2i(grid_y), i_max(0, f2i((py - max_radius) / (float)TILE_N)));
  *rect_max_x = i_min(f2i(grid_x), i_max(0, f2i((px + max_radius + (float)TILE_M - 1.0f) / (float)TILE_M)));
  *rect_max_y = i_min(f2i(grid_y), i_max(0, f2i((py + max_radius + (float)TILE_N - 1.0f) / (float)TILE_N)));
}

/* [Warp] wp.quat_to_matrix(q=(x,y,z,w)): columns are quat_rotate(q, e_k):
 *   quat_rotate(q,v) = v*(2w^2-1) + q.xyz*(2*dot(q.xyz,v)) + cross(q.xyz,v)*w*2
 * Equals the textbook rotation matrix only for unit q (quirk F4: forward never normalises).
 * Terms multiplied by the zero components of e_k are exact zeros and are dropped here. */
static inline void quat_to_matrix(float qx, float qy, float qz, float qw, float R[9]) {
  float c = 2.0f * qw * qw - 1.0f;
  float d;
  /* column 0: v = (1,0,0) */
  d = 2.0f * qx;
  R[0 * 3 + 0] = c + qx * d;
  R[1 * 3 + 0] = qy * d + qz * qw * 2.0f;
  R[2 * 3 + 0] = qz * d + (-qy) * qw * 2.0f;
  /* column 1: v = (0,1,0) */
  d = 2.0f * qy;
  R[0 * 3 + 1] = qx * d + (-qz) * qw * 2.0f;
  R[1 * 3 + 1] = c + qy * d;
  R[2 * 3 + 1] = qz * d + qx * qw * 2.0f;
  /* column 2: v = (0,0,1) */
  d = 2.0f * qz;
  R[0 * 3 + 2] = qx * d + qy * qw * 2.0f;
  R[1 * 3 + 2] = qy * d + (-qx) * qw * 2.0f;
  R[2 * 3 + 2] = c + qz * d;
}

/* forward.py:146-186 compute_cov3d -> VEC6 (xx,xy,xz,yy,yz,zz) */
static inline void compute_cov3d(const float scale[3], float scale_mod, const float rot[4], float cov6[6]) {
  float S[9] = {scale_mod * scale[0], 0.0f, 0.0f, 0.0f, scale_mod * scale[1], 0.0f, 0.0f, 0.0f, scale_mod * scale[2]};
  float R[9], M[9], Mt[9], sigma[9];
  quat_to_matrix(rot[0], rot[1], rot[2], rot[3], R);
  mat33_mul(R, S, M);
  mat33_transpose(M, Mt);
  mat33_mul(M, Mt, sigma);
  cov6[0] = sigma[0];
  cov6[1] = sigma[1];
  cov6[2] = sigma[2];
  cov6[3] = sigma[4];
  cov6[4] = sigma[5];
  cov6[5] = sigma[8];
}

/* forward.py:79-144 compute_cov2d.  Quirk G1: W = view_matrix[:3,:3] as stored (row-vector
 * convention), i.e. the transpose of the rotation the maths needs.  Reproduced literally. */
static inline void compute_cov2d(const float p[3], const float cov3d[6], const float* V, float tan_fovx, float tan_fovy,
                                 float width, float height, float out[3]) {
  float pv[4] = {p[0], p[1], p[2], 1.0f};
  float t[4];
  vec4_mul_mat44(pv, V, t);
  float limx = 1.3f * tan_fovx;
  float limy = 1.3f * tan_fovy;
  float txtz = t[0] / t[2];
  float tytz = t[1] / t[2];
  t[0] = f_min(limx, f_max(-limx, txtz)) * t[2];
  t[1] = f_min(limy, f_max(-limy, tytz)) * t[2];
  float focal_x = width / (2.0f * tan_fovx);
  float focal_y = height / (2.0f * tan_fovy);
  float J[9] = {focal_x / t[2], 0.0f, -(focal_x * t[0]) / (t[2] * t[2]), 0.0f, focal_y / t[2],
                -(focal_y * t[1]) / (t[2] * t[2]), 0.0f, 0.0f, 0.0f};
  float W[9] = {V[0], V[1], V[2], V[4], V[5], V[6], V[8], V[9], V[10]};
  float T[9], Tt[9], A[9], cov[9];
  mat33_mul(J, W, T);
  float Vrk[9] = {cov3d[0], cov3d[1], cov3d[2], cov3d[1], cov3d[3], cov3d[4], cov3d[2], cov3d[4], cov3d[5]};
  float Vrkt[9];
  mat33_transpose(Vrk, Vrkt);
  mat33_transpose(T, Tt);
  mat33_mul(T, Vrkt, A); /* cov = T * transpose(Vrk) * transpose(T), left to right */
  mat33_mul(A, Tt, cov);
  out[0] = cov[0];
  out[1] = cov[1];
  out[2] = cov[4];
}

/* SH constants, forward.py:44-45, 330-344; backward.py:158-162,190-196 */
#define SH_C0 0.28209479177387814f
#define SH_C1 0.4886025119029199f
#define C2_0 1.0925484305920792f
#define C2_1 (-1.0925484305920792f)
#define C2_2 0.31539156525252005f
#define C2_3 (-1.0925484305920792f)
#define C2_4 0.5462742152960396f
#define C3_0 (-0.5900435899266435f)
#define C3_1 2.890611442640554f
#define C3_2 (-0.4570457994644658f)
#define C3_3 0.3731763325901154f
#define C3_4 (-0.4570457994644658f)
#define C3_5 1.445305721320277f
#define C3_6 (-0.5900435899266435f)

/* forward.py:189-382 wp_preprocess, one Gaussian.  Outputs must be zero-initialised by the
 * caller (forward.py:703-710): culled Gaussians keep zeros (cov3D excepted, line 260). */
static void preprocess_one(int i, const float* means, const float* scales, float scale_modifier, const float* rots,
                           const float* opac, const float* shs, int degree, int clamped, const float* V, const float* P,
                           const float* campos, int W, int H, float tan_fovx, float tan_fovy, float grid_x, float grid_y,
                           int* radii, float* xy, float* depths, float* cov3Ds, float* rgb, float* conic_opacity,
                           int* tiles_touched, float* clamped_state) {
  const float* p_orig = means + 3 * i;
  float ph[4] = {p_orig[0], p_orig[1], p_orig[2], 1.0f};
  float p_view[4];
  vec4_mul_mat44(ph, V, p_view);
  if (p_view[2] < 0.2f) return; /* forward.py:250 */

  float p_hom[4];
  vec4_mul_mat44(ph, P, p_hom);
  float p_w = 1.0f / (p_hom[3] + 0.0000001f);
  float p_proj[3] = {p_hom[0] * p_w, p_hom[1] * p_w, p_hom[2] * p_w};

  float cov3d[6];
  compute_cov3d(scales + 3 * i, scale_modifier, rots + 4 * i, cov3d);
  for (int k = 0; k < 6; ++k) cov3Ds[6 * i + k] = cov3d[k]; /* forward.py:260 */

  float cov2d[3];
  compute_cov2d(p_orig, cov3d, V, tan_fovx, tan_fovy, (float)W, (float)H, cov2d);

  float h_var = 0.3f;
  float W_float = (float)W;
  float H_float = (float)H;
  float cb0 = cov2d[0] + h_var, cb1 = cov2d[1], cb2 = cov2d[2] + h_var;
  float det = cb0 * cb2 - cb1 * cb1;
  if (det == 0.0f) return; /* forward.py:278 */

  float det_inv = 1.0f / det;
  float conic[3] = {cb2 * det_inv, -cb1 * det_inv, cb0 * det_inv};
  float mid = 0.5f * (cb0 + cb2);
  float lambda1 = mid + sqrtf(f_max(0.1f, mid * mid - det));
  float lambda2 = mid - sqrtf(f_max(0.1f, mid * mid - det));
  float my_radius = ceilf(3.0f * sqrtf(f_max(lambda1, lambda2)));
  float pix = ndc2pix(p_proj[0], W_float);
  float piy = ndc2pix(p_proj[1], H_float);

  int rminx, rminy, rmaxx, rmaxy;
  get_rect(pix, piy, my_radius, grid_x, grid_y, &rminx, &rminy, &rmaxx, &rmaxy);
  if ((rmaxx - rminx) * (rmaxy - rminy) == 0) return; /* forward.py:301 */

  /* SH -> RGB, forward.py:304-346 */
  float dir_orig[3] = {p_orig[0] - campos[0], p_orig[1] - campos[1], p_orig[2] - campos[2]};
  float len = length3(dir_orig);
  float dir[3] = {0.0f, 0.0f, 0.0f}; /* [Warp] normalize: v/len if len > 0 else 0 */
  if (len > 0.0f) {
    dir[0] = dir_orig[0] / len;
    dir[1] = dir_orig[1] / len;
    dir[2] = dir_orig[2] / len;
  }
  float x = dir[0], y = dir[1], z = dir[2];
  const float* sh = shs + (size_t)i * 48;
  float result[3];
  for (int c = 0; c < 3; ++c) {
#define SHK(k) sh[(k) * 3 + c]
    float r = SH_C0 * SHK(0);
    if (degree > 0) {
      r = r - SH_C1 * y * SHK(1) + SH_C1 * z * SHK(2) - SH_C1 * x * SHK(3);
      if (degree > 1) {
        float xx = x * x, yy = y * y, zz = z * z, xy_ = x * y, yz = y * z, xz = x * z;
        r = r + C2_0 * xy_ * SHK(4);
        r = r + C2_1 * yz * SHK(5);
        r = r + C2_2 * (2.0f * zz - xx - yy) * SHK(6);
        r = r + C2_3 * xz * SHK(7);
        r = r + C2_4 * (xx - yy) * SHK(8);
        if (degree > 2) {
          r = r + C3_0 * y * (3.0f * xx - yy) * SHK(9);
          r = r + C3_1 * xy_ * z * SHK(10);
          r = r + C3_2 * y * (4.0f * zz - xx - yy) * SHK(11);
          r = r + C3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * SHK(12);
          r = r + C3_4 * x * (4.0f * zz - xx - yy) * SHK(13);
          r = r + C3_5 * z * (xx - yy) * SHK(14);
          r = r + C3_6 * x * (xx - 3.0f * yy) * SHK(15);
        }
      }
    }
#undef SHK
    r = r + 0.5f;
    result[c] = r;
  }
  for (int c = 0; c < 3; ++c) clamped_state[3 * i + c] = (result[c] < 0.0f) ? 1.0f : 0.0f; /* forward.py:351-362 */
  if (clamped)
    for (int c = 0; c < 3; ++c) result[c] = f_max(result[c], 0.0f);
  for (int c = 0; c < 3; ++c) rgb[3 * i + c] = result[c];

  depths[i] = p_view[2];
  radii[i] = f2i(my_radius);
  xy[2 * i + 0] = pix;
  xy[2 * i + 1] = piy;
  conic_opacity[4 * i + 0] = conic[0];
  conic_opacity[4 * i + 1] = conic[1];
  conic_opacity[4 * i + 2] = conic[2];
  conic_opacity[4 * i + 3] = opac[i];
  tiles_touched[i] = (rmaxy - rminy) * (rmaxx - rminx);
}

typedef struct {
  const float *means, *scales, *rots, *opac, *shs, *view, *proj, *campos;
  float scale_modifier, tan_fovx, tan_fovy, grid_x, grid_y;
  int degree, clamped, W, H;
  int *radii, *tiles_touched;
  float *xy, *depths, *cov3Ds, *rgb, *conic_opacity, *clamped_state;
} pre_ctx;
static void pre_body(int b, int e, void* p, long* acc) {
  pre_ctx* c = (pre_ctx*)p;
  (void)acc;
  for (int i = b; i < e; ++i)
    preprocess_one(i, c->means, c->scales, c->scale_modifier, c->rots, c->opac, c->shs, c->degree, c->clamped, c->view,
                   c->proj, c->campos, c->W, c->H, c->tan_fovx, c->tan_fovy, c->grid_x, c->grid_y, c->radii, c->xy,
                   c->depths, c->cov3Ds, c->rgb, c->conic_opacity, c->tiles_touched, c->clamped_state);
}
GSO_API void gso_preprocess(int N, const float* means, const float* scales, float scale_modifier, const float* rots,
                            const float* opac, const float* shs, int degree, int clamped, const float* view,
                            const float* proj, const float* campos, int W, int H, float tan_fovx, float tan_fovy,
                            int* radii, float* xy, float* depths, float* cov3Ds, float* rgb, float* conic_opacity,
                            int* tiles_touched, float* clamped_state) {
  /* tile grid as floats, forward.py:698-700 */
  pre_ctx c = {means, scales, rots, opac, shs, view, proj, campos, scale_modifier, tan_fovx, tan_fovy,
               (float)((W + TILE_M - 1) / TILE_M), (float)((H + TILE_N - 1) / TILE_N), degree, clamped, W, H,
               radii, tiles_touched, xy, depths, cov3Ds, rgb, conic_opacity, clamped_state};
  parallel_for(N, 1024, pre_body, &c);
}

/* utils/wp_utils.py:46-60 wp_prefix_sum: serial inclusive scan */
GSO_API void gso_prefix_sum(int N, const int* in, int* out) {
  if (N <= 0) return;
  out[0] = in[0];
  for (int i = 1; i < N; ++i) out[i] = out[i - 1] + in[i];
}

/* forward.py:517-558 wp_duplicate_with_keys */
typedef struct {
  const float *xy, *depths;
  const int *point_offsets, *radii;
  float grid_x, grid_y;
  int64_t* keys;
  int* vals;
} dup_ctx;
static void dup_body(int b, int e, void* p, long* acc) {
  dup_ctx* c = (dup_ctx*)p;
  (void)acc;
  for (int tid = b; tid < e; ++tid) {
    int r = c->radii[tid];
    if (r <= 0) continue;
    int offset = 0;
    if (tid > 0) offset = c->point_offsets[tid - 1];
    int rminx, rminy, rmaxx, rmaxy;
    get_rect(c->xy[2 * tid], c->xy[2 * tid + 1], (float)r, c->grid_x, c->grid_y, &rminx, &rminy, &rmaxx, &rmaxy);
    float depth_val = c->depths[tid];
    for (int y = rminy; y < rmaxy; ++y)
      for (int x = rminx; x < rmaxx; ++x) {
        int tile_id = y * f2i(c->grid_x) + x;
        int64_t key = ((int64_t)tile_id << 32) | (int64_t)float_bits(depth_val);
        c->keys[offset] = key;
        c->vals[offset] = tid;
        offset += 1;
      }
  }
}
GSO_API void gso_duplicate_with_keys(int N, const float* xy, const float* depths, const int* point_offsets,
                                     const int* radii, int W, int H, int64_t* keys, int* vals) {
  dup_ctx c = {xy, depths, point_offsets, radii, (float)((W + TILE_M - 1) / TILE_M),
               (float)((H + TILE_N - 1) / TILE_N), keys, vals};
  parallel_for(N, 1024, dup_body, &c);
}

/* forward.py:799-803 wp.utils.radix_sort_pairs [Warp]: ascending, STABLE, all 64 key bits.
 * Restated as a stable LSD radix sort (8 passes of 8 bits).  tmp buffers are caller-provided. */
GSO_API void gso_sort_pairs(int64_t* keys, int* vals, int n, int64_t* tmp_keys, int* tmp_vals) {
  int64_t* ks = keys;
  int* vs = vals;
  int64_t* kd = tmp_keys;
  int* vd = tmp_vals;
  for (int pass = 0; pass < 8; ++pass) {
    size_t hist[257];
    memset(hist, 0, sizeof(hist));
    int shift = pass * 8;
    for (int i = 0; i < n; ++i) hist[(((uint64_t)ks[i]) >> shift & 0xFF) + 1]++;
    if (hist[1] == (size_t)n) continue; /* all digits zero: pass is the identity */
    for (int d = 0; d < 256; ++d) hist[d + 1] += hist[d];
    for (int i = 0; i < n; ++i) {
      size_t dst = hist[((uint64_t)ks[i]) >> shift & 0xFF]++;
      kd[dst] = ks[i];
      vd[dst] = vs[i];
    }
    int64_t* tk = ks;
    ks = kd;
    kd = tk;
    int* tv = vs;
    vs = vd;
    vd = tv;
  }
  if (ks != keys) {
    memcpy(keys, ks, sizeof(int64_t) * (size_t)n);
    memcpy(vals, vs, sizeof(int) * (size_t)n);
  }
}

/* forward.py:560-586 wp_identify_tile_ranges; ranges is (start,end) int32 pairs, pre-zeroed */
GSO_API void gso_identify_tile_ranges(int num_rendered, const int64_t* keys, int* ranges) {
  for (int idx = 0; idx < num_rendered; ++idx) {
    int curr_tile = (int)(keys[idx] >> 32);
    if (idx == 0) {
      ranges[2 * curr_tile + 0] = 0;
    } else {
      int prev_tile = (int)(keys[idx - 1] >> 32);
      if (curr_tile != prev_tile) {
        ranges[2 * prev_tile + 1] = idx;
        ranges[2 * curr_tile + 0] = idx;
      }
    }
    if (idx == num_rendered - 1) ranges[2 * curr_tile + 1] = num_rendered;
  }
}

/* forward.py:384-515 wp_render_gaussians, one pixel.  Returns the number of (pixel,Gaussian)
 * pairs evaluated (work counter K_fwd). */
static inline long render_pixel(int pix_x, int pix_y, int tile_id, const int* ranges, const int* point_list, int W,
                                const float* xy, const float* colors, const float* conic_opacity, const float* depths,
                                const float* bg, float* image, float* depth_image, float* final_Ts, int* n_contrib) {
  float pixf_x = (float)pix_x;
  float pixf_y = (float)pix_y;
  int range_start = ranges[2 * tile_id + 0];
  int range_end = ranges[2 * tile_id + 1];
  float T = 1.0f;
  float r = 0.0f, g = 0.0f, b = 0.0f;
  float expected_inv_depth = 0.0f;
  int contributor_count = 0;
  int last_contributor = 0;
  long pairs = 0;
  for (int i = range_start; i < range_end; ++i) {
    int gid = point_list[i];
    float gx = xy[2 * gid], gy = xy[2 * gid + 1];
    const float* con_o = conic_opacity + 4 * gid;
    const float* color = colors + 3 * gid;
    float d_x = gx - pixf_x;
    float d_y = gy - pixf_y;
    contributor_count += 1;
    pairs += 1;
    float power = -0.5f * (con_o[0] * d_x * d_x + con_o[2] * d_y * d_y) - con_o[1] * d_x * d_y;
    if (power > 0.0f) continue;
    float alpha = f_min(0.99f, con_o[3] * gs_exp(power));
    if (alpha < (1.0f / 255.0f)) continue;
    float test_T = T * (1.0f - alpha);
    if (test_T < 0.0001f) break;
    r += color[0] * alpha * T;
    g += color[1] * alpha * T;
    b += color[2] * alpha * T;
    expected_inv_depth += (1.0f / depths[gid]) * alpha * T;
    T = test_T;
    last_contributor = contributor_count;
  }
  size_t pidx = (size_t)pix_y * W + pix_x;
  final_Ts[pidx] = T;
  n_contrib[pidx] = last_contributor;
  image[3 * pidx + 0] = r + T * bg[0];
  image[3 * pidx + 1] = g + T * bg[1];
  image[3 * pidx + 2] = b + T * bg[2];
  depth_image[pidx] = expected_inv_depth;
  return pairs;
}

typedef struct {
  int W, H, gx, gy;
  const int *ranges, *point_list;
  const float *xy, *colors, *conic_opacity, *depths, *bg;
  float *image, *depth_image, *final_Ts;
  int* n_contrib;
} ren_ctx;
/* [Warp] launch dim (gx,gy,16,16), last index fastest: tile_x outermost ... tid_y innermost.
 * The linear tile index t = tile_x*gy + tile_y reproduces that order serially. */
static void ren_body(int b, int e, void* p, long* acc) {
  ren_ctx* c = (ren_ctx*)p;
  for (int t = b; t < e; ++t) {
    int tile_x = t / c->gy, tile_y = t % c->gy;
    for (int tid_x = 0; tid_x < TILE_M; ++tid_x)
      for (int tid_y = 0; tid_y < TILE_N; ++tid_y) {
        int pix_x = tile_x * TILE_M + tid_x, pix_y = tile_y * TILE_N + tid_y;
        if (!(pix_x < c->W && pix_y < c->H)) continue;
        *acc += render_pixel(pix_x, pix_y, tile_y * c->gx + tile_x, c->ranges, c->point_list, c->W, c->xy, c->colors,
                             c->conic_opacity, c->depths, c->bg, c->image, c->depth_image, c->final_Ts, c->n_contrib);
      }
  }
}
GSO_API long gso_render(int W, int H, const int* ranges, const int* point_list, const float* xy, const float* colors,
                        const float* conic_opacity, const float* depths, const float* bg, float* image,
                        float* depth_image, float* final_Ts, int* n_contrib) {
  int gx = (W + TILE_M - 1) / TILE_M, gy = (H + TILE_N - 1) / TILE_N;
  ren_ctx c = {W, H, gx, gy, ranges, point_list, xy, colors, conic_opacity, depths, bg, image, depth_image, final_Ts,
               n_contrib};
  return parallel_for(gx * gy, 1, ren_body, &c);
}

/* ----------------------------------------------------------------------------------------- */
/* backward                                                                                   */
/* ----------------------------------------------------------------------------------------- */
static inline void atomic_addf(float* p, float v) {
  if (g_threads > 1) { /* CAS loop; only taken by the multi-threaded cpu_baseline leg */
    _Atomic uint32_t* a = (_Atomic uint32_t*)p;
    uint32_t old = atomic_load_explicit(a, memory_order_relaxed);
    for (;;) {
      float nv = bits_float(old) + v;
      if (atomic_compare_exchange_weak_explicit(a, &old, float_bits(nv), memory_order_relaxed, memory_order_relaxed))
        return;
    }
  }
  *p += v;
}

/* backward.py:558-706 wp_render_backward_kernel, one pixel */
static inline long render_backward_pixel(int pix_x, int pix_y, int tile_id, const int* ranges, const int* point_list,
                                         int W, int H, const float* bg, const float* xy, const float* conic_opacity,
                                         const float* colors, const float* final_Ts, const int* n_contrib,
                                         const float* dL_dpixels, float* dL_dmean2D, float* dL_dconic2D,
                                         float* dL_dopacity, float* dL_dcolors) {
  float pixf_x = (float)pix_x;
  float pixf_y = (float)pix_y;
  int range_start = ranges[2 * tile_id + 0];
  int range_end = ranges[2 * tile_id + 1];
  size_t pidx = (size_t)pix_y * W + pix_x;
  float T_final = final_Ts[pidx];
  int last_contributor = n_contrib[pidx];
  int last_kept = i_min(range_end, range_start + last_contributor);
  float T = T_final;
  float accum_rec[3] = {0.0f, 0.0f, 0.0f};
  float last_alpha = 0.0f;
  float last_color[3] = {0.0f, 0.0f, 0.0f};
  const float* dL_dpixel = dL_dpixels + 3 * pidx;
  float ddelx_dx = 0.5f * (float)W;
  float ddely_dy = 0.5f * (float)H;
  long pairs = 0;
  for (int i = last_kept - 1; i > range_start - 1; --i) {
    int gid = point_list[i];
    float gx = xy[2 * gid], gy = xy[2 * gid + 1];
    const float* con_o = conic_opacity + 4 * gid;
    const float* color = colors + 3 * gid;
    float d_x = gx - pixf_x;
    float d_y = gy - pixf_y;
    pairs += 1;
    float power = -0.5f * (con_o[0] * d_x * d_x + con_o[2] * d_y * d_y) - con_o[1] * d_x * d_y;
    if (power > 0.0f) continue;
    float G = gs_exp(power);
    float alpha = f_min(0.99f, con_o[3] * G);
    if (alpha < (1.0f / 255.0f)) continue;
    T = T / (1.0f - alpha);
    float dchannel_dcolor = alpha * T;
    float dL_dalpha = 0.0f;
    for (int c = 0; c < 3; ++c) accum_rec[c] = last_alpha * last_color[c] + (1.0f - last_alpha) * accum_rec[c];
    for (int c = 0; c < 3; ++c) last_color[c] = color[c];
    {
      float diff[3] = {color[0] - accum_rec[0], color[1] - accum_rec[1], color[2] - accum_rec[2]};
      dL_dalpha = dot3(diff, dL_dpixel);
    }
    for (int c = 0; c < 3; ++c) atomic_addf(&dL_dcolors[3 * gid + c], dchannel_dcolor * dL_dpixel[c]);
    dL_dalpha *= T;
    last_alpha = alpha;
    float bg_dot_dpixel = dot3(bg, dL_dpixel);
    dL_dalpha += (-T_final / (1.0f - alpha)) * bg_dot_dpixel;
    float dL_dG = con_o[3] * dL_dalpha;
    float gdx = G * d_x;
    float gdy = G * d_y;
    float dG_ddelx = -gdx * con_o[0] - gdy * con_o[1];
    float dG_ddely = -gdy * con_o[2] - gdx * con_o[1];
    atomic_addf(&dL_dmean2D[3 * gid + 0], dL_dG * dG_ddelx * ddelx_dx);
    atomic_addf(&dL_dmean2D[3 * gid + 1], dL_dG * dG_ddely * ddely_dy);
    /* z component: += 0.0 (backward.py:694) */
    atomic_addf(&dL_dconic2D[4 * gid + 0], -0.5f * gdx * d_x * dL_dG);
    atomic_addf(&dL_dconic2D[4 * gid + 1], -0.5f * gdx * d_y * dL_dG);
    /* component 2: += 0.0 (backward.py:701) */
    atomic_addf(&dL_dconic2D[4 * gid + 3], -0.5f * gdy * d_y * dL_dG);
    atomic_addf(&dL_dopacity[gid], G * dL_dalpha);
  }
  (void)H;
  return pairs;
}

typedef struct {
  int W, H, gx, gy;
  const int *ranges, *point_list, *n_contrib;
  const float *bg, *xy, *conic_opacity, *colors, *final_Ts, *dL_dpixels;
  float *dL_dmean2D, *dL_dconic2D, *dL_dopacity, *dL_dcolors;
} rbw_ctx;
static void rbw_body(int b, int e, void* p, long* acc) {
  rbw_ctx* c = (rbw_ctx*)p;
  for (int t = b; t < e; ++t) {
    int tile_x = t / c->gy, tile_y = t % c->gy;
    for (int tid_x = 0; tid_x < TILE_M; ++tid_x)
      for (int tid_y = 0; tid_y < TILE_N; ++tid_y) {
        int pix_x = tile_x * TILE_M + tid_x, pix_y = tile_y * TILE_N + tid_y;
        if (!(pix_x < c->W && pix_y < c->H)) continue;
        *acc += render_backward_pixel(pix_x, pix_y, tile_y * c->gx + tile_x, c->ranges, c->point_list, c->W, c->H,
                                      c->bg, c->xy, c->conic_opacity, c->colors, c->final_Ts, c->n_contrib,
                                      c->dL_dpixels, c->dL_dmean2D, c->dL_dconic2D, c->dL_dopacity, c->dL_dcolors);
      }
  }
}
GSO_API long gso_render_backward(int W, int H, const int* ranges, const int* point_list, const float* bg,
                                 const float* xy, const float* conic_opacity, const float* colors, const float* final_Ts,
                                 const int* n_contrib, const float* dL_dpixels, float* dL_dmean2D, float* dL_dconic2D,
                                 float* dL_dopacity, float* dL_dcolors) {
  int gx = (W + TILE_M - 1) / TILE_M, gy = (H + TILE_N - 1) / TILE_N;
  rbw_ctx c = {W, H, gx, gy, ranges, point_list, n_contrib, bg, xy, conic_opacity, colors, final_Ts, dL_dpixels,
               dL_dmean2D, dL_dconic2D, dL_dopacity, dL_dcolors};
  return parallel_for(gx * gy, 1, rbw_body, &c);
}

/* backward.py:258-435 compute_cov2d_backward_kernel, one Gaussian */
static void cov2d_backward_one(int idx, const float* means, const float* cov3Ds, const int* radii, float h_x, float h_y,
                               float tan_fovx, float tan_fovy, const float* V, const float* dL_dconics,
                               float* dL_dmeans, float* dL_dcov3Ds) {
  if (radii[idx] <= 0) {
    for (int k = 0; k < 6; ++k) dL_dcov3Ds[6 * idx + k] = 0.0f;
    return;
  }
  const float* mean = means + 3 * idx;
  const float* c3 = cov3Ds + 6 * idx;
  float dL_dconic[3] = {dL_dconics[4 * idx + 0], dL_dconics[4 * idx + 1], dL_dconics[4 * idx + 3]};
  float mh[4] = {mean[0], mean[1], mean[2], 1.0f};
  float t[4];
  vec4_mul_mat44(mh, V, t);
  float limx = 1.3f * tan_fovx;
  float limy = 1.3f * tan_fovy;
  float tz = t[2];
  float inv_tz = 1.0f / tz;
  float txtz = t[0] * inv_tz;
  float tytz = t[1] * inv_tz;
  int x_clamped_flag = (txtz < -limx) || (txtz > limx);
  int y_clamped_flag = (tytz < -limy) || (tytz > limy);
  float x_grad_mul = 1.0f - (float)x_clamped_flag;
  float y_grad_mul = 1.0f - (float)y_clamped_flag;
  float tx = f_min(limx, f_max(-limx, txtz)) * tz;
  float ty = f_min(limy, f_max(-limy, tytz)) * tz;
  float inv_tz2 = inv_tz * inv_tz;
  float inv_tz3 = inv_tz2 * inv_tz;
  float J00 = h_x * inv_tz;
  float J11 = h_y * inv_tz;
  float J02 = -h_x * tx * inv_tz2;
  float J12 = -h_y * ty * inv_tz2;
  /* J = transpose(mat33(J00,0,J02, 0,J11,J12, 0,0,0)) */
  float J[9] = {J00, 0.0f, 0.0f, 0.0f, J11, 0.0f, J02, J12, 0.0f};
  float Wm[9] = {V[0], V[1], V[2], V[4], V[5], V[6], V[8], V[9], V[10]};
  float T[9];
  mat33_mul(Wm, J, T);
  float c0 = c3[0], c1 = c3[1], c2 = c3[2], c11 = c3[3], c12 = c3[4], c22 = c3[5];
  float Vrk[9] = {c0, c1, c2, c1, c11, c12, c2, c12, c22};
  float Tt[9], Vrkt[9], A[9], cov2D[9];
  mat33_transpose(T, Tt);
  mat33_transpose(Vrk, Vrkt);
  mat33_mul(Tt, Vrkt, A); /* transpose(T) * transpose(Vrk) * T */
  mat33_mul(A, T, cov2D);
  float a = cov2D[0] + 0.3f;
  float b = cov2D[1];
  float c = cov2D[4] + 0.3f;
  float denom = a * c - b * b;
  float dL_da = 0.0f, dL_db = 0.0f, dL_dc = 0.0f;
  if (denom != 0.0f) {
    float denom2inv = 1.0f / (denom * denom + 1e-7f);
    dL_da = denom2inv * (-c * c * dL_dconic[0] + 2.0f * b * c * dL_dconic[1] + (denom - a * c) * dL_dconic[2]);
    dL_dc = denom2inv * (-a * a * dL_dconic[2] + 2.0f * a * b * dL_dconic[1] + (denom - a * c) * dL_dconic[0]);
    dL_db = denom2inv * 2.0f * (b * c * dL_dconic[0] - (denom + 2.0f * b * b) * dL_dconic[1] + a * b * dL_dconic[2]);
  }
#define T_(i, j) T[(i) * 3 + (j)]
#define V_(i, j) Vrk[(i) * 3 + (j)]
  float* o = dL_dcov3Ds + 6 * idx;
  o[0] = T_(0, 0) * T_(0, 0) * dL_da + T_(0, 0) * T_(0, 1) * dL_db + T_(0, 1) * T_(0, 1) * dL_dc;
  o[1] = 2.0f * T_(0, 0) * T_(1, 0) * dL_da + (T_(0, 0) * T_(1, 1) + T_(1, 0) * T_(0, 1)) * dL_db +
         2.0f * T_(0, 1) * T_(1, 1) * dL_dc;
  o[2] = 2.0f * T_(0, 0) * T_(2, 0) * dL_da + (T_(0, 0) * T_(2, 1) + T_(2, 0) * T_(0, 1)) * dL_db +
         2.0f * T_(0, 1) * T_(2, 1) * dL_dc;
  o[3] = T_(1, 0) * T_(1, 0) * dL_da + T_(1, 0) * T_(1, 1) * dL_db + T_(1, 1) * T_(1, 1) * dL_dc;
  o[4] = 2.0f * T_(2, 0) * T_(1, 0) * dL_da + (T_(1, 0) * T_(2, 1) + T_(2, 0) * T_(1, 1)) * dL_db +
         2.0f * T_(1, 1) * T_(2, 1) * dL_dc;
  o[5] = T_(2, 0) * T_(2, 0) * dL_da + T_(2, 0) * T_(2, 1) * dL_db + T_(2, 1) * T_(2, 1) * dL_dc;

  float dL_dT00 = 2.0f * (T_(0, 0) * V_(0, 0) + T_(1, 0) * V_(1, 0) + T_(2, 0) * V_(2, 0)) * dL_da +
                  (T_(0, 1) * V_(0, 0) + T_(1, 1) * V_(1, 0) + T_(2, 1) * V_(2, 0)) * dL_db;
  float dL_dT01 = 2.0f * (T_(0, 0) * V_(0, 1) + T_(1, 0) * V_(1, 1) + T_(2, 0) * V_(2, 1)) * dL_da +
                  (T_(0, 1) * V_(0, 1) + T_(1, 1) * V_(1, 1) + T_(2, 1) * V_(2, 1)) * dL_db;
  float dL_dT02 = 2.0f * (T_(0, 0) * V_(0, 2) + T_(1, 0) * V_(1, 2) + T_(2, 0) * V_(2, 2)) * dL_da +
                  (T_(0, 1) * V_(0, 2) + T_(1, 1) * V_(1, 2) + T_(2, 1) * V_(2, 2)) * dL_db;
  float dL_dT10 = 2.0f * (T_(0, 1) * V_(0, 0) + T_(1, 1) * V_(1, 0) + T_(2, 1) * V_(2, 0)) * dL_dc +
                  (T_(0, 0) * V_(0, 0) + T_(1, 0) * V_(1, 0) + T_(2, 0) * V_(2, 0)) * dL_db;
  float dL_dT11 = 2.0f * (T_(0, 1) * V_(0, 1) + T_(1, 1) * V_(1, 1) + T_(2, 1) * V_(2, 1)) * dL_dc +
                  (T_(0, 0) * V_(0, 1) + T_(1, 0) * V_(1, 1) + T_(2, 0) * V_(2, 1)) * dL_db;
  float dL_dT12 = 2.0f * (T_(0, 1) * V_(0, 2) + T_(1, 1) * V_(1, 2) + T_(2, 1) * V_(2, 2)) * dL_dc +
                  (T_(0, 0) * V_(0, 2) + T_(1, 0) * V_(1, 2) + T_(2, 0) * V_(2, 2)) * dL_db;
#undef T_
#undef V_
#define W_(i, j) Wm[(i) * 3 + (j)]
  float dL_dJ00 = W_(0, 0) * dL_dT00 + W_(1, 0) * dL_dT01 + W_(2, 0) * dL_dT02;
  float dL_dJ02 = W_(0, 2) * dL_dT00 + W_(1, 2) * dL_dT01 + W_(2, 2) * dL_dT02;
  float dL_dJ11 = W_(0, 1) * dL_dT10 + W_(1, 1) * dL_dT11 + W_(2, 1) * dL_dT12;
  float dL_dJ12 = W_(0, 2) * dL_dT10 + W_(1, 2) * dL_dT11 + W_(2, 2) * dL_dT12;
#undef W_
  float dL_dtx = -h_x * inv_tz2 * dL_dJ02;
  float dL_dty = -h_y * inv_tz2 * dL_dJ12;
  float dL_dtz = -h_x * inv_tz2 * dL_dJ00 - h_y * inv_tz2 * dL_dJ11 + 2.0f * h_x * tx * inv_tz3 * dL_dJ02 +
                 2.0f * h_y * ty * inv_tz3 * dL_dJ12;
  float dL_dt[4] = {dL_dtx * x_grad_mul, dL_dty * y_grad_mul, dL_dtz, 1.0f}; /* stray w = 1.0, backward.py:434 */
  /* vec4 * transpose(view_matrix): r[j] = sum_i v[i] * V[j][i] */
  for (int j = 0; j < 3; ++j) {
    float s = V[j * 4 + 0] * dL_dt[0];
    s = s + V[j * 4 + 1] * dL_dt[1];
    s = s + V[j * 4 + 2] * dL_dt[2];
    s = s + V[j * 4 + 3] * dL_dt[3];
    dL_dmeans[3 * idx + j] += s;
  }
}

/* backward.py:708-768 compute_projection_backward_kernel, one Gaussian */
static void projection_backward_one(int idx, const float* means, const int* radii, const float* P,
                                    const float* dL_dmean2D, float* dL_dmeans) {
  if (radii[idx] <= 0) return;
  const float* m = means + 3 * idx;
  const float* g2 = dL_dmean2D + 3 * idx;
  float mh[4] = {m[0], m[1], m[2], 1.0f};
  float m_hom[4];
  vec4_mul_mat44(mh, P, m_hom);
  float m_w = 1.0f / (m_hom[3] + 0.0000001f);
#define P_(i, j) P[(i) * 4 + (j)]
  float mul1 = (P_(0, 0) * m[0] + P_(1, 0) * m[1] + P_(2, 0) * m[2] + P_(3, 0)) * m_w * m_w;
  float mul2 = (P_(0, 1) * m[0] + P_(1, 1) * m[1] + P_(2, 1) * m[2] + P_(3, 1)) * m_w * m_w;
  float d0 = (P_(0, 0) * m_w - P_(0, 3) * mul1) * g2[0] + (P_(0, 1) * m_w - P_(0, 3) * mul2) * g2[1];
  float d1 = (P_(1, 0) * m_w - P_(1, 3) * mul1) * g2[0] + (P_(1, 1) * m_w - P_(1, 3) * mul2) * g2[1];
  float d2 = (P_(2, 0) * m_w - P_(2, 3) * mul1) * g2[0] + (P_(2, 1) * m_w - P_(2, 3) * mul2) * g2[1];
#undef P_
  dL_dmeans[3 * idx + 0] += d0;
  dL_dmeans[3 * idx + 1] += d1;
  dL_dmeans[3 * idx + 2] += d2;
}

/* backward.py:42-64 dnormvdv */
static inline void dnormvdv(const float v[3], const float dv[3], float out[3]) {
  float sum2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  if (sum2 < 1e-10f) {
    out[0] = out[1] = out[2] = 0.0f;
    return;
  }
  float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
  out[0] = ((sum2 - v[0] * v[0]) * dv[0] - v[1] * v[0] * dv[1] - v[2] * v[0] * dv[2]) * invsum32;
  out[1] = (-v[0] * v[1] * dv[0] + (sum2 - v[1] * v[1]) * dv[1] - v[2] * v[1] * dv[2]) * invsum32;
  out[2] = (-v[0] * v[2] * dv[0] - v[1] * v[2] * dv[1] + (sum2 - v[2] * v[2]) * dv[2]) * invsum32;
}

/* backward.py:68-255 sh_backward_kernel, one Gaussian.  dL_dshs has stride 16 always. */
static void sh_backward_one(int idx, int degree, const float* means, const float* shs, const int* radii,
                            const float* campos, const float* clamped_state, const float* dL_dcolor, float* dL_dmeans,
                            float* dL_dshs) {
  if (radii[idx] <= 0) return;
  const float* mean = means + 3 * idx;
  float dir_orig[3] = {mean[0] - campos[0], mean[1] - campos[1], mean[2] - campos[2]};
  float dir_len = length3(dir_orig);
  if (dir_len < 1e-8f) return;
  float x = dir_orig[0] / dir_len, y = dir_orig[1] / dir_len, z = dir_orig[2] / dir_len;
  float dL_dRGB[3];
  for (int c = 0; c < 3; ++c) /* mask is applied regardless of the forward `clamped` flag (Note G) */
    dL_dRGB[c] = dL_dcolor[3 * idx + c] * (1.0f + (-1.0f * clamped_state[3 * idx + c]));
  const float* sh = shs + (size_t)idx * 48;
  float* dsh = dL_dshs + (size_t)idx * 48;
  float dRGBdx[3] = {0.0f, 0.0f, 0.0f}, dRGBdy[3] = {0.0f, 0.0f, 0.0f}, dRGBdz[3] = {0.0f, 0.0f, 0.0f};
#define SHV(k, c) sh[(k) * 3 + (c)]
#define DSH(k, val)                                         \
  do {                                                      \
    float _b = (val);                                       \
    for (int c = 0; c < 3; ++c) dsh[(k) * 3 + c] = _b * dL_dRGB[c]; \
  } while (0)
  DSH(0, SH_C0);
  if (degree > 0) {
    DSH(1, -SH_C1 * y);
    DSH(2, SH_C1 * z);
    DSH(3, -SH_C1 * x);
    for (int c = 0; c < 3; ++c) {
      dRGBdx[c] = -SH_C1 * SHV(3, c);
      dRGBdy[c] = -SH_C1 * SHV(1, c);
      dRGBdz[c] = SH_C1 * SHV(2, c);
    }
    if (degree > 1) {
      float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      DSH(4, C2_0 * xy);
      DSH(5, C2_1 * yz);
      DSH(6, C2_2 * (2.0f * zz - xx - yy));
      DSH(7, C2_3 * xz);
      DSH(8, C2_4 * (xx - yy));
      for (int c = 0; c < 3; ++c) {
        float sh4 = SHV(4, c), sh5 = SHV(5, c), sh6 = SHV(6, c), sh7 = SHV(7, c), sh8 = SHV(8, c);
        dRGBdx[c] += C2_0 * y * sh4 + C2_2 * 2.0f * -x * sh6 + C2_3 * z * sh7 + C2_4 * 2.0f * x * sh8;
        dRGBdy[c] += C2_0 * x * sh4 + C2_1 * z * sh5 + C2_2 * 2.0f * -y * sh6 + C2_4 * 2.0f * -y * sh8;
        dRGBdz[c] += C2_1 * y * sh5 + C2_2 * 2.0f * 2.0f * z * sh6 + C2_3 * x * sh7;
      }
      if (degree > 2) {
        DSH(9, C3_0 * y * (3.0f * xx - yy));
        DSH(10, C3_1 * xy * z);
        DSH(11, C3_2 * y * (4.0f * zz - xx - yy));
        DSH(12, C3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy));
        DSH(13, C3_4 * x * (4.0f * zz - xx - yy));
        DSH(14, C3_5 * z * (xx - yy));
        DSH(15, C3_6 * x * (xx - 3.0f * yy));
        for (int c = 0; c < 3; ++c) {
          float sh9 = SHV(9, c), sh10 = SHV(10, c), sh11 = SHV(11, c), sh12 = SHV(12, c), sh13 = SHV(13, c),
                sh14 = SHV(14, c), sh15 = SHV(15, c);
          dRGBdx[c] += (C3_0 * sh9 * 3.0f * 2.0f * xy + C3_1 * sh10 * yz + C3_2 * sh11 * -2.0f * xy +
                        C3_3 * sh12 * -3.0f * 2.0f * xz + C3_4 * sh13 * (-3.0f * xx + 4.0f * zz - yy) +
                        C3_5 * sh14 * 2.0f * xz + C3_6 * sh15 * 3.0f * (xx - yy));
          dRGBdy[c] += (C3_0 * sh9 * 3.0f * (xx - yy) + C3_1 * sh10 * xz + C3_2 * sh11 * (-3.0f * yy + 4.0f * zz - xx) +
                        C3_3 * sh12 * -3.0f * 2.0f * yz + C3_4 * sh13 * -2.0f * xy + C3_5 * sh14 * -2.0f * yz +
                        C3_6 * sh15 * -3.0f * 2.0f * xy);
          dRGBdz[c] += (C3_1 * sh10 * xy + C3_2 * sh11 * 4.0f * 2.0f * yz + C3_3 * sh12 * 3.0f * (2.0f * zz - xx - yy) +
                        C3_4 * sh13 * 4.0f * 2.0f * xz + C3_5 * sh14 * (xx - yy));
        }
      }
    }
  }
#undef SHV
#undef DSH
  float dL_ddir[3] = {dot3(dRGBdx, dL_dRGB), dot3(dRGBdy, dL_dRGB), dot3(dRGBdz, dL_dRGB)};
  float dm[3];
  dnormvdv(dir_orig, dL_ddir, dm);
  for (int c = 0; c < 3; ++c) dL_dmeans[3 * idx + c] += dm[c];
}

/* backward.py:438-556 compute_cov3d_backward_kernel, one Gaussian.  Quirk G2 (a column-major
 * formula applied to row-major matrices) is reproduced literally. */
static void cov3d_backward_one(int idx, const float* scales, const float* rots, const int* radii, float scale_modifier,
                               const float* dL_dcov3Ds, float* dL_dscales, float* dL_drots) {
  if (radii[idx] <= 0) {
    for (int k = 0; k < 3; ++k) dL_dscales[3 * idx + k] = 0.0f;
    for (int k = 0; k < 4; ++k) dL_drots[4 * idx + k] = 0.0f;
    return;
  }
  const float* sv = scales + 3 * idx;
  const float* q = rots + 4 * idx;
  float r = q[3], x = q[0], y = q[1], z = q[2];
  float R[9] = {1.0f - 2.0f * (y * y + z * z), 2.0f * (x * y - r * z),        2.0f * (x * z + r * y),
                2.0f * (x * y + r * z),        1.0f - 2.0f * (x * x + z * z), 2.0f * (y * z - r * x),
                2.0f * (x * z - r * y),        2.0f * (y * z + r * x),        1.0f - 2.0f * (x * x + y * y)};
  float s_vec[3] = {scale_modifier * sv[0], scale_modifier * sv[1], scale_modifier * sv[2]};
  float S[9] = {s_vec[0], 0.0f, 0.0f, 0.0f, s_vec[1], 0.0f, 0.0f, 0.0f, s_vec[2]};
  float M[9];
  mat33_mul(S, R, M);
  const float* g = dL_dcov3Ds + 6 * idx;
  float dL_dSigma[9] = {g[0], 0.5f * g[1], 0.5f * g[2], 0.5f * g[1], g[3], 0.5f * g[4], 0.5f * g[2], 0.5f * g[4], g[5]};
  float M2[9], dL_dM[9], Rt[9], dL_dMt[9];
  for (int k = 0; k < 9; ++k) M2[k] = 2.0f * M[k]; /* 2.0 * M * dL_dSigma, left to right */
  mat33_mul(M2, dL_dSigma, dL_dM);
  mat33_transpose(R, Rt);
  mat33_transpose(dL_dM, dL_dMt);
  float dL_dscale[3] = {dot3(Rt + 0, dL_dMt + 0), dot3(Rt + 3, dL_dMt + 3), dot3(Rt + 6, dL_dMt + 6)};
  for (int k = 0; k < 3; ++k) dL_dscales[3 * idx + k] = dL_dscale[k] * scale_modifier;
  float D[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) D[i * 3 + j] = dL_dMt[i * 3 + j] * s_vec[i];
#define D_(i, j) D[(i) * 3 + (j)]
  float dL_dr = 2.0f * (z * (D_(0, 1) - D_(1, 0)) + y * (D_(2, 0) - D_(0, 2)) + x * (D_(1, 2) - D_(2, 1)));
  float dL_dx = 2.0f * (y * (D_(1, 0) + D_(0, 1)) + z * (D_(2, 0) + D_(0, 2)) + r * (D_(1, 2) - D_(2, 1))) -
                4.0f * x * (D_(2, 2) + D_(1, 1));
  float dL_dy = 2.0f * (x * (D_(1, 0) + D_(0, 1)) + r * (D_(2, 0) - D_(0, 2)) + z * (D_(1, 2) + D_(2, 1))) -
                4.0f * y * (D_(2, 2) + D_(0, 0));
  float dL_dz = 2.0f * (r * (D_(0, 1) - D_(1, 0)) + x * (D_(2, 0) + D_(0, 2)) + y * (D_(1, 2) + D_(2, 1))) -
                4.0f * z * (D_(1, 1) + D_(0, 0));
#undef D_
  dL_drots[4 * idx + 0] = dL_dx;
  dL_drots[4 * idx + 1] = dL_dy;
  dL_drots[4 * idx + 2] = dL_dz;
  dL_drots[4 * idx + 3] = dL_dr;
}

/* backward.py:770-888 backward_preprocess: cov2d-bwd -> projection-bwd -> SH-bwd -> cov3d-bwd.
 * Quirk G3: cov3d backward always runs with scale_modifier = 1.0 (backward.py:805,1155-1182).
 * dL_dcov3D_scratch is the internal buffer of backward.py:812 (the dict's dL_dcov3D stays 0). */
typedef struct {
  const float *means, *shs, *scales, *rots, *view, *proj, *cov3Ds, *campos, *clamped_state, *dL_dmean2D, *dL_dconic,
      *dL_dcolors;
  const int* radii;
  float tan_fovx, tan_fovy, focal_x, focal_y;
  int degree;
  float *dL_dmeans, *dL_dsh, *dL_dscales, *dL_drots, *dL_dcov3D;
} bpp_ctx;
static void bpp_body(int b, int e, void* p, long* acc) {
  bpp_ctx* c = (bpp_ctx*)p;
  (void)acc;
  for (int i = b; i < e; ++i) {
    cov2d_backward_one(i, c->means, c->cov3Ds, c->radii, c->focal_x, c->focal_y, c->tan_fovx, c->tan_fovy, c->view,
                       c->dL_dconic, c->dL_dmeans, c->dL_dcov3D);
    projection_backward_one(i, c->means, c->radii, c->proj, c->dL_dmean2D, c->dL_dmeans);
    sh_backward_one(i, c->degree, c->means, c->shs, c->radii, c->campos, c->clamped_state, c->dL_dcolors, c->dL_dmeans,
                    c->dL_dsh);
    cov3d_backward_one(i, c->scales, c->rots, c->radii, 1.0f, c->dL_dcov3D, c->dL_dscales, c->dL_drots);
  }
}
GSO_API void gso_backward_preprocess(int N, const float* means, const int* radii, const float* shs, const float* scales,
                                     const float* rots, const float* view, const float* proj, float tan_fovx,
                                     float tan_fovy, float focal_x, float focal_y, const float* cov3Ds,
                                     const float* campos, const float* clamped_state, const float* dL_dmean2D,
                                     const float* dL_dconic, const float* dL_dcolors, int degree, float* dL_dmeans,
                                     float* dL_dsh, float* dL_dscales, float* dL_drots, float* dL_dcov3D_scratch) {
  bpp_ctx c = {means, shs, scales, rots, view, proj, cov3Ds, campos, clamped_state, dL_dmean2D, dL_dconic, dL_dcolors,
               radii, tan_fovx, tan_fovy, focal_x, focal_y, degree, dL_dmeans, dL_dsh, dL_dscales, dL_drots,
               dL_dcov3D_scratch};
  parallel_for(N, 1024, bpp_body, &c);
}

/* ----------------------------------------------------------------------------------------- */
/* optimizer.py                                                                               */
/* ----------------------------------------------------------------------------------------- */

/* vec3-style update: p -= lr * ( (m/bc1) / ((sqrt(v/bc2) + eps) + 1e-9) ), optimizer.py:51-59
 * with utils/wp_utils.py:15-20 (the extra 1e-9) */
static inline float adam_vec3_elem(float* m, float* v, float g, float beta1, float beta2, float bc1, float bc2,
                                   float eps, float lr) {
  *m = beta1 * (*m) + (1.0f - beta1) * g;
  *v = beta2 * (*v) + (1.0f - beta2) * (g * g);
  float mc = (*m) / bc1;
  float vc = (*v) / bc2;
  float denom = sqrtf(vc) + eps;
  float safe = denom + 1e-9f;
  return lr * (mc / safe);
}
/* scalar-style update used for rotations and opacity: (lr * m^) / (sqrt(v^) + eps) */
static inline float adam_scalar_elem(float* m, float* v, float g, float beta1, float beta2, float bc1, float bc2,
                                     float eps, float lr) {
  *m = beta1 * (*m) + (1.0f - beta1) * g;
  *v = beta2 * (*v) + (1.0f - beta2) * (g * g);
  float mc = (*m) / bc1;
  float vc = (*v) / bc2;
  return lr * mc / (sqrtf(vc) + eps);
}

typedef struct {
  const float *pos_g, *scale_g, *rot_g, *opac_g, *sh_g;
  float lr_pos, lr_scale, lr_rot, lr_opac, lr_sh, beta1, beta2, epsilon, bc1, bc2;
  float *pos, *scales, *rots, *opac, *shs, *m_pos, *m_scale, *m_rot, *m_opac, *m_sh, *v_pos, *v_scale, *v_rot, *v_opac,
      *v_sh;
} adam_ctx;
static void adam_body(int b, int e_, void* p, long* acc) {
  adam_ctx* c = (adam_ctx*)p;
  (void)acc;
  float beta1 = c->beta1, beta2 = c->beta2, bc1 = c->bc1, bc2 = c->bc2, epsilon = c->epsilon;
  for (int i = b; i < e_; ++i) {
    for (int k = 0; k < 3; ++k) {
      int e = 3 * i + k;
      c->pos[e] = c->pos[e] - adam_vec3_elem(&c->m_pos[e], &c->v_pos[e], c->pos_g[e], beta1, beta2, bc1, bc2, epsilon,
                                             c->lr_pos);
    }
    for (int k = 0; k < 3; ++k) {
      int e = 3 * i + k;
      float upd = adam_vec3_elem(&c->m_scale[e], &c->v_scale[e], c->scale_g[e], beta1, beta2, bc1, bc2, epsilon,
                                 c->lr_scale);
      c->scales[e] = f_max(c->scales[e] - upd, 0.001f);
    }
    for (int k = 0; k < 4; ++k) {
      int e = 4 * i + k;
      c->rots[e] = c->rots[e] - adam_scalar_elem(&c->m_rot[e], &c->v_rot[e], c->rot_g[e], beta1, beta2, bc1, bc2,
                                                 epsilon, c->lr_rot);
    }
    {
      float* q = c->rots + 4 * i;
      float quat_length = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
      if (quat_length > 0.0f)
        for (int k = 0; k < 4; ++k) q[k] = q[k] / quat_length;
    }
    {
      float upd = adam_scalar_elem(&c->m_opac[i], &c->v_opac[i], c->opac_g[i], beta1, beta2, bc1, bc2, epsilon,
                                   c->lr_opac);
      c->opac[i] = f_max(f_min(c->opac[i] - upd, 1.0f), 0.0f);
    }
    for (int j = 0; j < 48; ++j) {
      size_t e = (size_t)i * 48 + j;
      c->shs[e] = c->shs[e] - adam_vec3_elem(&c->m_sh[e], &c->v_sh[e], c->sh_g[e], beta1, beta2, bc1, bc2, epsilon,
                                             c->lr_sh);
    }
  }
}
/* optimizer.py:6-139 adam_update */
GSO_API void gso_adam_update(int N, const float* pos_g, const float* scale_g, const float* rot_g, const float* opac_g,
                             const float* sh_g, float lr_pos, float lr_scale, float lr_rot, float lr_opac, float lr_sh,
                             float beta1, float beta2, float epsilon, int iteration, float* pos, float* scales,
                             float* rots, float* opac, float* shs, float* m_pos, float* m_scale, float* m_rot,
                             float* m_opac, float* m_sh, float* v_pos, float* v_scale, float* v_rot, float* v_opac,
                             float* v_sh) {
  float bc1 = 1.0f - powf(beta1, (float)(iteration + 1)); /* optimizer.py:47-48 */
  float bc2 = 1.0f - powf(beta2, (float)(iteration + 1));
  adam_ctx c = {pos_g, scale_g, rot_g, opac_g, sh_g, lr_pos, lr_scale, lr_rot, lr_opac, lr_sh, beta1, beta2, epsilon,
                bc1, bc2, pos, scales, rots, opac, shs, m_pos, m_scale, m_rot, m_opac, m_sh, v_pos, v_scale, v_rot,
                v_opac, v_sh};
  parallel_for(N, 1024, adam_body, &c);
}

/* [Warp] wp.randf(uint32 state): PCG hash then 24-bit mantissa.  Unpinned (cannot run Warp). */
static inline uint32_t rand_pcg(uint32_t state) {
  uint32_t b = state * 747796405u + 2891336453u;
  uint32_t c = ((b >> ((b >> 28u) + 4u)) ^ b) * 277803737u;
  return (c >> 22u) ^ c;
}
GSO_API float gso_randf(uint32_t state) { return (float)(rand_pcg(state) >> 8) * (1.0f / 16777216.0f); }

/* train.py:36-92 init_gaussian_params */
GSO_API void gso_init_gaussian_params(int N, float init_scale, float* pos, float* scales, float* rots, float* opac,
                                      float* shs) {
  for (int i = 0; i < N; ++i) {
    for (int k = 0; k < 3; ++k) pos[3 * i + k] = gso_randf((uint32_t)(i * 3 + k)) * 2.6f - 1.3f;
    for (int k = 0; k < 3; ++k) scales[3 * i + k] = init_scale;
    rots[4 * i + 0] = 1.0f;
    rots[4 * i + 1] = 0.0f;
    rots[4 * i + 2] = 0.0f;
    rots[4 * i + 3] = 0.0f;
    opac[i] = 0.1f;
    for (int j = 0; j < 48; ++j) shs[(size_t)i * 48 + j] = (j < 3) ? -0.007f : 0.0f;
  }
}

/* train.py:398-405 compute_grad_norms */
GSO_API void gso_grad_norms(int N, const float* pos_grad, float* norms) {
  for (int i = 0; i < N; ++i) norms[i] = length3(pos_grad + 3 * i);
}

/* optimizer.py:212-242 (clone: <=) and 180-210 (split: >).  n_grads = length of the grads
 * array: train.py:479-492 reads a stale shorter array after cloning (quirk G4); entries past
 * n_grads are treated as 0.0 (never marked) -- the reference reads out of bounds there. */
GSO_API void gso_mark_candidates(int N, int n_grads, const float* grads, const float* scales, float grad_threshold,
                                 float scene_extent, float percent_dense, int want_split, int* mask) {
  for (int i = 0; i < N; ++i) {
    float gr = (i < n_grads) ? grads[i] : 0.0f;
    int high_grad = gr >= grad_threshold;
    float max_scale = f_max(f_max(scales[3 * i], scales[3 * i + 1]), scales[3 * i + 2]);
    float scale_threshold = percent_dense * scene_extent;
    int sel = want_split ? (max_scale > scale_threshold) : (max_scale <= scale_threshold);
    mask[i] = (high_grad && sel) ? 1 : 0;
  }
}

/* [Warp] wp.utils.array_scan(inclusive=False) */
GSO_API void gso_exclusive_scan(int N, const int* in, int* out) {
  int s = 0;
  for (int i = 0; i < N; ++i) {
    out[i] = s;
    s += in[i];
  }
}

static inline void copy_gaussian(int src, int dst, const float* pos, const float* scales, const float* rots,
                                 const float* opac, const float* shs, float* o_pos, float* o_scales, float* o_rots,
                                 float* o_opac, float* o_shs) {
  for (int k = 0; k < 3; ++k) o_pos[3 * dst + k] = pos[3 * src + k];
  for (int k = 0; k < 3; ++k) o_scales[3 * dst + k] = scales[3 * src + k];
  for (int k = 0; k < 4; ++k) o_rots[4 * dst + k] = rots[4 * src + k];
  o_opac[dst] = opac[src];
  memcpy(o_shs + (size_t)dst * 48, shs + (size_t)src * 48, 48 * sizeof(float));
}

/* optimizer.py:312-362 clone_gaussians.  out arrays hold new_N entries; writes at index
 * >= new_N (possible through quirk G5: the exclusive-scan total drops the last flag) are
 * skipped instead of written out of bounds. */
GSO_API void gso_clone_gaussians(int N, int new_N, const int* clone_mask, const int* prefix_sum, const float* pos,
                                 const float* scales, const float* rots, const float* opac, const float* shs,
                                 float noise_scale, float* o_pos, float* o_scales, float* o_rots, float* o_opac,
                                 float* o_shs) {
  int offset = N;
  for (int i = 0; i < N; ++i) {
    copy_gaussian(i, i, pos, scales, rots, opac, shs, o_pos, o_scales, o_rots, o_opac, o_shs);
    if (clone_mask[i] == 1) {
      int base_idx = prefix_sum[i] + offset;
      if (base_idx >= new_N) continue;
      copy_gaussian(i, base_idx, pos, scales, rots, opac, shs, o_pos, o_scales, o_rots, o_opac, o_shs);
      for (int k = 0; k < 3; ++k)
        o_pos[3 * base_idx + k] = pos[3 * i + k] + gso_randf((uint32_t)(i * 3 + k)) * noise_scale;
    }
  }
}

/* optimizer.py:244-309 split_gaussians */
GSO_API void gso_split_gaussians(int N, int new_N, const int* split_mask, const int* prefix_sum, const float* pos,
                                 const float* scales, const float* rots, const float* opac, const float* shs,
                                 int N_split, float scale_factor, float* o_pos, float* o_scales, float* o_rots,
                                 float* o_opac, float* o_shs) {
  int offset = N;
  for (int i = 0; i < N; ++i) {
    copy_gaussian(i, i, pos, scales, rots, opac, shs, o_pos, o_scales, o_rots, o_opac, o_shs);
    if (split_mask[i] == 1) {
      int split_idx = prefix_sum[i];
      for (int j = 0; j < N_split; ++j) {
        int new_idx = offset + split_idx * N_split + j;
        if (new_idx < new_N) {
          copy_gaussian(i, new_idx, pos, scales, rots, opac, shs, o_pos, o_scales, o_rots, o_opac, o_shs);
          for (int k = 0; k < 3; ++k) o_scales[3 * new_idx + k] = scales[3 * i + k] * scale_factor;
          for (int k = 0; k < 3; ++k)
            o_pos[3 * new_idx + k] =
                pos[3 * i + k] + ((gso_randf((uint32_t)(new_idx * 3 + k))) * 2.0f - 1.0f) * 0.01f;
        }
      }
    }
  }
}

/* train.py:547-573: valid = 1 - (i < offset && split_mask[i] == 1).  split_mask has `offset`
 * entries. */
GSO_API void gso_split_valid_mask(int num_points, int offset, const int* split_mask, int* valid) {
  for (int i = 0; i < num_points; ++i) {
    int prune = (i < offset && split_mask[i] == 1) ? 1 : 0;
    valid[i] = 1 - prune;
  }
}

/* optimizer.py:364-382 prune_gaussians */
GSO_API void gso_prune_mask(int N, const float* opac, float threshold, int* valid) {
  for (int i = 0; i < N; ++i) valid[i] = (opac[i] > threshold) ? 1 : 0;
}

/* optimizer.py:384-415 compact_gaussians; out arrays hold out_N entries (writes past the end,
 * quirk G5, are skipped) */
GSO_API void gso_compact_gaussians(int N, int out_N, const int* valid_mask, const int* prefix_sum, const float* pos,
                                   const float* scales, const float* rots, const float* opac, const float* shs,
                                   float* o_pos, float* o_scales, float* o_rots, float* o_opac, float* o_shs) {
  for (int i = 0; i < N; ++i) {
    if (valid_mask[i] == 0) continue;
    int new_i = prefix_sum[i];
    if (new_i >= out_N) continue;
    copy_gaussian(i, new_i, pos, scales, rots, opac, shs, o_pos, o_scales, o_rots, o_opac, o_shs);
  }
}

/* ----------------------------------------------------------------------------------------- */
/* loss.py ("next" row 8f-1)                                                                  */
/* ----------------------------------------------------------------------------------------- */

/* loss.py:11-30 + 148-176: sum |r-t| over pixels (serial fp32 accumulation in thread order
 * i (x) outer, j (y) inner [Warp dim=(width,height)]), then float(sum)/(W*H*3) in double */
GSO_API double gso_l1_loss(int W, int H, const float* rendered, const float* target) {
  float acc = 0.0f;
  for (int i = 0; i < W; ++i)
    for (int j = 0; j < H; ++j) {
      size_t p = ((size_t)j * W + i) * 3;
      float d0 = fabsf(rendered[p + 0] - target[p + 0]);
      float d1 = fabsf(rendered[p + 1] - target[p + 1]);
      float d2 = fabsf(rendered[p + 2] - target[p + 2]);
      float l1 = d0 + d1 + d2;
      acc += l1;
    }
  return (double)acc / ((double)W * (double)H * 3.0);
}

/* loss.py:121-146 backprop_l1_pixel_gradients.  l1_weight = (1 - lambda_dssim)/(H*W*3.0) is
 * computed by the host in double (loss.py:236) and arrives as a float kernel argument.
 * [Warp] wp.sign(x) = (x < 0) ? -1 : +1, so sign(0) = +1. */
GSO_API void gso_l1_grad(int W, int H, const float* rendered, const float* target, float l1_weight,
                         float* pixel_grad) {
  size_t n = (size_t)W * H * 3;
  for (size_t e = 0; e < n; ++e) {
    float d = rendered[e] - target[e];
    float s = (d < 0.0f) ? -1.0f : 1.0f;
    pixel_grad[e] = l1_weight * s;
  }
}

/* loss.py:33-45 gaussian_kernel + loss.py:47-119 ssim_kernel + loss.py:178-215 ssim(): per pixel an
 * 11x11 window truncated at the image border, weights exp(-x^2 / (2 sigma^2)) (sigma = 1.5) per axis,
 * normalised by the weights actually summed; the per-pixel value is the mean of the three channels'
 * SSIM; serial fp32 accumulation in thread order i (x) outer, j (y) inner, then float(sum)/(W*H). */
GSO_API double gso_ssim(int W, int H, const float* rendered, const float* target) {
  const int window_size = 11, half_window = window_size / 2;
  const float sigma = 1.5f;
  float gw[11];
  for (int i = 0; i < window_size; ++i) {
    int x = i - window_size / 2;
    gw[i] = gs_exp(-1.0f * (float)(x * x) / (2.0f * sigma * sigma));
  }
  const float c1 = 0.01f * 0.01f, c2 = 0.03f * 0.03f;
  float acc = 0.0f;
  for (int i = 0; i < W; ++i)
    for (int j = 0; j < H; ++j) {
      float mu1[3] = {0, 0, 0}, mu2[3] = {0, 0, 0}, s1[3] = {0, 0, 0}, s2[3] = {0, 0, 0}, s12[3] = {0, 0, 0};
      float weight_sum = 0.0f;
      int y0 = j - half_window < 0 ? 0 : j - half_window, y1 = j + half_window + 1 > H ? H : j + half_window + 1;
      int x0 = i - half_window < 0 ? 0 : i - half_window, x1 = i + half_window + 1 > W ? W : i + half_window + 1;
      for (int y = y0; y < y1; ++y)
        for (int x = x0; x < x1; ++x) {
          int wy = y - j < 0 ? j - y : y - j, wx = x - i < 0 ? i - x : x - i;
          /* loss.py:84: weights are indexed by |offset| (0..5), i.e. the LEFT half of the kernel array:
           * gaussian_weights[k] = exp(-(k-5)^2 / 4.5), so the centre tap gets exp(-25/4.5) and the
           * farthest tap 1.0 -- reproduced as written */
          float w = gw[wx] * gw[wy];
          size_t p = ((size_t)y * W + x) * 3;
          for (int c = 0; c < 3; ++c) {
            float p1 = rendered[p + c], p2 = target[p + c];
            mu1[c] = mu1[c] + p1 * w;
            mu2[c] = mu2[c] + p2 * w;
            s1[c] = s1[c] + (p1 * p1) * w;
            s2[c] = s2[c] + (p2 * p2) * w;
            s12[c] = s12[c] + (p1 * p2) * w;
          }
          weight_sum = weight_sum + w;
        }
      float ssim_c[3];
      for (int c = 0; c < 3; ++c) {
        float m1 = mu1[c], m2 = mu2[c], v1 = s1[c], v2 = s2[c], v12 = s12[c];
        if (weight_sum > 0.0f) {
          m1 = m1 / weight_sum;
          m2 = m2 / weight_sum;
          v1 = v1 / weight_sum;
          v2 = v2 / weight_sum;
          v12 = v12 / weight_sum;
        }
        v1 = v1 - m1 * m1;
        v2 = v2 - m2 * m2;
        v12 = v12 - m1 * m2;
        ssim_c[c] = ((2.0f * m1 * m2 + c1) * (2.0f * v12 + c2)) / ((m1 * m1 + m2 * m2 + c1) * (v1 + v2 + c2));
      }
      float ssim_val = (ssim_c[0] + ssim_c[1] + ssim_c[2]) / 3.0f;
      acc += ssim_val;
    }
  return (double)acc / ((double)W * (double)H);
}

/* loss.py:248-268 depth_loss_kernel + 270-306 depth_loss(): sum |rendered - target| * mask, / (W*H) */
GSO_API double gso_depth_loss(int W, int H, const float* rendered_depth, const float* target_depth,
                              const float* depth_mask) {
  float acc = 0.0f;
  for (int i = 0; i < W; ++i)
    for (int j = 0; j < H; ++j) {
      size_t p = (size_t)j * W + i;
      acc += fabsf(rendered_depth[p] - target_depth[p]) * depth_mask[p];
    }
  return (double)acc / ((double)W * (double)H);
}

/* forward.py:589-627 track_pixel_stats.  Provably a no-op after wp_render_gaussians (SURVEY
 * 8a F12); restated so that a test can assert exactly that.  Returns #elements modified. */
GSO_API int gso_track_pixel_stats(int W, int H, const float* image, const float* bg, float* final_Ts, int* n_contrib) {
  int changed = 0;
  for (int x = 0; x < W; ++x)
    for (int y = 0; y < H; ++y) {
      size_t p = (size_t)y * W + x;
      float diff_r = fabsf(image[3 * p + 0] - bg[0]);
      float diff_g = fabsf(image[3 * p + 1] - bg[1]);
      float diff_b = fabsf(image[3 * p + 2] - bg[2]);
      int has_content = (diff_r > 0.01f) || (diff_g > 0.01f) || (diff_b > 0.01f);
      if (has_content) {
        if (final_Ts[p] == 0.0f) {
          float max_diff = f_max(diff_r, f_max(diff_g, diff_b));
          final_Ts[p] = 1.0f - f_min(0.99f, max_diff);
          changed++;
        }
        if (n_contrib[p] == 0) {
          n_contrib[p] = 1;
          changed++;
        }
      }
    }
  return changed;
}
