/*
 * gsr_oracle.c - CPU restatement of the reference render path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker.  The product (libgsr_b200.so) never links,
 * imports or falls back to it.
 *
 * What it restates (citations are into /root/reference/):
 *   - Sigma = R S S^T R^T from log-scales and quaternions          rasterize.py:41-56, 89-120
 *   - camera-space / clip-space / pixel-space means, z<0.2 cull    rasterize.py:80-86, 371-382, 391
 *   - EWA projection of the covariance, +0.3 low-pass              rasterize.py:201-252, 384-388
 *   - eigen radius, tile rect, pixel bbox, conic, skip rule        rasterize.py:154-198, 395-420, 441
 *   - SH degree 0..3 colour                                        spherical_harmonics.py:27-73
 *   - global stable ascending depth order                          rasterize.py:424-425 (pinned stable)
 *   - one-Gaussian-at-a-time front-to-back blend                   rasterize.py:255-305, 436-446
 *
 * Arithmetic: every float operation is written as ONE rounded f32 operation in the order the
 * reference's torch-CPU execution performs it: sgemm-backed products are
 * fmaf chains in ascending k with a separately rounded "+ row"; 3x3 bmm products are plain
 * multiply/add in ascending k; scalar/tensor divisions are reciprocal-then-multiply.
 * Build with -ffp-contract=off (see oracle/build.py).  exp-derived values agree with the reference to
 * <= 1 ulp, not always bitwise (see exp_scale / exp_sleef_u10 below); everything else is bit-exact.
 *
 * PINNING: tests/test_oracle_cpu.py checks this file against golden vectors captured from the
 * unmodified reference (tests/golden/make_golden.py) - depth bit-exact, images to 1e-5.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define BLOCK 16
#define SPREAD 3.0f
#define MAX_DENSITY 0.99f
#define MIN_ALPHA ((float)(1.0 / 255.0))

/* Same field order as gsr_camera in include/gsr_b200.h, so one ctypes struct serves both. */
typedef struct {
  float w2c_t[16];      /* get_world_to_camera_matrix(...).T, row-major      rasterize.py:361 */
  float full_t[16];     /* w2c_t @ proj_t                                    rasterize.py:364 */
  float cam_center[3];  /* inverse(w2c_t)[3,:3]                              spherical_harmonics.py:35 */
  float focal_x, focal_y; /* f32(fx/2), f32(fy/2)                            rasterize.py:216 */
  float limx, limy;     /* f32(1.3*tan_fov)                                  rasterize.py:210-211 */
  int32_t width, height;
  int32_t sh_degree;
  int32_t reserved;
} gsr_camera;

static const float SH0 = 0.28209479177387814f;
static const float SH1 = 0.4886025119029199f;
static const float SH2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                             -1.0925484305920792f, 0.5462742152960396f};
static const float SH3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                             -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

/*
 * exp as the reference's torch-CPU build evaluates it:
 *  - torch.exp on a contiguous f32 tensor (rasterize.py:97) goes to MKL VML (high-accuracy mode, closed
 *    source).  It differs from the correctly rounded result in ~1.1 % of inputs (measured), so the best
 *    restatement is the correctly rounded value: double exp, rounded once to f32.
 *  - torch.sigmoid (rasterize.py:358) is 1/(1+e) with e from the vectorised Sleef expf (u10, FMA form).
 *    The sequence below reproduces it bit-for-bit on the vector body (2M-sample check: only the <16-element
 *    scalar tails of each thread chunk differ, by 1 ulp, because torch runs those through libm).
 */
static inline float exp_scale(float x) { return (float)exp((double)x); }

static inline float exp_sleef_u10(float d) {
  float qf = rintf(d * 1.442695040888963407359924681001892137426645954152985934135449406931f);
  int q = (int)qf;
  float s = fmaf(qf, -0.693145751953125f, d);
  s = fmaf(qf, -1.428606765330187045e-06f, s);
  float u = 0.000198527617612853646278381f;
  u = fmaf(u, s, 0.00139304355252534151077271f);
  u = fmaf(u, s, 0.00833336077630519866943359f);
  u = fmaf(u, s, 0.0416664853692054748535156f);
  u = fmaf(u, s, 0.166666671633720397949219f);
  u = fmaf(u, s, 0.5f);
  u = 1.0f + fmaf(s * s, u, s);
  union { float f; int32_t i; } a, b;
  a.i = ((q >> 1) + 0x7f) << 23;
  b.i = ((q - (q >> 1)) + 0x7f) << 23;
  u = (u * a.f) * b.f;
  if (d < -104.0f) u = 0.0f;
  if (d > 100.0f) u = INFINITY;
  return u;
}

static inline float sigmoid_ref(float x) { return 1.0f / (1.0f + exp_sleef_u10(-x)); }

static inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int64_t clampi(int64_t v, int64_t lo, int64_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* p @ M[:3,:] + M[3,:] for one output column c of a row-major 4x4 (sgemm: fma chain, then a separate add). */
static inline float row_affine(const float *m, int c, float x, float y, float z) {
  float acc = x * m[0 * 4 + c];
  acc = fmaf(y, m[1 * 4 + c], acc);
  acc = fmaf(z, m[2 * 4 + c], acc);
  return acc + m[3 * 4 + c];
}

/* spherical_harmonics.py:36-71 for one Gaussian.  sh(k,c) = sh[k*sk + c*sc]. */
static void sh_colour(const float *xyz, const float *sh, int64_t sk, int64_t sc, const float *cc, int degree,
                      float *rgb) {
  float dx = xyz[0] - cc[0], dy = xyz[1] - cc[1], dz = xyz[2] - cc[2];
  float nrm = sqrtf((dx * dx + dy * dy) + dz * dz);
  float x = dx / nrm, y = dy / nrm, z = dz / nrm;
  for (int c = 0; c < 3; ++c) {
#define C(k) sh[(k) * sk + c * sc]
    float col = C(0) * SH0;
    if (degree > 0) { /* each "col +=" adds one left-to-right sum, as the reference's statements do */
      float t = ((-SH1) * y) * C(1);
      t = t + (SH1 * z) * C(2);
      t = t - (SH1 * x) * C(3);
      col += t;
      if (degree > 1) {
        float xx = x * x, yy = y * y;
        t = ((SH2[0] * x) * y) * C(4);
        t = t + ((SH2[1] * y) * z) * C(5);
        t = t + (SH2[2] * (((2.0f * z) * z - xx) - yy)) * C(6);
        t = t + ((SH2[3] * x) * z) * C(7);
        t = t + (SH2[4] * (xx - yy)) * C(8);
        col += t;
        if (degree > 2) {
          t = ((SH3[0] * y) * ((3.0f * x) * x - yy)) * C(9);
          t = t + (((SH3[1] * x) * y) * z) * C(10);
          t = t + ((SH3[2] * y) * (((4.0f * z) * z - xx) - yy)) * C(11);
          t = t + ((SH3[3] * z) * (((2.0f * z) * z - (3.0f * x) * x) - (3.0f * y) * y)) * C(12);
          t = t + ((SH3[4] * x) * (((4.0f * z) * z - xx) - yy)) * C(13);
          t = t + ((SH3[5] * z) * (xx - yy)) * C(14);
          t = t + ((SH3[6] * x) * (xx - (3.0f * y) * y)) * C(15);
          col += t;
        }
      }
    }
#undef C
    col += 0.5f;
    rgb[c] = clampf(col, 0.0f, 1.0f);
  }
}

void gsr_oracle_sh(int64_t n, const float *xyz, const float *sh, int64_t sn, int64_t sk, int64_t sc,
                   const float *cam_center, int degree, float *rgb) {
  for (int64_t i = 0; i < n; ++i) sh_colour(xyz + 3 * i, sh + i * sn, sk, sc, cam_center, degree, rgb + 3 * i);
}

/* rasterize.py:41-56,97-120 : Sigma (row-major 3x3) from log-scales and an un-normalised quaternion. */
static void covariance3d(const float *ls, const float *rot, float *S) {
  float s[3] = {exp_scale(ls[0]), exp_scale(ls[1]), exp_scale(ls[2])};
  float nrm = sqrtf(((rot[0] * rot[0] + rot[1] * rot[1]) + rot[2] * rot[2]) + rot[3] * rot[3]);
  float den = nrm > 1e-12f ? nrm : 1e-12f;
  float w = rot[0] / den, x = rot[1] / den, y = rot[2] / den, z = rot[3] / den;
  float R[3][3];
  R[0][0] = (1.0f - 2.0f * (y * y)) - 2.0f * (z * z);
  R[0][1] = (2.0f * x) * y - (2.0f * z) * w;
  R[0][2] = (2.0f * x) * z + (2.0f * y) * w;
  R[1][0] = (2.0f * x) * y + (2.0f * z) * w;
  R[1][1] = (1.0f - 2.0f * (x * x)) - 2.0f * (z * z);
  R[1][2] = (2.0f * y) * z - (2.0f * x) * w;
  R[2][0] = (2.0f * x) * z - (2.0f * y) * w;
  R[2][1] = (2.0f * y) * z + (2.0f * x) * w;
  R[2][2] = (1.0f - 2.0f * (x * x)) - 2.0f * (y * y);
  float M[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) M[i][j] = R[i][j] * s[j];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) S[i * 3 + j] = (M[i][0] * M[j][0] + M[i][1] * M[j][1]) + M[i][2] * M[j][2];
}

void gsr_oracle_cov3d(int64_t n, const float *log_scales, const float *rots, float *cov3d) {
  for (int64_t i = 0; i < n; ++i) covariance3d(log_scales + 3 * i, rots + 4 * i, cov3d + 9 * i);
}

/* rasterize.py:201-252 for one Gaussian; returns c00,c01,c10,c11 BEFORE the cull zeroing of :388. */
static void covariance2d(const float *S, const float *cam, const gsr_camera *k, float *c) {
  const float *w = k->w2c_t;
  float z = cam[2];
  float txtz = cam[0] / z, tytz = cam[1] / z;
  float tx = fminf(k->limx, fmaxf(-k->limx, txtz)) * z;
  float ty = fminf(k->limy, fmaxf(-k->limy, tytz)) * z;
  float rz = 1.0f / z;
  float J00 = rz * k->focal_x, J02 = (-(k->focal_x * tx)) / (z * z);
  float J11 = rz * k->focal_y, J12 = (-(k->focal_y * ty)) / (z * z);
  float T[3][3];
  for (int i = 0; i < 3; ++i) { /* T = J . R_cam, bmm ascending k, no fma; zero terms kept */
    T[0][i] = (w[i * 4 + 0] * J00 + w[i * 4 + 1] * 0.0f) + w[i * 4 + 2] * J02;
    T[1][i] = (w[i * 4 + 0] * 0.0f + w[i * 4 + 1] * J11) + w[i * 4 + 2] * J12;
    T[2][i] = (w[i * 4 + 0] * 0.0f + w[i * 4 + 1] * 0.0f) + w[i * 4 + 2] * 0.0f;
  }
  float V[3][3] = {{S[0], S[1], S[2]}, {S[1], S[4], S[5]}, {S[2], S[5], S[8]}}; /* upper triangle, symmetrised */
  float A[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A[i][j] = (T[i][0] * V[0][j] + T[i][1] * V[1][j]) + T[i][2] * V[2][j];
  float P[2][2];
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) P[i][j] = (A[i][0] * T[j][0] + A[i][1] * T[j][1]) + A[i][2] * T[j][2];
  c[0] = P[0][0] + 0.3f;
  c[1] = P[0][1];
  c[2] = P[1][0];
  c[3] = P[1][1] + 0.3f;
}

/*
 * Everything the reference computes per Gaussian before its blend loop (rasterize.py:354-420).
 * Outputs (any may be NULL):
 *   cov3d[N,9] cam_means[N,3] mean2d[N,2] cov2d[N,4] (after cull zeroing) conic[N,3]=(a,c,b)
 *   tile_rect[N,4] i32 (tile units)  pixel_bbox[N,4] i32  rgb[N,3]  opacity[N]  depth[N]  radius[N]
 *   skip[N] u8 : 1 when the loop at rasterize.py:441 skips the Gaussian
 */
void gsr_oracle_preprocess(int64_t n, const float *xyz, const float *log_scales, const float *rots,
                           const float *opacity_logit, const float *sh, int64_t sn, int64_t sk, int64_t sc,
                           const gsr_camera *k, float *cov3d, float *cam_means, float *mean2d, float *cov2d,
                           float *conic, int32_t *tile_rect, int32_t *pixel_bbox, float *rgb, float *opacity,
                           float *depth, float *radius, uint8_t *skip) {
  const float Wf = (float)k->width, Hf = (float)k->height;
  const float Wm1 = (float)(k->width - 1), Hm1 = (float)(k->height - 1);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const float *p = xyz + 3 * i;
    float S[9];
    covariance3d(log_scales + 3 * i, rots + 4 * i, S);
    float cam[3], clip[4];
    for (int c = 0; c < 3; ++c) cam[c] = row_affine(k->w2c_t, c, p[0], p[1], p[2]);
    for (int c = 0; c < 4; ++c) clip[c] = row_affine(k->full_t, c, p[0], p[1], p[2]);
    int culled = cam[2] < 0.2f;
    if (culled) clip[0] = clip[1] = clip[2] = clip[3] = 0.0f;
    float pw = 1.0f / (clip[3] + 0.0000001f);
    float ndx = clip[0] * pw, ndy = clip[1] * pw;
    float mx = ((ndx + 1.0f) * Wf - 1.0f) / 2.0f;
    float my = ((ndy + 1.0f) * Hf - 1.0f) / 2.0f;
    float c2[4];
    covariance2d(S, cam, k, c2);
    if (culled) c2[0] = c2[1] = c2[2] = c2[3] = 0.0f;
    /* compute_covering_bbox, rasterize.py:163-197 */
    float det = c2[0] * c2[3] - c2[2] * c2[1];
    float trace = c2[0] + c2[3];
    float disc = fmaxf((trace * trace) / 4.0f - det, 0.1f);
    float sq = sqrtf(disc);
    float l1 = trace / 2.0f + sq, l2 = trace / 2.0f - sq;
    float r = ceilf(SPREAD * sqrtf(l1 > l2 ? l1 : l2));
    float ext = (r + (float)BLOCK) - 1.0f;
    int64_t rect[4];
    rect[0] = (int64_t)floorf(clampf((mx - r) / (float)BLOCK, 0.0f, Wm1));
    rect[1] = (int64_t)floorf(clampf((my - r) / (float)BLOCK, 0.0f, Hm1));
    rect[2] = (int64_t)floorf(clampf((mx + ext) / (float)BLOCK, 0.0f, Wm1));
    rect[3] = (int64_t)floorf(clampf((my + ext) / (float)BLOCK, 0.0f, Hm1));
    /* conic, rasterize.py:395-411 */
    float det_inv = det == 0.0f ? 0.0f : 1.0f / det;
    float ca = c2[3] * det_inv, cc = c2[0] * det_inv, cb = (-c2[1]) * det_inv;
    /* pixel bbox, rasterize.py:415-420 */
    int64_t bb[4];
    bb[0] = clampi(rect[0] * BLOCK, 0, k->width - 1);
    bb[1] = clampi(rect[1] * BLOCK, 0, k->height - 1);
    bb[2] = clampi(rect[2] * BLOCK, 0, k->width - 1);
    bb[3] = clampi(rect[3] * BLOCK, 0, k->height - 1);
    int64_t area = (bb[2] - bb[0]) * (bb[3] - bb[1]);
    if (cov3d) memcpy(cov3d + 9 * i, S, sizeof S);
    if (cam_means) memcpy(cam_means + 3 * i, cam, sizeof cam);
    if (mean2d) { mean2d[2 * i] = mx; mean2d[2 * i + 1] = my; }
    if (cov2d) memcpy(cov2d + 4 * i, c2, sizeof c2);
    if (conic) { conic[3 * i] = ca; conic[3 * i + 1] = cc; conic[3 * i + 2] = cb; }
    if (tile_rect) for (int j = 0; j < 4; ++j) tile_rect[4 * i + j] = (int32_t)rect[j];
    if (pixel_bbox) for (int j = 0; j < 4; ++j) pixel_bbox[4 * i + j] = (int32_t)bb[j];
    if (rgb) sh_colour(p, sh + i * sn, sk, sc, k->cam_center, k->sh_degree, rgb + 3 * i);
    if (opacity) opacity[i] = sigmoid_ref(opacity_logit[i]);
    if (depth) depth[i] = cam[2];
    if (radius) radius[i] = r;
    if (skip) skip[i] = (uint8_t)(area == 0 || ca == 0.0f || cc == 0.0f || cb == 0.0f);
  }
}

/* Stable ascending argsort of f32 depths (bottom-up merge sort; a<b comparator like torch.sort(stable=True)). */
void gsr_oracle_argsort(int64_t n, const float *key, int64_t *order) {
  int64_t *tmp = (int64_t *)malloc((size_t)n * sizeof(int64_t));
  int64_t *src = order, *dst = tmp;
  for (int64_t i = 0; i < n; ++i) order[i] = i;
  for (int64_t width = 1; width < n; width *= 2) {
#pragma omp parallel for schedule(static)
    for (int64_t lo = 0; lo < n; lo += 2 * width) {
      int64_t mid = lo + width < n ? lo + width : n, hi = lo + 2 * width < n ? lo + 2 * width : n;
      int64_t a = lo, b = mid, o = lo;
      while (a < mid && b < hi) dst[o++] = (key[src[b]] < key[src[a]]) ? src[b++] : src[a++];
      while (a < mid) dst[o++] = src[a++];
      while (b < hi) dst[o++] = src[b++];
    }
    int64_t *t = src; src = dst; dst = t;
  }
  if (src != order) memcpy(order, src, (size_t)n * sizeof(int64_t));
  free(tmp);
}

/* rasterize.py:255-305 restricted to pixel columns [xlo,xhi): one Gaussian blended into screen[W,H,3], T[W,H]. */
static inline void blend_one(int64_t g, const int32_t *bb, const float *mean2d, const float *conic, const float *rgb,
                             const float *opacity, int H, float *screen, float *T, int xlo, int xhi) {
  const int32_t *b = bb + 4 * g;
  int x0 = b[0] > xlo ? b[0] : xlo, x1 = b[2] < xhi ? b[2] : xhi;
  const float a = conic[3 * g], c = conic[3 * g + 1], bxy = conic[3 * g + 2];
  const float mx = mean2d[2 * g], my = mean2d[2 * g + 1], o = opacity[g];
  const float *col = rgb + 3 * g;
  for (int x = x0; x < x1; ++x) {
    float dx = mx - (float)x;
    for (int y = b[1]; y < b[3]; ++y) {
      float dy = my - (float)y;
      float power = -0.5f * (a * (dx * dx) + c * (dy * dy)) - (bxy * dx) * dy;
      float alpha = fminf(o * expf(power), MAX_DENSITY);
      if (!(alpha > MIN_ALPHA && power <= 0.0f)) continue;
      size_t px = (size_t)x * H + y;
      float t = T[px];
      screen[3 * px + 0] += (alpha * col[0]) * t;
      screen[3 * px + 1] += (alpha * col[1]) * t;
      screen[3 * px + 2] += (alpha * col[2]) * t;
      T[px] = t * (1.0f - alpha);
    }
  }
}

/*
 * The loop of rasterize.py:436-446.  screen[W,H,3] and T[W,H] are x-major like the reference and must
 * be initialised by the caller (zeros / ones).  Blends order[first..last) only (progressive renders).
 * Returns the number of Gaussians actually blended.  Parallel over disjoint pixel-column strips:
 * every strip sees every Gaussian in the same order, so the result is identical to the serial loop.
 */
int64_t gsr_oracle_blend(int64_t n_order, const int64_t *order, const uint8_t *skip, const int32_t *pixel_bbox,
                         const float *mean2d, const float *conic, const float *rgb, const float *opacity, int W,
                         int H, float *screen, float *T, int64_t first, int64_t last) {
  int64_t blended = 0;
  if (last > n_order) last = n_order;
  for (int64_t i = first; i < last; ++i) blended += !skip[order[i]];
  int strips = 1;
#ifdef _OPENMP
  strips = omp_get_max_threads();
#endif
  if (strips > W / 32) strips = W / 32 > 0 ? W / 32 : 1;
#pragma omp parallel for schedule(static, 1) num_threads(strips)
  for (int s = 0; s < strips; ++s) {
    int xlo = (int)((int64_t)W * s / strips), xhi = (int)((int64_t)W * (s + 1) / strips);
    for (int64_t i = first; i < last; ++i) {
      int64_t g = order[i];
      if (skip[g]) continue;
      blend_one(g, pixel_bbox, mean2d, conic, rgb, opacity, H, screen, T, xlo, xhi);
    }
  }
  return blended;
}

/*
 * For (Gaussian, tile) pairs: does ANY pixel of the 16x16 tile pass the reference's keep test (rasterize.py:285-291)
 * for that Gaussian?  Used to check that the pairs the CUDA binning drops ("exact tile culling") could not have
 * changed the image.  Same arithmetic as blend_one; pixels x >= W-1 or y >= H-1 are never written (Q8).
 */
void gsr_oracle_pair_hits(int64_t n_pairs, const int64_t *gids, const int64_t *tiles, int tiles_x, int W, int H,
                          const float *mean2d, const float *conic, const float *opacity, uint8_t *hits) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t p = 0; p < n_pairs; ++p) {
    const int64_t g = gids[p];
    const int tx = (int)(tiles[p] % tiles_x), ty = (int)(tiles[p] / tiles_x);
    const float a = conic[3 * g], c = conic[3 * g + 1], bxy = conic[3 * g + 2];
    const float mx = mean2d[2 * g], my = mean2d[2 * g + 1], o = opacity[g];
    uint8_t hit = 0;
    for (int x = tx * BLOCK; x < tx * BLOCK + BLOCK && x < W - 1 && !hit; ++x) {
      float dx = mx - (float)x;
      for (int y = ty * BLOCK; y < ty * BLOCK + BLOCK && y < H - 1; ++y) {
        float dy = my - (float)y;
        float power = -0.5f * (a * (dx * dx) + c * (dy * dy)) - (bxy * dx) * dy;
        float alpha = fminf(o * expf(power), MAX_DENSITY);
        if (alpha > MIN_ALPHA && power <= 0.0f) { hit = 1; break; }
      }
    }
    hits[p] = hit;
  }
}

/* Sum over Gaussians of bbox pixels = number of Gaussian-pixel evaluations the reference performs. */
int64_t gsr_oracle_count_evals(int64_t n, const uint8_t *skip, const int32_t *bb) {
  int64_t tot = 0;
  for (int64_t i = 0; i < n; ++i)
    if (!skip[i]) tot += (int64_t)(bb[4 * i + 2] - bb[4 * i]) * (bb[4 * i + 3] - bb[4 * i + 1]);
  return tot;
}

int gsr_oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
