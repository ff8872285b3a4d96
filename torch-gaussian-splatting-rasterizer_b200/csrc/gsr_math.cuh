// gsr_math.cuh - per-Gaussian device math shared by the fused preprocess kernel and the staged helpers.
// Every function states the reference lines it follows (/root/reference/...).  Arithmetic is written one
// rounded f32 operation at a time (gsr_common.cuh) in the order torch-CPU executes the reference:
//   [N,3]@[3,k]+row  -> fma chain ascending k, separately rounded add        (sgemm,   rasterize.py:86,374)
//   3x3 bmm          -> plain multiply/add ascending k, zero terms kept      (rasterize.py:118,120,232,245)
//   scalar / tensor  -> reciprocal(tensor) * scalar                          (rasterize.py:225,227,381)
#pragma once

#include "gsr_common.cuh"

namespace gsr {

// SH basis constants, spherical_harmonics.py:4-24 (rounded to f32 as torch does for python scalars)
#define kSH0 0.28209479177387814f
#define kSH1 0.4886025119029199f
#define kSH2_0 1.0925484305920792f
#define kSH2_1 -1.0925484305920792f
#define kSH2_2 0.31539156525252005f
#define kSH2_3 -1.0925484305920792f
#define kSH2_4 0.5462742152960396f
#define kSH3_0 -0.5900435899266435f
#define kSH3_1 2.890611442640554f
#define kSH3_2 -0.4570457994644658f
#define kSH3_3 0.3731763325901154f
#define kSH3_4 -0.4570457994644658f
#define kSH3_5 1.445305721320277f
#define kSH3_6 -0.5900435899266435f

// torch.exp on the contiguous scale tensor (rasterize.py:97) runs in MKL VML's high-accuracy mode, which is
// within ~1 % of inputs of the correctly rounded result; the closest reproducible value is the correctly
// rounded one.  Four of these per Gaussian: FP64 cost is irrelevant next to the 236 bytes read.
__device__ __forceinline__ float exp_scale(float x) { return (float)exp((double)x); }

// torch.sigmoid (rasterize.py:358) is 1/(1+e), e from the vectorised Sleef expf u10 in its FMA form; this
// sequence reproduces it bit for bit (checked on 2M samples against torch 2.11 CPU).
__device__ __forceinline__ float exp_sleef_u10(float d) {
  float qf = rintf(fmul(d, 1.442695040888963407359924681001892137426645954152985934135449406931f));
  int q = (int)qf;
  float s = ffma(qf, -0.693145751953125f, d);
  s = ffma(qf, -1.428606765330187045e-06f, s);
  float u = 0.000198527617612853646278381f;
  u = ffma(u, s, 0.00139304355252534151077271f);
  u = ffma(u, s, 0.00833336077630519866943359f);
  u = ffma(u, s, 0.0416664853692054748535156f);
  u = ffma(u, s, 0.166666671633720397949219f);
  u = ffma(u, s, 0.5f);
  u = fadd(1.0f, ffma(fmul(s, s), u, s));
  float a = __int_as_float(((q >> 1) + 0x7f) << 23);
  float b = __int_as_float(((q - (q >> 1)) + 0x7f) << 23);
  u = fmul(fmul(u, a), b);
  if (d < -104.0f) u = 0.0f;
  if (d > 100.0f) u = __int_as_float(0x7f800000);
  return u;
}
__device__ __forceinline__ float sigmoid_ref(float x) { return frcp(fadd(1.0f, exp_sleef_u10(-x))); }

// p @ M[:3,:] + M[3,:], one output column (m is a row-major 4x4)
__device__ __forceinline__ float row_affine(const float *m, int c, float x, float y, float z) {
  float acc = fmul(x, m[0 * 4 + c]);
  acc = ffma(y, m[1 * 4 + c], acc);
  acc = ffma(z, m[2 * 4 + c], acc);
  return fadd(acc, m[3 * 4 + c]);
}

// quaternion_to_rotation_matrix, rasterize.py:52-54 (no normalisation here)
__device__ __forceinline__ void quat_to_rot(float w, float x, float y, float z, float R[3][3]) {
  R[0][0] = fsub(fsub(1.0f, fmul(2.0f, fmul(y, y))), fmul(2.0f, fmul(z, z)));
  R[0][1] = fsub(fmul(fmul(2.0f, x), y), fmul(fmul(2.0f, z), w));
  R[0][2] = fadd(fmul(fmul(2.0f, x), z), fmul(fmul(2.0f, y), w));
  R[1][0] = fadd(fmul(fmul(2.0f, x), y), fmul(fmul(2.0f, z), w));
  R[1][1] = fsub(fsub(1.0f, fmul(2.0f, fmul(x, x))), fmul(2.0f, fmul(z, z)));
  R[1][2] = fsub(fmul(fmul(2.0f, y), z), fmul(fmul(2.0f, x), w));
  R[2][0] = fsub(fmul(fmul(2.0f, x), z), fmul(fmul(2.0f, y), w));
  R[2][1] = fadd(fmul(fmul(2.0f, y), z), fmul(fmul(2.0f, x), w));
  R[2][2] = fsub(fsub(1.0f, fmul(2.0f, fmul(x, x))), fmul(2.0f, fmul(y, y)));
}

// get_covariance_matrix_from_mesh, rasterize.py:97-120: Sigma = (R diag(s)) (R diag(s))^T, row-major 3x3
__device__ __forceinline__ void covariance3d(const float ls[3], const float rot[4], float S[9]) {
  float s[3] = {exp_scale(ls[0]), exp_scale(ls[1]), exp_scale(ls[2])};
  float nrm = fsqrt(fadd(fadd(fadd(fmul(rot[0], rot[0]), fmul(rot[1], rot[1])), fmul(rot[2], rot[2])),
                         fmul(rot[3], rot[3])));
  float den = nrm > 1e-12f ? nrm : 1e-12f;  // F.normalize eps, rasterize.py:112
  float R[3][3];
  quat_to_rot(fdiv(rot[0], den), fdiv(rot[1], den), fdiv(rot[2], den), fdiv(rot[3], den), R);
  float M[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) M[i][j] = fmul(R[i][j], s[j]);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      S[i * 3 + j] = fadd(fadd(fmul(M[i][0], M[j][0]), fmul(M[i][1], M[j][1])), fmul(M[i][2], M[j][2]));
}

// compute_2d_covariance, rasterize.py:201-252: c = (c00, c01, c10, c11), before the cull zeroing of :388
__device__ __forceinline__ void covariance2d(const float S[9], const float cam[3], const gsr_camera &k, float c[4]) {
  const float *w = k.w2c_t;
  float z = cam[2];
  float txtz = fdiv(cam[0], z), tytz = fdiv(cam[1], z);
  float tx = fmul(fminf(k.limx, fmaxf(-k.limx, txtz)), z);
  float ty = fmul(fminf(k.limy, fmaxf(-k.limy, tytz)), z);
  float rz = frcp(z), zz = fmul(z, z);
  float J00 = fmul(rz, k.focal_x), J02 = fdiv(-fmul(k.focal_x, tx), zz);
  float J11 = fmul(rz, k.focal_y), J12 = fdiv(-fmul(k.focal_y, ty), zz);
  float T[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {  // T = J . R_cam with the structural zeros of J multiplied through
    T[0][i] = fadd(fadd(fmul(w[i * 4 + 0], J00), fmul(w[i * 4 + 1], 0.0f)), fmul(w[i * 4 + 2], J02));
    T[1][i] = fadd(fadd(fmul(w[i * 4 + 0], 0.0f), fmul(w[i * 4 + 1], J11)), fmul(w[i * 4 + 2], J12));
    T[2][i] = fadd(fadd(fmul(w[i * 4 + 0], 0.0f), fmul(w[i * 4 + 1], 0.0f)), fmul(w[i * 4 + 2], 0.0f));
  }
  float V[3][3] = {{S[0], S[1], S[2]}, {S[1], S[4], S[5]}, {S[2], S[5], S[8]}};  // upper triangle symmetrised :234-243
  float A[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      A[i][j] = fadd(fadd(fmul(T[i][0], V[0][j]), fmul(T[i][1], V[1][j])), fmul(T[i][2], V[2][j]));
  float P[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
      P[i][j] = fadd(fadd(fmul(A[i][0], T[j][0]), fmul(A[i][1], T[j][1])), fmul(A[i][2], T[j][2]));
  c[0] = fadd(P[0][0], 0.3f);
  c[1] = P[0][1];
  c[2] = P[1][0];
  c[3] = fadd(P[1][1], 0.3f);
}

// same, from the upper triangle S6 = (S00 S01 S02 S11 S12 S22) the reference symmetrises from (:234-243)
__device__ __forceinline__ void covariance2d_sym(const float S6[6], const float cam[3], const gsr_camera &k, float c[4]) {
  const float S[9] = {S6[0], S6[1], S6[2], S6[1], S6[3], S6[4], S6[2], S6[4], S6[5]};
  covariance2d(S, cam, k, c);
}

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

// compute_covering_bbox, rasterize.py:163-197: radius and the reference's tile rect (tile units, as i64 there)
__device__ __forceinline__ void covering_rect(float mx, float my, const float c[4], int width, int height,
                                              float &radius, int rect[4]) {
  float det = fsub(fmul(c[0], c[3]), fmul(c[2], c[1]));
  float trace = fadd(c[0], c[3]);
  float half = fmul(trace, 0.5f);  // x / 2^k == x * 2^-k bit for bit: the reference's divisions by 2, 4 and 16 cost one FMUL
  float disc = fmaxf(fsub(fmul(fmul(trace, trace), 0.25f), det), 0.1f);
  float sq = fsqrt(disc);
  float l1 = fadd(half, sq), l2 = fsub(half, sq);
  radius = ceilf(fmul(3.0f, fsqrt(l1 > l2 ? l1 : l2)));
  float ext = fsub(fadd(radius, (float)GSR_TILE), 1.0f);
  const float wm1 = (float)(width - 1), hm1 = (float)(height - 1), inv_blk = 1.0f / (float)GSR_TILE;
  rect[0] = (int)floorf(clampf(fmul(fsub(mx, radius), inv_blk), 0.0f, wm1));
  rect[1] = (int)floorf(clampf(fmul(fsub(my, radius), inv_blk), 0.0f, hm1));
  rect[2] = (int)floorf(clampf(fmul(fadd(mx, ext), inv_blk), 0.0f, wm1));
  rect[3] = (int)floorf(clampf(fmul(fadd(my, ext), inv_blk), 0.0f, hm1));
}

// conic ("sigmas"), rasterize.py:395-411: (a, c, b) = (c11, c00, -c01) * det_inv
__device__ __forceinline__ void conic_from_cov(const float c[4], float &a, float &cc, float &b) {
  float det = fsub(fmul(c[0], c[3]), fmul(c[2], c[1]));
  float det_inv = det == 0.0f ? 0.0f : frcp(det);
  a = fmul(c[3], det_inv);
  cc = fmul(c[0], det_inv);
  b = fmul(-c[1], det_inv);
}

// sh_to_rgb, spherical_harmonics.py:36-71.  `coef(k)` returns SH coefficient k of the current channel.
template <typename Coef>
__device__ __forceinline__ float sh_channel(float x, float y, float z, int degree, Coef coef) {
  float col = fmul(coef(0), kSH0);
  if (degree > 0) {
    float t = fmul(fmul(-kSH1, y), coef(1));
    t = fadd(t, fmul(fmul(kSH1, z), coef(2)));
    t = fsub(t, fmul(fmul(kSH1, x), coef(3)));
    col = fadd(col, t);
    if (degree > 1) {
      float xx = fmul(x, x), yy = fmul(y, y);
      t = fmul(fmul(fmul(kSH2_0, x), y), coef(4));
      t = fadd(t, fmul(fmul(fmul(kSH2_1, y), z), coef(5)));
      t = fadd(t, fmul(fmul(kSH2_2, fsub(fsub(fmul(fmul(2.0f, z), z), xx), yy)), coef(6)));
      t = fadd(t, fmul(fmul(fmul(kSH2_3, x), z), coef(7)));
      t = fadd(t, fmul(fmul(kSH2_4, fsub(xx, yy)), coef(8)));
      col = fadd(col, t);
      if (degree > 2) {
        float zz4 = fmul(fmul(4.0f, z), z), zz2 = fmul(fmul(2.0f, z), z);
        float xx3 = fmul(fmul(3.0f, x), x), yy3 = fmul(fmul(3.0f, y), y);
        t = fmul(fmul(fmul(kSH3_0, y), fsub(xx3, yy)), coef(9));
        t = fadd(t, fmul(fmul(fmul(fmul(kSH3_1, x), y), z), coef(10)));
        t = fadd(t, fmul(fmul(fmul(kSH3_2, y), fsub(fsub(zz4, xx), yy)), coef(11)));
        t = fadd(t, fmul(fmul(fmul(kSH3_3, z), fsub(fsub(zz2, xx3), yy3)), coef(12)));
        t = fadd(t, fmul(fmul(fmul(kSH3_4, x), fsub(fsub(zz4, xx), yy)), coef(13)));
        t = fadd(t, fmul(fmul(fmul(kSH3_5, z), fsub(xx, yy)), coef(14)));
        t = fadd(t, fmul(fmul(fmul(kSH3_6, x), fsub(xx, yy3)), coef(15)));
        col = fadd(col, t);
      }
    }
  }
  col = fadd(col, 0.5f);
  return clampf(col, 0.0f, 1.0f);
}

// ---- exact tile culling ------------------------------------------------------------------------------------
// A (Gaussian, tile) pair can only change the image if some pixel of the tile passes the reference's keep test
// alpha = o*exp(power) > 1/255 (rasterize.py:291), i.e. power >= thr.  power is a concave quadratic of the pixel
// position, so its maximum over the tile's pixel box is 0 when the mean lies inside and otherwise sits on one of
// the four edges, where it is a clamped 1-D parabola.  Pairs whose maximum stays below thr (minus a rounding
// margin) are dropped from the tile lists: the image is unchanged, binning and blending get ~35 % less work.
// Used by BOTH the count (preprocess) and the emission (binning) so they always agree.
struct CullGaussian {
  float mx, my, A, C, B, kv, ku, limit;  // kv = -B/(2C), ku = -B/(2A); limit = thr - margin(thr)
};
__device__ __forceinline__ CullGaussian make_cull(float mx, float my, float A, float C, float B, float thr) {
  CullGaussian g;
  g.mx = mx; g.my = my; g.A = A; g.C = C; g.B = B;
  // approximate divisions (2 ulp): the apex is only a parameter at which the exact parabola is evaluated, and the
  // keep test carries a 2^-16 relative margin
  g.kv = __fdividef(-B, 2.0f * C);
  g.ku = __fdividef(-B, 2.0f * A);
  g.limit = thr - 1e-4f * (1.0f + fabsf(thr));
  return g;
}
__device__ __forceinline__ float parabola_max(float q2, float q1, float q0, float t_opt, float lo, float hi) {
  const float t = fminf(fmaxf(t_opt, lo), hi);
  return fmaf(fmaf(q2, t, q1), t, q0);
}
// true when no pixel with x in [px0,px1], y in [py0,py1] can pass the keep test
__device__ __forceinline__ bool tile_is_empty(const CullGaussian &g, float px0, float py0, float px1, float py1) {
  const float u0 = px0 - g.mx, u1 = px1 - g.mx, v0 = py0 - g.my, v1 = py1 - g.my;
  if (u0 <= 0.0f && u1 >= 0.0f && v0 <= 0.0f && v1 >= 0.0f) return false;  // mean inside: power reaches 0
  {  // quick accept: the power at the box point nearest to the mean already reaches the limit (most kept tiles)
    const float cu = fminf(fmaxf(0.0f, u0), u1), cv = fminf(fmaxf(0.0f, v0), v1);
    if (fmaf(g.A * cu, cu, fmaf(g.C * cv, cv, g.B * cu * cv)) >= g.limit) return false;
  }
  // Only the edges that FACE the mean can hold the maximum.  At the maximiser p* of a concave f over the box the gradient
  // lies in the normal cone of the box; for a negative definite quadratic grad f(p) . p = 2 f(p) < 0, so an outward normal
  // (+1, 0) at p* needs u1 < 0, (-1, 0) needs u0 > 0, likewise in v: the edge u = u0 qualifies only when the box lies to
  // the right of the mean, u = u1 only when it lies to the left (a corner maximiser belongs to a qualifying edge too).
  // At most one vertical and one horizontal edge qualify (the mean is not inside): two clamped parabolas instead of four.
  // edge u = eu: f(v) = C v^2 + (B eu) v + A eu^2 ; edge v = ev: f(u) = A u^2 + (B ev) u + C ev^2
  const float eu = u0 > 0.0f ? u0 : u1, ev = v0 > 0.0f ? v0 : v1;
  const float fu = parabola_max(g.C, g.B * eu, g.A * eu * eu, g.kv * eu, v0, v1);
  const float fv = parabola_max(g.A, g.B * ev, g.C * ev * ev, g.ku * ev, u0, u1);
  const float ninf = __int_as_float(0xff800000);
  float best = fmaxf((u0 > 0.0f || u1 < 0.0f) ? fu : ninf, (v0 > 0.0f || v1 < 0.0f) ? fv : ninf);
  if (!(fu == fu) || !(fv == fv)) best = __int_as_float(0x7fc00000);  // any NaN: keep the tile (the comparison below is false)
  // rounding margin proportional to the largest term that can cancel inside `power` on this box
  const float U = fmaxf(fabsf(u0), fabsf(u1)), V = fmaxf(fabsf(v0), fabsf(v1));
  const float mag = fabsf(g.A) * U * U + fabsf(g.C) * V * V + fabsf(g.B) * U * V;
  return best < g.limit - 1.52587890625e-05f * mag;  // any NaN compares false -> keep
}
constexpr int kCullMaxTiles = 64;  // rects up to 64 tiles carry a 64-bit keep mask; larger ones are not culled

// position of the n-th (0-based) set bit of a 64-bit mask (lo, hi); n < popc(mask)
__device__ __forceinline__ uint32_t nth_set_bit(uint32_t lo, uint32_t hi, uint32_t n) {
  uint32_t c = __popc(lo), m = lo, base = 0;
  if (n >= c) { n -= c; m = hi; base = 32; }
  c = __popc(m & 0xFFFFu); if (n >= c) { n -= c; base += 16; m >>= 16; }
  c = __popc(m & 0xFFu);   if (n >= c) { n -= c; base += 8;  m >>= 8; }
  c = __popc(m & 0xFu);    if (n >= c) { n -= c; base += 4;  m >>= 4; }
  c = __popc(m & 0x3u);    if (n >= c) { n -= c; base += 2;  m >>= 2; }
  if (n >= (m & 1u)) base += 1;
  return base;
}

__device__ __forceinline__ void view_direction(const float p[3], const float cc[3], float &x, float &y, float &z) {
  float dx = fsub(p[0], cc[0]), dy = fsub(p[1], cc[1]), dz = fsub(p[2], cc[2]);
  float nrm = fsqrt(fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz)));
  x = fdiv(dx, nrm);
  y = fdiv(dy, nrm);
  z = fdiv(dz, nrm);
}

}  // namespace gsr
