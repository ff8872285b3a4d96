// blend_common.cuh - the per-(pixel, Gaussian) evaluation shared by the forward and backward blend kernels, so that
// both take bit-identical keep decisions (the backward pass replays the forward's contributor list).
#pragma once

#include "gsr_common.cuh"

namespace gsr {

constexpr float kMinAlpha = (float)(1.0 / 255.0);   // MIN_ALPHA, rasterize.py:38 (compared in f32 like torch does)
constexpr float kMaxDensity = 0.99f;                // MAX_GAUSSIAN_DENSITY, rasterize.py:36
constexpr float kStopT = 1e-4f;                     // a pixel stops AFTER the blend that takes T below this (Q12)

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// (A dx^2 + C dy^2) + (B dx) dy == -0.5 (a dx^2 + c dy^2) - (b dx) dy of rasterize.py:280-283, bit for bit
__device__ __forceinline__ float blend_power(float A, float C, float B, float dx, float dy) {
  return fadd(fadd(fmul(A, fmul(dx, dx)), fmul(C, fmul(dy, dy))), fmul(fmul(B, dx), dy));
}

// alpha = min(o*exp(power), 0.99) (:285-288) as one FFMA + MUFU.EX2.  Its VALUE may differ from the reference's by
// a few 1e-7 (continuous, harmless); the DECISION alpha > 1/255 (:291) is discontinuous, so inside a +-1e-5 band
// around the threshold it is re-taken with the reference's own arithmetic.  Call only when power >= thr.
__device__ __forceinline__ bool blend_alpha(float power, float opacity, float log2_opacity, float &alpha) {
  alpha = fminf(ex2_approx(fmaf(power, 1.4426950408889634f, log2_opacity)), kMaxDensity);
  if (fabsf(alpha - kMinAlpha) < 3.9e-8f) alpha = fminf(fmul(opacity, expf(power)), kMaxDensity);  // 1e-5 relative
  return alpha > kMinAlpha && power <= 0.0f;  // :291
}


// ---- packed FP32 (Blackwell FFMA2 / FMUL2 / FADD2: two independent IEEE single operations per instruction) --------
// One issue slot does the work of two, each half rounded exactly like the scalar operation, so results are bit-identical
// to the scalar code.  ptxas contracts mul.rn.f32x2 followed by add.rn.f32x2 into one FFMA2 (it honours neither the
// .rn qualifiers nor -fmad=false for the packed forms), which would drop a rounding the reference performs: products
// that feed an addition are therefore written as fma(a, b, -0), which cannot be contracted and equals round(a*b).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 v;
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi));
  return v;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 v;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(v) : "l"(a), "l"(b));
  return v;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 v;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(v) : "l"(a), "l"(b));
  return v;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {  // only where the product does not feed an add/sub
  f32x2 v;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(v) : "l"(a), "l"(b));
  return v;
}
__device__ __forceinline__ f32x2 mul2_exact(f32x2 a, f32x2 b, f32x2 neg_zero) {  // round(a*b), never contracted
  f32x2 v;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(v) : "l"(a), "l"(b), "l"(neg_zero));
  return v;
}

// power of four pixels of one column (shared dx), two packed pairs: (A dx^2 + C dy^2) + (B dx) dy in the reference's
// operation order (rasterize.py:280-283), each half rounded like the scalar blend_power.  `neg_zero2` must be a
// packed -0.0f the compiler cannot see through (see above).
__device__ __forceinline__ void blend_power4(float my, float C, float adx2, float bdx, const float fy[4],
                                             f32x2 neg_zero2, float dy_out[4], float power[4]) {
  const f32x2 my2 = pack2(my, my), c2 = pack2(C, C), a2 = pack2(adx2, adx2), b2 = pack2(bdx, bdx);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const f32x2 dy = sub2(my2, pack2(fy[2 * h], fy[2 * h + 1]));
    const f32x2 cdy2 = mul2_exact(c2, mul2(dy, dy), neg_zero2);
    const f32x2 p = add2(add2(a2, cdy2), mul2_exact(b2, dy, neg_zero2));
    unpack2(dy, dy_out[2 * h], dy_out[2 * h + 1]);
    unpack2(p, power[2 * h], power[2 * h + 1]);
  }
}

// cp.async (LDGSTS) staging helpers shared by the forward and backward blend kernels
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

}  // namespace gsr
