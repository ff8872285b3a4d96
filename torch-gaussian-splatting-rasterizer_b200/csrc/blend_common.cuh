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
// a few 1e-7 (continuous, harmless); the DECISION alpha > 1/255 (:291) is discontinuous, so inside a +-1e-5 (relative)
// band around the threshold it is re-taken with the reference's own arithmetic, o * expf(power).  Call only when
// power >= thr.
//   fast path (everything but a thin shell at the rim of the splat): e > 1/255 + 4e-8 is above the band, so the
//   decision is `power <= 0` and the value min(e, 0.99) - two compares, no subtraction, no second exp;
//   band path (kept out of the hot instruction stream): the exact re-evaluation.  alpha - 1/255 is exact there
//   (Sterbenz), so the band test means what it says.
//   The re-evaluation uses a CORRECTLY ROUNDED exp (double exp rounded once to f32) - CUDA's expf is only good to
//   2 ulp, and one ulp decides here: at BASELINE config 2 (2594x1681) a Gaussian with o * e^power = 1/255 + 0.2 ulp was
//   kept by the expf version and dropped by the reference (torch's exp is correctly rounded on ~99 % of inputs), a
//   2.3e-3 error at one pixel (profiles/r02_pixel_attribution_C2_full.json).
constexpr float kBandHalfWidth = 3.9e-8f;
constexpr float kBandHi = kMinAlpha + 4.0e-8f;  // any rounding of this sum (< 2.4e-10) keeps it above 1/255 + 3.9e-8
// exp(x) rounded once to f32, for the band path (x = ln(1/(255 o)) there, i.e. -5.6 <= x <= 0; valid for |x| < 700).
// Inline FP64 - argument reduction by ln 2 (hi + lo), degree-13 Taylor polynomial on |r| <= 0.347 (truncation 4e-18),
// exact scaling by 2^k - instead of a call to exp(): an ABI call inside the blend loop cost 6 registers per thread.
__device__ __forceinline__ float exp_correctly_rounded(float xf) {
  const double x = (double)xf;
  const double k = rint(x * 1.4426950408889634074);
  double r = fma(k, -6.93147180369123816490e-01, x);
  r = fma(k, -1.90821492927058770002e-10, r);
  double p = 1.6059043836821613e-10;   // 1/13!
  p = fma(p, r, 2.08767569878681e-09);   // 1/12!
  p = fma(p, r, 2.505210838544172e-08);  // 1/11!
  p = fma(p, r, 2.755731922398589e-07);  // 1/10!
  p = fma(p, r, 2.7557319223985893e-06); // 1/9!
  p = fma(p, r, 2.48015873015873e-05);   // 1/8!
  p = fma(p, r, 1.984126984126984e-04);  // 1/7!
  p = fma(p, r, 1.388888888888889e-03);  // 1/6!
  p = fma(p, r, 8.333333333333333e-03);  // 1/5!
  p = fma(p, r, 4.1666666666666664e-02); // 1/4!
  p = fma(p, r, 1.6666666666666666e-01); // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double scale = __longlong_as_double(((long long)k + 1023ll) << 52);  // 2^k, exact for -1022 <= k <= 1023
  return (float)(p * scale);
}
// Out of line on purpose: inlined, the sixteen copies (4 records x 4 pixel slots) of the FP64 polynomial made the blend
// loop 1720 instructions instead of 1008 (instruction-fetch stalls); the call costs nothing where it matters because
// the band is reached by about one evaluation in 10^5.  -DGSR_BAND_INLINE=1 restores the inlined form (A/B builds).
#ifndef GSR_BAND_INLINE
#define GSR_BAND_INLINE 0
#endif
#if GSR_BAND_INLINE
__device__ __forceinline__ float blend_alpha_exact(float power, float opacity) {
#else
static __device__ __noinline__ float blend_alpha_exact(float power, float opacity) {
#endif
  return fminf(fmul(opacity, exp_correctly_rounded(power)), kMaxDensity);
}
__device__ __forceinline__ bool blend_alpha_band(float power, float opacity, float &alpha) {
  if (fabsf(alpha - kMinAlpha) < kBandHalfWidth) alpha = blend_alpha_exact(power, opacity);
  return alpha > kMinAlpha && power <= 0.0f;  // :291
}
__device__ __forceinline__ bool blend_alpha(float power, float opacity, float log2_opacity, float &alpha) {
  const float e = ex2_approx(fmaf(power, 1.4426950408889634f, log2_opacity));
  alpha = fminf(e, kMaxDensity);
  if (e > kBandHi) return power <= 0.0f;
  return blend_alpha_band(power, opacity, alpha);
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
__device__ __forceinline__ uint32_t smem_addr32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {  // after cp_async_wait / mbar_wait (volatile asm keeps the order)
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- bulk async copies (TMA unit, UBLKCP) + mbarrier completion: the alternative staging path of the blend kernel ----
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_arrive(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(a), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

}  // namespace gsr
