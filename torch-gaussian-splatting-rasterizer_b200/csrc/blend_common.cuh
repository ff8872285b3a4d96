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
  if (fabsf(fmaf(alpha, 255.0f, -1.0f)) < 1e-5f) alpha = fminf(fmul(opacity, expf(power)), kMaxDensity);
  return alpha > kMinAlpha && power <= 0.0f;  // :291
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
