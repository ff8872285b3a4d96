// gsr_common.cuh - shared host/device plumbing for libgsr_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/gsr_b200.h"

namespace gsr {

// ---- error reporting across the C ABI (no exceptions) -------------------------------------------
void set_error(const char *fmt, ...);
#define GSR_CHECK_CUDA(expr)                                                                  \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      gsr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return GSR_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)
#define GSR_REQUIRE(cond, msg)                              \
  do {                                                      \
    if (!(cond)) {                                          \
      gsr::set_error("invalid argument: %s", msg);          \
      return GSR_ERR_INVALID_ARGUMENT;                      \
    }                                                       \
  } while (0)
#define GSR_LAUNCH_CHECK() GSR_CHECK_CUDA(cudaGetLastError())

// ---- per-device state ------------------------------------------------------------------------------------------
// One process may drive several GPUs (and several host threads): nothing device-specific is cached process-wide.
constexpr int kMaxDevices = 64;
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev >= 0 && dev < kMaxDevices ? dev : 0;
}
// Makes `device` current for the lifetime of the object and restores the caller's device afterwards: an object bound to
// a device (gsr_host_renderer) may be driven by a thread whose current device is another one, and a library call must not
// leave the caller - torch included, which reads cudaGetDevice - on a different device than it came in with.
struct DeviceScope {
  int before = -1;
  cudaError_t status = cudaSuccess;
  explicit DeviceScope(int device) {
    status = cudaGetDevice(&before);
    if (status == cudaSuccess && before != device) status = cudaSetDevice(device);
    else before = -1;  // nothing to restore
  }
  ~DeviceScope() {
    if (before >= 0) cudaSetDevice(before);
  }
  DeviceScope(const DeviceScope &) = delete;
  DeviceScope &operator=(const DeviceScope &) = delete;
};
inline int sm_count() {  // SM count of the CURRENT device
  static std::atomic<int> cache[kMaxDevices];
  const int dev = current_device();
  int v = cache[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cache[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}
// Opt a kernel into > 48 KB of dynamic shared memory, once per DEVICE (the attribute lives in the device's context;
// a process-wide flag would leave the second GPU of a process unconfigured).  Racing first calls are harmless: both
// set the same value.
template <auto Kernel>
inline cudaError_t ensure_dynamic_smem(int bytes) {
  static std::atomic<int> done[kMaxDevices];
  const int dev = current_device();
  if (done[dev].load(std::memory_order_acquire) >= bytes) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done[dev].store(bytes, std::memory_order_release);
  return e;
}

// ---- single-rounding float ops ---------------------------------------------------------------
// Decisions downstream (depth order, ceil of the radius, floor of the tile rect, alpha > 1/255) are
// discontinuous, so the geometry follows the reference's torch-CPU arithmetic operation by operation
// (SURVEY.md App. B.8).  These intrinsics are never contracted into FMAs by nvcc.
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float frcp(float a) { return __fdiv_rn(1.0f, a); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }

// order-preserving u32 image of an f32 (ascending)
__device__ __forceinline__ uint32_t float_to_key(float f) {
  uint32_t b = __float_as_uint(f);
  return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Tile grid.  tiles_x/tiles_y cover every pixel; *_eff is the part the reference can ever write:
// pixel column W-1 and row H-1 are never written, so a tile whose first column is >= W-1
// can hold no Gaussian.
struct TileGrid {
  int tiles_x, tiles_y, tiles_x_eff, tiles_y_eff;
};
__host__ __device__ __forceinline__ TileGrid make_grid(int width, int height) {
  TileGrid g;
  g.tiles_x = ceil_div(width, GSR_TILE);
  g.tiles_y = ceil_div(height, GSR_TILE);
  g.tiles_x_eff = ceil_div(width - 1, GSR_TILE);
  g.tiles_y_eff = ceil_div(height - 1, GSR_TILE);
  return g;
}

// ---- one frame's workspace, carved (api.cu) ----------------------------------------------------------------------
struct FrameBuffers {
  uint32_t *depth_key, *tiles_touched, *tile_span, *order, *blend_rank, *depth_hist, *tile_ranges, *tile_gids;
  float *rec;
  void *bin, *zero_a;
  size_t bin_bytes, zero_a_bytes, zero_b_bytes;
};
int frame_buffers(int64_t n, int width, int height, int64_t max_dup, void *workspace, size_t workspace_bytes,
                  FrameBuffers *fb);
int frame_begin(const FrameBuffers &fb, cudaStream_t stream);
int frame_finish(const FrameBuffers &fb, int64_t n, const gsr_camera *cam, int64_t max_dup, float *image,
                 gsr_frame_info *info_dev, cudaStream_t stream, float *final_T = nullptr, uint32_t *n_contrib = nullptr,
                 uint32_t first_rank = 0, uint32_t last_rank = 0xFFFFFFFFu, int32_t tile_row_begin = 0,
                 int32_t tile_row_end = -1);

// ---- stage launchers shared between translation units ------------------------------------------
struct SortScratch {   // carved from a caller workspace
  uint32_t *hist;      // [4][256] global digit counts
  uint32_t *status;    // [passes][tiles][256] decoupled look-back state
  uint32_t *tickets;   // [4] dynamic block ids
  size_t status_stride;  // elements per pass
};
size_t sort_scratch_bytes(int64_t n);
// zeroes and carves; returns GSR_OK or error
int sort_scratch_carve(void *workspace, size_t bytes, int64_t n, SortScratch *out);
int preprocess_raw(const gsr_gaussians *g, const gsr_camera *cam, int32_t row_begin, int32_t row_end,
                   uint32_t *depth_key, uint32_t *tiles_touched, uint32_t *tile_span, float *rec, float *cov3d, float *cov2d,
                   int32_t *tile_rect_raw, float *mean2d_all, uint32_t *depth_hist, cudaStream_t stream, bool hist_zeroed);
int launch_preprocess_prepared(const void *prepared, int64_t n, const gsr_camera *cam, int row_begin, int row_end,
                               uint32_t *depth_key, uint32_t *tiles_touched, uint32_t *tile_span, float *rec,
                               uint32_t *depth_hist, cudaStream_t stream);
int launch_sort_pairs(const uint32_t *keys_in, const uint32_t *values_in, uint32_t *keys_out, uint32_t *values_out,
                      uint32_t *keys_tmp, uint32_t *values_tmp, int64_t n, const uint32_t *n_dev, int begin_bit,
                      int end_bit, const SortScratch &s, cudaStream_t stream, bool skip_final_keys,
                      const uint32_t *ready_hist, const uint32_t *live_flags = nullptr, uint32_t *n_live_out = nullptr,
                      bool scratch_zeroed = false, bool adaptive = false);
// Where an ADAPTIVE sort (launch_sort_pairs(..., adaptive = true)) left its result: 0 = (keys_tmp, values_tmp),
// 1 = (keys_out, values_out).  hist = the digit histogram rows of all `passes` passes, n_total = number of sorted keys.
// Called by all threads of a CTA of at least 256 threads.
__device__ __forceinline__ int sorted_buffer_index(const uint32_t *__restrict__ hist, int passes, uint32_t n_total) {
  int executed = 1;
  for (int q = 1; q < passes; ++q) {
    const uint32_t h = threadIdx.x < 256 ? __ldg(hist + q * 256 + threadIdx.x) : 0u;
    executed += __syncthreads_or(h == n_total) == 0;
  }
  return (executed & 1) ? 0 : 1;
}
// gsr_bin with the option of a workspace whose zero-initialised head (bin_zero_bytes) the caller has already cleared
int bin_impl(int64_t n, const uint32_t *depth_key, const uint32_t *depth_hist, const uint32_t *tiles_touched,
             const uint32_t *tile_span, int32_t width, int32_t height, int64_t max_dup, uint32_t *depth_order,
             uint32_t *blend_rank, uint32_t *tile_gids, uint32_t *tile_ranges, gsr_frame_info *info_dev, void *workspace,
             size_t workspace_bytes, cudaStream_t stream, bool zeroed);
size_t bin_zero_bytes(int64_t n, int64_t max_dup);

}  // namespace gsr
