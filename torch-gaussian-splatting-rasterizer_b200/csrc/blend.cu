// blend.cu - per-16x16-tile front-to-back alpha blend.
//
// Replaces the reference's one-Gaussian-at-a-time loop (/root/reference/rasterize.py:436-446) and its body
// rasterize_gaussian (:255-305).  The reference visits Gaussians in global depth order and, for each, the
// pixels of its tile-aligned bbox; a pixel therefore sees exactly the Gaussians whose tile rect contains its
// tile, in depth order - which is the tile list built by binning.cu.  Per pixel and Gaussian:
//     d      = mean2d - pixel                                   (:279, integer pixel coordinates, Q10)
//     power  = -0.5 (a dx^2 + c dy^2) - b dx dy                 (:280-283)
//     alpha  = min(opacity * exp(power), 0.99)                  (:285-288)
//     keep   = alpha > 1/255  and  power <= 0                   (:291)
//     C     += (alpha * rgb) * T ;  T *= (1 - alpha)            (:295-303)
// The arithmetic is kept in the reference's operation order (single-rounded ops, gsr_common.cuh) because the
// keep test is discontinuous.  The reference never terminates a pixel; here a pixel stops AFTER the blend that
// takes T below 1e-4 (what is left can change the colour by < 1e-4; stopping before that blend, as 3DGS does,
// measured 4.7e-3 - SURVEY.md Q12).  Pixel column W-1 and row H-1 are never written by the reference (Q8).
//
// Work decomposition: one CTA per tile, one thread per pixel; the tile's list is staged through shared memory
// in batches of 256 records (each thread gathers one 48-byte record), all threads then walk the batch with
// broadcast shared-memory reads.  A CTA leaves as soon as every pixel has terminated.
#include "gsr_common.cuh"

namespace gsr {

constexpr int kBlendThreads = GSR_TILE * GSR_TILE;
constexpr float kMinAlpha = (float)(1.0 / 255.0);   // MIN_ALPHA, rasterize.py:38 (compared in f32 like torch does)
constexpr float kMaxDensity = 0.99f;                // MAX_GAUSSIAN_DENSITY, rasterize.py:36
constexpr float kStopT = 1e-4f;

template <bool COUNT>
__global__ void __launch_bounds__(kBlendThreads)
blend_kernel(const float4 *__restrict__ rec, const uint32_t *__restrict__ gids, const uint2 *__restrict__ ranges,
             int width, int height, int tiles_x, int row_begin, const uint32_t *__restrict__ blend_rank,
             uint32_t first_rank, uint32_t last_rank, float *__restrict__ image, float *__restrict__ final_T,
             uint32_t *__restrict__ n_contrib, gsr_frame_info *info) {
  __shared__ float4 s_a[kBlendThreads];  // mean.x mean.y conic.a conic.c
  __shared__ float4 s_b[kBlendThreads];  // conic.b opacity r g
  __shared__ float s_c[kBlendThreads];   // b

  const int tid = threadIdx.x;
  const int tile_x = blockIdx.x, tile_y = blockIdx.y + row_begin;
  const int px = tile_x * GSR_TILE + (tid & (GSR_TILE - 1));
  const int py = tile_y * GSR_TILE + (tid / GSR_TILE);
  const bool inside = px < width && py < height;
  const bool writable = px < width - 1 && py < height - 1;
  const uint2 range = ranges[tile_y * tiles_x + tile_x];
  const float fx = (float)px, fy = (float)py;

  float T = 1.0f, C0 = 0.0f, C1 = 0.0f, C2 = 0.0f;
  bool done = !writable;
  uint32_t last = 0, seen = 0;
  unsigned long long my_evals = 0;

  for (uint32_t start = range.x; start < range.y; start += kBlendThreads) {
    if (__syncthreads_count(!done) == 0) break;
    const uint32_t idx = start + tid;
    if (idx < range.y) {
      const uint32_t g = __ldg(gids + idx);
      float4 a = __ldg(rec + (size_t)g * 3), b = __ldg(rec + (size_t)g * 3 + 1);
      const float c = __ldg((const float *)(rec + (size_t)g * 3 + 2));
      if (blend_rank) {
        const uint32_t r = __ldg(blend_rank + g);
        if (r < first_rank || r >= last_rank) b.y = 0.0f;  // opacity 0 => alpha 0 => never kept
      }
      s_a[tid] = a;
      s_b[tid] = b;
      s_c[tid] = c;
    }
    __syncthreads();
    const int count = min((uint32_t)kBlendThreads, range.y - start);
    if (!done) {
      for (int j = 0; j < count; ++j) {
        const float4 a = s_a[j];
        const float4 b = s_b[j];
        ++seen;
        if (COUNT) ++my_evals;
        const float dx = fsub(a.x, fx), dy = fsub(a.y, fy);
        const float power = fsub(fmul(-0.5f, fadd(fmul(a.z, fmul(dx, dx)), fmul(a.w, fmul(dy, dy)))),
                                 fmul(fmul(b.x, dx), dy));
        if (power > 0.0f) continue;
        const float alpha = fminf(fmul(b.y, expf(power)), kMaxDensity);
        if (!(alpha > kMinAlpha)) continue;
        C0 = fadd(C0, fmul(fmul(alpha, b.z), T));
        C1 = fadd(C1, fmul(fmul(alpha, b.w), T));
        C2 = fadd(C2, fmul(fmul(alpha, s_c[j]), T));
        T = fmul(T, fsub(1.0f, alpha));
        last = seen;
        if (T < kStopT) {
          done = true;
          break;
        }
      }
    }
  }
  if (COUNT) {
    // warp-granular count: a warp issues as many evaluation slots as its busiest lane consumed
    unsigned long long mx = my_evals;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, mx, o);
      mx = other > mx ? other : mx;
    }
    unsigned long long sum = my_evals;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    if ((tid & 31) == 0 && info) {
      atomicAdd((unsigned long long *)&info->warp_evals, mx * 32ull);
      atomicAdd((unsigned long long *)&info->pixel_evals, sum);
    }
  }
  if (inside) {
    const size_t p = (size_t)py * width + px;
    image[p * 3 + 0] = C0;
    image[p * 3 + 1] = C1;
    image[p * 3 + 2] = C2;
    if (final_T) final_T[p] = T;
    if (n_contrib) n_contrib[p] = last;
  }
}

// rasterize_gaussian (:255-305): one Gaussian into the x-major screen, one thread per bbox pixel
__global__ void blend_one_kernel(int x0, int y0, int x1, int y1, float mx, float my, float a, float c, float b,
                                 float r, float g, float bl, float opacity, int height, float *screen, float *T) {
  const int w = x1 - x0, h = y1 - y0;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w * h) return;
  const int x = x0 + i / h, y = y0 + i % h;
  const float dx = fsub(mx, (float)x), dy = fsub(my, (float)y);
  const float power = fsub(fmul(-0.5f, fadd(fmul(a, fmul(dx, dx)), fmul(c, fmul(dy, dy)))), fmul(fmul(b, dx), dy));
  const float alpha = fminf(fmul(opacity, expf(power)), kMaxDensity);
  if (!(alpha > kMinAlpha && power <= 0.0f)) return;
  const size_t p = (size_t)x * height + y;
  const float t = T[p];
  screen[p * 3 + 0] = fadd(screen[p * 3 + 0], fmul(fmul(alpha, r), t));
  screen[p * 3 + 1] = fadd(screen[p * 3 + 1], fmul(fmul(alpha, g), t));
  screen[p * 3 + 2] = fadd(screen[p * 3 + 2], fmul(fmul(alpha, bl), t));
  T[p] = fmul(t, fsub(1.0f, alpha));
}

}  // namespace gsr

using namespace gsr;

extern "C" int gsr_blend_fwd(const float *rec, const uint32_t *tile_gids, const uint32_t *tile_ranges, int32_t width,
                             int32_t height, int32_t tile_row_begin, int32_t tile_row_end, const uint32_t *blend_rank,
                             uint32_t first_rank, uint32_t last_rank, float *image, float *final_T,
                             uint32_t *n_contrib, gsr_frame_info *info_dev, void *stream) {
  GSR_REQUIRE(rec && tile_gids && tile_ranges && image, "null pointer");
  GSR_REQUIRE(width > 0 && height > 0, "bad image size");
  GSR_REQUIRE(((uintptr_t)rec & 15) == 0, "rec must be 16-byte aligned");
  const TileGrid tg = make_grid(width, height);
  if (tile_row_begin < 0) tile_row_begin = 0;
  if (tile_row_end > tg.tiles_y || tile_row_end < 0) tile_row_end = tg.tiles_y;
  if (tile_row_end <= tile_row_begin) return GSR_OK;
  if (first_rank == 0 && last_rank == 0xFFFFFFFFu) blend_rank = nullptr;
  dim3 grid(tg.tiles_x, tile_row_end - tile_row_begin);
  if (info_dev)
    blend_kernel<true><<<grid, kBlendThreads, 0, (cudaStream_t)stream>>>(
        (const float4 *)rec, tile_gids, (const uint2 *)tile_ranges, width, height, tg.tiles_x, tile_row_begin,
        blend_rank, first_rank, last_rank, image, final_T, n_contrib, info_dev);
  else
    blend_kernel<false><<<grid, kBlendThreads, 0, (cudaStream_t)stream>>>(
        (const float4 *)rec, tile_gids, (const uint2 *)tile_ranges, width, height, tg.tiles_x, tile_row_begin,
        blend_rank, first_rank, last_rank, image, final_T, n_contrib, info_dev);
  GSR_LAUNCH_CHECK();
  return GSR_OK;
}

extern "C" int gsr_blend_one(const int64_t *bbox4_host, const float *mean2d_host2, const float *conic_host3,
                             const float *rgb_host3, float opacity, int32_t width, int32_t height, float *screen,
                             float *opacity_buffer, void *stream) {
  GSR_REQUIRE(bbox4_host && mean2d_host2 && conic_host3 && rgb_host3 && screen && opacity_buffer, "null pointer");
  const int x0 = (int)bbox4_host[0], y0 = (int)bbox4_host[1], x1 = (int)bbox4_host[2], y1 = (int)bbox4_host[3];
  GSR_REQUIRE(x0 >= 0 && y0 >= 0 && x1 <= width && y1 <= height, "bbox outside the screen");
  if (x1 <= x0 || y1 <= y0) return GSR_OK;
  const int n = (x1 - x0) * (y1 - y0);
  blend_one_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      x0, y0, x1, y1, mean2d_host2[0], mean2d_host2[1], conic_host3[0], conic_host3[1], conic_host3[2], rgb_host3[0],
      rgb_host3[1], rgb_host3[2], opacity, height, screen, opacity_buffer);
  GSR_LAUNCH_CHECK();
  return GSR_OK;
}
