// api.cu - C-ABI glue: error strings, device info, whole-frame driver and the host-buffer entry point.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "gsr_common.cuh"

namespace gsr {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof g_error, fmt, ap);
  va_end(ap);
}

// One frame's workspace.  Offsets of everything a backward pass replays (tiles_touched, rec, tile_ranges, tile_gids)
// depend on n and the image size only - NOT on max_dup - so a forward and its backward agree on them whatever pair
// capacity each was called with; buffers whose SIZE depends on max_dup come last.  The zero-initialised buffers are
// two contiguous runs: depth_hist | tile_ranges, and the head of the binning workspace.
struct FrameLayout {
  size_t depth_key, tiles_touched, tile_span, rec, order, blend_rank, depth_hist, tile_ranges, tile_gids, bin, total;
  size_t bin_bytes, zero_a_bytes, zero_b_bytes;
};

static int frame_layout(int64_t n, int width, int height, int64_t max_dup, FrameLayout *L) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t at = off;
    off = align_up(off + bytes, 256);
    return at;
  };
  const size_t n1 = (size_t)(n > 0 ? n : 1), d1 = (size_t)(max_dup > 0 ? max_dup : 1);
  const TileGrid tg = make_grid(width, height);
  L->depth_key = take(4 * n1);
  L->tiles_touched = take(4 * n1);
  L->tile_span = take(16 * n1);
  L->rec = take(sizeof(float) * GSR_REC_FLOATS * n1);
  L->order = take(4 * n1);
  L->blend_rank = take(4 * n1);
  L->depth_hist = take(4 * 1024);
  L->tile_ranges = take(8 * (size_t)tg.tiles_x * tg.tiles_y);
  L->tile_gids = take(4 * d1);
  L->zero_a_bytes = L->tile_gids - L->depth_hist;
  int rc = gsr_bin_workspace_bytes(n, width, height, max_dup, &L->bin_bytes);
  if (rc) return rc;
  L->bin = take(L->bin_bytes);
  L->zero_b_bytes = bin_zero_bytes(n, max_dup);
  L->total = off;
  return GSR_OK;
}

// FP32 FMA throughput probe: the denominator of the blend kernel's roofline (MEASURED_PEAKS.json has no FP32 figure).
__global__ void __launch_bounds__(256) fma_probe_kernel(float *out, int iters, float a, float b) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.0f, x2 = x0 + 2.0f, x3 = x0 + 3.0f, x4 = x0 + 4.0f, x5 = x0 + 5.0f,
        x6 = x0 + 6.0f, x7 = x0 + 7.0f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  const float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456f) out[0] = s;  // never true; keeps the chains alive
}

}  // namespace gsr

using namespace gsr;

extern "C" int gsr_fp32_fma_peak(double *tflops, double *sm_mhz_effective) {
  GSR_REQUIRE(tflops, "null pointer");
  int dev = 0, sms = 0, khz = 0;
  GSR_CHECK_CUDA(cudaGetDevice(&dev));
  GSR_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  GSR_CHECK_CUDA(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  (void)khz;
  float *d = nullptr;
  GSR_CHECK_CUDA(cudaMalloc(&d, 16));
  cudaEvent_t e0, e1;
  GSR_CHECK_CUDA(cudaEventCreate(&e0));
  GSR_CHECK_CUDA(cudaEventCreate(&e1));
  const int iters = 4096, blocks = sms * 8;
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    GSR_CHECK_CUDA(cudaEventRecord(e0));
    fma_probe_kernel<<<blocks, 256>>>(d, iters, 0.999f, 0.001f);
    GSR_CHECK_CUDA(cudaEventRecord(e1));
    GSR_CHECK_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    GSR_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flop = 2.0 * 64.0 * iters * 256.0 * blocks;
    const double t = flop / (ms * 1e-3) / 1e12;
    if (rep > 0 && t > best) best = t;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops = best;
  if (sm_mhz_effective) *sm_mhz_effective = best * 1e12 / (2.0 * 128.0 * sms) / 1e6;
  return GSR_OK;
}

extern "C" const char *gsr_last_error(void) { return g_error; }
extern "C" int gsr_abi_version(void) { return GSR_ABI_VERSION; }

extern "C" int gsr_device_info(int *sm_count, int *cc_major, int *cc_minor, size_t *l2_bytes) {
  int dev = 0;
  GSR_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  GSR_CHECK_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (l2_bytes) *l2_bytes = (size_t)p.l2CacheSize;
  return GSR_OK;
}

extern "C" int gsr_render_workspace_bytes(int64_t n, int32_t width, int32_t height, int64_t max_dup, size_t *bytes) {
  GSR_REQUIRE(bytes && n >= 0 && width > 0 && height > 0 && max_dup > 0, "bad arguments");
  FrameLayout L;
  int rc = frame_layout(n, width, height, max_dup, &L);
  if (rc) return rc;
  *bytes = L.total;
  return GSR_OK;
}

// ---- a frame in three steps (shared by the single-call entry points and by the host pipeline, host.cu) ----
int gsr::frame_buffers(int64_t n, int width, int height, int64_t max_dup, void *workspace, size_t workspace_bytes,
                       FrameBuffers *fb) {
  GSR_REQUIRE(workspace && fb, "null pointer");
  GSR_REQUIRE(max_dup > 0, "max_dup must be positive");
  FrameLayout L;
  int rc = frame_layout(n, width, height, max_dup, &L);
  if (rc) return rc;
  if (workspace_bytes < L.total || ((uintptr_t)workspace & 255)) {
    set_error("render workspace too small or not 256-byte aligned (%zu < %zu)", workspace_bytes, L.total);
    return GSR_ERR_WORKSPACE_TOO_SMALL;
  }
  char *ws = (char *)workspace;
  fb->depth_key = (uint32_t *)(ws + L.depth_key);
  fb->tiles_touched = (uint32_t *)(ws + L.tiles_touched);
  fb->tile_span = (uint32_t *)(ws + L.tile_span);
  fb->rec = (float *)(ws + L.rec);
  fb->order = (uint32_t *)(ws + L.order);
  fb->blend_rank = (uint32_t *)(ws + L.blend_rank);
  fb->depth_hist = (uint32_t *)(ws + L.depth_hist);
  fb->tile_ranges = (uint32_t *)(ws + L.tile_ranges);
  fb->tile_gids = (uint32_t *)(ws + L.tile_gids);
  fb->bin = ws + L.bin;
  fb->bin_bytes = L.bin_bytes;
  fb->zero_a = ws + L.depth_hist;
  fb->zero_a_bytes = L.zero_a_bytes;
  fb->zero_b_bytes = L.zero_b_bytes;
  return GSR_OK;
}

// zero-initialised parts of the workspace: depth_hist | tile_ranges, and the head of the binning workspace
int gsr::frame_begin(const FrameBuffers &fb, cudaStream_t stream) {
  GSR_CHECK_CUDA(cudaMemsetAsync(fb.zero_a, 0, fb.zero_a_bytes, stream));
  GSR_CHECK_CUDA(cudaMemsetAsync(fb.bin, 0, fb.zero_b_bytes, stream));
  return GSR_OK;
}

// binning + blend of a frame whose per-Gaussian outputs (depth_key .. rec, depth_hist) are in place
int gsr::frame_finish(const FrameBuffers &fb, int64_t n, const gsr_camera *cam, int64_t max_dup, float *image,
                      gsr_frame_info *info_dev, cudaStream_t stream, float *final_T, uint32_t *n_contrib,
                      uint32_t first_rank, uint32_t last_rank, int32_t tile_row_begin, int32_t tile_row_end) {
  const bool windowed = !(first_rank == 0 && last_rank == 0xFFFFFFFFu);
  int rc = bin_impl(n, fb.depth_key, fb.depth_hist, fb.tiles_touched, fb.tile_span, cam->width, cam->height, max_dup,
                    fb.order, windowed ? fb.blend_rank : nullptr, fb.tile_gids, fb.tile_ranges, info_dev, fb.bin,
                    fb.bin_bytes, stream, true);
  if (rc) return rc;
  return gsr_blend_fwd(fb.rec, fb.tile_gids, fb.tile_ranges, cam->width, cam->height, tile_row_begin, tile_row_end,
                       windowed ? fb.blend_rank : nullptr, first_rank, last_rank, image, final_T, n_contrib, nullptr,
                       stream);
}

// One frame on `stream`: 2 memsets + 10 kernels, no host synchronisation.  Exactly one of g / prepared is given.
// info_dev is NOT cleared here: status / n_dup_peak are sticky across frames (include/gsr_b200.h).
static int render_frame(const gsr_gaussians *g, const void *prepared, int64_t n, const gsr_camera *cam, int64_t max_dup,
                        void *workspace, size_t workspace_bytes, float *image, float *final_T, uint32_t *n_contrib,
                        uint32_t first_rank, uint32_t last_rank, int32_t tile_row_begin, int32_t tile_row_end,
                        gsr_frame_info *info_dev, cudaStream_t stream) {
  GSR_REQUIRE((g || prepared) && cam && workspace && image, "null pointer");
  FrameBuffers fb;
  int rc = frame_buffers(n, cam->width, cam->height, max_dup, workspace, workspace_bytes, &fb);
  if (rc) return rc;
  const TileGrid tg = make_grid(cam->width, cam->height);
  if (tile_row_begin < 0) tile_row_begin = 0;
  if (tile_row_end < 0 || tile_row_end > tg.tiles_y) tile_row_end = tg.tiles_y;
  rc = frame_begin(fb, stream);
  if (rc) return rc;
  if (n > 0) {
    if (prepared) {
      GSR_REQUIRE(cam->sh_degree >= 0 && cam->sh_degree <= 3, "sh_degree must be 0..3");
      GSR_REQUIRE(cam->width > 0 && cam->height > 0 && cam->width <= 65535 * GSR_TILE && cam->height <= 65535 * GSR_TILE,
                  "image size out of range");
      GSR_REQUIRE(((uintptr_t)prepared & 255) == 0, "prepared model must be 256-byte aligned");
      rc = launch_preprocess_prepared(prepared, n, cam, tile_row_begin, tile_row_end, fb.depth_key, fb.tiles_touched,
                                      fb.tile_span, fb.rec, fb.depth_hist, stream);
    } else {
      rc = preprocess_raw(g, cam, tile_row_begin, tile_row_end, fb.depth_key, fb.tiles_touched, fb.tile_span, fb.rec,
                          nullptr, nullptr, nullptr, nullptr, fb.depth_hist, stream, true);
    }
    if (rc) return rc;
  }
  return frame_finish(fb, n, cam, max_dup, image, info_dev, stream, final_T, n_contrib, first_rank, last_rank,
                      tile_row_begin, tile_row_end);
}

extern "C" int gsr_render_fwd(const gsr_gaussians *g, const gsr_camera *cam, int64_t max_dup, void *workspace,
                              size_t workspace_bytes, float *image, float *final_T, uint32_t *n_contrib,
                              uint32_t first_rank, uint32_t last_rank, int32_t tile_row_begin, int32_t tile_row_end,
                              gsr_frame_info *info_dev, void *stream) {
  GSR_REQUIRE(g, "null gaussians");
  return render_frame(g, nullptr, g->n, cam, max_dup, workspace, workspace_bytes, image, final_T, n_contrib, first_rank,
                      last_rank, tile_row_begin, tile_row_end, info_dev, (cudaStream_t)stream);
}

extern "C" int gsr_render_prepared_fwd(const void *prepared, int64_t n, const gsr_camera *cam, int64_t max_dup,
                                       void *workspace, size_t workspace_bytes, float *image, float *final_T,
                                       uint32_t *n_contrib, uint32_t first_rank, uint32_t last_rank,
                                       int32_t tile_row_begin, int32_t tile_row_end, gsr_frame_info *info_dev,
                                       void *stream) {
  GSR_REQUIRE(prepared && n >= 0 && n < (1ll << 31), "null prepared model / n out of range");
  return render_frame(nullptr, prepared, n, cam, max_dup, workspace, workspace_bytes, image, final_T, n_contrib,
                      first_rank, last_rank, tile_row_begin, tile_row_end, info_dev, (cudaStream_t)stream);
}

// A batch of views of one resident model: view v is enqueued on streams[v % n_lanes] with workspaces[v % n_lanes] and
// info_dev[v % n_lanes] (frames of one lane are stream-ordered, lanes overlap: the latency-bound binning kernels of one
// frame run under the issue-bound blend of another).  The caller owns streams and workspaces and joins the streams.
extern "C" int gsr_render_views_fwd(const void *prepared, int64_t n, const gsr_camera *cams_host, int32_t n_views,
                                    int64_t max_dup, void *const *workspaces, size_t workspace_bytes, int32_t n_lanes,
                                    void *const *streams, float *images, int64_t image_stride_floats,
                                    gsr_frame_info *info_dev) {
  GSR_REQUIRE(prepared && cams_host && workspaces && streams && images, "null pointer");
  GSR_REQUIRE(n_views >= 0 && n_lanes >= 1, "bad view / lane count");
  for (int32_t v = 0; v < n_views; ++v) {
    const int lane = v % n_lanes;
    const gsr_camera *cam = cams_host + v;
    GSR_REQUIRE(image_stride_floats >= (int64_t)3 * cam->width * cam->height, "image stride smaller than a frame");
    int rc = render_frame(nullptr, prepared, n, cam, max_dup, workspaces[lane], workspace_bytes,
                          images + (size_t)v * image_stride_floats, nullptr, nullptr, 0, 0xFFFFFFFFu, 0, -1,
                          info_dev ? info_dev + lane : nullptr, (cudaStream_t)streams[lane]);
    if (rc) return rc;
  }
  return GSR_OK;
}

// Gradients of one frame.  Must follow gsr_render_fwd of the SAME frame with the SAME workspace (the records and
// tile lists of the forward pass are replayed, not rebuilt).
extern "C" int gsr_render_bwd(const gsr_gaussians *g, const gsr_camera *cam, int64_t max_dup, void *workspace,
                              size_t workspace_bytes, const float *final_T, const uint32_t *n_contrib,
                              const float *dL_dimage, float *grad_rec, float *d_xyz, float *d_log_scales, float *d_rots,
                              float *d_opacity_logit, float *d_sh, int64_t dsh_sn, int64_t dsh_sk, int64_t dsh_sc,
                              int32_t tile_row_begin, int32_t tile_row_end, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GSR_REQUIRE(g && cam && workspace && final_T && n_contrib && dL_dimage && grad_rec, "null pointer");
  FrameLayout L;
  int rc = frame_layout(g->n, cam->width, cam->height, max_dup, &L);
  if (rc) return rc;
  if (workspace_bytes < L.total || ((uintptr_t)workspace & 255)) {
    set_error("render workspace too small or not 256-byte aligned (%zu < %zu)", workspace_bytes, L.total);
    return GSR_ERR_WORKSPACE_TOO_SMALL;
  }
  char *ws = (char *)workspace;
  GSR_CHECK_CUDA(cudaMemsetAsync(grad_rec, 0, sizeof(float) * 9 * (size_t)(g->n > 0 ? g->n : 1), stream));
  rc = gsr_blend_bwd((const float *)(ws + L.rec), (const uint32_t *)(ws + L.tile_gids),
                     (const uint32_t *)(ws + L.tile_ranges), cam->width, cam->height, tile_row_begin, tile_row_end,
                     final_T, n_contrib, dL_dimage, grad_rec, stream);
  if (rc) return rc;
  return gsr_preprocess_bwd(g, cam, (const uint32_t *)(ws + L.tiles_touched), grad_rec, d_xyz, d_log_scales, d_rots,
                            d_opacity_logit, d_sh, dsh_sn, dsh_sk, dsh_sc, stream);
}

// Host-buffer frame: stands where the body of the reference's run_rasterization stands (rasterize.py:354-452).
// Convenience entry (allocates and frees its device buffers on every call); a caller that renders more than one frame
// keeps the model resident instead (gsr_prepare + gsr_render_prepared_fwd).
extern "C" int gsr_render_host(const gsr_gaussians *gh, const gsr_camera *cam, int64_t max_dup, float *image_host,
                               gsr_frame_info *info_host) {
  GSR_REQUIRE(gh && cam && image_host, "null pointer");
  GSR_REQUIRE(gh->n >= 0 && gh->n < (1ll << 31) && max_dup > 0, "bad sizes");
  GSR_REQUIRE(cam->width > 0 && cam->height > 0, "bad image size");
  const int64_t n = gh->n;
  const size_t img_bytes = sizeof(float) * 3 * (size_t)cam->width * cam->height;
  if (info_host) memset(info_host, 0, sizeof *info_host);
  if (n == 0) {  // nothing to upload, nothing to blend: the reference's screen stays zero
    memset(image_host, 0, img_bytes);
    return GSR_OK;
  }
  GSR_REQUIRE(gh->xyz && gh->log_scales && gh->rots && gh->opacity_logit && gh->sh, "null attribute pointer");
  // Every attribute is uploaded as the smallest host span that contains its strided view (strides are kept), so all
  // spans are computed and validated BEFORE anything is allocated or copied.
  struct Attr {
    const float *src;
    int64_t sn, s1, c1, s2, c2, span;
  } attr[5] = {{gh->xyz, gh->xyz_sn, gh->xyz_sc, 3, 0, 1, 0},
               {gh->log_scales, gh->scale_sn, gh->scale_sc, 3, 0, 1, 0},
               {gh->rots, gh->rot_sn, gh->rot_sc, 4, 0, 1, 0},
               {gh->opacity_logit, gh->opacity_sn, 0, 1, 0, 1, 0},
               {gh->sh, gh->sh_sn, gh->sh_sk, 16, gh->sh_sc, 3, 0}};
  const int64_t dense = n * 59;
  int64_t total = 0;
  for (Attr &a : attr) {
    GSR_REQUIRE(a.sn >= 0 && a.s1 >= 0 && a.s2 >= 0, "negative strides are not supported by the host-buffer entry");
    a.span = 1 + (n - 1) * a.sn + (a.c1 - 1) * a.s1 + (a.c2 - 1) * a.s2;
    a.span = (a.span + 3) & ~(int64_t)3;  // keep every attribute 16-byte aligned on the device
    total += a.span;
  }
  if (total > 2 * dense + 64) {  // column views of a wider table etc.: refuse rather than upload the whole table 5 times
    set_error("strided host views span %lld floats for %lld parameters; pass contiguous or planar arrays",
              (long long)total, (long long)dense);
    return GSR_ERR_INVALID_ARGUMENT;
  }
  size_t ws_bytes = 0;
  int rc = gsr_render_workspace_bytes(n, cam->width, cam->height, max_dup, &ws_bytes);
  if (rc) return rc;
  float *d_params = nullptr, *d_image = nullptr;
  void *d_ws = nullptr;
  gsr_frame_info *d_info = nullptr;
  cudaStream_t stream = nullptr;
  auto cleanup = [&]() {
    if (d_params) cudaFree(d_params);
    if (d_image) cudaFree(d_image);
    if (d_ws) cudaFree(d_ws);
    if (d_info) cudaFree(d_info);
    if (stream) cudaStreamDestroy(stream);
  };
#define GSR_HOST_CHECK(expr)                                                                    \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      set_error("%s failed: %s", #expr, cudaGetErrorString(e__));                               \
      cleanup();                                                                                \
      return GSR_ERR_CUDA;                                                                      \
    }                                                                                           \
  } while (0)
  GSR_HOST_CHECK(cudaStreamCreate(&stream));
  GSR_HOST_CHECK(cudaMalloc(&d_params, sizeof(float) * (size_t)total));
  GSR_HOST_CHECK(cudaMalloc(&d_image, img_bytes));
  GSR_HOST_CHECK(cudaMalloc(&d_ws, ws_bytes));
  GSR_HOST_CHECK(cudaMalloc(&d_info, sizeof(gsr_frame_info)));
  GSR_HOST_CHECK(cudaMemsetAsync(d_info, 0, sizeof(gsr_frame_info), stream));
  gsr_gaussians gd = *gh;
  const float **dst[5] = {&gd.xyz, &gd.log_scales, &gd.rots, &gd.opacity_logit, &gd.sh};
  float *cursor = d_params;
  for (int k = 0; k < 5; ++k) {
    const int64_t exact = 1 + (n - 1) * attr[k].sn + (attr[k].c1 - 1) * attr[k].s1 + (attr[k].c2 - 1) * attr[k].s2;
    GSR_HOST_CHECK(cudaMemcpyAsync(cursor, attr[k].src, sizeof(float) * (size_t)exact, cudaMemcpyHostToDevice, stream));
    *dst[k] = cursor;
    cursor += attr[k].span;
  }
  rc = gsr_render_fwd(&gd, cam, max_dup, d_ws, ws_bytes, d_image, nullptr, nullptr, 0, 0xFFFFFFFFu, 0, -1, d_info, stream);
  if (rc) {
    cleanup();
    return rc;
  }
  GSR_HOST_CHECK(cudaMemcpyAsync(image_host, d_image, img_bytes, cudaMemcpyDeviceToHost, stream));
  gsr_frame_info info;
  GSR_HOST_CHECK(cudaMemcpyAsync(&info, d_info, sizeof info, cudaMemcpyDeviceToHost, stream));
  GSR_HOST_CHECK(cudaStreamSynchronize(stream));
  if (info_host) *info_host = info;
  cleanup();
  if (info.status != GSR_OK) {
    set_error("binning produced %u (Gaussian, tile) pairs but max_dup is %lld", info.n_dup, (long long)max_dup);
    return (int)info.status;
  }
  return GSR_OK;
}
