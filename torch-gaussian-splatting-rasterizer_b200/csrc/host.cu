// host.cu - host-buffer frames as a pipeline: gsr_host_renderer.
//
// Stands where the body of the reference's run_rasterization stands (/root/reference/rasterize.py:354-452) for a caller
// whose Gaussians live in HOST memory (the reference's own situation: it builds its tensors from a PLY on the CPU and
// moves them with .to(device), :353-358).  A frame from host buffers is bounded by PCIe: 59 floats per Gaussian up
// (708 MB at 3 M), one image down.  What can be hidden is everything else:
//   * the upload is cut into chunks of Gaussians; as soon as a chunk has landed the per-Gaussian preprocess kernel runs
//     on it (on another stream) while the next chunk is still on the bus;
//   * binning + blending start when the last chunk is done, the image goes down on a third stream;
//   * `depth` frames are in flight (each with its own device buffers), so the upload of frame k+1 overlaps the render
//     and the read-back of frame k - PCIe is full duplex.
// All device memory, streams and events are created once, by gsr_host_renderer_create.
#include <stdlib.h>
#include <string.h>

#include <new>

#include "gsr_common.cuh"

namespace gsr {

constexpr int kHostMaxDepth = 4;
constexpr int kHostChunks = 8;

struct HostSlot {
  float *d_params = nullptr;   // xyz | scales | rots | opacity | sh, dense, each 16-byte aligned
  float *d_image = nullptr;
  void *d_ws = nullptr;
  gsr_frame_info *d_info = nullptr;
  gsr_frame_info *h_info = nullptr;  // pinned
  cudaEvent_t chunk_up[kHostChunks] = {};
  cudaEvent_t params_free = nullptr;  // the frame's preprocess kernels have consumed d_params
  cudaEvent_t rendered = nullptr, done = nullptr;
  bool busy = false;
  int64_t max_dup = 0;
  int32_t ticket = -1;
};

}  // namespace gsr

struct gsr_host_renderer {
  int64_t n_max = 0, max_dup = 0;
  int32_t width = 0, height = 0, depth = 0, device = 0;
  size_t ws_bytes = 0, param_floats = 0;
  cudaStream_t up = nullptr, compute = nullptr, down = nullptr;
  gsr::HostSlot slot[gsr::kHostMaxDepth];
  int32_t next_ticket = 0;
};

using namespace gsr;

static size_t pad4(int64_t v) { return (size_t)((v + 3) & ~(int64_t)3); }

static void host_renderer_free(gsr_host_renderer *h) {
  if (!h) return;
  for (int s = 0; s < kHostMaxDepth; ++s) {
    HostSlot &sl = h->slot[s];
    if (sl.d_params) cudaFree(sl.d_params);
    if (sl.d_image) cudaFree(sl.d_image);
    if (sl.d_ws) cudaFree(sl.d_ws);
    if (sl.d_info) cudaFree(sl.d_info);
    if (sl.h_info) cudaFreeHost(sl.h_info);
    for (cudaEvent_t e : sl.chunk_up)
      if (e) cudaEventDestroy(e);
    if (sl.params_free) cudaEventDestroy(sl.params_free);
    if (sl.rendered) cudaEventDestroy(sl.rendered);
    if (sl.done) cudaEventDestroy(sl.done);
  }
  if (h->up) cudaStreamDestroy(h->up);
  if (h->compute) cudaStreamDestroy(h->compute);
  if (h->down) cudaStreamDestroy(h->down);
  delete h;
}

#define GSR_HR_CHECK(expr)                                                             \
  do {                                                                                 \
    cudaError_t e__ = (expr);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      host_renderer_free(h);                                                           \
      return GSR_ERR_CUDA;                                                             \
    }                                                                                  \
  } while (0)

extern "C" int gsr_host_renderer_create(int64_t n_max, int32_t width, int32_t height, int64_t max_dup, int32_t depth,
                                        gsr_host_renderer **out) {
  GSR_REQUIRE(out, "null pointer");
  *out = nullptr;
  GSR_REQUIRE(n_max >= 0 && n_max < (1ll << 31) && width > 0 && height > 0 && max_dup > 0, "bad sizes");
  GSR_REQUIRE(depth >= 1 && depth <= kHostMaxDepth, "depth must be 1..4 frames in flight");
  gsr_host_renderer *h = new (std::nothrow) gsr_host_renderer();
  GSR_REQUIRE(h, "out of host memory");
  h->n_max = n_max; h->max_dup = max_dup; h->width = width; h->height = height; h->depth = depth;
  GSR_HR_CHECK(cudaGetDevice(&h->device));
  int rc = gsr_render_workspace_bytes(n_max, width, height, max_dup, &h->ws_bytes);
  if (rc) {
    host_renderer_free(h);
    return rc;
  }
  const int64_t n1 = n_max > 0 ? n_max : 1;
  h->param_floats = pad4(3 * n1) + pad4(3 * n1) + pad4(4 * n1) + pad4(n1) + pad4(48 * n1);
  GSR_HR_CHECK(cudaStreamCreateWithFlags(&h->up, cudaStreamNonBlocking));
  GSR_HR_CHECK(cudaStreamCreateWithFlags(&h->compute, cudaStreamNonBlocking));
  GSR_HR_CHECK(cudaStreamCreateWithFlags(&h->down, cudaStreamNonBlocking));
  for (int s = 0; s < depth; ++s) {
    HostSlot &sl = h->slot[s];
    GSR_HR_CHECK(cudaMalloc(&sl.d_params, sizeof(float) * h->param_floats));
    GSR_HR_CHECK(cudaMalloc(&sl.d_image, sizeof(float) * 3 * (size_t)width * height));
    GSR_HR_CHECK(cudaMalloc(&sl.d_ws, h->ws_bytes));
    GSR_HR_CHECK(cudaMalloc(&sl.d_info, sizeof(gsr_frame_info)));
    GSR_HR_CHECK(cudaMemset(sl.d_info, 0, sizeof(gsr_frame_info)));
    GSR_HR_CHECK(cudaHostAlloc(&sl.h_info, sizeof(gsr_frame_info), cudaHostAllocDefault));
    for (cudaEvent_t &e : sl.chunk_up) GSR_HR_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    GSR_HR_CHECK(cudaEventCreateWithFlags(&sl.params_free, cudaEventDisableTiming));
    GSR_HR_CHECK(cudaEventCreateWithFlags(&sl.rendered, cudaEventDisableTiming));
    GSR_HR_CHECK(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    sl.max_dup = max_dup;
  }
  *out = h;
  return GSR_OK;
}
#undef GSR_HR_CHECK

extern "C" int gsr_host_renderer_destroy(gsr_host_renderer *h) {
  if (!h) return GSR_OK;
  DeviceScope scope(h->device);
  cudaDeviceSynchronize();
  host_renderer_free(h);
  return GSR_OK;
}

// How an attribute [n, inner] lies in host memory: dense rows (row stride == inner count, unit inner stride) can be cut
// into chunks of rows; the planar SH block [3,16,N] viewed as [N,16,3] (strides 1, N, 16 N) is a 48 x N matrix whose
// column ranges are chunks (one 2-D copy each).
static bool dense_rows(int64_t sn, int64_t sc, int64_t inner) { return sn == inner && (inner == 1 || sc == 1); }

extern "C" int gsr_host_renderer_submit(gsr_host_renderer *h, const gsr_gaussians *gh, const gsr_camera *cam,
                                        float *image_host, int32_t *ticket) {
  GSR_REQUIRE(h && gh && cam && image_host && ticket, "null pointer");
  const int64_t n = gh->n;
  GSR_REQUIRE(n >= 0 && n <= h->n_max, "n exceeds the renderer's capacity");
  GSR_REQUIRE(cam->width == h->width && cam->height == h->height, "camera image size differs from the renderer's");
  GSR_REQUIRE(n == 0 || (gh->xyz && gh->log_scales && gh->rots && gh->opacity_logit && gh->sh), "null attribute pointer");
  DeviceScope scope(h->device);  // the caller's current device is restored on every return path
  GSR_CHECK_CUDA(scope.status);
  const int32_t t = h->next_ticket;
  HostSlot &sl = h->slot[t % h->depth];
  // the frame that used this slot `depth` submissions ago must have been waited for: recycling it silently would lose
  // its status (an overflowed frame would go unnoticed)
  GSR_REQUIRE(!sl.busy, "all `depth` frames are in flight: gsr_host_renderer_wait() for the oldest ticket first");
  const size_t img_bytes = sizeof(float) * 3 * (size_t)h->width * h->height;
  const int64_t n1 = n > 0 ? n : 1;
  float *d_xyz = sl.d_params, *d_sc = d_xyz + pad4(3 * n1), *d_rot = d_sc + pad4(3 * n1), *d_op = d_rot + pad4(4 * n1),
        *d_sh = d_op + pad4(n1);
  const bool sh_planar = gh->sh_sn == 1 && gh->sh_sk == n && gh->sh_sc == 16 * n;
  const bool sh_rows = gh->sh_sn == 48 && gh->sh_sk == 3 && gh->sh_sc == 1;
  const bool chunkable = dense_rows(gh->xyz_sn, gh->xyz_sc, 3) && dense_rows(gh->scale_sn, gh->scale_sc, 3) &&
                         dense_rows(gh->rot_sn, gh->rot_sc, 4) && gh->opacity_sn == 1 && (sh_planar || sh_rows);
  GSR_REQUIRE(n == 0 || chunkable,
              "host renderer takes dense [n,3] / [n,3] / [n,4] / [n] arrays and SH as planar [3,16,n] or dense [n,16,3]");
  // device-side views: the layouts are kept (planar stays planar)
  gsr_gaussians gd;
  gd.n = n;
  gd.xyz = d_xyz; gd.xyz_sn = 3; gd.xyz_sc = 1;
  gd.log_scales = d_sc; gd.scale_sn = 3; gd.scale_sc = 1;
  gd.rots = d_rot; gd.rot_sn = 4; gd.rot_sc = 1;
  gd.opacity_logit = d_op; gd.opacity_sn = 1;
  gd.sh = d_sh; gd.sh_sn = gh->sh_sn; gd.sh_sk = gh->sh_sk; gd.sh_sc = gh->sh_sc;

  // frame begin on the compute stream: zero-initialised parts of the workspace
  FrameBuffers fb;
  int rc = frame_buffers(n, cam->width, cam->height, sl.max_dup, sl.d_ws, h->ws_bytes, &fb);
  if (rc) return rc;
  rc = frame_begin(fb, h->compute);
  if (rc) return rc;
  // the previous frame of this slot has consumed d_params (stream-ordered: same compute stream, but the UPLOAD stream
  // must not overwrite them earlier)
  GSR_CHECK_CUDA(cudaStreamWaitEvent(h->up, sl.params_free, 0));
  const int chunks = n >= 64 * kHostChunks ? kHostChunks : 1;
  for (int c = 0; c < chunks && n > 0; ++c) {
    const int64_t c0 = (n * c / chunks) & ~(int64_t)31, c1 = c == chunks - 1 ? n : ((n * (c + 1) / chunks) & ~(int64_t)31);
    const int64_t m = c1 - c0;
    if (m <= 0) continue;
    GSR_CHECK_CUDA(cudaMemcpyAsync(d_xyz + 3 * c0, gh->xyz + 3 * c0, sizeof(float) * 3 * m, cudaMemcpyHostToDevice, h->up));
    GSR_CHECK_CUDA(cudaMemcpyAsync(d_sc + 3 * c0, gh->log_scales + 3 * c0, sizeof(float) * 3 * m, cudaMemcpyHostToDevice, h->up));
    GSR_CHECK_CUDA(cudaMemcpyAsync(d_rot + 4 * c0, gh->rots + 4 * c0, sizeof(float) * 4 * m, cudaMemcpyHostToDevice, h->up));
    GSR_CHECK_CUDA(cudaMemcpyAsync(d_op + c0, gh->opacity_logit + c0, sizeof(float) * m, cudaMemcpyHostToDevice, h->up));
    if (sh_planar) {
      GSR_CHECK_CUDA(cudaMemcpy2DAsync(d_sh + c0, sizeof(float) * n, gh->sh + c0, sizeof(float) * n, sizeof(float) * m, 48,
                                       cudaMemcpyHostToDevice, h->up));
    } else {
      GSR_CHECK_CUDA(cudaMemcpyAsync(d_sh + 48 * c0, gh->sh + 48 * c0, sizeof(float) * 48 * m, cudaMemcpyHostToDevice, h->up));
    }
    GSR_CHECK_CUDA(cudaEventRecord(sl.chunk_up[c], h->up));
    GSR_CHECK_CUDA(cudaStreamWaitEvent(h->compute, sl.chunk_up[c], 0));
    // per-Gaussian preprocess of the chunk: views and outputs offset by c0, histograms accumulate over the chunks
    gsr_gaussians gc = gd;
    gc.n = m;
    gc.xyz += 3 * c0; gc.log_scales += 3 * c0; gc.rots += 4 * c0; gc.opacity_logit += c0;
    gc.sh += sh_planar ? c0 : 48 * c0;
    rc = preprocess_raw(&gc, cam, 0, 1 << 30, fb.depth_key + c0, fb.tiles_touched + c0, fb.tile_span + 4 * c0,
                        fb.rec + GSR_REC_FLOATS * c0, nullptr, nullptr, nullptr, nullptr, fb.depth_hist, h->compute, true);
    if (rc) return rc;
  }
  GSR_CHECK_CUDA(cudaEventRecord(sl.params_free, h->compute));
  rc = frame_finish(fb, n, cam, sl.max_dup, sl.d_image, sl.d_info, h->compute);
  if (rc) return rc;
  GSR_CHECK_CUDA(cudaEventRecord(sl.rendered, h->compute));
  GSR_CHECK_CUDA(cudaStreamWaitEvent(h->down, sl.rendered, 0));
  GSR_CHECK_CUDA(cudaMemcpyAsync(image_host, sl.d_image, img_bytes, cudaMemcpyDeviceToHost, h->down));
  GSR_CHECK_CUDA(cudaMemcpyAsync(sl.h_info, sl.d_info, sizeof(gsr_frame_info), cudaMemcpyDeviceToHost, h->down));
  GSR_CHECK_CUDA(cudaMemsetAsync(sl.d_info, 0, sizeof(gsr_frame_info), h->down));  // sticky status handed over: clear
  GSR_CHECK_CUDA(cudaEventRecord(sl.done, h->down));
  // the next frame of this slot may not start its memsets / kernels before this image has left (same d_image / d_ws)
  GSR_CHECK_CUDA(cudaStreamWaitEvent(h->compute, sl.done, 0));
  sl.busy = true;
  sl.ticket = t;
  *ticket = t;
  h->next_ticket = t + 1;
  return GSR_OK;
}

extern "C" int gsr_host_renderer_wait(gsr_host_renderer *h, int32_t ticket, gsr_frame_info *info_host) {
  GSR_REQUIRE(h, "null pointer");
  GSR_REQUIRE(ticket >= 0 && ticket < h->next_ticket && ticket >= h->next_ticket - h->depth, "unknown or recycled ticket");
  HostSlot &sl = h->slot[ticket % h->depth];
  GSR_REQUIRE(sl.ticket == ticket, "ticket already recycled");
  DeviceScope scope(h->device);
  GSR_CHECK_CUDA(scope.status);
  GSR_CHECK_CUDA(cudaEventSynchronize(sl.done));
  sl.busy = false;
  if (info_host) *info_host = *sl.h_info;
  if (sl.h_info->status == GSR_ERR_DUP_OVERFLOW) {
    set_error("binning needed %u (Gaussian, tile) pairs but max_dup is %lld; create the renderer with a larger max_dup",
              sl.h_info->n_dup_peak, (long long)sl.max_dup);
    return GSR_ERR_DUP_OVERFLOW;
  }
  return (int)sl.h_info->status;
}

extern "C" int gsr_host_renderer_render(gsr_host_renderer *h, const gsr_gaussians *g_host, const gsr_camera *cam,
                                        float *image_host, gsr_frame_info *info_host) {
  int32_t t = -1;
  int rc = gsr_host_renderer_submit(h, g_host, cam, image_host, &t);
  if (rc) return rc;
  return gsr_host_renderer_wait(h, t, info_host);
}
