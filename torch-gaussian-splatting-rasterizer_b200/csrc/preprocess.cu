// preprocess.cu - fused per-Gaussian preprocess + SH colour, and the staged helpers built from the same
// device functions.  One thread per Gaussian; with the planar [3,16,N] SH layout every load instruction of a
// warp reads one full 128-byte line (HBM-bound stage: 236 B in / 64 B out per Gaussian at SH degree 3).
//
// Replaces /root/reference/rasterize.py:97-120, 358, 371-420, 441 and spherical_harmonics.py:27-73.
#include "gsr_math.cuh"

namespace gsr {

constexpr int kPreThreads = 256;

template <typename T>
__device__ __forceinline__ T ldg_stream(const T *p) {
  return __ldg(p);
}

__device__ __forceinline__ uint32_t
preprocess_one(const int64_t i, const gsr_gaussians &g, const gsr_camera &cam, int row_begin, int row_end,
               uint32_t *__restrict__ depth_key, uint32_t *__restrict__ tiles_touched,
               ushort4 *__restrict__ rect_out, float4 *__restrict__ rec, float *__restrict__ cov3d_out,
               float *__restrict__ cov2d_out, int *__restrict__ rect_raw_out, float *__restrict__ mean2d_out) {

  // ---- loads: everything this Gaussian needs is requested up front (59 independent 4-byte loads) ----
  float p[3], ls[3], rot[4];
#pragma unroll
  for (int c = 0; c < 3; ++c) p[c] = __ldg(g.xyz + i * g.xyz_sn + c * g.xyz_sc);
#pragma unroll
  for (int c = 0; c < 3; ++c) ls[c] = __ldg(g.log_scales + i * g.scale_sn + c * g.scale_sc);
#pragma unroll
  for (int c = 0; c < 4; ++c) rot[c] = __ldg(g.rots + i * g.rot_sn + c * g.rot_sc);
  const float op_logit = __ldg(g.opacity_logit + i * g.opacity_sn);
  const int degree = cam.sh_degree;
  const int ncoef = (degree + 1) * (degree + 1);
  float sh[3][16];
  const float *shp = g.sh + i * g.sh_sn;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int k = 0; k < 16; ++k) sh[c][k] = (k < ncoef) ? __ldg(shp + k * g.sh_sk + c * g.sh_sc) : 0.0f;

  // ---- geometry: rasterize.py:371-420 ----
  float camp[3], clip[4];
#pragma unroll
  for (int c = 0; c < 3; ++c) camp[c] = row_affine(cam.w2c_t, c, p[0], p[1], p[2]);
#pragma unroll
  for (int c = 0; c < 4; ++c) clip[c] = row_affine(cam.full_t, c, p[0], p[1], p[2]);
  const bool culled = camp[2] < 0.2f;  // :377
  if (culled) clip[0] = clip[1] = clip[3] = 0.0f;
  const float pw = frcp(fadd(clip[3], 0.0000001f));  // :381
  const float mx = fdiv(fsub(fmul(fadd(fmul(clip[0], pw), 1.0f), (float)cam.width), 1.0f), 2.0f);   // :391
  const float my = fdiv(fsub(fmul(fadd(fmul(clip[1], pw), 1.0f), (float)cam.height), 1.0f), 2.0f);

  float S[9], c2[4];
  covariance3d(ls, rot, S);
  covariance2d(S, camp, cam, c2);
  if (culled) c2[0] = c2[1] = c2[2] = c2[3] = 0.0f;  // :388
  float radius;
  int rr[4];
  covering_rect(mx, my, c2, cam.width, cam.height, radius, rr);
  float ca, cc, cb;
  conic_from_cov(c2, ca, cc, cb);

  // tile rect clipped to the writable tile grid and to this rank's band of tile rows
  const TileGrid tg = make_grid(cam.width, cam.height);
  int x0 = min(rr[0], tg.tiles_x_eff), x1 = min(rr[2], tg.tiles_x_eff);
  int y0 = min(rr[1], tg.tiles_y_eff), y1 = min(rr[3], tg.tiles_y_eff);
  y0 = max(y0, row_begin);
  y1 = min(y1, row_end);
  const bool skip = (ca == 0.0f) | (cc == 0.0f) | (cb == 0.0f);  // :441 ("any conic component is zero")
  uint32_t tiles = 0;
  if (!skip && x1 > x0 && y1 > y0) tiles = (uint32_t)(x1 - x0) * (uint32_t)(y1 - y0);

  const uint32_t key = float_to_key(camp[2]);
  depth_key[i] = key;
  tiles_touched[i] = tiles;
  rect_out[i] = make_ushort4((unsigned short)x0, (unsigned short)y0, (unsigned short)x1, (unsigned short)y1);

  if (cov3d_out) {
#pragma unroll
    for (int k = 0; k < 9; ++k) cov3d_out[i * 9 + k] = S[k];
  }
  if (cov2d_out) {
#pragma unroll
    for (int k = 0; k < 4; ++k) cov2d_out[i * 4 + k] = c2[k];
  }
  if (rect_raw_out) {
#pragma unroll
    for (int k = 0; k < 4; ++k) rect_raw_out[i * 4 + k] = rr[k];
  }
  if (mean2d_out) {
    mean2d_out[i * 2] = mx;
    mean2d_out[i * 2 + 1] = my;
  }

  if (tiles == 0 && !cov3d_out) return key;  // invisible: the blend never reads its record

  // ---- colour + opacity (spherical_harmonics.py:36-71, rasterize.py:358) ----
  float dx, dy, dz;
  view_direction(p, cam.cam_center, dx, dy, dz);
  float rgb[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) rgb[c] = sh_channel(dx, dy, dz, degree, [&](int k) { return sh[c][k]; });
  const float opacity = sigmoid_ref(op_logit);

  // Blend record (include/gsr_b200.h).  A = -a/2, C = -c/2, B = -b are exact rescalings of the reference's conic,
  // so power = (A dx^2 + C dy^2) + (B dx) dy reproduces rasterize.py:280-283 bit for bit.  `thr` is a lower bound
  // on the power at which alpha = o*exp(power) can exceed 1/255: below it the blend kernel skips the exp.
  const float thr = opacity > 0.0f ? -logf(255.0f * opacity) - 1e-4f : __int_as_float(0x7f800000);
  rec[i * 3 + 0] = make_float4(mx, my, fmul(-0.5f, ca), fmul(-0.5f, cc));
  rec[i * 3 + 1] = make_float4(-cb, thr, opacity, rgb[0]);
  rec[i * 3 + 2] = make_float4(rgb[1], rgb[2], opacity > 0.0f ? log2f(opacity) : -1e30f, radius);
  return key;
}


// Persistent grid-stride kernel.  Besides the per-Gaussian outputs it accumulates the four 8-bit digit histograms of
// the depth keys (the one-sweep sort's global digit counts) in shared memory - the atomics ride on the otherwise
// idle LSU pipe of this memory-bound kernel - and flushes them once per CTA.
__global__ void __launch_bounds__(kPreThreads, 2)
preprocess_kernel(const gsr_gaussians g, const __grid_constant__ gsr_camera cam, int row_begin, int row_end,
                  uint32_t *__restrict__ depth_key, uint32_t *__restrict__ tiles_touched,
                  ushort4 *__restrict__ rect_out, float4 *__restrict__ rec, float *__restrict__ cov3d_out,
                  float *__restrict__ cov2d_out, int *__restrict__ rect_raw_out, float *__restrict__ mean2d_out,
                  uint32_t *__restrict__ depth_hist) {
  __shared__ uint32_t s_hist[4][256];
  if (depth_hist) {
    for (int k = threadIdx.x; k < 1024; k += kPreThreads) (&s_hist[0][0])[k] = 0;
    __syncthreads();
  }
  const int64_t stride = (int64_t)gridDim.x * kPreThreads;
  for (int64_t i = (int64_t)blockIdx.x * kPreThreads + threadIdx.x; i < g.n; i += stride) {
    const uint32_t key = preprocess_one(i, g, cam, row_begin, row_end, depth_key, tiles_touched, rect_out, rec,
                                        cov3d_out, cov2d_out, rect_raw_out, mean2d_out);
    if (depth_hist) {
#pragma unroll
      for (int p = 0; p < 4; ++p) atomicAdd(&s_hist[p][(key >> (8 * p)) & 255u], 1u);
    }
  }
  if (depth_hist) {
    __syncthreads();
    for (int k = threadIdx.x; k < 1024; k += kPreThreads) {
      const uint32_t v = (&s_hist[0][0])[k];
      if (v) atomicAdd(depth_hist + k, v);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// staged helpers
// ---------------------------------------------------------------------------------------------
__global__ void sh_kernel(int64_t n, const float *xyz, int64_t xn, int64_t xc, const float *sh, int64_t sn, int64_t sk,
                          int64_t sc, float3 centre, int degree, float *rgb) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float p[3] = {xyz[i * xn], xyz[i * xn + xc], xyz[i * xn + 2 * xc]};
  float cc[3] = {centre.x, centre.y, centre.z};
  float dx, dy, dz;
  view_direction(p, cc, dx, dy, dz);
  const float *shp = sh + i * sn;
#pragma unroll
  for (int c = 0; c < 3; ++c)
    rgb[i * 3 + c] = sh_channel(dx, dy, dz, degree, [&](int k) { return __ldg(shp + k * sk + c * sc); });
}

__global__ void cov3d_kernel(int64_t n, const float *ls, int64_t sn, int64_t sc, const float *rots, int64_t rn,
                             int64_t rc, float *cov3d) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float l[3] = {ls[i * sn], ls[i * sn + sc], ls[i * sn + 2 * sc]};
  float r[4] = {rots[i * rn], rots[i * rn + rc], rots[i * rn + 2 * rc], rots[i * rn + 3 * rc]};
  float S[9];
  covariance3d(l, r, S);
#pragma unroll
  for (int k = 0; k < 9; ++k) cov3d[i * 9 + k] = S[k];
}

__global__ void quat_kernel(int64_t n, const float *q, int64_t qc, int64_t qn, float *rot) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float R[3][3];
  quat_to_rot(q[i * qn], q[i * qn + qc], q[i * qn + 2 * qc], q[i * qn + 3 * qc], R);
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) rot[(a * 3 + b) * n + i] = R[a][b];  // [3,3,n]
}

struct Mat4 {
  float m[16];
};
__global__ void affine_kernel(int64_t n, const float *xyz, int64_t sn, int64_t sc, const __grid_constant__ Mat4 M,
                              int cols, float *out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x = xyz[i * sn], y = xyz[i * sn + sc], z = xyz[i * sn + 2 * sc];
  for (int c = 0; c < cols; ++c) out[i * cols + c] = row_affine(M.m, c, x, y, z);
}

__global__ void cov2d_kernel(int64_t n, const float *cov3d, const float *cam_means, const __grid_constant__ gsr_camera cam,
                             float *cov2d) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float S[9], cm[3], c2[4];
#pragma unroll
  for (int k = 0; k < 9; ++k) S[k] = cov3d[i * 9 + k];
#pragma unroll
  for (int k = 0; k < 3; ++k) cm[k] = cam_means[i * 3 + k];
  covariance2d(S, cm, cam, c2);
#pragma unroll
  for (int k = 0; k < 4; ++k) cov2d[i * 4 + k] = c2[k];
}

__global__ void bbox_kernel(int64_t n, const float *mean2d, const float *cov2d, int width, int height, int64_t *rect) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float c2[4] = {cov2d[i * 4], cov2d[i * 4 + 1], cov2d[i * 4 + 2], cov2d[i * 4 + 3]};
  float radius;
  int rr[4];
  covering_rect(mean2d[i * 2], mean2d[i * 2 + 1], c2, width, height, radius, rr);
#pragma unroll
  for (int k = 0; k < 4; ++k) rect[i * 4 + k] = rr[k];
}

static inline unsigned blocks_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace gsr

using namespace gsr;

extern "C" int gsr_preprocess_fwd(const gsr_gaussians *g, const gsr_camera *cam, int32_t row_begin, int32_t row_end,
                                  uint32_t *depth_key, uint32_t *tiles_touched, uint16_t *rect, float *rec,
                                  float *cov3d, float *cov2d, int32_t *tile_rect_raw, float *mean2d_all,
                                  uint32_t *depth_hist, void *stream) {
  GSR_REQUIRE(g && cam, "null gaussians/camera");
  GSR_REQUIRE(g->n >= 0 && g->n < (1ll << 31), "n out of range");
  GSR_REQUIRE(cam->sh_degree >= 0 && cam->sh_degree <= 3, "sh_degree must be 0..3");
  GSR_REQUIRE(cam->width > 0 && cam->height > 0 && cam->width <= 65535 * GSR_TILE && cam->height <= 65535 * GSR_TILE,
              "image size out of range");
  GSR_REQUIRE(depth_key && tiles_touched && rect && rec, "null output");
  GSR_REQUIRE(((uintptr_t)rec & 15) == 0 && ((uintptr_t)rect & 7) == 0, "rec must be 16-byte and rect 8-byte aligned");
  if (depth_hist) GSR_CHECK_CUDA(cudaMemsetAsync(depth_hist, 0, sizeof(uint32_t) * 1024, (cudaStream_t)stream));
  if (g->n == 0) return GSR_OK;
  unsigned blocks = blocks_for(g->n, kPreThreads);
  const unsigned resident = 148 * 4;  // persistent: a few CTAs per SM, each loops over its share
  if (blocks > resident) blocks = resident;
  preprocess_kernel<<<blocks, kPreThreads, 0, (cudaStream_t)stream>>>(
      *g, *cam, row_begin, row_end, depth_key, tiles_touched, (ushort4 *)rect, (float4 *)rec, cov3d, cov2d,
      tile_rect_raw, mean2d_all, depth_hist);
  GSR_LAUNCH_CHECK();
  return GSR_OK;
}

extern "C" int gsr_sh_fwd(int64_t n, const float *xyz, int64_t xyz_sn, int64_t xyz_sc, const float *sh, int64_t sh_sn,
                          int64_t sh_sk, int64_t sh_sc, const float *cam_center_host3, int32_t degree, float *rgb,
                          void *stream) {
  GSR_REQUIRE(degree >= 0 && degree <= 3, "sh degree must be 0..3");
  GSR_REQUIRE(xyz && sh && rgb && cam_center_host3, "null pointer");
  if (n == 0) return GSR_OK;
  float3 c = make_float3(cam_center_host3[0], cam_center_host3[1], cam_center_host3[2]);
  sh_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, xyz, xyz_sn, xyz_sc, sh, sh_sn, sh_sk, sh_sc, c,
                                                                 degree, rgb);
  GSR_LAUNCH_CHECK();
  return GSR_OK;
}

extern "C" int gsr_cov3d_fwd(int64_t n, const float *log_scales, int64_t s_sn, int64_t s_sc, const float *rots,
                             int64_t r_sn, int64_t r_sc, float *cov3d, void *stream) {
  GSR_REQUIRE(log_scales && rots && cov3d, "null pointer");
  if (n == 0) return GSR_OK;
  cov3d_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, log_scales, s_sn, s_sc, rots, r_sn, r_sc, cov3d);
  GSR_LAUNCH_CHECK();
  return GSR_OK;
}

extern "C" int gsr_quat_to_rotmat(int64_t n, const float *q, int64_t q_sc, int64_t q_sn, float *rot, void *stream) {
  GSR_REQUIRE(q && rot, "null pointer");
  if (n == 0) return GSR_OK;
  quat_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, q, q_sc, q_sn, rot);
  GSR_LAUNCH_CHECK();
  return GSR_OK;
}

extern "C" int gsr_affine_fwd(int64_t n, const float *xyz, int64_t sn, int64_t sc, const float *m_host16, int32_t cols,
                              float *out, void *stream) {
  GSR_REQUIRE(xyz && m_host16 && out, "null pointer");
  GSR_REQUIRE(cols >= 1 && cols <= 4, "cols must be 1..4");
  if (n == 0) return GSR_OK;
  Mat4 M;
  for (int k = 0; k < 16; ++k) M.m[k] = m_host16[k];
  affine_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, xyz, sn, sc, M, cols, out);
  GSR_LAUNCH_CHECK();
  return GSR_OK;
}

extern "C" int gsr_cov2d_fwd(int64_t n, const float *cov3d, const float *cam_means, const gsr_camera *cam, float *cov2d,
                             void *stream) {
  GSR_REQUIRE(cov3d && cam_means && cam && cov2d, "null pointer");
  if (n == 0) return GSR_OK;
  cov2d_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, cov3d, cam_means, *cam, cov2d);
  GSR_LAUNCH_CHECK();
  return GSR_OK;
}

extern "C" int gsr_bbox_fwd(int64_t n, const float *mean2d, const float *cov2d, int32_t width, int32_t height,
                            int64_t *rect, void *stream) {
  GSR_REQUIRE(mean2d && cov2d && rect, "null pointer");
  if (n == 0) return GSR_OK;
  bbox_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, mean2d, cov2d, width, height, rect);
  GSR_LAUNCH_CHECK();
  return GSR_OK;
}
