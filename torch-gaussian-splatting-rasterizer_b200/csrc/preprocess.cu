// preprocess.cu - fused per-Gaussian preprocess + SH colour, and the staged helpers built from the same
// device functions.  One thread per Gaussian; with the planar [3,16,N] SH layout every load instruction of a
// warp reads one full 128-byte line (HBM-bound stage: 236 B in / 64 B out per Gaussian at SH degree 3).
//
// Replaces /root/reference/rasterize.py:97-120, 358, 371-420, 441 and spherical_harmonics.py:27-73.
#include <stdlib.h>

#include "gsr_math.cuh"

namespace gsr {

constexpr int kPreThreads = 256;
#ifndef GSR_PRE_PREFETCH
#define GSR_PRE_PREFETCH 1   // L2 prefetch of the block's SH rows and the next block's geometry: 0.185 -> 0.169 ms (r02_notes)
#endif
#ifndef GSR_PRE_CTAS
#define GSR_PRE_CTAS 3   // resident CTAs per SM the prepared kernel is compiled for (79 registers; 4 would need <= 64)
#endif

// ---- the per-Gaussian body shared by the two preprocess kernels ------------------------------------------------
// preprocess_kernel         reads the RAW strided parameters (any torch view) and computes everything per view;
// preprocess_prepared_kernel reads the PREPARED model (prepare.cu: Sigma, sigmoid(opacity), thr, log2 o computed once per
//                           model, everything in 16-byte blocked rows) - the resident-scene render path.
// Both call pre_body() with the same values in the same order, so their outputs are bit-identical
// (tests/test_gpu_parity.py::test_prepared_preprocess_is_bit_identical).
//   A  per lane : projection, EWA covariance, radius, tile rect, conic                        (rasterize.py:371-420)
//   B  per warp : exact tile culling of the warp's 32 rects, candidate tiles dealt evenly to the lanes (a lane
//                 looping over its own rect would make the warp wait for its largest Gaussian)
//   C  per lane : SH colour + blend record, only for Gaussians that keep at least one tile
// It also accumulates the four 8-bit digit histograms of the depth keys (global digit counts of the one-sweep
// sort) in shared memory; the kernels flush them once per CTA.
struct PreShared {
  uint32_t hist[4][256];
  CullGaussian cg[kPreThreads / 32][32];
  int4 rect[kPreThreads / 32][32];      // x0, y0, width, 2^16/width rounded up (0 when the rect is empty)
  uint32_t mask[kPreThreads / 32][32][2];
  int pref[kPreThreads / 32][33];
};

struct PreDebug {  // optional outputs of the raw kernel (tests); all null in the prepared kernel
  float *cov3d, *cov2d, *mean2d;
  int *rect_raw;
};

// S6 = upper triangle of Sigma (S00 S01 S02 S11 S12 S22).  `colour(dx, dy, dz, rgb)` evaluates sh_to_rgb for this
// lane's Gaussian; it is only called for Gaussians that keep at least one tile (or when debug outputs are requested).
// `warm()` is called as soon as the Gaussian is known to keep a non-empty rect in this rank's band - the per-view
// kernel uses it to prefetch the SH rows of survivors only when it renders a band of the frame.
template <bool HIST, typename Colour, typename Warm>
__device__ __forceinline__ void pre_body(PreShared &sm, const gsr_camera &cam, const TileGrid &tg, int row_begin,
                                         int row_end, bool active, int64_t i, const float p[3], const float S6[6],
                                         float opacity, float thr, float log2o, uint32_t *__restrict__ depth_key,
                                         uint32_t *__restrict__ tiles_touched, uint4 *__restrict__ span_out,
                                         float4 *__restrict__ rec, const PreDebug &dbg, Colour colour, Warm warm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ---- A: geometry, rasterize.py:371-420 ----
  float camp[3], clip[4];
#pragma unroll
  for (int c = 0; c < 3; ++c) camp[c] = row_affine(cam.w2c_t, c, p[0], p[1], p[2]);
#pragma unroll
  for (int c = 0; c < 4; ++c) clip[c] = row_affine(cam.full_t, c, p[0], p[1], p[2]);
  const bool culled = camp[2] < 0.2f;  // :377
  if (culled) clip[0] = clip[1] = clip[3] = 0.0f;
  const float pw = frcp(fadd(clip[3], 0.0000001f));  // :381
  const float mx = fmul(fsub(fmul(fadd(fmul(clip[0], pw), 1.0f), (float)cam.width), 1.0f), 0.5f);   // :391 (/ 2)
  const float my = fmul(fsub(fmul(fadd(fmul(clip[1], pw), 1.0f), (float)cam.height), 1.0f), 0.5f);
  float c2[4];
  covariance2d_sym(S6, camp, cam, c2);
  if (culled) c2[0] = c2[1] = c2[2] = c2[3] = 0.0f;  // :388
  float radius;
  int rr[4];
  covering_rect(mx, my, c2, cam.width, cam.height, radius, rr);
  float ca, cc, cb;
  conic_from_cov(c2, ca, cc, cb);
  const bool skip = (ca == 0.0f) | (cc == 0.0f) | (cb == 0.0f);  // :441 ("any conic component is zero")
  // A = -a/2, C = -c/2, B = -b are exact rescalings of the reference's conic, so
  // power = (A dx^2 + C dy^2) + (B dx) dy reproduces rasterize.py:280-283 bit for bit.  `thr` is a lower bound on
  // the power at which alpha = o*exp(power) can exceed 1/255: below it the blend kernel skips the exp.
  const float A = fmul(-0.5f, ca), C = fmul(-0.5f, cc), B = -cb;
  // reference tile rect, clipped to the writable tile grid and to this rank's band of tile rows
  int x0 = min(rr[0], tg.tiles_x_eff), x1 = min(rr[2], tg.tiles_x_eff);
  int y0 = max(min(rr[1], tg.tiles_y_eff), row_begin), y1 = min(min(rr[3], tg.tiles_y_eff), row_end);
  const bool blendable = active && !skip && x1 > x0 && y1 > y0;  // the reference's loop would blend it
  if (blendable) warm();

  // ---- B: drop the (Gaussian, tile) pairs in which no pixel can pass alpha > 1/255 (the image is unchanged) ----
  // B1 (per lane): the pixels that can pass lie in the ellipse power >= limit of the conic THE BLEND KERNEL USES
  // (A, C, B as rounded above - not the covariance they came from: for long thin splats det = c00 c11 - c01^2 cancels
  // and the rounded conic's ellipse is measurably larger or smaller than the covariance's).  With
  // D = 4AC - B^2 > 0 its bounding box is mean +- sqrt(L (-4C) / D), sqrt(L (-4A) / D); D is computed with an
  // explicit rounding bound and the shrink (and the per-tile test, which assumes a concave power) is skipped when the
  // rounded form is not safely negative definite or is ill-conditioned (4AC / D > 50: a thin splat lying diagonally,
  // where the terms of `power` cancel and its rounding error near the ellipse can exceed the margin inside `limit`).
  // Such Gaussians keep the reference's whole rect.
  const CullGaussian cg = make_cull(mx, my, A, C, B, thr);
  bool concave = false;
  if (blendable) {
    const float L = -cg.limit;
    if (!(L > 0.0f)) {
      x1 = x0;  // opacity <= 1/255: no pixel ever passes; the rect becomes empty
    } else if (L < 1e30f) {
      const float ac4 = 4.0f * A * C, b2 = B * B;
      const float D = ac4 - b2 - 4.8e-7f * (fabsf(ac4) + b2);  // lower bound on 4AC - B^2 (4 ulp of the larger terms)
      if (A < 0.0f && C < 0.0f && D > 0.02f * ac4) {
        concave = true;
        const float k = __fdividef(4.0f * L, D);
        const float ex = sqrtf(-C * k) * 1.0001f + 1e-3f, ey = sqrtf(-A * k) * 1.0001f + 1e-3f;
        x0 = max(x0, (int)floorf((mx - ex) * (1.0f / GSR_TILE)));
        x1 = min(x1, (int)floorf((mx + ex) * (1.0f / GSR_TILE)) + 1);
        y0 = max(y0, (int)floorf((my - ey) * (1.0f / GSR_TILE)));
        y1 = min(y1, (int)floorf((my + ey) * (1.0f / GSR_TILE)) + 1);
      }
    }
  }
  // B2 (per warp): rects of at least 2x2 tiles still have corners the ellipse may miss: exact per-tile test, the
  // candidate tiles of the warp's 32 rects dealt evenly to the lanes.  (A one-tile-wide strip under the ellipse's
  // bounding box is always touched, so it needs no test.)
  x1 = max(x1, x0);  // an emptied rect is stored with zero extent
  y1 = max(y1, y0);
  const int w = x1 - x0, h = y1 - y0;
  const int area = blendable ? w * h : 0;
  const int cand = (concave && w >= 2 && h >= 2 && area <= kCullMaxTiles) ? area : 0;  // larger rects are kept whole
  int inc = cand;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int up = __shfl_up_sync(0xFFFFFFFFu, inc, o);
    if (lane >= o) inc += up;
  }
  const int total = __shfl_sync(0xFFFFFFFFu, inc, 31);
  sm.pref[warp][lane] = inc - cand;
  sm.rect[warp][lane] = make_int4(x0, y0, w, w > 0 ? (65536 + w - 1) / w : 0);  // .w: 2^16/width rounded up
  sm.cg[warp][lane] = cg;
  sm.mask[warp][lane][0] = 0;
  sm.mask[warp][lane][1] = 0;
  __syncwarp();
  {
    const float xmax = (float)(cam.width - 2), ymax = (float)(cam.height - 2);  // last column / row never written
    for (int e0 = 0; e0 < total; e0 += 32) {  // warp-uniform trip count: every lane takes part in the votes
      const bool valid = e0 + lane < total;
      const int e = valid ? e0 + lane : total - 1;
      int lo = 0, hi = 31;  // owner: last lane whose exclusive prefix is <= e
#pragma unroll
      for (int it = 0; it < 5; ++it) {
        const int mid = (lo + hi + 1) >> 1;
        if (sm.pref[warp][mid] <= e) lo = mid; else hi = mid - 1;
      }
      const int4 r = sm.rect[warp][lo];
      const int local = e - sm.pref[warp][lo];
      const int row = (local * r.w) >> 16, col = local - row * r.z;  // exact for local < 64, width <= 64
      const float px0 = (float)((r.x + col) * GSR_TILE), py0 = (float)((r.y + row) * GSR_TILE);
      const float px1 = fminf(px0 + (float)(GSR_TILE - 1), xmax), py1 = fminf(py0 + (float)(GSR_TILE - 1), ymax);
      // Lanes that test candidates of the same owner form one contiguous run of the warp (candidates are ordered by
      // owner); the first lane of each run ORs the run's keep bits into the owner's mask - no shared atomics.
      const bool keep = valid && !tile_is_empty(sm.cg[warp][lo], px0, py0, px1, py1);
      const uint32_t kept = __ballot_sync(0xFFFFFFFFu, keep);
      const int prev_lo = __shfl_up_sync(0xFFFFFFFFu, lo, 1);
      const uint32_t heads = __ballot_sync(0xFFFFFFFFu, lane == 0 || prev_lo != lo);
      if ((heads >> lane) & 1u) {
        const uint32_t later = heads & ~((2u << lane) - 1u);                  // heads after this lane
        const int end = later ? __ffs(later) - 1 : 32;                         // one past the run's last lane
        const uint32_t run = (kept >> lane) & ((end - lane) >= 32 ? 0xFFFFFFFFu : ((1u << (end - lane)) - 1u));
        const unsigned long long bits = (unsigned long long)run << local;     // kept bits of a run fit: local + len <= 64
        sm.mask[warp][lo][0] |= (uint32_t)bits;
        sm.mask[warp][lo][1] |= (uint32_t)(bits >> 32);
      }
      __syncwarp();
    }
  }
  __syncwarp();
  uint32_t mask_lo = 0xFFFFFFFFu, mask_hi = 0xFFFFFFFFu, tiles = (uint32_t)area;  // untested rects keep every tile
  if (cand) {
    mask_lo = sm.mask[warp][lane][0];
    mask_hi = sm.mask[warp][lane][1];
    tiles = (uint32_t)(__popc(mask_lo) + __popc(mask_hi));
  }
  __syncwarp();  // shared arrays are rewritten by the next iteration

  const uint32_t key = float_to_key(camp[2]);
  if (active) {
    depth_key[i] = key;
    tiles_touched[i] = tiles | (blendable ? 0x80000000u : 0u);  // bit 31: counted by the reference's loop (:441)
    // tile span: rect (x0 | y0 << 16, x1 | y1 << 16) + keep mask, one 16-byte word - the only thing the binning
    // stage gathers per Gaussian
    span_out[i] = make_uint4((uint32_t)x0 | ((uint32_t)y0 << 16), (uint32_t)x1 | ((uint32_t)y1 << 16), mask_lo, mask_hi);
    if (HIST && blendable) {  // only the Gaussians the reference's loop blends are depth-sorted
#pragma unroll
      for (int q = 0; q < 4; ++q) atomicAdd(&sm.hist[q][(key >> (8 * q)) & 255u], 1u);
    }
    if (dbg.cov2d) {
#pragma unroll
      for (int k = 0; k < 4; ++k) dbg.cov2d[i * 4 + k] = c2[k];
    }
    if (dbg.rect_raw) {
#pragma unroll
      for (int k = 0; k < 4; ++k) dbg.rect_raw[i * 4 + k] = rr[k];
    }
    if (dbg.mean2d) {
      dbg.mean2d[i * 2] = mx;
      dbg.mean2d[i * 2 + 1] = my;
    }
  }

  // ---- C: colour (spherical_harmonics.py:36-71) + blend record (include/gsr_b200.h) ----
  if (active && (tiles != 0 || dbg.cov3d)) {
    float dx, dy, dz;
    view_direction(p, cam.cam_center, dx, dy, dz);
    float rgb[3];
    colour(dx, dy, dz, rgb);
    rec[i * 3 + 0] = make_float4(mx, my, A, C);
    rec[i * 3 + 1] = make_float4(B, thr, opacity, rgb[0]);
    rec[i * 3 + 2] = make_float4(rgb[1], rgb[2], log2o, radius);
  }
}

// view-independent scalars of the blend record (also computed once per model by prepare.cu - same expressions)
__device__ __forceinline__ void opacity_terms(float op_logit, float &opacity, float &thr, float &log2o) {
  opacity = sigmoid_ref(op_logit);                                                             // rasterize.py:358
  thr = opacity > 0.0f ? -logf(255.0f * opacity) - 1e-4f : __int_as_float(0x7f800000);
  log2o = opacity > 0.0f ? log2f(opacity) : -1e30f;
}

// Persistent grid-stride kernel over the RAW parameters, one Gaussian per thread and iteration.
template <int DEGREE>
__global__ void __launch_bounds__(kPreThreads, 2)
preprocess_kernel(const gsr_gaussians g, const __grid_constant__ gsr_camera cam, int row_begin, int row_end,
                  uint32_t *__restrict__ depth_key, uint32_t *__restrict__ tiles_touched,
                  uint4 *__restrict__ span_out, float4 *__restrict__ rec,
                  float *__restrict__ cov3d_out, float *__restrict__ cov2d_out, int *__restrict__ rect_raw_out,
                  float *__restrict__ mean2d_out, uint32_t *__restrict__ depth_hist) {
  __shared__ PreShared sm;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (depth_hist) {
    for (int k = threadIdx.x; k < 1024; k += kPreThreads) (&sm.hist[0][0])[k] = 0;
    __syncthreads();
  }
  constexpr int degree = DEGREE;
  constexpr int ncoef = (degree + 1) * (degree + 1);
  const TileGrid tg = make_grid(cam.width, cam.height);
  const PreDebug dbg = {cov3d_out, cov2d_out, mean2d_out, rect_raw_out};
  const int64_t stride = (int64_t)gridDim.x * kPreThreads;
  for (int64_t base = (int64_t)blockIdx.x * kPreThreads + warp * 32; base < g.n; base += stride) {
    const bool active = base + lane < g.n;                 // the whole warp stays in the loop (phase B is cooperative)
    const int64_t i = active ? base + lane : g.n - 1;
    // loads - everything this Gaussian needs is requested up front (59 independent 4-byte loads)
    float p[3], ls[3], rot[4];
#pragma unroll
    for (int c = 0; c < 3; ++c) p[c] = __ldg(g.xyz + i * g.xyz_sn + c * g.xyz_sc);
#pragma unroll
    for (int c = 0; c < 3; ++c) ls[c] = __ldg(g.log_scales + i * g.scale_sn + c * g.scale_sc);
#pragma unroll
    for (int c = 0; c < 4; ++c) rot[c] = __ldg(g.rots + i * g.rot_sn + c * g.rot_sc);
    const float op_logit = __ldg(g.opacity_logit + i * g.opacity_sn);
    float sh[3][ncoef];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float *shp = g.sh + i * g.sh_sn + c * g.sh_sc;  // walk the coefficients by pointer: one add per load
#pragma unroll
      for (int k = 0; k < ncoef; ++k) {
        sh[c][k] = __ldg(shp);
        shp += g.sh_sk;
      }
    }
    float S[9];
    covariance3d(ls, rot, S);
    const float S6[6] = {S[0], S[1], S[2], S[4], S[5], S[8]};
    float opacity, thr, log2o;
    opacity_terms(op_logit, opacity, thr, log2o);
    if (active && cov3d_out) {
#pragma unroll
      for (int k = 0; k < 9; ++k) cov3d_out[i * 9 + k] = S[k];
    }
    auto colour = [&](float dx, float dy, float dz, float rgb[3]) {
#pragma unroll
      for (int c = 0; c < 3; ++c) rgb[c] = sh_channel(dx, dy, dz, degree, [&](int k) { return sh[c][k < ncoef ? k : 0]; });
    };
    if (depth_hist)
      pre_body<true>(sm, cam, tg, row_begin, row_end, active, i, p, S6, opacity, thr, log2o, depth_key, tiles_touched,
                     span_out, rec, dbg, colour, [] {});
    else
      pre_body<false>(sm, cam, tg, row_begin, row_end, active, i, p, S6, opacity, thr, log2o, depth_key, tiles_touched,
                      span_out, rec, dbg, colour, [] {});
  }
  if (depth_hist) {
    __syncthreads();
    for (int k = threadIdx.x; k < 1024; k += kPreThreads) {
      const uint32_t v = (&sm.hist[0][0])[k];
      if (v) atomicAdd(depth_hist + k, v);
    }
  }
}

// ---- prepared model --------------------------------------------------------------------------------------------
// Everything in the per-Gaussian set-up that does not depend on the camera - Sigma = R S S^T R^T (rasterize.py:97-120:
// three exp, a quaternion normalisation, two 3x3 products), sigmoid(opacity) (:358) and the two logarithms of the blend
// record - is computed ONCE per model by prepare_kernel, with the very functions the raw kernel uses (bit-identical by
// construction), and stored with the positions and the SH coefficients in a layout made for the per-view kernel:
// blocks of 32 Gaussians, every row of a block 32 x float4 = 512 contiguous bytes,
//     geometry [blocks][3][32] float4 :  x y z opacity | S00 S01 S02 S11 | S12 S22 thr log2(opacity)
//     SH       [blocks][12][32] float4:  row 4c + j = coefficients 4j .. 4j+3 of channel c
// so a warp reads its block with 15 fully coalesced 16-byte loads at immediate offsets (the raw layout costs 59 4-byte
// loads and their 64-bit strided address arithmetic per Gaussian) and the view kernel holds no SH registers until it
// knows the Gaussian is visible.  240 B per Gaussian resident (720 MB at 3 M) next to the raw parameters.
constexpr int kPrepGeomRows = 3, kPrepShRows = 12;
static inline int64_t prepared_blocks(int64_t n) { return (n + 31) / 32; }
static inline size_t prepared_geom_bytes(int64_t n) {
  return align_up((size_t)prepared_blocks(n) * kPrepGeomRows * 32 * sizeof(float4), 256);
}
static inline size_t prepared_total_bytes(int64_t n) {
  return prepared_geom_bytes(n) + align_up((size_t)prepared_blocks(n) * kPrepShRows * 32 * sizeof(float4), 256);
}

__global__ void __launch_bounds__(256)
prepare_kernel(const gsr_gaussians g, float4 *__restrict__ geom, float4 *__restrict__ shb) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t blk = i >> 5;
  const int lane = threadIdx.x & 31;
  if (blk >= (g.n + 31) / 32) return;
  float4 *gp = geom + blk * (kPrepGeomRows * 32) + lane;
  float4 *sp = shb + blk * (kPrepShRows * 32) + lane;
  if (i >= g.n) {  // padding lanes of the last block: a Gaussian that can never be seen
    gp[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    gp[32] = make_float4(0.f, 0.f, 0.f, 0.f);
    gp[64] = make_float4(0.f, 0.f, __int_as_float(0x7f800000), -1e30f);
#pragma unroll
    for (int r = 0; r < kPrepShRows; ++r) sp[r * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  float p[3], ls[3], rot[4];
#pragma unroll
  for (int c = 0; c < 3; ++c) p[c] = __ldg(g.xyz + i * g.xyz_sn + c * g.xyz_sc);
#pragma unroll
  for (int c = 0; c < 3; ++c) ls[c] = __ldg(g.log_scales + i * g.scale_sn + c * g.scale_sc);
#pragma unroll
  for (int c = 0; c < 4; ++c) rot[c] = __ldg(g.rots + i * g.rot_sn + c * g.rot_sc);
  float S[9], opacity, thr, log2o;
  covariance3d(ls, rot, S);
  opacity_terms(__ldg(g.opacity_logit + i * g.opacity_sn), opacity, thr, log2o);
  gp[0] = make_float4(p[0], p[1], p[2], opacity);
  gp[32] = make_float4(S[0], S[1], S[2], S[4]);
  gp[64] = make_float4(S[5], S[8], thr, log2o);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float *shp = g.sh + i * g.sh_sn + c * g.sh_sc;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      sp[(c * 4 + j) * 32] = make_float4(__ldg(shp + (4 * j) * g.sh_sk), __ldg(shp + (4 * j + 1) * g.sh_sk),
                                         __ldg(shp + (4 * j + 2) * g.sh_sk), __ldg(shp + (4 * j + 3) * g.sh_sk));
  }
}

// Per-view kernel over the prepared model: one warp per block of 32 Gaussians, persistent.
template <int DEGREE>
__global__ void __launch_bounds__(kPreThreads, GSR_PRE_CTAS)
preprocess_prepared_kernel(const float4 *__restrict__ geom, const float4 *__restrict__ shb, int64_t n,
                           const __grid_constant__ gsr_camera cam, int row_begin, int row_end,
                           uint32_t *__restrict__ depth_key, uint32_t *__restrict__ tiles_touched,
                           uint4 *__restrict__ span_out, float4 *__restrict__ rec, uint32_t *__restrict__ depth_hist,
                           int late_sh /* prefetch SH rows only for Gaussians that keep a rect (band renders) */) {
  __shared__ PreShared sm;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < 1024; k += kPreThreads) (&sm.hist[0][0])[k] = 0;
  __syncthreads();
  constexpr int ncoef = (DEGREE + 1) * (DEGREE + 1);
  constexpr int rows = (ncoef + 3) / 4;  // float4 rows per channel that hold coefficients of this degree
  const TileGrid tg = make_grid(cam.width, cam.height);
  const PreDebug dbg = {nullptr, nullptr, nullptr, nullptr};
  const int64_t blocks = (n + 31) >> 5;
#if GSR_PRE_PREFETCH == 2
  float4 nq0 = make_float4(0.f, 0.f, 0.f, 0.f), nq1 = nq0, nq2 = nq0;
  {
    const int64_t b0 = (int64_t)blockIdx.x * (kPreThreads / 32) + warp;
    if (b0 < blocks) {
      const float4 *g0 = geom + b0 * (kPrepGeomRows * 32) + lane;
      nq0 = __ldg(g0); nq1 = __ldg(g0 + 32); nq2 = __ldg(g0 + 64);
    }
  }
#endif
  for (int64_t blk = (int64_t)blockIdx.x * (kPreThreads / 32) + warp; blk < blocks; blk += (int64_t)gridDim.x * (kPreThreads / 32)) {
    const int64_t i = blk * 32 + lane;
    const bool active = i < n;
    const float4 *gp = geom + blk * (kPrepGeomRows * 32) + lane;
#if GSR_PRE_PREFETCH == 2
    const float4 q0 = nq0, q1 = nq1, q2 = nq2;  // loaded during the previous iteration
    {
      const int64_t nb = blk + (int64_t)gridDim.x * (kPreThreads / 32);
      if (nb < blocks) {
        const float4 *gn = geom + nb * (kPrepGeomRows * 32) + lane;
        nq0 = __ldg(gn); nq1 = __ldg(gn + 32); nq2 = __ldg(gn + 64);
      }
    }
#else
    const float4 q0 = __ldg(gp), q1 = __ldg(gp + 32), q2 = __ldg(gp + 64);
#endif
#if GSR_PRE_PREFETCH
    {  // L2 prefetch: this block's SH rows (wanted ~400 instructions from now, and fetched from DRAM at sector granularity
       // whether or not the neighbours turn out visible) and the next block's geometry rows - no registers held
      const float4 *spf = shb + blk * (kPrepShRows * 32) + lane;
      if (!late_sh) {
#pragma unroll
        for (int r = 0; r < 3 * rows; ++r) {
          const int c = r / rows, j = r % rows;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(spf + (c * 4 + j) * 32));
        }
      }
#if GSR_PRE_PREFETCH == 1
      const int64_t nb = blk + (int64_t)gridDim.x * (kPreThreads / 32);
      if (nb < blocks) {
        const float4 *gn = geom + nb * (kPrepGeomRows * 32) + lane;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(gn));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(gn + 32));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(gn + 64));
      }
#endif
    }
#endif
    const float p[3] = {q0.x, q0.y, q0.z};
    const float S6[6] = {q1.x, q1.y, q1.z, q1.w, q2.x, q2.y};
    const float4 *sp = shb + blk * (kPrepShRows * 32) + lane;
    auto colour = [&](float dx, float dy, float dz, float rgb[3]) {
      float4 v[3][rows];
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int j = 0; j < rows; ++j) v[c][j] = __ldg(sp + (c * 4 + j) * 32);
#pragma unroll
      for (int c = 0; c < 3; ++c)
        rgb[c] = sh_channel(dx, dy, dz, DEGREE, [&](int k) {
          const float4 q = v[c][k < ncoef ? k >> 2 : 0];
          const int e = k & 3;
          return e == 0 ? q.x : (e == 1 ? q.y : (e == 2 ? q.z : q.w));
        });
    };
    auto warm = [&]() {
#if GSR_PRE_PREFETCH
      if (late_sh) {
#pragma unroll
        for (int r = 0; r < 3 * rows; ++r) {
          const int c = r / rows, j = r % rows;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(sp + (c * 4 + j) * 32));
        }
      }
#endif
    };
    pre_body<true>(sm, cam, tg, row_begin, row_end, active, i, p, S6, q0.w, q2.z, q2.w, depth_key, tiles_touched, span_out,
                   rec, dbg, colour, warm);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < 1024; k += kPreThreads) {
    const uint32_t v = (&sm.hist[0][0])[k];
    if (v) atomicAdd(depth_hist + k, v);
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

// launch only (depth_hist already zeroed by the caller)
int launch_preprocess_prepared(const void *prepared, int64_t n, const gsr_camera *cam, int row_begin, int row_end,
                               uint32_t *depth_key, uint32_t *tiles_touched, uint32_t *tile_span, float *rec,
                               uint32_t *depth_hist, cudaStream_t stream) {
  const float4 *geom = (const float4 *)prepared;
  const float4 *shb = (const float4 *)((const char *)prepared + prepared_geom_bytes(n));
  unsigned blocks = blocks_for(prepared_blocks(n), kPreThreads / 32);
  // persistent: GSR_PRE_GRID CTAs per SM (default 6 = two rounds of the 3 that are resident), each loops over its share
  static const int per_sm = [] {
    const char *e = getenv("GSR_PRE_GRID");
    const int v = e ? atoi(e) : 6;
    return (v < 1 || v > 64) ? 6 : v;
  }();
  const unsigned resident = (unsigned)sm_count() * (unsigned)per_sm;
  if (blocks > resident) blocks = resident;
  // A band render keeps a rect for a fraction of the Gaussians only: their SH rows (192 of the 240 bytes per Gaussian)
  // are then prefetched once the rect is known instead of up front for everybody.  GSR_PRE_LATE_SH=0/1 forces the mode.
  static const int late_env = [] {
    const char *e = getenv("GSR_PRE_LATE_SH");
    return e ? atoi(e) : -1;
  }();
  const TileGrid tgh = make_grid(cam->width, cam->height);
  const int late_sh = late_env >= 0 ? late_env : ((row_begin > 0 || row_end < tgh.tiles_y_eff) ? 1 : 0);
#define GSR_PRE(D)                                                                                                  \
  preprocess_prepared_kernel<D><<<blocks, kPreThreads, 0, stream>>>(geom, shb, n, *cam, row_begin, row_end, depth_key, \
                                                                    tiles_touched, (uint4 *)tile_span, (float4 *)rec, \
                                                                    depth_hist, late_sh)
  switch (cam->sh_degree) {
    case 0: GSR_PRE(0); break;
    case 1: GSR_PRE(1); break;
    case 2: GSR_PRE(2); break;
    default: GSR_PRE(3); break;
  }
#undef GSR_PRE
  GSR_LAUNCH_CHECK();
  return GSR_OK;
}

}  // namespace gsr

using namespace gsr;

extern "C" int gsr_preprocess_fwd(const gsr_gaussians *g, const gsr_camera *cam, int32_t row_begin, int32_t row_end,
                                  uint32_t *depth_key, uint32_t *tiles_touched, uint32_t *tile_span, float *rec,
                                  float *cov3d, float *cov2d, int32_t *tile_rect_raw, float *mean2d_all,
                                  uint32_t *depth_hist, void *stream) {
  return preprocess_raw(g, cam, row_begin, row_end, depth_key, tiles_touched, tile_span, rec, cov3d, cov2d, tile_rect_raw,
                        mean2d_all, depth_hist, (cudaStream_t)stream, false);
}

int gsr::preprocess_raw(const gsr_gaussians *g, const gsr_camera *cam, int32_t row_begin, int32_t row_end,
                        uint32_t *depth_key, uint32_t *tiles_touched, uint32_t *tile_span, float *rec, float *cov3d,
                        float *cov2d, int32_t *tile_rect_raw, float *mean2d_all, uint32_t *depth_hist,
                        cudaStream_t stream, bool hist_zeroed) {
  GSR_REQUIRE(g && cam, "null gaussians/camera");
  GSR_REQUIRE(g->n >= 0 && g->n < (1ll << 31), "n out of range");
  GSR_REQUIRE(cam->sh_degree >= 0 && cam->sh_degree <= 3, "sh_degree must be 0..3");
  GSR_REQUIRE(cam->width > 0 && cam->height > 0 && cam->width <= 65535 * GSR_TILE && cam->height <= 65535 * GSR_TILE,
              "image size out of range");
  GSR_REQUIRE(depth_key && tiles_touched && tile_span && rec, "null output");
  GSR_REQUIRE(((uintptr_t)rec & 15) == 0 && ((uintptr_t)tile_span & 15) == 0, "rec and tile_span must be 16-byte aligned");
  if (depth_hist && !hist_zeroed) GSR_CHECK_CUDA(cudaMemsetAsync(depth_hist, 0, sizeof(uint32_t) * 1024, stream));
  if (g->n == 0) return GSR_OK;
  unsigned blocks = blocks_for(g->n, kPreThreads);
  const unsigned resident = (unsigned)sm_count() * 4;  // persistent: a few CTAs per SM, each loops over its share
  if (blocks > resident) blocks = resident;
#define GSR_PRE(D)                                                                                                \
  preprocess_kernel<D><<<blocks, kPreThreads, 0, stream>>>(*g, *cam, row_begin, row_end, depth_key,                   \
                                                                        tiles_touched, (uint4 *)tile_span,           \
                                                                        (float4 *)rec, cov3d, cov2d, tile_rect_raw,  \
                                                                        mean2d_all, depth_hist)
  switch (cam->sh_degree) {  // the SH degree fixes the number of loads and the colour polynomial at compile time
    case 0: GSR_PRE(0); break;
    case 1: GSR_PRE(1); break;
    case 2: GSR_PRE(2); break;
    default: GSR_PRE(3); break;
  }
#undef GSR_PRE
  GSR_LAUNCH_CHECK();
  return GSR_OK;
}

extern "C" int gsr_prepared_bytes(int64_t n, size_t *bytes) {
  GSR_REQUIRE(bytes && n >= 0 && n < (1ll << 31), "bad arguments");
  *bytes = prepared_total_bytes(n > 0 ? n : 1);
  return GSR_OK;
}

extern "C" int gsr_prepare(const gsr_gaussians *g, void *prepared, size_t prepared_bytes, void *stream) {
  GSR_REQUIRE(g && prepared, "null pointer");
  GSR_REQUIRE(g->n >= 0 && g->n < (1ll << 31), "n out of range");
  GSR_REQUIRE(((uintptr_t)prepared & 255) == 0, "prepared buffer must be 256-byte aligned");
  if (prepared_bytes < prepared_total_bytes(g->n > 0 ? g->n : 1)) {
    set_error("prepared buffer too small (%zu < %zu)", prepared_bytes, prepared_total_bytes(g->n > 0 ? g->n : 1));
    return GSR_ERR_WORKSPACE_TOO_SMALL;
  }
  if (g->n == 0) return GSR_OK;
  float4 *geom = (float4 *)prepared;
  float4 *shb = (float4 *)((char *)prepared + prepared_geom_bytes(g->n));
  prepare_kernel<<<blocks_for(prepared_blocks(g->n) * 32, 256), 256, 0, (cudaStream_t)stream>>>(*g, geom, shb);
  GSR_LAUNCH_CHECK();
  return GSR_OK;
}

extern "C" int gsr_preprocess_prepared_fwd(const void *prepared, int64_t n, const gsr_camera *cam, int32_t row_begin,
                                           int32_t row_end, uint32_t *depth_key, uint32_t *tiles_touched,
                                           uint32_t *tile_span, float *rec, uint32_t *depth_hist, void *stream) {
  GSR_REQUIRE(prepared && cam, "null prepared model / camera");
  GSR_REQUIRE(n >= 0 && n < (1ll << 31), "n out of range");
  GSR_REQUIRE(cam->sh_degree >= 0 && cam->sh_degree <= 3, "sh_degree must be 0..3");
  GSR_REQUIRE(cam->width > 0 && cam->height > 0 && cam->width <= 65535 * GSR_TILE && cam->height <= 65535 * GSR_TILE,
              "image size out of range");
  GSR_REQUIRE(depth_key && tiles_touched && tile_span && rec && depth_hist, "null output");
  GSR_REQUIRE(((uintptr_t)rec & 15) == 0 && ((uintptr_t)tile_span & 15) == 0 && ((uintptr_t)prepared & 255) == 0,
              "rec / tile_span must be 16-byte aligned, prepared 256-byte aligned");
  GSR_CHECK_CUDA(cudaMemsetAsync(depth_hist, 0, sizeof(uint32_t) * 1024, (cudaStream_t)stream));
  if (n == 0) return GSR_OK;
  return launch_preprocess_prepared(prepared, n, cam, row_begin, row_end, depth_key, tiles_touched, tile_span, rec,
                                    depth_hist, (cudaStream_t)stream);
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
