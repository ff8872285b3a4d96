// backward.cu - gradients of the render path: blend -> preprocess -> SH (north_star item 4).
//
// The reference has no backward; its forward is differentiable torch code, so the contract is "what torch.autograd
// computes through /root/reference/rasterize.py:255-305, 371-420 and spherical_harmonics.py:27-73" (tests compare
// with autograd through the reference's own functions, tests/golden/grad_small.npz, and through oracle/torch_port.py).
//
// Non-differentiable pieces follow autograd: the keep mask (alpha > 1/255, power <= 0) selects which pixels a
// Gaussian touches but carries no gradient; min(alpha, 0.99) passes gradient only on the unclamped side;
// clamp(colour, 0, 1) only inside [0,1]; the clamp of x/z in the EWA Jacobian only inside the limits; radius,
// tile rect and the depth order are integer-valued.  Culled Gaussians (z < 0.2) and Gaussians that reach no tile get
// zero gradients.
//
//   blend_bwd      : one CTA per tile with the forward kernel's decomposition (64 threads, 4 pixels of one column per
//                    thread, 128-record batches staged with cp.async), the tile list replayed BACK TO FRONT.  T before
//                    contributor i is recovered as T_{i+1}/(1-alpha_i) from the forward's final T.  A thread sums the
//                    raw moments of dL/dpower over its 4 pixels; a warp then reduces 8 of the 9 sums with a
//                    recursive-halving butterfly (15 shuffles instead of 45) that leaves each sum in a different lane
//                    group, so ONE atomic instruction adds all nine values: 2 warp reductions per (tile, Gaussian)
//                    instead of 8, ~4x fewer instructions than the first, one-pixel-per-thread version.
//   preprocess_bwd : one thread per Gaussian, forward intermediates recomputed from the raw parameters.
#include "blend_common.cuh"
#include "gsr_math.cuh"

namespace gsr {

constexpr int kBwdPX = 4;                        // pixels per thread (one column, rows r, r+4, r+8, r+12)
constexpr int kBwdThreads = GSR_TILE * GSR_TILE / kBwdPX;
constexpr int kBwdBatch = 128;                   // records per staged batch
// grad_rec row: sums over the pixels a Gaussian was blended into of dL/dpower x (dx, dy, dx^2, dy^2, dx dy, 1), then
// dL/d(r, g, b).  Everything the blend contributes is linear in these, so the conversion to d mean2d, d conic and
// d opacity happens once per Gaussian in preprocess_bwd instead of once per (warp, Gaussian) here.
constexpr int kGradRec = 9;
constexpr float kBwdParkedY = 1e18f;

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

// Sums v[0..7] over the warp by recursive halving: after the exchange with lane^16 a lane keeps four of the sums,
// after lane^8 two, after lane^4 one; two plain butterfly steps finish it.  Returns sum number
// 4*bit4(lane) + 2*bit3(lane) + bit2(lane), complete in every lane.
__device__ __forceinline__ float warp_sum8_scattered(const float v[8], int lane) {
  const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4;
  float r[4], s[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = h4 ? v[i] : v[i + 4], keep = h4 ? v[i + 4] : v[i];
    r[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = h3 ? r[i] : r[i + 2], keep = h3 ? r[i + 2] : r[i];
    s[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, 8);
  }
  const float send = h2 ? s[0] : s[1], keep = h2 ? s[1] : s[0];
  float t = keep + __shfl_xor_sync(0xFFFFFFFFu, send, 4);
  t += __shfl_xor_sync(0xFFFFFFFFu, t, 2);
  t += __shfl_xor_sync(0xFFFFFFFFu, t, 1);
  return t;
}

__global__ void __launch_bounds__(kBwdThreads)
blend_bwd_kernel(const float4 *__restrict__ rec, const uint32_t *__restrict__ gids, const uint2 *__restrict__ ranges,
                 int width, int height, int tiles_x, int row_begin, const float *__restrict__ final_T,
                 const uint32_t *__restrict__ n_contrib, const float *__restrict__ dL_dimage,
                 float *__restrict__ grad_rec, float neg_zero) {
  static_assert(kBwdPX == 4, "blend_power4 evaluates four pixels");
  const f32x2 neg_zero2 = pack2(neg_zero, neg_zero);  // -0.0f from the host: see blend_common.cuh
  constexpr int PX = kBwdPX, THREADS = kBwdThreads, kBatch = kBwdBatch;
  constexpr int PER = kBatch / THREADS;
  constexpr int R = GSR_TILE / PX, COLS = 32 / R;  // same lane -> pixel map as the forward kernel
  __shared__ __align__(16) float s_rec[2][kBatch * GSR_REC_FLOATS];
  __shared__ uint32_t s_gid[2][kBatch];
  __shared__ uint32_t s_max_last;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile_x = blockIdx.x, tile_y = blockIdx.y + row_begin;
  const int px = tile_x * GSR_TILE + warp * COLS + (lane % COLS);
  const int py0 = tile_y * GSR_TILE + (lane / COLS);
  const uint2 range = ranges[tile_y * tiles_x + tile_x];
  const uint32_t total = range.y - range.x;
  const float fx = (float)px;

  float fy_real[PX], fy[PX], T[PX], g0[PX], g1[PX], g2[PX], behind[PX];
  uint32_t last[PX];
  uint32_t my_max = 0;
#pragma unroll
  for (int k = 0; k < PX; ++k) {
    const int py = py0 + k * R;
    fy_real[k] = (float)py;
    fy[k] = kBwdParkedY;
    T[k] = 1.0f;
    g0[k] = g1[k] = g2[k] = 0.0f;
    behind[k] = 0.0f;  // sum over contributors behind the current one of alpha_j T_j (c_j . g)
    last[k] = 0;
    if (px < width - 1 && py < height - 1) {  // Q8: the reference never writes the last column / row
      const size_t p = (size_t)py * width + px;
      T[k] = final_T[p];
      last[k] = min(n_contrib[p], total);
      g0[k] = dL_dimage[p * 3 + 0];
      g1[k] = dL_dimage[p * 3 + 1];
      g2[k] = dL_dimage[p * 3 + 2];
    }
    my_max = max(my_max, last[k]);
  }
  if (tid == 0) s_max_last = 0;
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) my_max = max(my_max, __shfl_xor_sync(0xFFFFFFFFu, my_max, o));
  if (lane == 0) atomicMax(&s_max_last, my_max);
  __syncthreads();
  const uint32_t max_last = s_max_last;  // contributors past it touched no pixel of this tile
  const int num_batches = (int)((max_last + kBatch - 1) / kBatch);
  if (num_batches == 0) return;

  // which of the nine sums this lane adds to grad_rec after the warp reduction
  const int my_sum = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  const bool adds = (lane & 3) == 0 || lane == 1;  // lanes 0,4,..,28 add outputs 0..7, lane 1 adds output 8

  // batches are visited last to first: step s handles batch num_batches-1-s
  uint32_t gnext[PER];
  auto load_ids = [&](int s) {
#pragma unroll
    for (int r = 0; r < PER; ++r) {
      const uint32_t idx = (uint32_t)(num_batches - 1 - s) * kBatch + r * THREADS + tid;
      gnext[r] = (s < num_batches && idx < max_last) ? __ldg(gids + range.x + idx) : 0xFFFFFFFFu;
    }
  };
  auto issue = [&](int buf) {
#pragma unroll
    for (int r = 0; r < PER; ++r) {
      const int slot = r * THREADS + tid;
      float *dst = &s_rec[buf][slot * GSR_REC_FLOATS];
      if (gnext[r] != 0xFFFFFFFFu) {
        const float4 *src = rec + (size_t)gnext[r] * 3;
        cp_async16(dst, src);
        cp_async16(dst + 4, src + 1);
        cp_async16(dst + 8, src + 2);
        s_gid[buf][slot] = gnext[r];
      } else {
        dst[5] = __int_as_float(0x7f800000);  // slot past the end: thr = +inf, no pixel passes the test
      }
    }
    cp_async_commit();
  };
  load_ids(0);
  issue(0);
  load_ids(1);

  for (int s = 0; s < num_batches; ++s) {
    const int buf = s & 1;
    const uint32_t first_idx = (uint32_t)(num_batches - 1 - s) * kBatch;  // list index of slot 0
    issue(buf ^ 1);  // batch of step s+1 (empty group when past the end)
    load_ids(s + 2);
    cp_async_wait<1>();
    __syncthreads();
    const int count = (int)min((uint32_t)kBatch, max_last - first_idx);
    const int count4 = (count + 3) & ~3;
    const float *base = s_rec[buf];
    for (int j = count4 - 4; j >= 0; j -= 4) {
      // a pixel joins when the walk reaches its last contributor (list index last-1); until then it is parked
#pragma unroll
      for (int k = 0; k < PX; ++k) fy[k] = (first_idx + (uint32_t)j < last[k]) ? fy_real[k] : kBwdParkedY;
#pragma unroll
      for (int u = 3; u >= 0; --u) {
        const float *r = base + (j + u) * GSR_REC_FLOATS;
        const uint32_t lidx = first_idx + (uint32_t)(j + u);
        const float4 q = *reinterpret_cast<const float4 *>(r);      // mx my A C
        const float2 w = *reinterpret_cast<const float2 *>(r + 4);  // B thr
        const float dx = fsub(q.x, fx);
        const float adx2 = fmul(q.z, fmul(dx, dx));
        const float bdx = fmul(w.x, dx);
        float dy[PX], power[PX];
        bool any = false;
        blend_power4(q.y, q.w, adx2, bdx, fy, neg_zero2, dy, power);  // packed FP32, == blend_power bit for bit
#pragma unroll
        for (int k = 0; k < PX; ++k) any |= power[k] >= w.y;
        // sums over this thread's pixels: dL/dpower times (dx, dy, dx^2, dy^2, dx dy), dL/dpower itself, dL/d(r,g,b)
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, dblue = 0.f;
        bool hit = false;
        if (any) {
          const float2 orr = *reinterpret_cast<const float2 *>(r + 6);  // opacity r
          const float4 gbl = *reinterpret_cast<const float4 *>(r + 8);  // g b log2(opacity) .
#pragma unroll
          for (int k = 0; k < PX; ++k) {
            if (power[k] >= w.y && lidx < last[k]) {
              float alpha;
              if (blend_alpha(power[k], orr.x, gbl.z, alpha)) {  // the forward's keep decision, bit for bit
                hit = true;
                const float inv = rcp_approx(1.0f - alpha);  // 1 ulp; T is a product of up to a few hundred of these
                T[k] *= inv;  // transmittance in front of this contributor
                const float wgt = alpha * T[k];
                v[6] = fmaf(wgt, g0[k], v[6]);
                v[7] = fmaf(wgt, g1[k], v[7]);
                dblue = fmaf(wgt, g2[k], dblue);
                const float cdotg = fmaf(orr.y, g0[k], fmaf(gbl.x, g1[k], gbl.y * g2[k]));
                const float dL_dalpha = fmaf(T[k], cdotg, -behind[k] * inv);
                behind[k] = fmaf(wgt, cdotg, behind[k]);
                // min(o e^p, 0.99): gradient only on the unclamped side; d alpha / d power = alpha
                const float dLp = alpha < kMaxDensity ? alpha * dL_dalpha : 0.0f;
                const float ux = dx * dLp, uy = dy[k] * dLp;
                v[0] += ux;
                v[1] += uy;
                v[2] = fmaf(dx, ux, v[2]);
                v[3] = fmaf(dy[k], uy, v[3]);
                v[4] = fmaf(dx, uy, v[4]);
                v[5] += dLp;
              }
            }
          }
        }
        if (__any_sync(0xFFFFFFFFu, hit)) {  // converged here: reduce over the warp's 128 pixels
          const float t = warp_sum8_scattered(v, lane);
          dblue = warp_sum(dblue);
          const float out = lane == 1 ? dblue : t;
          if (adds) atomicAdd(grad_rec + (size_t)s_gid[buf][j + u] * kGradRec + (lane == 1 ? 8 : my_sum), out);
        }
      }
    }
    __syncthreads();  // everyone is done with s_rec[buf] before the batch of step s+2 is gathered into it
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------------------------
#ifndef GSR_PREBWD_PREFETCH
#define GSR_PREBWD_PREFETCH 0   // L2 prefetch of the SH coefficients at the top of preprocess_bwd_kernel (A/B builds)
#endif
#ifndef GSR_PREBWD_CTAS
#define GSR_PREBWD_CTAS 1   // resident CTAs per SM the kernel is compiled for (A/B builds; see profiles/r02_notes.md)
#endif
template <int DEGREE>
__global__ void __launch_bounds__(256, GSR_PREBWD_CTAS)
preprocess_bwd_kernel(const gsr_gaussians g, const __grid_constant__ gsr_camera cam,
                      const uint32_t *__restrict__ tiles_touched, const float *__restrict__ grad_rec,
                      float *__restrict__ d_xyz, float *__restrict__ d_scales, float *__restrict__ d_rots,
                      float *__restrict__ d_opacity, float *__restrict__ d_sh, int64_t dsh_sn, int64_t dsh_sk,
                      int64_t dsh_sc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.n) return;
  constexpr int degree = DEGREE;  // compile-time: the basis tables below stay in registers
  constexpr int ncoef = (degree + 1) * (degree + 1);
#if GSR_PREBWD_PREFETCH
  // L2 prefetch of this Gaussian's SH coefficients (the planar block: one 128-byte line per coefficient and warp).  They
  // are wanted a few hundred instructions from now, after the tiles_touched / parameter round trip; the kernel waits on
  // `long_scoreboard` at 2 CTAs/SM (profiles/r02_backward_ncu.txt) and prefetches hold no registers.  Issued for every
  // Gaussian: DRAM delivers 32-byte sectors (8 neighbours), so skipping the untouched ones would save nothing.
  if (g.sh_sn == 1) {
    const float *spf = g.sh + i;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int k = 0; k < ncoef; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(spf + c * g.sh_sc + k * g.sh_sk));
  }
#endif
  float gp[3] = {0.f, 0.f, 0.f}, gs[3] = {0.f, 0.f, 0.f}, gq[4] = {0.f, 0.f, 0.f, 0.f}, gop = 0.f;
  const bool touched = (__ldg(tiles_touched + i) & 0x7FFFFFFFu) != 0;  // blended somewhere: everything else gets zeros
  float *dsh = d_sh + i * dsh_sn;
  if (!touched) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float *out = dsh + c * dsh_sc;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        *out = 0.0f;
        out += dsh_sk;
      }
    }
  } else {
    float p[3], ls[3], rot[4];
#pragma unroll
    for (int c = 0; c < 3; ++c) p[c] = __ldg(g.xyz + i * g.xyz_sn + c * g.xyz_sc);
#pragma unroll
    for (int c = 0; c < 3; ++c) ls[c] = __ldg(g.log_scales + i * g.scale_sn + c * g.scale_sc);
#pragma unroll
    for (int c = 0; c < 4; ++c) rot[c] = __ldg(g.rots + i * g.rot_sn + c * g.rot_sc);
    const float logit = __ldg(g.opacity_logit + i * g.opacity_sn);
    const float *gr = grad_rec + i * kGradRec;
    const float m_x = gr[0], m_y = gr[1], m_xx = gr[2], m_yy = gr[3], m_xy = gr[4], m_1 = gr[5];  // moments of dL/dpower
    const float drgb[3] = {gr[6], gr[7], gr[8]};

    // ---- opacity = sigmoid(logit); alpha = o e^power, so dL/do = sum(dL/dpower) / o ----
    const float o = sigmoid_ref(logit);
    gop = m_1 * (1.0f - o);

    // ---- colour: rgb_c = clamp(sum_k basis_k(dir) sh[k][c] + 0.5, 0, 1) ----
    float dirv[3] = {p[0] - cam.cam_center[0], p[1] - cam.cam_center[1], p[2] - cam.cam_center[2]};
    const float dlen = sqrtf(dirv[0] * dirv[0] + dirv[1] * dirv[1] + dirv[2] * dirv[2]);
    const float x = dirv[0] / dlen, y = dirv[1] / dlen, z = dirv[2] / dlen;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    float basis[16], bx[16], by[16], bz[16];  // value and d/d(x,y,z) of every basis function
#pragma unroll
    for (int k = 0; k < 16; ++k) basis[k] = bx[k] = by[k] = bz[k] = 0.0f;
    basis[0] = kSH0;
    if (degree > 0) {
      basis[1] = -kSH1 * y; by[1] = -kSH1;
      basis[2] = kSH1 * z;  bz[2] = kSH1;
      basis[3] = -kSH1 * x; bx[3] = -kSH1;
    }
    if (degree > 1) {
      basis[4] = kSH2_0 * xy; bx[4] = kSH2_0 * y; by[4] = kSH2_0 * x;
      basis[5] = kSH2_1 * yz; by[5] = kSH2_1 * z; bz[5] = kSH2_1 * y;
      basis[6] = kSH2_2 * (2.0f * zz - xx - yy); bx[6] = kSH2_2 * -2.0f * x; by[6] = kSH2_2 * -2.0f * y; bz[6] = kSH2_2 * 4.0f * z;
      basis[7] = kSH2_3 * xz; bx[7] = kSH2_3 * z; bz[7] = kSH2_3 * x;
      basis[8] = kSH2_4 * (xx - yy); bx[8] = kSH2_4 * 2.0f * x; by[8] = kSH2_4 * -2.0f * y;
    }
    if (degree > 2) {
      basis[9] = kSH3_0 * y * (3.0f * xx - yy); bx[9] = kSH3_0 * 6.0f * xy; by[9] = kSH3_0 * (3.0f * xx - 3.0f * yy);
      basis[10] = kSH3_1 * xy * z; bx[10] = kSH3_1 * yz; by[10] = kSH3_1 * xz; bz[10] = kSH3_1 * xy;
      basis[11] = kSH3_2 * y * (4.0f * zz - xx - yy); bx[11] = kSH3_2 * -2.0f * xy; by[11] = kSH3_2 * (4.0f * zz - xx - 3.0f * yy); bz[11] = kSH3_2 * 8.0f * yz;
      basis[12] = kSH3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy); bx[12] = kSH3_3 * -6.0f * xz; by[12] = kSH3_3 * -6.0f * yz; bz[12] = kSH3_3 * (6.0f * zz - 3.0f * xx - 3.0f * yy);
      basis[13] = kSH3_4 * x * (4.0f * zz - xx - yy); bx[13] = kSH3_4 * (4.0f * zz - 3.0f * xx - yy); by[13] = kSH3_4 * -2.0f * xy; bz[13] = kSH3_4 * 8.0f * xz;
      basis[14] = kSH3_5 * z * (xx - yy); bx[14] = kSH3_5 * 2.0f * xz; by[14] = kSH3_5 * -2.0f * yz; bz[14] = kSH3_5 * (xx - yy);
      basis[15] = kSH3_6 * x * (xx - 3.0f * yy); bx[15] = kSH3_6 * (3.0f * xx - 3.0f * yy); by[15] = kSH3_6 * -6.0f * xy;
    }
    float ddir[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float raw = 0.5f, gx = 0.f, gy = 0.f, gz = 0.f;
      const float *shp = g.sh + i * g.sh_sn + c * g.sh_sc;
#pragma unroll
      for (int k = 0; k < ncoef; ++k) {
        const float coef = __ldg(shp);
        shp += g.sh_sk;
        raw += basis[k] * coef;
        gx += bx[k] * coef; gy += by[k] * coef; gz += bz[k] * coef;
      }
      const float draw = (raw >= 0.0f && raw <= 1.0f) ? drgb[c] : 0.0f;  // clamp(0,1) passes gradient inside only
      float *out = dsh + c * dsh_sc;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        *out = k < ncoef ? basis[k] * draw : 0.0f;
        out += dsh_sk;
      }
      ddir[0] += gx * draw; ddir[1] += gy * draw; ddir[2] += gz * draw;
    }
    {  // dir = d/|d|
      const float dot = x * ddir[0] + y * ddir[1] + z * ddir[2];
      gp[0] += (ddir[0] - x * dot) / dlen;
      gp[1] += (ddir[1] - y * dot) / dlen;
      gp[2] += (ddir[2] - z * dot) / dlen;
    }

    // ---- forward intermediates of the geometry ----
    float camp[3], clip[4];
#pragma unroll
    for (int c = 0; c < 3; ++c) camp[c] = row_affine(cam.w2c_t, c, p[0], p[1], p[2]);
#pragma unroll
    for (int c = 0; c < 4; ++c) clip[c] = row_affine(cam.full_t, c, p[0], p[1], p[2]);
    float S[9], c2[4];
    covariance3d(ls, rot, S);
    covariance2d(S, camp, cam, c2);

    // ---- power = -(a dx^2 + c dy^2)/2 - b dx dy with d = mean2d - pixel: the blend's moments become
    //      d mean2d = -(a m_x + b m_y, c m_y + b m_x), d(a, c, b) = (-m_xx/2, -m_yy/2, -m_xy) ----
    float dmean_x = 0.f, dmean_y = 0.f;
    const float da = -0.5f * m_xx, dc = -0.5f * m_yy, db = -m_xy;
    {
      float ca, cc, cb;
      conic_from_cov(c2, ca, cc, cb);
      dmean_x = -(ca * m_x + cb * m_y);
      dmean_y = -(cc * m_y + cb * m_x);
    }

    // ---- mean2d = ((clip.xy * pw + 1) * (W,H) - 1) / 2, pw = 1/(clip.w + 1e-7) ----
    float dclip[4] = {0.f, 0.f, 0.f, 0.f};
    {
      const float pw = 1.0f / (clip[3] + 0.0000001f);
      const float dndx = dmean_x * 0.5f * (float)cam.width, dndy = dmean_y * 0.5f * (float)cam.height;
      dclip[0] = dndx * pw;
      dclip[1] = dndy * pw;
      dclip[3] = -(dndx * clip[0] + dndy * clip[1]) * pw * pw;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
      gp[k] += cam.full_t[k * 4 + 0] * dclip[0] + cam.full_t[k * 4 + 1] * dclip[1] + cam.full_t[k * 4 + 3] * dclip[3];

    // ---- conic (a, c, b) = (c11, c00, -c01) / det ----
    const float det = c2[0] * c2[3] - c2[2] * c2[1];
    float dP[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    if (det != 0.0f) {
      const float inv = 1.0f / det;
      const float d_inv = da * c2[3] + dc * c2[0] - db * c2[1];
      const float d_det = -inv * inv * d_inv;
      dP[1][1] = da * inv + d_det * c2[0];
      dP[0][0] = dc * inv + d_det * c2[3];
      dP[0][1] = -db * inv - d_det * c2[2];
      dP[1][0] = -d_det * c2[1];
    }

    // ---- P = T V T^T with T = J Rc (rows 0,1), V = upper triangle of Sigma symmetrised ----
    const float zc = camp[2], rz = 1.0f / zc;
    const float u = camp[0] * rz, v = camp[1] * rz;
    const float uc = fminf(cam.limx, fmaxf(-cam.limx, u)), vc = fminf(cam.limy, fmaxf(-cam.limy, v));
    const float tx = uc * zc, ty = vc * zc;
    const float J00 = cam.focal_x * rz, J02 = -cam.focal_x * tx * rz * rz;
    const float J11 = cam.focal_y * rz, J12 = -cam.focal_y * ty * rz * rz;
    float Rc[3][3];  // Rc[k][i] = w2c_t[i][k]
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int ii = 0; ii < 3; ++ii) Rc[k][ii] = cam.w2c_t[ii * 4 + k];
    float Tm[2][3];
#pragma unroll
    for (int ii = 0; ii < 3; ++ii) {
      Tm[0][ii] = J00 * Rc[0][ii] + J02 * Rc[2][ii];
      Tm[1][ii] = J11 * Rc[1][ii] + J12 * Rc[2][ii];
    }
    const float V[3][3] = {{S[0], S[1], S[2]}, {S[1], S[4], S[5]}, {S[2], S[5], S[8]}};
    float TV[2][3], TVt[2][3];  // T V and T V^T (V symmetric: identical, kept separate for clarity)
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int cidx = 0; cidx < 3; ++cidx) {
        TV[r][cidx] = Tm[r][0] * V[0][cidx] + Tm[r][1] * V[1][cidx] + Tm[r][2] * V[2][cidx];
        TVt[r][cidx] = Tm[r][0] * V[cidx][0] + Tm[r][1] * V[cidx][1] + Tm[r][2] * V[cidx][2];
      }
    float dT[2][3], dV[3][3];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int cidx = 0; cidx < 3; ++cidx)
        dT[r][cidx] = dP[r][0] * TVt[0][cidx] + dP[r][1] * TVt[1][cidx] + dP[0][r] * TV[0][cidx] + dP[1][r] * TV[1][cidx];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int cidx = 0; cidx < 3; ++cidx)
        dV[r][cidx] = Tm[0][r] * (dP[0][0] * Tm[0][cidx] + dP[0][1] * Tm[1][cidx]) +
                      Tm[1][r] * (dP[1][0] * Tm[0][cidx] + dP[1][1] * Tm[1][cidx]);
    // J entries: dJ[j][k] = sum_i dT[j][i] Rc[k][i]
    const float dJ00 = dT[0][0] * Rc[0][0] + dT[0][1] * Rc[0][1] + dT[0][2] * Rc[0][2];
    const float dJ02 = dT[0][0] * Rc[2][0] + dT[0][1] * Rc[2][1] + dT[0][2] * Rc[2][2];
    const float dJ11 = dT[1][0] * Rc[1][0] + dT[1][1] * Rc[1][1] + dT[1][2] * Rc[1][2];
    const float dJ12 = dT[1][0] * Rc[2][0] + dT[1][1] * Rc[2][1] + dT[1][2] * Rc[2][2];
    float dcam[3] = {0.f, 0.f, 0.f};
    {
      const float rz2 = rz * rz, rz3 = rz2 * rz;
      dcam[2] += -cam.focal_x * rz2 * dJ00 - cam.focal_y * rz2 * dJ11;
      const float dtx = -cam.focal_x * rz2 * dJ02, dty = -cam.focal_y * rz2 * dJ12;
      dcam[2] += 2.0f * cam.focal_x * tx * rz3 * dJ02 + 2.0f * cam.focal_y * ty * rz3 * dJ12;
      // tx = clamp(x/z) * z
      const float duc = dtx * zc, dvc = dty * zc;
      dcam[2] += dtx * uc + dty * vc;
      const float du = (u >= -cam.limx && u <= cam.limx) ? duc : 0.0f;
      const float dv = (v >= -cam.limy && v <= cam.limy) ? dvc : 0.0f;
      dcam[0] += du * rz;
      dcam[1] += dv * rz;
      dcam[2] += -(du * camp[0] + dv * camp[1]) * rz2;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
      gp[k] += cam.w2c_t[k * 4 + 0] * dcam[0] + cam.w2c_t[k * 4 + 1] * dcam[1] + cam.w2c_t[k * 4 + 2] * dcam[2];

    // ---- V <- upper triangle of Sigma;  Sigma = M M^T, M = R diag(s) ----
    float dS[3][3] = {{dV[0][0], dV[0][1] + dV[1][0], dV[0][2] + dV[2][0]},
                      {0.f, dV[1][1], dV[1][2] + dV[2][1]},
                      {0.f, 0.f, dV[2][2]}};
    const float s[3] = {exp_scale(ls[0]), exp_scale(ls[1]), exp_scale(ls[2])};
    const float nrm = sqrtf(rot[0] * rot[0] + rot[1] * rot[1] + rot[2] * rot[2] + rot[3] * rot[3]);
    const float den = nrm > 1e-12f ? nrm : 1e-12f;
    const float qw = rot[0] / den, qx = rot[1] / den, qy = rot[2] / den, qz = rot[3] / den;
    float R[3][3];
    quat_to_rot(qw, qx, qy, qz, R);
    float M[3][3], dM[3][3], dR[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int cidx = 0; cidx < 3; ++cidx) M[r][cidx] = R[r][cidx] * s[cidx];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int cidx = 0; cidx < 3; ++cidx) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) acc += (dS[r][k] + dS[k][r]) * M[k][cidx];
        dM[r][cidx] = acc;
      }
#pragma unroll
    for (int cidx = 0; cidx < 3; ++cidx) {
      float acc = 0.f;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        dR[r][cidx] = dM[r][cidx] * s[cidx];
        acc += dM[r][cidx] * R[r][cidx];
      }
      gs[cidx] = acc * s[cidx];  // d/d log-scale
    }
    float dq[4];
    dq[0] = 2.0f * (-qz * dR[0][1] + qy * dR[0][2] + qz * dR[1][0] - qx * dR[1][2] - qy * dR[2][0] + qx * dR[2][1]);
    dq[1] = 2.0f * (qy * dR[0][1] + qz * dR[0][2] + qy * dR[1][0] - 2.0f * qx * dR[1][1] - qw * dR[1][2] + qz * dR[2][0] +
                    qw * dR[2][1] - 2.0f * qx * dR[2][2]);
    dq[2] = 2.0f * (-2.0f * qy * dR[0][0] + qx * dR[0][1] + qw * dR[0][2] + qx * dR[1][0] + qz * dR[1][2] - qw * dR[2][0] +
                    qz * dR[2][1] - 2.0f * qy * dR[2][2]);
    dq[3] = 2.0f * (-2.0f * qz * dR[0][0] - qw * dR[0][1] + qx * dR[0][2] + qw * dR[1][0] - 2.0f * qz * dR[1][1] + qy * dR[1][2] +
                    qx * dR[2][0] + qy * dR[2][1]);
    if (nrm > 1e-12f) {
      const float dot = qw * dq[0] + qx * dq[1] + qy * dq[2] + qz * dq[3];
      gq[0] = (dq[0] - qw * dot) / den;
      gq[1] = (dq[1] - qx * dot) / den;
      gq[2] = (dq[2] - qy * dot) / den;
      gq[3] = (dq[3] - qz * dot) / den;
    } else {
      gq[0] = dq[0] / den; gq[1] = dq[1] / den; gq[2] = dq[2] / den; gq[3] = dq[3] / den;
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) d_xyz[i * 3 + c] = gp[c];
#pragma unroll
  for (int c = 0; c < 3; ++c) d_scales[i * 3 + c] = gs[c];
#pragma unroll
  for (int c = 0; c < 4; ++c) d_rots[i * 4 + c] = gq[c];
  d_opacity[i] = gop;
}

}  // namespace gsr

using namespace gsr;

extern "C" int gsr_blend_bwd(const float *rec, const uint32_t *tile_gids, const uint32_t *tile_ranges, int32_t width,
                             int32_t height, int32_t tile_row_begin, int32_t tile_row_end, const float *final_T,
                             const uint32_t *n_contrib, const float *dL_dimage, float *grad_rec, void *stream) {
  GSR_REQUIRE(rec && tile_gids && tile_ranges && final_T && n_contrib && dL_dimage && grad_rec, "null pointer");
  GSR_REQUIRE(width > 0 && height > 0, "bad image size");
  const TileGrid tg = make_grid(width, height);
  if (tile_row_begin < 0) tile_row_begin = 0;
  if (tile_row_end > tg.tiles_y || tile_row_end < 0) tile_row_end = tg.tiles_y;
  if (tile_row_end <= tile_row_begin) return GSR_OK;
  dim3 grid(tg.tiles_x, tile_row_end - tile_row_begin);
  blend_bwd_kernel<<<grid, kBwdThreads, 0, (cudaStream_t)stream>>>((const float4 *)rec, tile_gids,
                                                                  (const uint2 *)tile_ranges, width, height, tg.tiles_x,
                                                                  tile_row_begin, final_T, n_contrib, dL_dimage, grad_rec, -0.0f);
  GSR_LAUNCH_CHECK();
  return GSR_OK;
}

extern "C" int gsr_preprocess_bwd(const gsr_gaussians *g, const gsr_camera *cam, const uint32_t *tiles_touched,
                                  const float *grad_rec, float *d_xyz, float *d_log_scales, float *d_rots,
                                  float *d_opacity_logit, float *d_sh, int64_t dsh_sn, int64_t dsh_sk, int64_t dsh_sc,
                                  void *stream) {
  GSR_REQUIRE(g && cam && tiles_touched && grad_rec && d_xyz && d_log_scales && d_rots && d_opacity_logit && d_sh,
              "null pointer");
  if (g->n == 0) return GSR_OK;
  GSR_REQUIRE(cam->sh_degree >= 0 && cam->sh_degree <= 3, "sh_degree must be 0..3");
#define GSR_PRE_BWD(D)                                                                                     \
  preprocess_bwd_kernel<D><<<(unsigned)((g->n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(              \
      *g, *cam, tiles_touched, grad_rec, d_xyz, d_log_scales, d_rots, d_opacity_logit, d_sh, dsh_sn, dsh_sk, dsh_sc)
  switch (cam->sh_degree) {
    case 0: GSR_PRE_BWD(0); break;
    case 1: GSR_PRE_BWD(1); break;
    case 2: GSR_PRE_BWD(2); break;
    default: GSR_PRE_BWD(3); break;
  }
#undef GSR_PRE_BWD
  GSR_LAUNCH_CHECK();
  return GSR_OK;
}
