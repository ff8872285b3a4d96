// ply_unpack.cu - 3DGS PLY vertex records (array of structs, as stored on disk) -> the kernel layout in HBM.
//
// The reference goes through `plyfile` and stacks 59 numpy columns on the host (rasterize.py:353-358, 97-109;
// utils.py:21-31).  Here the raw vertex block is uploaded once and de-interleaved on the GPU: xyz[n,3],
// log-scales[n,3], quaternions[n,4], opacity[n] and the planar SH block [3,16,n] that read_color_components builds.
// One warp handles 32 consecutive vertices: the records are read with coalesced loads through shared memory
// (32 x stride floats, contiguous on disk) and every output plane is written 128 bytes at a time.
#include "gsr_common.cuh"

namespace gsr {

struct PlyColumns {
  int32_t col[59];  // float index inside a vertex record of: x y z | scale_0..2 | rot_0..3 | opacity | f_dc_0..2 | f_rest_0..44
};

constexpr int kUnpackWarps = 4;

__global__ void __launch_bounds__(kUnpackWarps * 32)
ply_unpack_kernel(const float *__restrict__ vertices, int64_t n, int stride, const __grid_constant__ PlyColumns cols,
                  float *__restrict__ xyz, float *__restrict__ scales, float *__restrict__ rots,
                  float *__restrict__ opacity, float *__restrict__ sh_planar) {
  extern __shared__ float s_rec[];  // [kUnpackWarps][32 * stride]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float *mine = s_rec + (size_t)warp * 32 * stride;
  const int64_t first = ((int64_t)blockIdx.x * kUnpackWarps + warp) * 32;
  if (first >= n) return;
  const int count = (int)min((int64_t)32, n - first);
  const float *src = vertices + first * stride;
  for (int k = lane; k < count * stride; k += 32) mine[k] = __ldg(src + k);
  __syncwarp();
  if (lane < count) {
    const float *r = mine + lane * stride;  // stride is odd for the 62-float 3DGS record: conflict-free
    const int64_t i = first + lane;
#pragma unroll
    for (int c = 0; c < 3; ++c) xyz[i * 3 + c] = r[cols.col[c]];
#pragma unroll
    for (int c = 0; c < 3; ++c) scales[i * 3 + c] = r[cols.col[3 + c]];
#pragma unroll
    for (int c = 0; c < 4; ++c) rots[i * 4 + c] = r[cols.col[6 + c]];
    opacity[i] = r[cols.col[10]];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      sh_planar[((int64_t)c * 16) * n + i] = r[cols.col[11 + c]];  // sh[n,0,c] = f_dc_c
#pragma unroll
      for (int k = 1; k < 16; ++k) sh_planar[((int64_t)c * 16 + k) * n + i] = r[cols.col[14 + 15 * c + k - 1]];  // f_rest_{15c+k-1}
    }
  }
}

}  // namespace gsr

using namespace gsr;

extern "C" int gsr_ply_unpack(const float *vertices, int64_t n, int32_t stride_floats, const int32_t *columns_host59,
                              float *xyz, float *log_scales, float *rots, float *opacity_logit, float *sh_planar,
                              void *stream) {
  GSR_REQUIRE(vertices && columns_host59 && xyz && log_scales && rots && opacity_logit && sh_planar, "null pointer");
  GSR_REQUIRE(stride_floats >= 59 && stride_floats <= 256, "vertex stride must be 59..256 floats");
  PlyColumns cols;
  for (int k = 0; k < 59; ++k) {
    GSR_REQUIRE(columns_host59[k] >= 0 && columns_host59[k] < stride_floats, "column index outside the vertex record");
    cols.col[k] = columns_host59[k];
  }
  if (n == 0) return GSR_OK;
  const size_t smem = sizeof(float) * kUnpackWarps * 32 * (size_t)stride_floats;
  GSR_CHECK_CUDA(ensure_dynamic_smem<&ply_unpack_kernel>((int)(sizeof(float) * kUnpackWarps * 32 * 256)));
  const unsigned blocks = (unsigned)ceil_div64(n, kUnpackWarps * 32);
  ply_unpack_kernel<<<blocks, kUnpackWarps * 32, smem, (cudaStream_t)stream>>>(vertices, n, stride_floats, cols, xyz,
                                                                             log_scales, rots, opacity_logit, sh_planar);
  GSR_LAUNCH_CHECK();
  return GSR_OK;
}
