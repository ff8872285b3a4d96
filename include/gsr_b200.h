/*
 * gsr_b200.h - C ABI of libgsr_b200.so, the B200 (sm_100a) render path that replaces the
 * reference's per-Gaussian PyTorch loop.
 *
 * The reference (arnaudstiegler/torch-gaussian-splatting-rasterizer) has no FFI: its interface for
 * this path is a set of module-level Python functions.  Each entry point below names the reference
 * lines it replaces (citations into /root/reference/); INTEGRATION.md shows the ctypes stubs a
 * maintainer of the reference would add.
 *
 * Conventions
 *  - every function returns a gsr_status (0 = ok); no C++ exceptions cross the boundary;
 *    gsr_last_error() returns a thread-local message for the last non-zero status.
 *  - all data pointers are DEVICE pointers unless the name ends in _host; the caller owns every
 *    buffer (inputs, outputs, workspace).  Nothing is allocated and nothing synchronises inside the
 *    device entry points: work is enqueued on `stream` (a cudaStream_t passed as void*).
 *  - thread safety: the device entry points keep no state between calls (device-specific caches are per device and
 *    race-free), so different host threads may enqueue frames at the same time as long as each uses its own workspace,
 *    outputs and stream (tests/test_gpu_guards.py: two threads, two streams).  They act on the CURRENT device.
 *  - float tensors are f32; strides are in ELEMENTS, so any torch view can be passed without a copy.
 *  - Gaussian parameters are the raw 3DGS PLY values: log-scales, un-normalised quaternions (w,x,y,z),
 *    opacity logits, SH coefficients indexed [n, k(16), c(3)].
 */
#ifndef GSR_B200_H
#define GSR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSR_ABI_VERSION 2
#define GSR_TILE 16 /* BLOCK_SIZE, rasterize.py:34 */

typedef enum gsr_status {
  GSR_OK = 0,
  GSR_ERR_INVALID_ARGUMENT = 1,
  GSR_ERR_CUDA = 2,
  GSR_ERR_WORKSPACE_TOO_SMALL = 3,
  GSR_ERR_DUP_OVERFLOW = 4 /* reported through gsr_frame_info.status, see gsr_render_fwd */
} gsr_status;

/* Per-view constants, computed on the host exactly as the reference does (rasterize.py:336-364). */
typedef struct gsr_camera {
  float w2c_t[16];     /* get_world_to_camera_matrix(q,t).T, row-major (row-vector convention)  :361 */
  float full_t[16];    /* w2c_t @ get_projection_matrix(fovx,fovy).T                            :364 */
  float cam_center[3]; /* inverse(w2c_t)[3,:3]                                 spherical_harmonics.py:35 */
  float focal_x, focal_y; /* f32(fx/2), f32(fy/2)                                               :216 */
  float limx, limy;    /* f32(1.3*tan(fov/2))                                                   :210-211 */
  int32_t width, height; /* rendered image size                                                 :338 */
  int32_t sh_degree;   /* 0..3 (the reference hard-codes 3,                                     :368) */
  int32_t reserved;
} gsr_camera;

/* Strided views of the Gaussian parameters (element strides). */
typedef struct gsr_gaussians {
  int64_t n;
  const float *xyz;           int64_t xyz_sn, xyz_sc;       /* [n,3]                    :354-356 */
  const float *log_scales;    int64_t scale_sn, scale_sc;   /* [n,3]  scale_0..2        :97-99   */
  const float *rots;          int64_t rot_sn, rot_sc;       /* [n,4]  rot_0..3 (w,x,y,z) :100-109 */
  const float *opacity_logit; int64_t opacity_sn;           /* [n]    opacity           :358     */
  const float *sh;            int64_t sh_sn, sh_sk, sh_sc;  /* [n,16,3] utils.py:21-31 (planar [3,16,n] view is the fast layout) */
} gsr_gaussians;

/* Written by the device at the end of a frame (device memory; copy back when convenient).  The caller zeroes it
 * once.  `status` and `n_dup_peak` are STICKY: no entry point ever clears them, so one read-back after a whole
 * sequence of frames (an orbit, a batch of views) still sees an overflow that happened on any of them, and
 * `n_dup_peak` says which capacity would have sufficed.  The caller clears them after handling the error.
 * `n_visible` / `n_dup` describe the LAST frame; `warp_evals` / `pixel_evals` accumulate (gsr_blend_fwd only). */
typedef struct gsr_frame_info {
  uint32_t status;        /* STICKY gsr_status: GSR_ERR_DUP_OVERFLOW when n_dup exceeded max_dup (that frame's image is invalid) */
  uint32_t n_visible;     /* Gaussians the reference's loop blends (not culled, non-zero conic, non-empty tile rect; :441) */
  uint32_t n_dup;         /* (Gaussian, tile) pairs the last frame needed (may exceed max_dup) */
  uint32_t n_dup_peak;    /* STICKY maximum of n_dup over the frames since the caller last cleared it */
  uint64_t warp_evals;    /* Gaussian-pixel evaluations executed by blend, counted at warp granularity (x32) */
  uint64_t pixel_evals;   /* evaluations executed by pixels that had not yet terminated */
} gsr_frame_info;

/* Per-Gaussian record consumed by the blend kernel: 3 x float4, 16-byte aligned. */
#define GSR_REC_FLOATS 12
/*  [0] mean2d.x  [1] mean2d.y  [2] A = -a/2   [3] C = -c/2     (a, c, b) = sigmas, rasterize.py:404-411
 *  [4] B = -b    [5] thr       [6] opacity    [7] r            power = (A dx^2 + C dy^2) + (B dx) dy
 *  [8] g         [9] b(lue)    [10] log2(o)   [11] radius      thr = -ln(255 opacity) - 1e-4: alpha <= 1/255 below it */

const char *gsr_last_error(void);
int gsr_abi_version(void);
/* SM count etc. of the current device; returns GSR_ERR_CUDA when no device is usable. */
int gsr_device_info(int *sm_count, int *cc_major, int *cc_minor, size_t *l2_bytes);
/* Measured FP32 FMA throughput of the current device (TFLOP/s, FMA = 2) and the SM clock it implies: the
 * denominator of the blend stage's roofline.  Synchronises; diagnostics only. */
int gsr_fp32_fma_peak(double *tflops, double *sm_mhz_effective);
/* Diagnostics: choose the blend kernel's staging path (0 = cp.async per CTA, 1 = cp.async.bulk + mbarrier, 2 = cp.async
 * per warp, the default) at run time; stage < 0 returns to the built-in default (or the GSR_BLEND_STAGE environment
 * variable).  `pretest` is ignored (a variant that was measured and removed).  All variants produce bit-identical
 * images. */
int gsr_debug_blend_variant(int stage, int pretest);

/* ---------------------------------------------------------------------------------------------
 * Whole frame: rasterize.py:354-452 for one view.  image[H,W,3] is the reference's
 * screen.transpose(1,0) (:471); final_T[H,W] (may be NULL) is its opacity_buffer transposed;
 * n_contrib[H,W] (may be NULL) is the position in the pixel's tile list after which blending stopped.
 * first_rank/last_rank restrict blending to Gaussians whose global depth rank r satisfies
 * first_rank <= r < last_rank (progressive renders, :448-452); pass 0 and UINT32_MAX for everything.
 * tile_row_begin/end restrict the frame to a band of tile rows (multi-GPU band sharding); rows outside
 * the band are left untouched in image.  Frames of up to 65536 tiles (e.g. 4096x4096) sort their (Gaussian, tile) pairs
 * in two 8-bit passes, larger ones (8K: 512 x 270 tiles) in three.  info_dev (device, may be NULL) receives counters;
 * it is never cleared by
 * the library (see gsr_frame_info).  Workspace offsets of everything gsr_render_bwd replays do not depend on max_dup.
 * --------------------------------------------------------------------------------------------- */
int gsr_render_workspace_bytes(int64_t n, int32_t width, int32_t height, int64_t max_dup, size_t *bytes);
int gsr_render_fwd(const gsr_gaussians *g, const gsr_camera *cam, int64_t max_dup, void *workspace,
                   size_t workspace_bytes, float *image, float *final_T, uint32_t *n_contrib,
                   uint32_t first_rank, uint32_t last_rank, int32_t tile_row_begin, int32_t tile_row_end,
                   gsr_frame_info *info_dev, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Prepared model: the view-independent half of the per-Gaussian set-up, computed once per model.
 * Sigma = R S S^T R^T (:97-120), sigmoid(opacity) (:358) and the two logarithms of the blend record do not depend on
 * the camera; gsr_prepare evaluates them with the same device functions (same operation order, bit-identical results)
 * the per-view kernel of gsr_preprocess_fwd uses, and stores them with xyz and the SH coefficients in blocks of
 * 32 Gaussians whose rows are 32 x 16 bytes:
 *     geometry [blocks][3][32][4] f32 :  x y z opacity | S00 S01 S02 S11 | S12 S22 thr log2(opacity)
 *     SH       [blocks][12][32][4] f32:  row 4c + j = coefficients 4j .. 4j+3 of colour channel c
 * (geometry first, SH at the next 256-byte boundary; 240 B per Gaussian).  `prepared` must be 256-byte aligned and
 * gsr_prepared_bytes(n) large.  A resident scene (video / orbit / serving) is prepared once and rendered with
 * gsr_render_prepared_fwd / gsr_render_views_fwd; whoever changes a parameter calls gsr_prepare again.
 * --------------------------------------------------------------------------------------------- */
int gsr_prepared_bytes(int64_t n, size_t *bytes);
int gsr_prepare(const gsr_gaussians *g, void *prepared, size_t prepared_bytes, void *stream);
/* gsr_preprocess_fwd over a prepared model (no debug outputs; depth_hist is mandatory).  Bit-identical outputs. */
int gsr_preprocess_prepared_fwd(const void *prepared, int64_t n, const gsr_camera *cam, int32_t tile_row_begin,
                                int32_t tile_row_end, uint32_t *depth_key, uint32_t *tiles_touched, uint32_t *tile_span,
                                float *rec, uint32_t *depth_hist, void *stream);
/* gsr_render_fwd over a prepared model (same workspace, same outputs, bit-identical image). */
int gsr_render_prepared_fwd(const void *prepared, int64_t n, const gsr_camera *cam, int64_t max_dup, void *workspace,
                            size_t workspace_bytes, float *image, float *final_T, uint32_t *n_contrib,
                            uint32_t first_rank, uint32_t last_rank, int32_t tile_row_begin, int32_t tile_row_end,
                            gsr_frame_info *info_dev, void *stream);
/* A batch of views of one prepared model (the 256-view orbit of :448-466's video path, a serving batch): view v is
 * enqueued on streams[v % n_lanes] using workspaces[v % n_lanes] (each workspace_bytes large) and info_dev[v % n_lanes];
 * images + v * image_stride_floats receives frame v.  cams_host is a HOST array of n_views cameras (all the same image
 * size).  Nothing synchronises: the caller joins the streams, then reads info_dev[0..n_lanes) (status is sticky, so one
 * read-back covers every frame of the lane). */
int gsr_render_views_fwd(const void *prepared, int64_t n, const gsr_camera *cams_host, int32_t n_views, int64_t max_dup,
                         void *const *workspaces, size_t workspace_bytes, int32_t n_lanes, void *const *streams,
                         float *images, int64_t image_stride_floats, gsr_frame_info *info_dev);

/* Same frame with HOST buffers: uploads the parameters, renders, downloads the image, synchronises.
 * This is the call that stands where the reference's run_rasterization body stands (:354-452).  One-shot
 * convenience: it allocates and frees its device buffers on every call. */
int gsr_render_host(const gsr_gaussians *g_host, const gsr_camera *cam, int64_t max_dup, float *image_host,
                    gsr_frame_info *info_host);

/* Host-buffer frames as a pipeline (the reference keeps its Gaussians in host tensors and moves them with .to(device),
 * :353-358; a caller in that position renders through this object).  All device memory, three streams and the events
 * are created once.  A frame is PCIe-bound (59 floats per Gaussian up, one image down); the renderer hides the rest:
 * the upload is cut into chunks of Gaussians and the per-Gaussian preprocess kernel runs on each chunk as soon as it
 * has landed, binning + blend follow the last chunk, the image goes down on a third stream, and up to `depth` frames
 * are in flight so the upload of frame k+1 overlaps the render and read-back of frame k.
 *   g_host: dense host arrays xyz[n,3], log_scales[n,3], rots[n,4], opacity_logit[n] (unit strides) and SH either as
 *   the planar [3,16,n] block (strides 1, n, 16n - what read_color_components builds, utils.py:21-31) or dense [n,16,3].
 *   Pinned host memory gives asynchronous copies; pageable memory works but serialises.
 * submit() enqueues one frame and returns a ticket; the host buffers (inputs and image) must stay valid and unchanged
 * until wait(ticket) returns.  At most `depth` tickets may be outstanding: submit() refuses (GSR_ERR_INVALID_ARGUMENT)
 * when the slot it would reuse has not been waited for - nothing is recycled silently, every frame's status is read.
 * wait() returns GSR_ERR_DUP_OVERFLOW (image invalid) when the frame needed more than max_dup pairs; info_host (may be
 * NULL) receives the frame's counters.  render() = submit + wait.  Not thread-safe per object.  The object is bound to the
 * device that was current in create(); its other calls may be made with any device current and leave it unchanged. */
typedef struct gsr_host_renderer gsr_host_renderer;
int gsr_host_renderer_create(int64_t n_max, int32_t width, int32_t height, int64_t max_dup, int32_t depth,
                             gsr_host_renderer **out);
int gsr_host_renderer_destroy(gsr_host_renderer *h);
int gsr_host_renderer_submit(gsr_host_renderer *h, const gsr_gaussians *g_host, const gsr_camera *cam,
                             float *image_host, int32_t *ticket);
int gsr_host_renderer_wait(gsr_host_renderer *h, int32_t ticket, gsr_frame_info *info_host);
int gsr_host_renderer_render(gsr_host_renderer *h, const gsr_gaussians *g_host, const gsr_camera *cam,
                             float *image_host, gsr_frame_info *info_host);

/* ---------------------------------------------------------------------------------------------
 * Stages (each usable on its own; the tests drive them one by one)
 * --------------------------------------------------------------------------------------------- */

/* Fused per-Gaussian preprocess: :97-120 (Sigma), :358 (sigmoid), :371-420 (projection, cull, EWA, radius,
 * tile rect, conic, skip rule :441) and spherical_harmonics.py:27-73.
 *  depth_key[n]      u32, order-preserving bits of z_cam (:424)
 *  tiles_touched[n]  u32: bit 31 = the reference's loop would blend this Gaussian (not culled / skipped / off
 *                    screen, :441); bits 0..30 = number of tiles of its rect in which at least one pixel can pass
 *                    alpha > 1/255 (exact tile culling, see gsr_math.cuh) - the image is unchanged by the culling
 *  tile_span[n,4]    u32: x0 | y0 << 16, x1 | y1 << 16 (tile rect clipped to the tile grid and to
 *                    [row_begin,row_end)), then the 64-bit keep mask of the rect's tiles, row-major, for rects of
 *                    <= 64 tiles (larger rects keep every tile).  16-byte aligned: the binning stage gathers it
 *                    with one load per Gaussian
 *  rec[n,12]         blend record (written for visible Gaussians only)
 * optional debug outputs (may be NULL): cov3d[n,9], cov2d[n,4] (after cull zeroing), tile_rect_raw[n,4] i32
 * (the reference's compute_covering_bbox result), mean2d_all[n,2]; depth_hist[4*256] u32 receives the 8-bit digit
 * histograms of the depth keys of the Gaussians the reference's loop blends (input of gsr_bin's one-sweep sort). */
int gsr_preprocess_fwd(const gsr_gaussians *g, const gsr_camera *cam, int32_t tile_row_begin, int32_t tile_row_end,
                       uint32_t *depth_key, uint32_t *tiles_touched, uint32_t *tile_span, float *rec, float *cov3d,
                       float *cov2d, int32_t *tile_rect_raw, float *mean2d_all, uint32_t *depth_hist, void *stream);

/* spherical_harmonics.sh_to_rgb (spherical_harmonics.py:27-73): rgb[n,3] contiguous. */
int gsr_sh_fwd(int64_t n, const float *xyz, int64_t xyz_sn, int64_t xyz_sc, const float *sh, int64_t sh_sn,
               int64_t sh_sk, int64_t sh_sc, const float *cam_center_host3, int32_t degree, float *rgb, void *stream);

/* Stable LSB radix sort of (u32 key, u32 value) pairs on bits [begin_bit, end_bit).  values_in may be NULL
 * (then values are 0..n-1).  Result lands in keys_out/values_out.  n_dev (device u32, may be NULL) overrides n
 * with a device-side count <= n.  Scratch comes from workspace (gsr_sort_workspace_bytes). */
int gsr_sort_workspace_bytes(int64_t n, size_t *bytes);
int gsr_sort_pairs(const uint32_t *keys_in, const uint32_t *values_in, uint32_t *keys_out, uint32_t *values_out,
                   uint32_t *keys_tmp, uint32_t *values_tmp, int64_t n, const uint32_t *n_dev, int32_t begin_bit,
                   int32_t end_bit, void *workspace, size_t workspace_bytes, void *stream);

/* Binning: global stable depth sort (:424-425, ties by Gaussian index), duplication into (tile, Gaussian) pairs
 * in depth order, stable sort by tile, tile range extraction.  Equivalent to sorting 64-bit (tile|depth) keys.
 *  depth_hist[4*256]  u32 digit histograms of the blended Gaussians' depth keys as written by gsr_preprocess_fwd
 *                     (in, may be NULL: they are then counted here)
 *  tile_span[n,4]     u32 tile rect + keep mask as written by gsr_preprocess_fwd (in)
 *  depth_order[n]     u32 ids of the Gaussians the reference's loop blends, in ascending depth; only the first
 *                     info_dev->n_visible entries are written            (out, may be NULL: kept in the workspace)
 *  blend_rank[n]      u32 rank of each blended Gaussian in that order (out, may be NULL; feeds progressive
 *                     renders; entries of the other Gaussians are left untouched)
 *  tile_gids[max_dup] u32 Gaussian ids grouped by tile, depth-sorted      (out)
 *  tile_ranges[tiles,2] u32 [begin,end) into tile_gids; (0,0) for tiles without pairs (out)
 *  info_dev->n_visible / n_dup / status are filled (info_dev may be NULL). */
int gsr_bin_workspace_bytes(int64_t n, int32_t width, int32_t height, int64_t max_dup, size_t *bytes);
int gsr_bin(int64_t n, const uint32_t *depth_key, const uint32_t *depth_hist, const uint32_t *tiles_touched,
            const uint32_t *tile_span, int32_t width, int32_t height, int64_t max_dup, uint32_t *depth_order,
            uint32_t *blend_rank, uint32_t *tile_gids, uint32_t *tile_ranges, gsr_frame_info *info_dev, void *workspace,
            size_t workspace_bytes, void *stream);

/* Per-16x16-tile front-to-back alpha blend (:255-305, :436-446).  Pixels x >= W-1 or y >= H-1 are written
 * as zero (the reference never writes them).  When blend_rank is given, only Gaussians with
 * first_rank <= blend_rank < last_rank are blended (the state of the reference's screen after last_rank
 * loop iterations that were not skipped, :448-452). */
int gsr_blend_fwd(const float *rec, const uint32_t *tile_gids, const uint32_t *tile_ranges, int32_t width,
                  int32_t height, int32_t tile_row_begin, int32_t tile_row_end, const uint32_t *blend_rank,
                  uint32_t first_rank, uint32_t last_rank, float *image, float *final_T, uint32_t *n_contrib,
                  gsr_frame_info *info_dev, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Backward (the reference has none: the contract is torch.autograd through its forward, :255-305, :371-420 and
 * spherical_harmonics.py:27-73)
 * --------------------------------------------------------------------------------------------- */
/* Blend backward: the tile lists replayed back to front.  final_T / n_contrib are the forward's outputs,
 * dL_dimage[H,W,3] the incoming gradient; grad_rec[n,9] (zeroed by the caller) accumulates per Gaussian the sums
 * over its pixels of dL/dpower x (dx, dy, dx^2, dy^2, dx dy, 1) and dL/d(r,g,b) - everything the blend contributes
 * to d mean2d, d conic and d opacity is linear in them; gsr_preprocess_bwd does that conversion. */
int gsr_blend_bwd(const float *rec, const uint32_t *tile_gids, const uint32_t *tile_ranges, int32_t width,
                  int32_t height, int32_t tile_row_begin, int32_t tile_row_end, const float *final_T,
                  const uint32_t *n_contrib, const float *dL_dimage, float *grad_rec, void *stream);
/* Preprocess + SH backward: grad_rec -> gradients of the raw parameters (contiguous [n,3], [n,3], [n,4], [n]);
 * d_sh is written with the given element strides ([n,16,3] logical, e.g. the strides of the sh input). */
int gsr_preprocess_bwd(const gsr_gaussians *g, const gsr_camera *cam, const uint32_t *tiles_touched,
                       const float *grad_rec, float *d_xyz, float *d_log_scales, float *d_rots, float *d_opacity_logit,
                       float *d_sh, int64_t dsh_sn, int64_t dsh_sk, int64_t dsh_sc, void *stream);
/* Whole-frame backward.  Must follow gsr_render_fwd of the same frame with the same, untouched workspace.
 * grad_rec[n,9] is scratch.  In band mode every rank produces PARTIAL gradients (its band's pixels only); their
 * sum over the ranks is the full gradient. */
int gsr_render_bwd(const gsr_gaussians *g, const gsr_camera *cam, int64_t max_dup, void *workspace,
                   size_t workspace_bytes, const float *final_T, const uint32_t *n_contrib, const float *dL_dimage,
                   float *grad_rec, float *d_xyz, float *d_log_scales, float *d_rots, float *d_opacity_logit,
                   float *d_sh, int64_t dsh_sn, int64_t dsh_sk, int64_t dsh_sc, int32_t tile_row_begin,
                   int32_t tile_row_end, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Drop-in helpers behind the reference's module-level functions
 * --------------------------------------------------------------------------------------------- */
/* get_covariance_matrix_from_mesh (:89-120): cov3d[n,3,3]. */
int gsr_cov3d_fwd(int64_t n, const float *log_scales, int64_t s_sn, int64_t s_sc, const float *rots, int64_t r_sn,
                  int64_t r_sc, float *cov3d, void *stream);
/* quaternion_to_rotation_matrix (:41-56): q[4,n] strided -> R[3,3,n] contiguous, no normalisation. */
int gsr_quat_to_rotmat(int64_t n, const float *q, int64_t q_sc, int64_t q_sn, float *rot, void *stream);
/* project_to_camera_space (:80-86) and the clip-space product (:374): out[n,cols] = p @ M[:3,:cols] + M[3,:cols],
 * M a HOST row-major 4x4. */
int gsr_affine_fwd(int64_t n, const float *xyz, int64_t sn, int64_t sc, const float *m_host16, int32_t cols,
                   float *out, void *stream);
/* compute_2d_covariance (:201-252): cov3d[n,3,3] + camera-space means[n,3] -> cov2d[n,2,2]. */
int gsr_cov2d_fwd(int64_t n, const float *cov3d, const float *cam_means, const gsr_camera *cam, float *cov2d,
                  void *stream);
/* compute_covering_bbox (:154-198): -> tile rect [n,4] int64. */
int gsr_bbox_fwd(int64_t n, const float *mean2d, const float *cov2d, int32_t width, int32_t height, int64_t *rect,
                 void *stream);
/* PLY vertex block -> kernel layout (replaces the host-side column stacking of :353-358, :97-109 and utils.py:21-31).
 * vertices: device copy of the raw little-endian f32 vertex records, stride_floats floats per vertex;
 * columns_host59: float index inside a record of x y z, scale_0..2, rot_0..3, opacity, f_dc_0..2, f_rest_0..44.
 * Outputs: xyz[n,3], log_scales[n,3], rots[n,4], opacity_logit[n], sh_planar[3,16,n]. */
int gsr_ply_unpack(const float *vertices, int64_t n, int32_t stride_floats, const int32_t *columns_host59, float *xyz,
                   float *log_scales, float *rots, float *opacity_logit, float *sh_planar, void *stream);
/* rasterize_gaussian (:255-305): blend ONE Gaussian into the x-major screen[W,H,3] / opacity_buffer[W,H]. */
int gsr_blend_one(const int64_t *bbox4_host, const float *mean2d_host2, const float *conic_host3,
                  const float *rgb_host3, float opacity, int32_t width, int32_t height, float *screen,
                  float *opacity_buffer, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GSR_B200_H */
