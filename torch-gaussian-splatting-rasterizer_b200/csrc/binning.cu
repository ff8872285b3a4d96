// binning.cu - from per-Gaussian (depth key, tile rect) to per-tile, depth-ordered Gaussian lists.
//
// The textbook route sorts N_dup 64-bit (tile | depth) keys: ceil((32+log2 tiles)/8) = 6 radix passes over 12-byte
// pairs.  The same total order - (tile, depth bits, Gaussian index) - is produced here with a third of the
// traffic by sorting on the two fields separately, most significant last:
//   1. stable sort of the N Gaussians by their 32-bit depth key (4 passes over 8-byte pairs, N << N_dup);
//   2. walk the Gaussians in depth order, warp-shuffle prefix scan of tiles-touched with a decoupled look-back
//      across CTAs, and emit one (tile id, Gaussian id) pair per touched tile - pairs come out depth-ordered;
//   3. stable sort of the pairs by tile id only (ceil(log2 tiles / 8) = 2 passes at 1080p);
//   4. the [begin, end) range of every tile comes from the per-tile list lengths counted during emission, whose
//      marginals are also the digit histograms of the tile sort (no counting pass over the pairs).
// Stability of both sorts makes ties fall back to the Gaussian index, i.e. the result is bit-identical to a
// stable CPU sort of the 64-bit keys (tests/test_gpu_binning.py checks exactly that).
// Replaces torch.sort at /root/reference/rasterize.py:424-425 and the implicit "for every Gaussian, for every
// pixel of its bbox" visit order of :440-446.
//
// The number of pairs never travels to the host: it is left in device memory (clamped to the caller's capacity,
// overflow flagged in gsr_frame_info.status) and every later kernel reads it from there.
#include "gsr_math.cuh"

namespace gsr {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;

constexpr unsigned long long kScanAggregate = 1ull << 62;
constexpr unsigned long long kScanInclusive = 2ull << 62;
constexpr unsigned long long kScanValue = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct BinCounters {   // device words shared by the binning kernels
  uint32_t n_dup;      // pairs actually stored (<= capacity): the n of the tile sort and of range marking
  uint32_t ticket;     // dynamic CTA index of scan_emit
  uint32_t pad[2];
};

// packed scan element: low 31 bits = tiles touched, bits 31..61 = visible count
__device__ __forceinline__ unsigned long long pack_scan(uint32_t tiles, uint32_t vis) {
  return (unsigned long long)tiles | ((unsigned long long)vis << 31);
}

__global__ void __launch_bounds__(kScanThreads)
scan_emit_kernel(int64_t n, const uint32_t *__restrict__ order, const uint32_t *__restrict__ tiles_touched,
                 const ushort4 *__restrict__ rect, const uint2 *__restrict__ tile_mask, int tiles_x, uint32_t capacity,
                 unsigned long long *status,
                 BinCounters *ctr, uint32_t *__restrict__ tile_keys, uint32_t *__restrict__ tile_gids,
                 uint32_t *__restrict__ blend_rank, uint32_t *__restrict__ tile_count, gsr_frame_info *info) {
  __shared__ unsigned long long s_warp[kScanThreads / 32];
  __shared__ unsigned long long s_base;
  __shared__ uint32_t s_tile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(&ctr->ticket, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const int64_t first = (int64_t)tile * kScanTile + (int64_t)tid * kScanItems;

  uint32_t gid[kScanItems], cnt[kScanItems], vis[kScanItems];
  unsigned long long mine = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    const int64_t pos = first + j;
    gid[j] = 0;
    cnt[j] = 0;
    vis[j] = 0;
    if (pos < n) {
      gid[j] = __ldg(order + pos);
      const uint32_t t = __ldg(tiles_touched + gid[j]);
      cnt[j] = t & 0x7FFFFFFFu;  // tiles kept after culling
      vis[j] = t >> 31;          // the reference's loop would blend it (rank numbering of progressive renders)
    }
    mine += pack_scan(cnt[j], vis[j]);
  }
  // warp-shuffle inclusive scan, then across warps
  unsigned long long inc = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long up = __shfl_up_sync(0xFFFFFFFFu, inc, o);
    if (lane >= o) inc += up;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  unsigned long long pre = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w) {
    if (w < warp) pre += s_warp[w];
    total += s_warp[w];
  }
  // decoupled look-back by warp 0, 32 predecessors per round trip (lane l inspects CTA tile-1-l of the window);
  // the exclusive prefix is broadcast through shared memory
  if (warp == 0) {
    unsigned long long prefix = 0;
    if (tile == 0) {
      if (lane == 0) st_volatile_u64(status, kScanInclusive | total);
    } else {
      if (lane == 0) st_volatile_u64(status + tile, kScanAggregate | total);
      int64_t win = (int64_t)tile - 1;  // newest predecessor of the current window
      while (true) {
        const int64_t mine_idx = win - lane;
        unsigned long long v = kScanInclusive;  // before tile 0: an inclusive prefix of zero
        if (mine_idx >= 0) {
          v = ld_volatile_u64(status + mine_idx);
          while ((v >> 62) == 0) {  // not published yet: yield the issue slots to working warps
            __nanosleep(40);
            v = ld_volatile_u64(status + mine_idx);
          }
        }
        const uint32_t incl = __ballot_sync(0xFFFFFFFFu, (v & kScanInclusive) != 0);
        const int stop = incl ? __ffs(incl) - 1 : 32;  // nearest predecessor that already knows its inclusive prefix
        unsigned long long part = lane <= stop ? (v & kScanValue) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
        prefix += part;
        if (incl) break;
        win -= 32;
      }
      if (lane == 0) st_volatile_u64(status + tile, kScanInclusive | (prefix + total));
    }
    if (lane == 0) {
      s_base = prefix;
      const int64_t tiles_total = (n + kScanTile - 1) / kScanTile;
      if ((int64_t)tile == tiles_total - 1) {  // last CTA in depth order knows the totals
        const unsigned long long all = prefix + total;
        const uint32_t dup = (uint32_t)(all & 0x7FFFFFFFull), vis = (uint32_t)(all >> 31);
        ctr->n_dup = dup < capacity ? dup : capacity;
        if (info) {
          info->n_dup = dup;
          info->n_visible = vis;
          if (dup > capacity) info->status = GSR_ERR_DUP_OVERFLOW;
        }
      }
    }
  }
  __syncthreads();
  // ---- warp-cooperative emission: the 128 Gaussians of a warp own one contiguous output range; lanes walk it
  //      element by element (coalesced 128-byte stores) and find their Gaussian by binary search ----
  __shared__ uint32_t s_off[kScanThreads / 32][32 * kScanItems + 1];
  __shared__ uint32_t s_gid[kScanThreads / 32][32 * kScanItems];
  __shared__ ushort4 s_rect[kScanThreads / 32][32 * kScanItems];
  __shared__ uint2 s_mask[kScanThreads / 32][32 * kScanItems];
  const unsigned long long thread_excl = s_base + pre + inc - mine;  // exclusive prefix of this thread's first item
  const unsigned long long warp_excl = __shfl_sync(0xFFFFFFFFu, thread_excl, 0);
  const uint32_t warp_out = (uint32_t)(warp_excl & 0x7FFFFFFFull);
  {
    unsigned long long run = thread_excl;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
      const int slot = lane * kScanItems + j;
      s_off[warp][slot] = (uint32_t)(run & 0x7FFFFFFFull) - warp_out;
      s_gid[warp][slot] = gid[j];
      if (cnt[j]) {
        s_rect[warp][slot] = __ldg(rect + gid[j]);
        s_mask[warp][slot] = __ldg(tile_mask + gid[j]);
      }
      if (vis[j] && blend_rank) blend_rank[gid[j]] = (uint32_t)(run >> 31);  // rank among the Gaussians the reference blends
      run += pack_scan(cnt[j], vis[j]);
    }
    if (lane == 31) s_off[warp][32 * kScanItems] = (uint32_t)(run & 0x7FFFFFFFull) - warp_out;
  }
  __syncwarp();
  const uint32_t warp_total = s_off[warp][32 * kScanItems];
  for (uint32_t e = lane; e < warp_total; e += 32) {
    int lo = 0, hi = 32 * kScanItems;  // last slot with s_off[slot] <= e
#pragma unroll
    for (int it = 0; it < 7; ++it) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_off[warp][mid] <= e) lo = mid; else hi = mid - 1;
    }
    const ushort4 r = s_rect[warp][lo];
    uint32_t local = e - s_off[warp][lo];
    const uint32_t w = (uint32_t)(r.z - r.x);
    if (w * (uint32_t)(r.w - r.y) <= (uint32_t)kCullMaxTiles) {  // local-th KEPT tile of the keep mask (all ones if untested)
      const uint2 km = s_mask[warp][lo];
      local = nth_set_bit(km.x, km.y, local);
    }
    const uint32_t row = local / w, col = local - row * w;
    const uint32_t out = warp_out + e;
    if (out < capacity) {
      const uint32_t tile = (r.y + row) * (uint32_t)tiles_x + r.x + col;
      tile_keys[out] = tile;
      tile_gids[out] = s_gid[warp][lo];
      atomicAdd(tile_count + tile, 1u);  // fire-and-forget RED: list length of the tile
    }
  }
}

// One CTA: exclusive scan of the per-tile list lengths -> [begin, end) of every tile, plus the digit histograms of
// the tile-id sort (marginals of the same table), so the pairs are never read just to be counted.
constexpr int kOffsThreads = 1024;
__global__ void __launch_bounds__(kOffsThreads)
tile_offsets_kernel(const uint32_t *__restrict__ tile_count, int num_tiles, uint2 *__restrict__ ranges,
                    uint32_t *__restrict__ hist /*[passes][256]*/, int passes, uint32_t last_mask, uint32_t capacity,
                    BinCounters *ctr, gsr_frame_info *info) {
  __shared__ uint32_t s_warp[kOffsThreads / 32];
  __shared__ uint32_t s_hist[2][256];
  __shared__ uint32_t s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 512) (&s_hist[0][0])[tid] = 0;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < num_tiles; base += kOffsThreads) {
    const int t = base + tid;
    const uint32_t c = t < num_tiles ? tile_count[t] : 0u;
    uint32_t inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, inc, o);
      if (lane >= o) inc += up;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t pre = s_carry;
    for (int w = 0; w < warp; ++w) pre += s_warp[w];
    if (t < num_tiles) {
      ranges[t] = make_uint2(pre + inc - c, pre + inc);
      if (c) {
        atomicAdd(&s_hist[0][t & 255], c);
        if (passes > 1) atomicAdd(&s_hist[1][(t >> 8) & 255 & (passes == 2 ? last_mask : 255u)], c);
      }
    }
    __syncthreads();
    if (tid == kOffsThreads - 1) s_carry = pre + inc;
    __syncthreads();
  }
  if (passes > 0 && tid < 256) {
    hist[tid] = s_hist[0][tid];
    if (passes > 1) hist[256 + tid] = s_hist[1][tid];
  }
  if (ctr) {  // bucket path: this kernel is the first to know the number of pairs
    const uint32_t total = s_carry;
    const bool overflow = total > capacity;
    if (overflow)  // no list may point past the pair buffer: the frame is flagged and drawn empty
      for (int t = tid; t < num_tiles; t += kOffsThreads) ranges[t] = make_uint2(0u, 0u);
    if (tid == 0) {
      ctr->n_dup = overflow ? 0u : total;
      if (info) {
        info->n_dup = total;
        if (overflow) info->status = GSR_ERR_DUP_OVERFLOW;
      }
    }
  }
}

// ---- bucket path (frames without a rank window) ---------------------------------------------------------------
// With the list length of every tile known (counted by the preprocess kernel), each (Gaussian, tile) pair can be
// dropped straight into its tile's segment through an atomic cursor - no global depth sort, no scan - as the
// 64-bit word (depth key << 32 | Gaussian id).  The segment of a tile is then sorted ON CHIP by that word, which
// is exactly the order (depth bits, Gaussian index): the arbitrary arrival order of the atomics cannot show.
constexpr int kEmitThreads = 256;
__global__ void __launch_bounds__(kEmitThreads)
emit_pairs_kernel(int64_t n, const uint32_t *__restrict__ tiles_touched, const ushort4 *__restrict__ rect,
                  const uint2 *__restrict__ tile_mask, const uint32_t *__restrict__ depth_key,
                  const uint2 *__restrict__ ranges, uint32_t *__restrict__ cursor, const BinCounters *ctr,
                  unsigned long long *__restrict__ pairs, int tiles_x) {
  constexpr int kWarps = kEmitThreads / 32;
  __shared__ int s_pref[kWarps][33];
  __shared__ ushort4 s_rect[kWarps][32];
  __shared__ uint2 s_mask[kWarps][32];
  __shared__ uint32_t s_key[kWarps][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t stored = ctr->n_dup;  // 0 when the frame overflowed
  if (stored == 0) return;
  const int64_t stride = (int64_t)gridDim.x * kEmitThreads;
  for (int64_t base = (int64_t)blockIdx.x * kEmitThreads + warp * 32; base < n; base += stride) {
    const int64_t i = base + lane;
    const int cnt = i < n ? (int)(__ldg(tiles_touched + i) & 0x7FFFFFFFu) : 0;
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xFFFFFFFFu, inc, o);
      if (lane >= o) inc += up;
    }
    const int total = __shfl_sync(0xFFFFFFFFu, inc, 31);
    if (total == 0) continue;
    __syncwarp();
    s_pref[warp][lane] = inc - cnt;
    if (cnt) {
      s_rect[warp][lane] = __ldg(rect + i);
      s_mask[warp][lane] = __ldg(tile_mask + i);
      s_key[warp][lane] = __ldg(depth_key + i);
    }
    __syncwarp();
    for (int e = lane; e < total; e += 32) {
      int lo = 0, hi = 31;
#pragma unroll
      for (int it = 0; it < 5; ++it) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_pref[warp][mid] <= e) lo = mid; else hi = mid - 1;
      }
      const ushort4 r = s_rect[warp][lo];
      uint32_t local = (uint32_t)(e - s_pref[warp][lo]);
      const uint32_t w = (uint32_t)(r.z - r.x);
      if (w * (uint32_t)(r.w - r.y) <= (uint32_t)kCullMaxTiles) {
        const uint2 km = s_mask[warp][lo];
        local = nth_set_bit(km.x, km.y, local);
      }
      const uint32_t row = local / w, col = local - row * w;
      const uint32_t tile = (r.y + row) * (uint32_t)tiles_x + r.x + col;
      const uint32_t slot = __ldg(&ranges[tile].x) + atomicAdd(cursor + tile, 1u);
      if (slot < stored) pairs[slot] = ((unsigned long long)s_key[warp][lo] << 32) | (uint32_t)(base + lo);
    }
  }
}

// Bitonic sort of one tile's segment, every compare-exchange ascending (the first step of each merge mirrors the
// block instead of reversing a direction), so positions >= n behave as +infinity without being stored and any
// length works.  GLOBAL = false: segment staged in shared memory; true: sorted in place in global memory (lists
// too long for shared memory - slow but complete).
template <int THREADS, bool GLOBAL>
__global__ void __launch_bounds__(THREADS)
tile_sort_kernel(const uint2 *__restrict__ ranges, unsigned long long *pairs, uint32_t *__restrict__ tile_gids,
                 int num_tiles, uint32_t min_len, uint32_t max_len) {
  extern __shared__ unsigned long long s_seg[];
  const int tid = threadIdx.x;
  // grid-stride over tiles: the short-list class is launched with one CTA per tile, the long-list classes with a
  // few resident CTAs that skim the range table for their (rare) members
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const uint2 r = ranges[tile];
    const uint32_t n = r.y - r.x;
    if (n < min_len || n > max_len) continue;  // another launch owns this length class (uniform per CTA)
    unsigned long long *seg = GLOBAL ? pairs + r.x : s_seg;
    if (!GLOBAL) {
      for (uint32_t i = tid; i < n; i += THREADS) seg[i] = pairs[r.x + i];
    }
    __syncthreads();
    for (uint32_t lk = 1; (1u << (lk - 1)) < n; ++lk) {       // merge blocks of k = 2^lk
      for (int lj = (int)lk - 1; lj >= 0; --lj) {              // compare distance j = 2^lj
        const uint32_t j = 1u << lj, kmask = (1u << lk) - 1u;
        const bool mirror = lj == (int)lk - 1;
        for (uint32_t c = tid;; c += THREADS) {
          const uint32_t i = ((c >> lj) << (lj + 1)) | (c & (j - 1));  // lower element of compare-exchange c
          if (i >= n) break;
          const uint32_t p = mirror ? (i ^ kmask) : (i | j);
          if (p < n) {
            const unsigned long long a = seg[i], b = seg[p];
            if (a > b) {
              seg[i] = b;
              seg[p] = a;
            }
          }
        }
        __syncthreads();
      }
    }
    for (uint32_t i = tid; i < n; i += THREADS) tile_gids[r.x + i] = (uint32_t)seg[i];
    __syncthreads();  // the segment buffer is reused by the next tile of this CTA
  }
}

// list lengths when the preprocess kernel did not count them (stand-alone use of gsr_bin)
__global__ void count_tiles_kernel(int64_t n, const uint32_t *__restrict__ tiles_touched, const ushort4 *__restrict__ rect,
                                   const uint2 *__restrict__ tile_mask, uint32_t *__restrict__ tile_count, int tiles_x) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t cnt = __ldg(tiles_touched + i) & 0x7FFFFFFFu;
  if (!cnt) return;
  const ushort4 r = __ldg(rect + i);
  const uint32_t w = (uint32_t)(r.z - r.x), area = w * (uint32_t)(r.w - r.y);
  const uint2 km = area <= (uint32_t)kCullMaxTiles ? __ldg(tile_mask + i) : make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
  for (uint32_t k = 0; k < cnt; ++k) {
    const uint32_t local = area <= (uint32_t)kCullMaxTiles ? nth_set_bit(km.x, km.y, k) : k;
    const uint32_t row = local / w, col = local - row * w;
    atomicAdd(tile_count + (r.y + row) * (uint32_t)tiles_x + r.x + col, 1u);
  }
}

struct BinLayout {
  size_t counters, scan_status, tile_count, tile_cursor, tile_hist, depth_sort, tile_sort, dkeys_tmp, dkeys_out, dvals_tmp, tkeys_a,
      tkeys_b, tkeys_c, tgids_a, tgids_b, pairs, total;
};

static BinLayout bin_layout(int64_t n, int64_t max_dup, int num_tiles) {
  BinLayout L;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t at = off;
    off = align_up(off + bytes, 256);
    return at;
  };
  const int64_t n1 = n > 0 ? n : 1, d1 = max_dup > 0 ? max_dup : 1;
  L.counters = take(sizeof(BinCounters));
  L.scan_status = take(sizeof(unsigned long long) * (size_t)ceil_div64(n1, kScanTile));
  L.tile_count = take(sizeof(uint32_t) * (size_t)(num_tiles > 0 ? num_tiles : 1));
  L.tile_cursor = take(sizeof(uint32_t) * (size_t)(num_tiles > 0 ? num_tiles : 1));
  L.tile_hist = take(sizeof(uint32_t) * 4 * 256);
  L.depth_sort = take(sort_scratch_bytes(n1));
  L.tile_sort = take(sort_scratch_bytes(d1));
  L.dkeys_tmp = take(4 * (size_t)n1);
  L.dkeys_out = take(4 * (size_t)n1);
  L.dvals_tmp = take(4 * (size_t)n1);
  L.tkeys_a = take(4 * (size_t)d1);
  L.tkeys_b = take(4 * (size_t)d1);
  L.tkeys_c = take(4 * (size_t)d1);
  L.tgids_a = take(4 * (size_t)d1);
  L.tgids_b = take(4 * (size_t)d1);
  L.pairs = take(8 * (size_t)d1);
  L.total = off;
  return L;
}

}  // namespace gsr

using namespace gsr;

extern "C" int gsr_bin_workspace_bytes(int64_t n, int32_t width, int32_t height, int64_t max_dup, size_t *bytes) {
  GSR_REQUIRE(bytes && n >= 0 && max_dup >= 0 && width > 0 && height > 0, "bad arguments");
  const TileGrid tg0 = make_grid(width, height);
  *bytes = bin_layout(n, max_dup, tg0.tiles_x * tg0.tiles_y).total;
  return GSR_OK;
}

static int launch_tile_sorts(const uint2 *ranges, unsigned long long *pairs, uint32_t *tile_gids, int num_tiles,
                             cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    GSR_CHECK_CUDA(cudaFuncSetAttribute(tile_sort_kernel<1024, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        16384 * 8));
    configured = true;
  }
  // one launch per length class; a CTA whose tile belongs to another class retires at once
  tile_sort_kernel<256, false><<<num_tiles, 256, 1024 * 8, stream>>>(ranges, pairs, tile_gids, num_tiles, 1u, 1024u);
  tile_sort_kernel<512, false><<<148 * 2, 512, 4096 * 8, stream>>>(ranges, pairs, tile_gids, num_tiles, 1025u, 4096u);
  tile_sort_kernel<1024, false><<<148, 1024, 16384 * 8, stream>>>(ranges, pairs, tile_gids, num_tiles, 4097u, 16384u);
  tile_sort_kernel<1024, true><<<148, 1024, 0, stream>>>(ranges, pairs, tile_gids, num_tiles, 16385u, 0xFFFFFFFFu);
  GSR_LAUNCH_CHECK();
  return GSR_OK;
}

extern "C" int gsr_bin(int64_t n, const uint32_t *depth_key, const uint32_t *depth_hist, const uint32_t *tiles_touched,
                       const uint16_t *rect, const uint32_t *tile_mask, const uint32_t *tile_count_in, int32_t width,
                       int32_t height, int64_t max_dup, uint32_t *depth_order, uint32_t *blend_rank, uint32_t *tile_gids,
                       uint32_t *tile_ranges, gsr_frame_info *info_dev, void *workspace, size_t workspace_bytes,
                       void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GSR_REQUIRE(depth_key && tiles_touched && rect && tile_mask && tile_gids && tile_ranges && workspace, "null pointer");
  GSR_REQUIRE((depth_order == nullptr) == (blend_rank == nullptr) || depth_order, "blend_rank needs depth_order");
  GSR_REQUIRE(n >= 0 && n < (1ll << 30) && max_dup > 0 && max_dup < (1ll << 30), "n / max_dup out of range");
  GSR_REQUIRE(width > 0 && height > 0, "bad image size");
  const TileGrid tg = make_grid(width, height);
  const int num_tiles = tg.tiles_x * tg.tiles_y;
  const BinLayout L = bin_layout(n, max_dup, num_tiles);
  if (workspace_bytes < L.total || ((uintptr_t)workspace & 255)) {
    set_error("bin workspace too small or not 256-byte aligned (%zu < %zu)", workspace_bytes, L.total);
    return GSR_ERR_WORKSPACE_TOO_SMALL;
  }
  char *ws = (char *)workspace;
  BinCounters *ctr = (BinCounters *)(ws + L.counters);
  unsigned long long *scan_status = (unsigned long long *)(ws + L.scan_status);
  uint32_t *tile_count = (uint32_t *)(ws + L.tile_count), *tile_hist = (uint32_t *)(ws + L.tile_hist);
  // counters + scan status + tile counts + cursors + tile digit histograms are contiguous: one memset
  GSR_CHECK_CUDA(cudaMemsetAsync(ws + L.counters, 0, L.depth_sort - L.counters, stream));
  if (n == 0) {
    GSR_CHECK_CUDA(cudaMemsetAsync(tile_ranges, 0, sizeof(uint2) * (size_t)num_tiles, stream));
    return GSR_OK;
  }

  if (!depth_order) {
    // ---- bucket path: no global depth order wanted (every frame without a rank window) ----
    const uint32_t *counts = tile_count_in;
    if (!counts) {
      count_tiles_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, stream>>>(n, tiles_touched, (const ushort4 *)rect,
                                                                        (const uint2 *)tile_mask, tile_count, tg.tiles_x);
      GSR_LAUNCH_CHECK();
      counts = tile_count;
    }
    tile_offsets_kernel<<<1, kOffsThreads, 0, stream>>>(counts, num_tiles, (uint2 *)tile_ranges, tile_hist, 0, 0u,
                                                        (uint32_t)max_dup, ctr, info_dev);
    GSR_LAUNCH_CHECK();
    unsigned long long *pairs = (unsigned long long *)(ws + L.pairs);
    int64_t want = ceil_div64(n, kEmitThreads);
    emit_pairs_kernel<<<(unsigned)(want > 148 * 16 ? 148 * 16 : want), kEmitThreads, 0, stream>>>(
        n, tiles_touched, (const ushort4 *)rect, (const uint2 *)tile_mask, depth_key, (const uint2 *)tile_ranges,
        (uint32_t *)(ws + L.tile_cursor), ctr, pairs, tg.tiles_x);
    GSR_LAUNCH_CHECK();
    return launch_tile_sorts((const uint2 *)tile_ranges, pairs, tile_gids, num_tiles, stream);
  }

  // ---- ordered path: also yields the global depth order / blend ranks (rank-windowed frames) ----
  // 1. global depth order (ties by Gaussian index); digit counts come from the preprocess kernel when available
  SortScratch s1, s2;
  int rc = sort_scratch_carve(ws + L.depth_sort, L.tile_sort - L.depth_sort, n, &s1);
  if (rc) return rc;
  rc = launch_sort_pairs(depth_key, nullptr, (uint32_t *)(ws + L.dkeys_out), depth_order, (uint32_t *)(ws + L.dkeys_tmp),
                         (uint32_t *)(ws + L.dvals_tmp), n, nullptr, 0, 32, s1, stream, true, depth_hist);
  if (rc) return rc;

  // 2. prefix scan of tiles-touched in depth order + pair emission + per-tile list lengths
  uint32_t *tkeys_a = (uint32_t *)(ws + L.tkeys_a), *tgids_a = (uint32_t *)(ws + L.tgids_a);
  scan_emit_kernel<<<(unsigned)ceil_div64(n, kScanTile), kScanThreads, 0, stream>>>(
      n, depth_order, tiles_touched, (const ushort4 *)rect, (const uint2 *)tile_mask, tg.tiles_x, (uint32_t)max_dup,
      scan_status, ctr, tkeys_a, tgids_a, blend_rank, tile_count, info_dev);
  GSR_LAUNCH_CHECK();

  // 3. tile ranges and the digit histograms of the tile sort, both from the list lengths
  int tile_bits = 1;
  while ((1 << tile_bits) < num_tiles) ++tile_bits;
  GSR_REQUIRE(tile_bits <= 16, "more than 65536 tiles are not supported by the two-pass tile sort");
  const int passes = (tile_bits + 7) / 8;
  const uint32_t last_mask = (1u << (tile_bits - 8 * (passes - 1))) - 1u;
  tile_offsets_kernel<<<1, kOffsThreads, 0, stream>>>(tile_count, num_tiles, (uint2 *)tile_ranges, tile_hist, passes,
                                                      last_mask, (uint32_t)max_dup, nullptr, nullptr);
  GSR_LAUNCH_CHECK();

  // 4. stable sort of the pairs by tile id
  rc = sort_scratch_carve(ws + L.tile_sort, L.dkeys_tmp - L.tile_sort, max_dup, &s2);
  if (rc) return rc;
  rc = launch_sort_pairs(tkeys_a, tgids_a, (uint32_t *)(ws + L.tkeys_c), tile_gids, (uint32_t *)(ws + L.tkeys_b),
                         (uint32_t *)(ws + L.tgids_b), max_dup, &ctr->n_dup, 0, tile_bits, s2, stream, true, tile_hist);
  return rc;
}
