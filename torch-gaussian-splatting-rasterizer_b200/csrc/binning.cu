// binning.cu - from per-Gaussian (depth key, tile rect) to per-tile, depth-ordered Gaussian lists.
//
// The textbook route sorts N_dup 64-bit (tile | depth) keys: ceil((32+log2 tiles)/8) = 6 radix passes over 12-byte
// pairs.  The same total order - (tile, depth bits, Gaussian index) - is produced here with a third of the
// traffic by sorting on the two fields separately, most significant last:
//   1. stable sort of the Gaussians the reference's loop blends by their 32-bit depth key (4 passes over 8-byte
//      pairs, N_vis << N_dup; the first pass reads the per-Gaussian keys directly and drops all the others);
//   2. walk the Gaussians in depth order, warp-shuffle prefix scan of tiles-touched with a decoupled look-back
//      across CTAs, and emit one (tile id, Gaussian id) pair per touched tile - pairs come out depth-ordered;
//   3. stable sort of the pairs by tile id only (ceil(log2 tiles / 8) = 2 passes at 1080p);
//   4. the [begin, end) range of every tile is read off the sorted tile ids (one boundary-detection pass).
// The digit histograms the one-sweep sorts need are by-products: those of the depth keys come from the preprocess
// kernel, those of the tile ids from the emission (shared-memory counts, one flush per CTA) - no counting pass.
// Stability of both sorts makes ties fall back to the Gaussian index, i.e. the result is bit-identical to a
// stable CPU sort of the 64-bit keys (tests/test_gpu_parity.py::test_binning_equals_stable_sort_of_tile_depth_keys checks exactly that).
// Replaces torch.sort at /root/reference/rasterize.py:424-425 and the implicit "for every Gaussian, for every
// pixel of its bbox" visit order of :440-446.
//
// The number of pairs never travels to the host: it is left in device memory (clamped to the caller's capacity,
// overflow flagged in gsr_frame_info.status) and every later kernel reads it from there.
#include "gsr_math.cuh"

namespace gsr {

constexpr int kScanThreads = 256;
#ifndef GSR_SCAN_ITEMS
#define GSR_SCAN_ITEMS 4
#endif
constexpr int kScanItems = GSR_SCAN_ITEMS;  // Gaussians per thread (slot = lane * kScanItems + j must fit 7 bits)
constexpr int kScanTile = kScanThreads * kScanItems;
constexpr int kStageCap = 1024;  // pairs a warp can stage in shared memory (one word each)
constexpr int kStageWords = kStageCap + 32 * kScanItems + (32 * kScanItems + 1) + 3;  // + ids + offsets, 16-byte multiple
constexpr int kTileBitsMax = 24;  // up to three 8-bit passes of the tile sort; the packed staging word holds tile << 7 | slot

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
  uint32_t n_vis;      // Gaussians the reference's loop blends = entries of the compacted depth sort
  uint32_t reserved;
};

// Tile span of one Gaussian as the preprocess kernel stores it: x0 | y0 << 16, x1 | y1 << 16, keep mask (64 bits).
// The number of pairs it stands for needs nothing else: the set bits of the mask for rects of <= 64 tiles (untested
// rects carry an all-ones mask), the whole area for larger ones.
struct Span {
  uint32_t x0, y0, w, h, area;
  unsigned long long mask;
};
__device__ __forceinline__ Span decode_span(uint4 v) {
  Span s;
  s.x0 = v.x & 0xFFFFu;
  s.y0 = v.x >> 16;
  s.w = (v.y & 0xFFFFu) - s.x0;
  s.h = (v.y >> 16) - s.y0;
  s.area = s.w * s.h;
  s.mask = ((unsigned long long)v.w << 32) | v.z;
  if (s.area < 64) s.mask &= (1ull << s.area) - 1ull;
  return s;
}
__device__ __forceinline__ uint32_t span_pairs(const Span &s) {
  return s.area <= (uint32_t)kCullMaxTiles ? (uint32_t)__popcll(s.mask) : s.area;
}

// Walks the compacted depth order: prefix scan of the pair counts (decoupled look-back across CTAs), emission of the
// (tile id, Gaussian id) pairs in depth order, blend ranks, and the digit histograms of the tile ids.
#ifndef GSR_SCAN_MINCTAS
#define GSR_SCAN_MINCTAS 4
#endif
__global__ void __launch_bounds__(kScanThreads, GSR_SCAN_MINCTAS)
scan_emit_kernel(const uint32_t *__restrict__ n_dev, const uint32_t *__restrict__ order_x0, uint32_t *__restrict__ order_x1,
                 const uint32_t *__restrict__ depth_hist, const uint4 *__restrict__ span,
                 int tiles_x, unsigned long long capacity, unsigned long long *status, BinCounters *ctr,
                 uint32_t *__restrict__ tile_keys, uint32_t *__restrict__ tile_gids, uint32_t *__restrict__ blend_rank,
                 uint32_t *__restrict__ tile_hist /*[3][256]*/, int wide_tiles /* more than 65536 tiles */, gsr_frame_info *info) {
  __shared__ unsigned long long s_warp[kScanThreads / 32];
  __shared__ unsigned long long s_base;
  __shared__ uint32_t s_tile;
  __shared__ uint32_t s_hist[3][256];
  extern __shared__ uint32_t s_dyn[];  // per warp: [kStageCap] packed pairs | [128] Gaussian ids | [129] offsets
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(&ctr->ticket, 1u);
  s_hist[0][tid] = 0;
  s_hist[1][tid] = 0;
  s_hist[2][tid] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;
  const int64_t n = (int64_t)*n_dev;  // entries of `order`: the Gaussians the reference's loop blends
  if ((int64_t)tile * kScanTile >= n && tile != 0) return;  // grid sized for the capacity; tile 0 always runs (totals for n == 0)
  const int64_t first = (int64_t)tile * kScanTile + (int64_t)tid * kScanItems;
  // The adaptive depth sort leaves the order in X0 or X1 depending on how many of its passes were identities (sort.cu);
  // when it is X0 the ids are copied to X1 on the way, so that X1 (the caller's depth_order) always holds the order.
  const bool from_x0 = n > 0 && sorted_buffer_index(depth_hist, 4, (uint32_t)n) == 0;
  const uint32_t *order = from_x0 ? order_x0 : order_x1;

  uint32_t gid[kScanItems], cnt[kScanItems];
  uint4 sp[kScanItems];
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    gid[j] = first + j < n ? __ldg(order + first + j) : 0xFFFFFFFFu;
    sp[j] = make_uint4(0u, 0u, 0u, 0u);
  }
#pragma unroll
  for (int j = 0; j < kScanItems; ++j)
    if (gid[j] != 0xFFFFFFFFu) sp[j] = __ldg(span + gid[j]);  // the only gather: 16 bytes per Gaussian
  if (from_x0) {
#pragma unroll
    for (int j = 0; j < kScanItems; ++j)
      if (gid[j] != 0xFFFFFFFFu) order_x1[first + j] = gid[j];
  }
  unsigned long long mine = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    cnt[j] = span_pairs(decode_span(sp[j]));
    mine += cnt[j];
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
  // Order of business: (1) warp 0 PUBLISHES this CTA's aggregate at once, so that no successor waits for it longer than
  // necessary; (2) every warp expands its pairs into shared-memory staging - that needs only CTA-relative offsets;
  // (3) warp 0 walks back over its predecessors (decoupled look-back, 32 per round trip: lane l inspects CTA
  // tile-1-l of the window) - by now most of them have published, and the other warps were expanding instead of
  // waiting at a barrier (ncu on the previous order: 6.4 barrier-stall cycles per issue); (4) coalesced stores.
  if (warp == 0 && lane == 0)
    st_volatile_u64(status + tile, (tile == 0 ? kScanInclusive : kScanAggregate) | total);
  // ---- emission.  The 128 Gaussians of a warp own one contiguous output range.  Every lane expands ITS Gaussians
  //      (a walk over the set bits of the keep mask: a few instructions per pair) into the warp's staging buffer in
  //      shared memory, at the offset its prefix gives; the warp then copies the buffer out with coalesced 128-byte
  //      stores.  A staged pair is one word, tile << 7 | slot of the Gaussian in the warp.  (A first version let
  //      every lane binary-search the owner of each output element: ~170 thread instructions per pair.)  Warps whose
  //      range exceeds the buffer - a few huge splats - fall back to that search, element by element.
  uint32_t *stage = s_dyn + (size_t)warp * kStageWords;
  uint32_t *s_gid = stage + kStageCap;
  uint32_t *s_off = s_gid + 32 * kScanItems;
  const unsigned long long thread_rel = pre + inc - mine;  // pairs of this CTA before this thread's first Gaussian
  const unsigned long long warp_rel = __shfl_sync(0xFFFFFFFFu, thread_rel, 0);
  const unsigned long long warp_total = __shfl_sync(0xFFFFFFFFu, thread_rel + mine, 31) - warp_rel;
  const bool staged = warp_total != 0 && warp_total <= (unsigned long long)kStageCap;
  if (blend_rank) {  // every entry of the compacted order is blended: its rank is its position
#pragma unroll
    for (int j = 0; j < kScanItems; ++j)
      if (gid[j] != 0xFFFFFFFFu) blend_rank[gid[j]] = (uint32_t)(first + j);
  }
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) s_gid[lane * kScanItems + j] = gid[j];
  if (staged) {
    uint32_t off = (uint32_t)(thread_rel - warp_rel);
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
      if (cnt[j]) {
        const Span r = decode_span(sp[j]);
        const uint32_t slot = (uint32_t)(lane * kScanItems + j);
        const uint32_t origin = r.y0 * (uint32_t)tiles_x + r.x0;
        if (r.area <= (uint32_t)kCullMaxTiles) {
          // row = bit / w through a reciprocal: any inv in [65536/w, 65536/w + 2) gives the exact quotient for
          // bit < 64, w <= 64 (the excess bit*(inv - 65536/w)/65536 < 1/64 never carries), and the approximate
          // division is within 0.016 of 65536/w, whose fraction is 0 or >= 1/64 - so trunc + 1 lands in that interval.
          // The mask is walked as two 32-bit halves (64-bit ffs / clear-lowest cost 25 instructions per pair).
          const uint32_t inv = __float2uint_rz(__fdividef(65536.0f, (float)r.w)) + 1u;
          auto put = [&](uint32_t bit) {
            const uint32_t row = (bit * inv) >> 16, col = bit - row * r.w;
            stage[off++] = ((origin + row * (uint32_t)tiles_x + col) << 7) | slot;
          };
          uint32_t lo = (uint32_t)r.mask, hi = (uint32_t)(r.mask >> 32);
          while (lo) {
            put((uint32_t)__ffs((int)lo) - 1u);
            lo &= lo - 1u;
          }
          while (hi) {
            put((uint32_t)__ffs((int)hi) + 31u);
            hi &= hi - 1u;
          }
        } else {
          for (uint32_t ty = 0; ty < r.h; ++ty)
            for (uint32_t tx = 0; tx < r.w; ++tx) stage[off++] = ((origin + ty * (uint32_t)tiles_x + tx) << 7) | slot;
        }
      }
    }
  } else if (warp_total != 0) {
    unsigned long long run = thread_rel - warp_rel;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
      s_off[lane * kScanItems + j] = (uint32_t)(run < 0xFFFFFFFFull ? run : 0xFFFFFFFFull);
      run += cnt[j];
    }
  }
  if (warp == 0) {
    unsigned long long prefix = 0;
    if (tile != 0) {
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
      const int64_t tiles_total = n > 0 ? (n + kScanTile - 1) / kScanTile : 1;
      if ((int64_t)tile == tiles_total - 1) {  // last CTA in depth order knows the totals
        const unsigned long long dup = prefix + total;
        ctr->n_dup = (uint32_t)(dup < capacity ? dup : capacity);
        if (info) {
          const uint32_t dup32 = dup < 0xFFFFFFFFull ? (uint32_t)dup : 0xFFFFFFFFu;
          info->n_dup = dup32;
          info->n_visible = (uint32_t)n;
          atomicMax(&info->n_dup_peak, dup32);                       // sticky: frames of other streams may share the word
          if (dup > capacity) info->status = GSR_ERR_DUP_OVERFLOW;  // sticky: only the caller clears it
        }
      }
    }
  }
  __syncthreads();  // s_base is known; every warp's staging buffer (written by its own lanes) is complete
  const unsigned long long warp_out = s_base + warp_rel;
  if (staged) {
    const unsigned long long room = capacity > warp_out ? capacity - warp_out : 0ull;  // pairs beyond it are dropped
    const uint32_t limit = (uint32_t)(warp_total < room ? warp_total : room);          // warp-uniform, <= kStageCap
    uint32_t *keys_out = tile_keys + warp_out, *gids_out = tile_gids + warp_out;
    for (uint32_t e = lane; e < limit; e += 32) {
      const uint32_t pr = stage[e], tile_id = pr >> 7;
      keys_out[e] = tile_id;
      gids_out[e] = s_gid[pr & 127u];
      atomicAdd(&s_hist[0][tile_id & 255u], 1u);
      atomicAdd(&s_hist[1][(tile_id >> 8) & 255u], 1u);
      if (wide_tiles) atomicAdd(&s_hist[2][tile_id >> 16], 1u);  // third pass: frames above 65536 tiles (uniform branch)
    }
  } else if (warp_total != 0) {
    // ---- fallback for a warp with more pairs than the staging buffer holds ----
    const unsigned long long room = capacity > warp_out ? capacity - warp_out : 0ull;  // pairs beyond it are dropped
    const uint32_t limit = (uint32_t)(warp_total < room ? warp_total : room);          // <= capacity < 2^30
    for (uint32_t e = lane; e < limit; e += 32) {
      int lo = 0, hi = 32 * kScanItems - 1;  // last slot with s_off[slot] <= e
#pragma unroll
      for (int it = 0; it < 7; ++it) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_off[mid] <= e) lo = mid; else hi = mid - 1;
      }
      const uint32_t g_id = s_gid[lo];
      const Span r = decode_span(__ldg(span + g_id));
      uint32_t local = e - s_off[lo];
      if (r.area <= (uint32_t)kCullMaxTiles) local = nth_set_bit((uint32_t)r.mask, (uint32_t)(r.mask >> 32), local);
      const uint32_t row = local / r.w, col = local - row * r.w;
      const uint32_t tile_id = (r.y0 + row) * (uint32_t)tiles_x + r.x0 + col;
      tile_keys[warp_out + e] = tile_id;
      tile_gids[warp_out + e] = g_id;
      atomicAdd(&s_hist[0][tile_id & 255u], 1u);
      atomicAdd(&s_hist[1][(tile_id >> 8) & 255u], 1u);
      if (wide_tiles) atomicAdd(&s_hist[2][tile_id >> 16], 1u);
    }
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const uint32_t v = s_hist[q][tid];
    if (v) atomicAdd(tile_hist + q * 256 + tid, v);
  }
}

// [begin, end) of every tile's list, read off the sorted tile ids: position i opens the list of its tile when the
// id differs from its left neighbour's and closes it when it differs from the right one's.  Tiles without pairs keep
// the (0, 0) the caller memset.
constexpr int kRangeThreads = 256;
__global__ void __launch_bounds__(kRangeThreads)
tile_ranges_kernel(const uint32_t *__restrict__ sorted_tiles, const BinCounters *ctr, uint2 *__restrict__ ranges) {
  const uint32_t n = ctr->n_dup;
  const uint32_t stride = gridDim.x * kRangeThreads * 4;
  // four consecutive ids per thread (one 16-byte load) plus the two neighbours of the quad
  for (uint32_t i = (blockIdx.x * kRangeThreads + threadIdx.x) * 4; i < n; i += stride) {
    uint32_t t[6];
    if (i + 4 <= n) {
      const uint4 q = __ldg(reinterpret_cast<const uint4 *>(sorted_tiles + i));
      t[1] = q.x; t[2] = q.y; t[3] = q.z; t[4] = q.w;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) t[1 + k] = i + k < n ? __ldg(sorted_tiles + i + k) : 0xFFFFFFFFu;
    }
    t[0] = i > 0 ? __ldg(sorted_tiles + i - 1) : 0xFFFFFFFFu;       // no tile id equals the sentinel
    t[5] = i + 4 < n ? __ldg(sorted_tiles + i + 4) : 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i + k < n) {
        if (t[k] != t[k + 1]) ranges[t[k + 1]].x = i + k;
        if (t[k + 2] != t[k + 1]) ranges[t[k + 1]].y = i + k + 1;
      }
    }
  }
}

struct BinLayout {
  size_t counters, scan_status, tile_hist, depth_sort, tile_sort, order, dkeys_tmp,
      dkeys_out, dvals_tmp, tkeys_a, tkeys_b, tkeys_c, tgids_a, tgids_b, total;
};

static BinLayout bin_layout(int64_t n, int64_t max_dup) {
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
  L.tile_hist = take(sizeof(uint32_t) * 4 * 256);
  L.depth_sort = take(sort_scratch_bytes(n1));
  L.tile_sort = take(sort_scratch_bytes(d1));
  L.order = take(4 * (size_t)n1);
  L.dkeys_tmp = take(4 * (size_t)n1);
  L.dkeys_out = take(4 * (size_t)n1);
  L.dvals_tmp = take(4 * (size_t)n1);
  L.tkeys_a = take(4 * (size_t)d1);
  L.tkeys_b = take(4 * (size_t)d1);
  L.tkeys_c = take(4 * (size_t)d1);
  L.tgids_a = take(4 * (size_t)d1);
  L.tgids_b = take(4 * (size_t)d1);
  L.total = off;
  return L;
}

}  // namespace gsr

using namespace gsr;

extern "C" int gsr_bin_workspace_bytes(int64_t n, int32_t width, int32_t height, int64_t max_dup, size_t *bytes) {
  GSR_REQUIRE(bytes && n >= 0 && max_dup >= 0 && width > 0 && height > 0, "bad arguments");
  *bytes = bin_layout(n, max_dup).total;
  return GSR_OK;
}

extern "C" int gsr_bin(int64_t n, const uint32_t *depth_key, const uint32_t *depth_hist, const uint32_t *tiles_touched,
                       const uint32_t *tile_span, int32_t width, int32_t height, int64_t max_dup, uint32_t *depth_order,
                       uint32_t *blend_rank, uint32_t *tile_gids, uint32_t *tile_ranges, gsr_frame_info *info_dev,
                       void *workspace, size_t workspace_bytes, void *stream) {
  return bin_impl(n, depth_key, depth_hist, tiles_touched, tile_span, width, height, max_dup, depth_order, blend_rank,
                  tile_gids, tile_ranges, info_dev, workspace, workspace_bytes, (cudaStream_t)stream, false);
}

// the zero-initialised head of the bin workspace: counters, look-back status, tile histograms, both sort scratches
size_t gsr::bin_zero_bytes(int64_t n, int64_t max_dup) { return bin_layout(n, max_dup).order; }

int gsr::bin_impl(int64_t n, const uint32_t *depth_key, const uint32_t *depth_hist, const uint32_t *tiles_touched,
                  const uint32_t *tile_span, int32_t width, int32_t height, int64_t max_dup, uint32_t *depth_order,
                  uint32_t *blend_rank, uint32_t *tile_gids, uint32_t *tile_ranges, gsr_frame_info *info_dev,
                  void *workspace, size_t workspace_bytes, cudaStream_t stream, bool zeroed) {
  GSR_REQUIRE(depth_key && tiles_touched && tile_span && tile_gids && tile_ranges && workspace, "null pointer");
  GSR_REQUIRE(((uintptr_t)tile_span & 15) == 0, "tile_span must be 16-byte aligned");
  GSR_REQUIRE(n >= 0 && n < (1ll << 30) && max_dup > 0 && max_dup < (1ll << 30), "n / max_dup out of range");
  GSR_REQUIRE(width > 0 && height > 0, "bad image size");
  const TileGrid tg = make_grid(width, height);
  const int num_tiles = tg.tiles_x * tg.tiles_y;
  int tile_bits = 1;
  while ((1 << tile_bits) < num_tiles) ++tile_bits;
  GSR_REQUIRE(tile_bits <= kTileBitsMax, "more than 2^24 tiles are not supported");
  const BinLayout L = bin_layout(n, max_dup);
  if (workspace_bytes < L.total || ((uintptr_t)workspace & 255)) {
    set_error("bin workspace too small or not 256-byte aligned (%zu < %zu)", workspace_bytes, L.total);
    return GSR_ERR_WORKSPACE_TOO_SMALL;
  }
  char *ws = (char *)workspace;
  BinCounters *ctr = (BinCounters *)(ws + L.counters);
  unsigned long long *scan_status = (unsigned long long *)(ws + L.scan_status);
  uint32_t *tile_hist = (uint32_t *)(ws + L.tile_hist);
  // counters + look-back status words + tile digit histograms + the scratch of both sorts are contiguous: one memset
  // (the whole-frame driver clears them together with its own zero-initialised buffers)
  if (!zeroed) {
    GSR_CHECK_CUDA(cudaMemsetAsync(ws + L.counters, 0, L.order - L.counters, stream));
    GSR_CHECK_CUDA(cudaMemsetAsync(tile_ranges, 0, sizeof(uint2) * (size_t)num_tiles, stream));
  }
  if (n == 0) {  // no scan runs: the per-frame counters are cleared here (status / n_dup_peak stay as they are)
    if (info_dev) GSR_CHECK_CUDA(cudaMemsetAsync(&info_dev->n_visible, 0, 2 * sizeof(uint32_t), stream));
    return GSR_OK;
  }
  if (!depth_order) depth_order = (uint32_t *)(ws + L.order);

  // 1. depth order of the Gaussians the reference's loop blends (bit 31 of tiles_touched), ties by index.  The first
  //    radix pass reads the per-Gaussian keys directly and drops everything else (a compaction for free: the sort
  //    and the scan below never see the culled / off-screen ones); digit counts come from the preprocess kernel
  //    when given.
  SortScratch s1, s2;
  int rc = sort_scratch_carve(ws + L.depth_sort, L.tile_sort - L.depth_sort, n, &s1);
  if (rc) return rc;
  uint32_t *vis_count = &ctr->n_vis;
  rc = launch_sort_pairs(depth_key, nullptr, (uint32_t *)(ws + L.dkeys_out), depth_order, (uint32_t *)(ws + L.dkeys_tmp),
                         (uint32_t *)(ws + L.dvals_tmp), n, vis_count, 0, 32, s1, stream, true, depth_hist, tiles_touched,
                         vis_count, true, true);
  const uint32_t *depth_hist_used = depth_hist ? depth_hist : s1.hist;
  if (rc) return rc;

  // 2. prefix scan of the pair counts in depth order + pair emission + digit histograms of the tile ids
  uint32_t *tkeys_a = (uint32_t *)(ws + L.tkeys_a), *tgids_a = (uint32_t *)(ws + L.tgids_a);
  const size_t stage_bytes = sizeof(uint32_t) * kStageWords * (kScanThreads / 32);
  scan_emit_kernel<<<(unsigned)ceil_div64(n, kScanTile), kScanThreads, stage_bytes, stream>>>(
      vis_count, (const uint32_t *)(ws + L.dvals_tmp), depth_order, depth_hist_used, (const uint4 *)tile_span, tg.tiles_x,
      (unsigned long long)max_dup, scan_status, ctr, tkeys_a,
      tgids_a, blend_rank, tile_hist, tile_bits > 16 ? 1 : 0, info_dev);
  GSR_LAUNCH_CHECK();

  // 3. stable sort of the pairs by tile id
  rc = sort_scratch_carve(ws + L.tile_sort, L.order - L.tile_sort, max_dup, &s2);
  if (rc) return rc;
  uint32_t *sorted_tiles = (uint32_t *)(ws + L.tkeys_c);
  rc = launch_sort_pairs(tkeys_a, tgids_a, sorted_tiles, tile_gids, (uint32_t *)(ws + L.tkeys_b),
                         (uint32_t *)(ws + L.tgids_b), max_dup, &ctr->n_dup, 0, tile_bits, s2, stream, false, tile_hist, nullptr,
                         nullptr, true, false);
  if (rc) return rc;

  // 4. tile ranges from the sorted ids
  const int64_t want = ceil_div64(max_dup, kRangeThreads * 4);  // four ids per thread
  tile_ranges_kernel<<<(unsigned)(want > 148 * 16 ? 148 * 16 : (want > 0 ? want : 1)), kRangeThreads, 0, stream>>>(
      sorted_tiles, ctr, (uint2 *)tile_ranges);
  GSR_LAUNCH_CHECK();
  return GSR_OK;
}
