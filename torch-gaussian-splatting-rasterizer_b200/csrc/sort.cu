// sort.cu - stable least-significant-digit radix sort of (u32 key, u32 value) pairs, written for sm_100a.
//
// Design (one-sweep): ONE histogram kernel counts every 8-bit digit of every pass up front; each pass is then a
// single kernel in which a CTA ranks a 4096-pair tile in shared memory and learns the number of equal-digit
// pairs in all preceding tiles through a decoupled look-back over per-tile status words (no scan kernel, one
// read and one write of the data per pass).  CTAs take their tile index from an atomic ticket, so every tile a
// CTA waits on belongs to a CTA that is already resident: the look-back cannot dead-lock.
//
// Stability + determinism: ranks are assigned in memory order (warp-major, item, lane) from counts only, so the
// output is THE stable sort of the input whatever the scheduling - ties between equal depths keep ascending
// Gaussian index, which is what "bit-exact against a stable CPU sort" asks for (replaces torch.sort at
// /root/reference/rasterize.py:424-425).
//
// n may live on the device (n_dev): grids are sized for the capacity and surplus CTAs retire immediately, so the
// binning stage never synchronises with the host to learn the number of (Gaussian, tile) pairs.
#include <stdlib.h>

#include "gsr_common.cuh"

namespace gsr {

constexpr int kRadix = 256;
constexpr int kMaxPasses = 4;

constexpr uint32_t kFlagAggregate = 1u << 30;
constexpr uint32_t kFlagInclusive = 2u << 30;
constexpr uint32_t kValueMask = (1u << 30) - 1;

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t *p, uint32_t v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Lanes of the warp that hold the same 8-bit digit as the caller (8 ballots; cost independent of the number of
// distinct digits, unlike __match_any_sync).  `live` == false lanes only match each other.
__device__ __forceinline__ uint32_t match_digit(uint32_t d, bool live = true) {
  uint32_t peers = __ballot_sync(0xFFFFFFFFu, live);
  peers = live ? peers : ~peers;
  // Four instructions per bit (test into a predicate, vote, mask from the predicate, one LOP3: peers & ~(vote ^ mask)).
  // Written as PTX because the C form compiles to six (a shift and a second compare per bit) and these passes are bound
  // by the integer pipe.
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 t, v, m;\n\t"
        "and.b32 t, %1, %2;\n\t"
        "setp.ne.u32 p, t, 0;\n\t"
        "vote.sync.ballot.b32 v, p, 0xffffffff;\n\t"
        "selp.b32 m, 0xffffffff, 0, p;\n\t"
        "lop3.b32 %0, %0, v, m, 0x90;\n\t"
        "}"
        : "+r"(peers)
        : "r"(d), "r"(1u << b));
  }
  return peers;
}

// Shared memory through 32-bit addresses computed once per thread.  Indexing the __shared__ arrays directly made the
// compiler rebuild the shared-window base (S2R CgaCtaId, LEA ...) at every use inside the unrolled per-key loops - ten
// integer instructions per key in kernels that are bound by the integer pipe.
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((uint16_t)v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// ---- digit histograms of all passes, one read of the keys --------------------------------------------------
// Shared-memory atomics cost ~2 cycles per lane whatever the address pattern, which made an atomic histogram the
// slowest kernel of the sort.  Here each warp owns a private histogram and the lowest lane of every group of
// equal digits adds the group's size with a plain read-modify-write (groups have distinct digits, so no two
// lanes of the warp touch the same word).
constexpr int kHistWarps = 8;
__global__ void __launch_bounds__(kHistWarps * 32) radix_hist_kernel(const uint32_t *__restrict__ keys, int64_t n_host,
                                                                     const uint32_t *__restrict__ n_dev,
                                                                     const uint32_t *__restrict__ live_flags,
                                                                     uint32_t *__restrict__ hist, int begin_bit,
                                                                     int passes, uint32_t last_mask) {
  extern __shared__ uint32_t s_hist[];  // [kHistWarps][passes][kRadix]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < kHistWarps * passes * kRadix; k += blockDim.x) s_hist[k] = 0;
  __syncthreads();
  uint32_t *mine = s_hist + warp * passes * kRadix;
  const int64_t n = n_dev ? (int64_t)*n_dev : n_host;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x + warp * 32; base < n; base += stride) {
    const int64_t i = base + lane;
    const bool live = i < n && (!live_flags || (__ldg(live_flags + i) >> 31));
    const uint32_t key = live ? __ldg(keys + i) : 0u;
    for (int p = 0; p < passes; ++p) {
      uint32_t d = (key >> (begin_bit + 8 * p)) & 255u;
      if (p == passes - 1) d &= last_mask;
      const uint32_t peers = match_digit(d, live);
      if (live && lane == __ffs(peers) - 1) mine[p * kRadix + d] += (uint32_t)__popc(peers);
    }
    __syncwarp();
  }
  __syncthreads();
  for (int k = threadIdx.x; k < passes * kRadix; k += blockDim.x) {
    uint32_t v = 0;
#pragma unroll
    for (int w = 0; w < kHistWarps; ++w) v += s_hist[w * passes * kRadix + k];
    if (v) atomicAdd(&hist[k], v);
  }
}

// exclusive scan of one (x, y) pair per thread over the first 256 threads of the CTA (called by ALL threads)
__device__ __forceinline__ uint2 block_exclusive_scan2(uint2 v, uint2 *s_warp /*[8]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint2 inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t ax = __shfl_up_sync(0xFFFFFFFFu, inc.x, o), ay = __shfl_up_sync(0xFFFFFFFFu, inc.y, o);
    if (lane >= o) {
      inc.x += ax;
      inc.y += ay;
    }
  }
  if (lane == 31 && warp < kRadix / 32) s_warp[warp] = inc;
  __syncthreads();
  uint2 pre = make_uint2(0, 0);
  for (int w = 0; w < warp && w < kRadix / 32; ++w) {
    pre.x += s_warp[w].x;
    pre.y += s_warp[w].y;
  }
  __syncthreads();
  return make_uint2(pre.x + inc.x - v.x, pre.y + inc.y - v.y);
}

// ---- one pass -----------------------------------------------------------------------------------------------
// Big tiles: GSR_SORT_THREADS threads x ITEMS keys, the tile staged in (dynamic) shared memory.  A first version
// with 4096-key tiles spent most of each pass in the look-back: all CTAs of a wave publish their aggregates at about
// the same time, so CTA t walks back over t unresolved predecessors - ~300 with two small CTAs per SM, wave after
// wave.  With ITEMS chosen so that the whole sort is one or two waves of at most 148 CTAs the walk is a handful of
// 32-wide rounds and the number of look-backs drops 5x.
#ifndef GSR_SORT_THREADS
// Threads per CTA (tile = threads x ITEMS keys).  1024: one CTA owns the SM's whole register file.  512 (default): two CTAs
// per SM whose phases (load / rank / look-back / store) overlap, and a sort CTA of one frame can start beside resident
// blend CTAs of another.  Measured (profiles/r02_notes.md): C3 binning alone 0.249 -> 0.262 ms, but three frames in
// flight 0.679 -> 0.667 ms per frame, and C5 (20 M pairs, several waves) 0.667 -> 0.605 ms.  256: 0.678 pipelined.
#define GSR_SORT_THREADS 512
#endif
constexpr int kBigThreads = GSR_SORT_THREADS;
constexpr int kQuarters = kBigThreads / 256;              // groups of 256 threads: one per look-back window
constexpr int kBigCtasPerSm = 1024 / kBigThreads;
constexpr int kBigWarps = kBigThreads / 32;

// FILTER: the first pass of a sort whose input still holds entries that must not take part (bit 31 of live_flags[i]
// clear): they are ranked in a group of their own and never stored, so the pass also compacts - the separate
// compaction kernel (and its look-back) of the binning stage is gone.  The pass then runs over all n_host entries and
// writes the number of survivors (the sum of the digit counts) to *n_out for the passes that follow.
//
// dyn_pass > 0: ADAPTIVE pass of a sort whose passes may be skipped on the device.  A pass whose digit histogram has a
// single occupied bin is the identity permutation (all keys share the digit) - e.g. the top byte of the depth keys of a
// scene whose depths lie within one or two binades.  Such a pass returns at once; the passes that do run ping-pong
// between the buffers (keys_in, vals_in) = X1 and (keys_out, vals_out) = X0 in the order they actually execute, each
// deriving its source from the histograms of the passes before it (hist points at the rows of ALL passes then).  The
// consumer locates the result the same way (sorted_buffer_index below).  Pass 0 always runs and writes X0.
__device__ __forceinline__ bool pass_is_identity(const uint32_t *__restrict__ hist_row, uint32_t n_total) {
  const uint32_t h = threadIdx.x < kRadix ? __ldg(hist_row + threadIdx.x) : 0u;
  return __syncthreads_or(h == n_total) != 0;  // one bin holds every key (called by all threads of the CTA)
}

template <int ITEMS, bool FILTER>
__global__ void __launch_bounds__(kBigThreads, kBigCtasPerSm)
onesweep_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int64_t n_host,
                const uint32_t *__restrict__ n_dev, const uint32_t *__restrict__ hist, uint32_t *status,
                uint32_t *ticket, int shift, uint32_t mask, const uint32_t *__restrict__ live_flags,
                uint32_t *__restrict__ n_out, int dyn_pass) {
  constexpr int TILE = kBigThreads * ITEMS;
  extern __shared__ uint32_t s_dyn[];                      // s_keys[TILE] | s_vals[TILE]
  uint32_t *s_keys = s_dyn, *s_vals = s_dyn + TILE;
  __shared__ uint16_t s_hist[kBigWarps][kRadix];           // per-warp digit counts, then exclusive over warps (< TILE <= 65535)
  __shared__ uint32_t s_gbase[kRadix];                     // global slot of a digit's first pair minus its tile-local slot
  __shared__ uint2 s_scan[kRadix / 32];
  __shared__ uint32_t s_part[kQuarters][kRadix];
  __shared__ uint32_t s_open[kRadix];
  __shared__ uint32_t s_tile, s_live_total;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(ticket, 1u);
  for (int k = tid; k < kBigWarps * kRadix / 2; k += kBigThreads) reinterpret_cast<uint32_t *>(&s_hist[0][0])[k] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;
  const int64_t n = (!FILTER && n_dev) ? (int64_t)*n_dev : n_host;
  const int64_t num_tiles = (n + TILE - 1) / TILE;
  if ((int64_t)tile >= num_tiles) return;
  if (!FILTER && dyn_pass > 0) {
    int executed = 1;  // pass 0
    for (int q = 1; q < dyn_pass; ++q) executed += !pass_is_identity(hist + q * kRadix, (uint32_t)n);
    if (pass_is_identity(hist + dyn_pass * kRadix, (uint32_t)n)) return;
    hist += dyn_pass * kRadix;
    if (executed & 1) {  // the previous executed pass wrote X0 = (keys_out, vals_out): read it, write X1
      const uint32_t *ki = keys_out, *vi = vals_out;
      keys_out = const_cast<uint32_t *>(keys_in);
      vals_out = const_cast<uint32_t *>(vals_in);
      keys_in = ki;
      vals_in = vi;
    }
  }

  // ---- load: warp w owns the contiguous run [w*ITEMS*32, (w+1)*ITEMS*32) of the tile, item i of lane l at i*32+l.
  //      Indices are 32-bit: n < 2^30 and tile * TILE < n + TILE ----
  static_assert(ITEMS % 2 == 0, "ranks are packed two per register");
  const uint32_t base = tile * (uint32_t)TILE + (uint32_t)(warp * (ITEMS * 32) + lane);
  const int32_t left = (int32_t)(uint32_t)n - (int32_t)base;  // item i exists iff i * 32 < left
  const uint32_t *kp = keys_in + base;
  uint32_t key[ITEMS];
  uint32_t live_mask = 0xFFFFFFFFu;  // FILTER: bit i = item i takes part
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const bool in = i * 32 < left;
    key[i] = in ? __ldg(kp + i * 32) : 0xFFFFFFFFu;  // sentinels carry the largest digit and come last
    if (FILTER && !(in && (__ldg(live_flags + base + i * 32) >> 31))) live_mask &= ~(1u << i);
  }

  // ---- rank inside the warp, in (item, lane) order.  The highest lane of a group of equal digits leads it (one FLO);
  //      ranks are packed two per register and pinned there (left alone, the compiler kept the vote masks instead and
  //      redid the popcounts in the scatter phase: 48 live registers, spills) ----
  const uint32_t hist_row = smem_addr(&s_hist[warp][0]);
  const uint32_t lanes_below = (1u << lane) - 1u;
  uint32_t rank2[ITEMS / 2];
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint32_t d = (key[i] >> shift) & mask;
    const bool live = !FILTER || ((live_mask >> i) & 1u);
    const uint32_t peers = FILTER ? match_digit(d, live) : match_digit(d);
    const int leader = 31 - __clz(peers);
    const bool lead = lane == leader && live;
    const uint32_t slot = hist_row + 2u * d;
    uint32_t before = 0;
    if (lead) before = lds_u16(slot);
    before = __shfl_sync(0xFFFFFFFFu, before, leader);
    const uint32_t r = before + __popc(peers & lanes_below);
    if (lead) sts_u16(slot, before + __popc(peers));
    __syncwarp();
    if (i & 1)
      rank2[i / 2] |= r << 16;
    else
      rank2[i / 2] = r;
    asm volatile("" : "+r"(rank2[i / 2]));
  }
  __syncthreads();

  // ---- thread d (< 256) owns digit d: tile count, scan over the digits, then each warp's count is replaced by the
  //      tile-local slot of its first key of that digit (digit offset folded in: one shared load per key in the scatter
  //      phase instead of two - these passes queue on the shared-memory pipe) ----
  uint32_t count = 0;
  if (tid < kRadix) {
#pragma unroll 8
    for (int w = 0; w < kBigWarps; ++w) count += s_hist[w][tid];
    // published at once: successors look back for it while this CTA is still scanning
    st_volatile_u32(status + (size_t)tile * kRadix + tid, (tile == 0 ? kFlagInclusive : kFlagAggregate) | count);
  }
  const uint2 ex = block_exclusive_scan2(make_uint2(count, tid < kRadix ? __ldg(hist + tid) : 0u), s_scan);
  if (tid < kRadix) {
    uint32_t run = ex.x;
#pragma unroll 8
    for (int w = 0; w < kBigWarps; ++w) {
      const uint32_t c = s_hist[w][tid];
      s_hist[w][tid] = (uint16_t)run;  // < TILE <= 24576
      run += c;
    }
  }
  if (FILTER && tid == kRadix - 1) {
    s_live_total = ex.x + count;                                      // survivors of this tile
    if (tile == 0) *n_out = ex.y + __ldg(hist + tid);                 // survivors of the whole input
  }

  // ---- decoupled look-back, 32 predecessor tiles per round trip: the four quarters of the CTA each fetch 8
  //      predecessors of digit d with independent, coalesced loads and fold them in order up to the first inclusive
  //      entry; the quarters are combined through shared memory ----
  {
    const int d = tid & (kRadix - 1), quarter = tid >> 8;
    uint32_t *my_status = status + (size_t)tile * kRadix + d;
    if (quarter == 0) {
      s_open[d] = tile != 0;
    }
    uint32_t prefix = 0;
    int64_t prev = (int64_t)tile - 1;
    while (true) {
      __syncthreads();
      const bool open = s_open[d] != 0;
      if (__syncthreads_or(open) == 0) break;
      uint32_t part = 0, found = 0;
      if (open) {
        constexpr int kWin = 8;
        const int64_t first = prev - (int64_t)quarter * kWin;
        uint32_t v[kWin];
#pragma unroll
        for (int i = 0; i < kWin; ++i)
          v[i] = (first - i >= 0) ? ld_volatile_u32(status + (size_t)(first - i) * kRadix + d) : kFlagInclusive;
#pragma unroll
        for (int i = 0; i < kWin; ++i) {
          if (found) break;
          while ((v[i] >> 30) == 0) {  // predecessor not published yet
            __nanosleep(40);
            v[i] = ld_volatile_u32(status + (size_t)(first - i) * kRadix + d);
          }
          part += v[i] & kValueMask;
          found = v[i] >> 31;  // kFlagInclusive
        }
      }
      s_part[quarter][d] = part | (found << 31);
      __syncthreads();
      if (quarter == 0 && open) {
        bool done = false;
#pragma unroll
        for (int q = 0; q < kQuarters; ++q) {
          if (!done) {
            const uint32_t pq = s_part[q][d];
            prefix += pq & 0x7FFFFFFFu;
            done = pq >> 31;
          }
        }
        if (done) s_open[d] = 0;
      }
      prev -= kQuarters * 8;
    }
    if (quarter == 0) {
      if (tile != 0) st_volatile_u32(my_status, kFlagInclusive | ((prefix + count) & kValueMask));
      s_gbase[d] = ex.y + prefix - ex.x;
    }
  }
  __syncthreads();

  // ---- scatter to the tile-local sorted position.  Keys first (their registers are then free), then all value loads
  //      of the thread are issued back to back before the first one is stored: in a fused loop every shared-memory
  //      store waited for its own global load (a quarter of the kernel's stall samples) ----
  const uint32_t keys_base = smem_addr(s_keys), vals_base = smem_addr(s_vals);
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint32_t d = (key[i] >> shift) & mask;
    const uint32_t r = (i & 1) ? (rank2[i / 2] >> 16) : (rank2[i / 2] & 0xFFFFu);
    const uint32_t pos = lds_u16(hist_row + 2u * d) + r;  // < TILE <= 24576
    if (!FILTER || ((live_mask >> i) & 1u)) sts_u32(keys_base + 4u * pos, key[i]);
    if (i & 1)
      rank2[i / 2] = (rank2[i / 2] & 0xFFFFu) | (pos << 16);
    else
      rank2[i / 2] = (rank2[i / 2] & 0xFFFF0000u) | pos;
  }
  {
    uint32_t val[ITEMS];
    const uint32_t *vp = vals_in ? vals_in + base : nullptr;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) val[i] = i * 32 < left ? (vp ? __ldg(vp + i * 32) : base + i * 32) : 0u;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const uint32_t pos = (i & 1) ? (rank2[i / 2] >> 16) : (rank2[i / 2] & 0xFFFFu);
      if (!FILTER || ((live_mask >> i) & 1u)) sts_u32(vals_base + 4u * pos, val[i]);
    }
  }
  __syncthreads();
  // ---- runs of equal digits leave coalesced ----
  const int64_t remaining = n - (int64_t)tile * TILE;
  const int valid = FILTER ? (int)s_live_total : (remaining < TILE ? (int)remaining : TILE);
  const uint32_t gbase_base = smem_addr(s_gbase);
  for (int j = tid; j < valid; j += kBigThreads) {
    const uint32_t k = lds_u32(keys_base + 4u * j);
    const uint32_t dst = lds_u32(gbase_base + 4u * ((k >> shift) & mask)) + j;
    if (keys_out) keys_out[dst] = k;
    vals_out[dst] = lds_u32(vals_base + 4u * j);
  }
}

// smallest supported tile that keeps the sort within one wave of CTAs (or the largest one)
static int pick_items(int64_t n, int sms) {
  // GSR_SORT_MAX_ITEMS=4..24 caps the tile (smaller tiles leave shared memory to other streams); read once, thread-safe
  static const int cap = [] {
    const char *e = getenv("GSR_SORT_MAX_ITEMS");
    const int c = e ? atoi(e) : 24;
    return (c < 4 || c > 24) ? 24 : c;
  }();
  const int options[] = {4, 8, 12, 16, 20, 24};
  int best = 4;
  for (int it : options) {
    if (it > cap) break;
    best = it;
    if (ceil_div64(n, (int64_t)kBigThreads * it) <= (int64_t)sms * kBigCtasPerSm) return it;
  }
  return best;
}
// Look-back status words: one row of 256 per tile and pass, for the tile size the launcher will pick for this n on the
// current device (both are known on the host), so the per-frame memset covers kilobytes, not megabytes.
// GSR_SORT_REST_ITEMS=4..24: tile size of the passes AFTER a filtering first pass (they see only the survivors, so a
// smaller tile spreads them over more SMs); 0 / unset: the same tile as the first pass.  Experiment knob.
static int rest_items_cap() {
  static const int v = [] {
    const char *e = getenv("GSR_SORT_REST_ITEMS");
    const int c = e ? atoi(e) : 0;
    return (c < 4 || c > 24) ? 0 : (c / 4) * 4;
  }();
  return v;
}
static int64_t sort_tiles(int64_t n) {
  const int64_t n1 = n > 0 ? n : 1;
  int items = pick_items(n1, sm_count());
  if (rest_items_cap() && rest_items_cap() < items) items = rest_items_cap();
  return ceil_div64(n1, (int64_t)kBigThreads * items);
}

size_t sort_scratch_bytes(int64_t n) {
  return align_up(sizeof(uint32_t) * (kMaxPasses * kRadix + 8 + (size_t)kMaxPasses * sort_tiles(n) * kRadix), 256);
}

int sort_scratch_carve(void *workspace, size_t bytes, int64_t n, SortScratch *out) {
  if (bytes < sort_scratch_bytes(n) || ((uintptr_t)workspace & 15)) {
    set_error("sort workspace too small or misaligned (%zu < %zu)", bytes, sort_scratch_bytes(n));
    return GSR_ERR_WORKSPACE_TOO_SMALL;
  }
  const int64_t tiles = sort_tiles(n);
  uint32_t *p = (uint32_t *)workspace;
  out->hist = p;
  out->tickets = p + kMaxPasses * kRadix;
  out->status = p + kMaxPasses * kRadix + 8;
  out->status_stride = (size_t)tiles * kRadix;
  return GSR_OK;
}

int launch_sort_pairs(const uint32_t *keys_in, const uint32_t *values_in, uint32_t *keys_out, uint32_t *values_out,
                      uint32_t *keys_tmp, uint32_t *values_tmp, int64_t n, const uint32_t *n_dev, int begin_bit,
                      int end_bit, const SortScratch &s, cudaStream_t stream, bool skip_final_keys,
                      const uint32_t *ready_hist, const uint32_t *live_flags, uint32_t *n_live_out, bool scratch_zeroed,
                      bool adaptive) {
  GSR_REQUIRE(n >= 0 && n < (1ll << 30), "sort size out of range (must be < 2^30)");
  GSR_REQUIRE((live_flags == nullptr) == (n_live_out == nullptr), "live_flags and n_live_out go together");
  GSR_REQUIRE(!live_flags || n_dev == n_live_out, "a filtering sort counts its own survivors: n_dev must be n_live_out");
  GSR_REQUIRE(begin_bit >= 0 && end_bit <= 32 && end_bit > begin_bit, "bad bit range");
  GSR_REQUIRE(keys_in && keys_out && values_out && keys_tmp && values_tmp, "null sort buffer");
  if (n == 0) return GSR_OK;
  const int bits = end_bit - begin_bit;
  const int passes = (bits + 7) / 8;
  const int last_bits = bits - 8 * (passes - 1);
  const uint32_t last_mask = (1u << last_bits) - 1u;
  const int items0 = pick_items(n, sm_count());
  const int items_rest = (live_flags && rest_items_cap() && rest_items_cap() < items0) ? rest_items_cap() : items0;
  const size_t zero_bytes = sizeof(uint32_t) * (kMaxPasses * kRadix + 8 + (size_t)passes * s.status_stride);
  if (!scratch_zeroed) GSR_CHECK_CUDA(cudaMemsetAsync(s.hist, 0, zero_bytes, stream));
  const uint32_t *hist = ready_hist ? ready_hist : s.hist;  // digit counts may come from the producer of the keys
  if (!ready_hist) {
    int64_t want = ceil_div64(n, 256 * 16);
    const unsigned blocks = (unsigned)(want < 1 ? 1 : (want > 148 * 4 ? 148 * 4 : want));
    const size_t smem = sizeof(uint32_t) * kHistWarps * passes * kRadix;
    radix_hist_kernel<<<blocks, kHistWarps * 32, smem, stream>>>(keys_in, n, live_flags ? nullptr : n_dev, live_flags, s.hist,
                                                                 begin_bit, passes, last_mask);
    GSR_LAUNCH_CHECK();
  }
  GSR_REQUIRE(!adaptive || live_flags, "adaptive pass skipping is implemented for filtering sorts");
  const uint32_t *kin = keys_in, *vin = values_in;
  for (int p = 0; p < passes; ++p) {
    const bool to_out = ((passes - 1 - p) % 2) == 0;
    uint32_t *kout = to_out ? keys_out : keys_tmp;
    uint32_t *vout = to_out ? values_out : values_tmp;
    if (p == passes - 1 && skip_final_keys && !adaptive) kout = nullptr;  // caller only wants the permutation
    const uint32_t mask = (p == passes - 1) ? last_mask : 255u;
    const int items = p == 0 ? items0 : items_rest;
    const int64_t tiles = ceil_div64(n, (int64_t)kBigThreads * items);
    const size_t smem = sizeof(uint32_t) * 2 * (size_t)kBigThreads * items;
    const bool filter = p == 0 && live_flags != nullptr;
    const uint32_t *hist_p = hist + p * kRadix;
    int dyn_pass = 0;
    if (adaptive) {  // X0 = (keys_tmp, values_tmp) is what pass 0 writes, X1 = (keys_out, values_out)
      if (p == 0) {
        kout = keys_tmp;
        vout = values_tmp;
      } else {  // the kernel swaps the two when an odd number of passes ran before it
        kin = keys_out;
        vin = values_out;
        kout = keys_tmp;
        vout = values_tmp;
        hist_p = hist;
        dyn_pass = p;
      }
    }
#define GSR_SWEEP_LAUNCH(IT, F)                                                                                   \
  do {                                                                                                            \
    GSR_CHECK_CUDA((ensure_dynamic_smem<&onesweep_kernel<IT, F>>((int)(sizeof(uint32_t) * 2 * kBigThreads * IT))));  \
    onesweep_kernel<IT, F><<<(unsigned)tiles, kBigThreads, smem, stream>>>(                                        \
        kin, vin, kout, vout, n, n_dev, hist_p, s.status + p * s.status_stride, s.tickets + p,                     \
        begin_bit + 8 * p, mask, live_flags, n_live_out, dyn_pass);                                               \
  } while (0)
#define GSR_SWEEP(IT)                                                                                             \
  case IT: {                                                                                                      \
    if (filter) GSR_SWEEP_LAUNCH(IT, true); else GSR_SWEEP_LAUNCH(IT, false);                                     \
  } break;
    switch (items) {
      GSR_SWEEP(4)
      GSR_SWEEP(8)
      GSR_SWEEP(12)
      GSR_SWEEP(16)
      GSR_SWEEP(20)
      GSR_SWEEP(24)
    }
#undef GSR_SWEEP
#undef GSR_SWEEP_LAUNCH
    GSR_LAUNCH_CHECK();
    kin = kout;
    vin = vout;
  }
  return GSR_OK;
}

}  // namespace gsr

using namespace gsr;

extern "C" int gsr_sort_workspace_bytes(int64_t n, size_t *bytes) {
  GSR_REQUIRE(bytes && n >= 0, "bad arguments");
  *bytes = sort_scratch_bytes(n);
  return GSR_OK;
}

extern "C" int gsr_sort_pairs(const uint32_t *keys_in, const uint32_t *values_in, uint32_t *keys_out,
                              uint32_t *values_out, uint32_t *keys_tmp, uint32_t *values_tmp, int64_t n,
                              const uint32_t *n_dev, int32_t begin_bit, int32_t end_bit, void *workspace,
                              size_t workspace_bytes, void *stream) {
  SortScratch s;
  int rc = sort_scratch_carve(workspace, workspace_bytes, n, &s);
  if (rc) return rc;
  return launch_sort_pairs(keys_in, values_in, keys_out, values_out, keys_tmp, values_tmp, n, n_dev, begin_bit,
                           end_bit, s, (cudaStream_t)stream, false, nullptr, nullptr, nullptr, false, false);
}
