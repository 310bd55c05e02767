// Tile binning.  The projection kernel counted, per band tile, the splats that reach it; here
//   bds_bin_count  scans those counts into per-tile start offsets (+ the record total), and
//   bds_bin_sort   (1) emits every (tile, splat) pair into its tile's segment through a per-tile atomic
//                  cursor (unordered inside the segment) as ONE 64-bit word (fp32 depth bits << 32 | splat
//                  slot), (2) sorts each segment with ONE CTA per tile - block merge sort in shared memory
//                  for segments up to kTileSortCap records, bitonic in place in global memory beyond -
//                  restores gsplat's Gaussian-id order among exact depth ties, and gathers the packed
//                  48-byte splat records into that order, so that the composite kernels read contiguous
//                  chunks (TMA bulk copies).
// No global sort: a record is written once as an 8-byte word and once as its 48-byte record.
//
// Replaces gsplat's isect_tiles + torch.cumsum + cub::DeviceRadixSort + isect_offset_encode for the
// reference call at models/trainers/base.py:393-408.  Ordering contract = gsplat's: per (camera, tile)
// ascending depth bits, ties in Gaussian-id order (what a stable sort over id-ordered emission gives).
//
// All kernels are HBM-bound integer / copy work: coalesced loads where the data allows, grid sized to
// the data.
#include "big_splats.cuh"

namespace bds {

// band helpers shared with projection.cu (same arithmetic, kept local to avoid a link dependency)
BDS_HD void band_rows2(const bds_render_desc& d, int tile_h, int c, int& ty0, int& ty1) {
  int g0 = c * tile_h, g1 = g0 + tile_h;
  int lo = d.row_begin > g0 ? d.row_begin : g0;
  int hi = d.row_end < g1 ? d.row_end : g1;
  if (hi <= lo) { ty0 = ty1 = 0; return; }
  ty0 = lo - g0;
  ty1 = hi - g0;
}

// ---------------------------------------------------------------------------------------------
// exclusive scan (reduce-then-scan, 3 kernels, no spin-waits): TIn -> int64 prefix
// ---------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;  // 4096 elements per block

template <typename T>
BDS_D T block_exclusive_scan(T v, T* smem /*[kScanThreads/32]*/, T& block_total) {
  // warp inclusive scan
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    T w = lane < kScanThreads / 32 ? smem[lane] : T(0);
    T winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      T t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < kScanThreads / 32) smem[lane] = winc - w;  // exclusive warp offsets
    if (lane == kScanThreads / 32 - 1) smem[kScanThreads / 32] = winc;
  }
  __syncthreads();
  T res = smem[warp] + inc - v;
  block_total = smem[kScanThreads / 32];
  __syncthreads();
  return res;
}

template <typename TIn>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const TIn* __restrict__ in, int64_t n,
                                                                   int64_t* __restrict__ block_sums) {
  __shared__ int64_t sm[kScanThreads / 32 + 1];
  int64_t base = (int64_t)blockIdx.x * kScanTile;
  int64_t acc = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int64_t i = base + (int64_t)k * kScanThreads + threadIdx.x;
    if (i < n) acc += (int64_t)in[i];
  }
  int64_t tot;
  block_exclusive_scan<int64_t>(acc, sm, tot);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

// single block: exclusive scan of block_sums in place; total written to *total
__global__ void __launch_bounds__(kScanThreads) scan_spine_kernel(int64_t* __restrict__ block_sums, int nblocks,
                                                                  int64_t* __restrict__ total) {
  __shared__ int64_t sm[kScanThreads / 32 + 1];
  int64_t carry = 0;
  for (int base = 0; base < nblocks; base += kScanThreads) {
    int i = base + threadIdx.x;
    int64_t v = i < nblocks ? block_sums[i] : 0;
    int64_t tot;
    int64_t ex = block_exclusive_scan<int64_t>(v, sm, tot);
    if (i < nblocks) block_sums[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0 && total) *total = carry;
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const TIn* __restrict__ in, int64_t n,
                                                                  const int64_t* __restrict__ block_sums,
                                                                  TOut* __restrict__ out) {
  __shared__ int64_t sm[kScanThreads / 32 + 1];
  // blocked arrangement: thread t owns items [t*kScanItems, (t+1)*kScanItems)
  int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int64_t vals[kScanItems];
  int64_t acc = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int64_t i = base + k;
    vals[k] = i < n ? (int64_t)in[i] : 0;
    acc += vals[k];
  }
  int64_t tot;
  int64_t ex = block_exclusive_scan<int64_t>(acc, sm, tot) + block_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int64_t i = base + k;
    if (i < n) out[i] = (TOut)ex;
    ex += vals[k];
  }
}

static size_t scan_workspace_bytes(int64_t n) { return align_up((size_t)(ceil_div(n, kScanTile) + 1) * sizeof(int64_t), 256); }

template <typename TIn, typename TOut>
static int exclusive_scan(const TIn* in, TOut* out, int64_t n, int64_t* total_dev, void* ws, cudaStream_t stream) {
  int nblocks = ceil_div(n, kScanTile);
  int64_t* sums = static_cast<int64_t*>(ws);
  if (n == 0) {
    if (total_dev) BDS_CHECK_CUDA(cudaMemsetAsync(total_dev, 0, sizeof(int64_t), stream));
    return 0;
  }
  scan_reduce_kernel<TIn><<<nblocks, kScanThreads, 0, stream>>>(in, n, sums);
  BDS_CHECK_LAUNCH();
  scan_spine_kernel<<<1, kScanThreads, 0, stream>>>(sums, nblocks, total_dev);
  BDS_CHECK_LAUNCH();
  scan_apply_kernel<TIn, TOut><<<nblocks, kScanThreads, 0, stream>>>(in, n, sums, out);
  BDS_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// emission: one thread per visible splat (compaction order), (tile, splat) pairs through per-tile cursors
// ---------------------------------------------------------------------------------------------
struct EmitParams {
  bds_render_desc d;
  int tile_w, tile_h;
  const int32_t* radii;
  const int32_t* tile_offsets;  // [n_tiles + 1]
  int32_t* cursors;             // [n_tiles], zero on entry
  int n_slots;
  const float* splats;
  uint64_t* keys;               // [n_isect] (fp32 depth bits << 32) | splat slot: ONE word per record, key and payload
  int32_t* big_q;               // [0] = count, [1..] = slots of the very large splats (big_splats.cuh)
};

__global__ void __launch_bounds__(256) emit_pairs_kernel(EmitParams p) {
  const int N = p.d.n_gauss;
  const int lane = threadIdx.x & 31;
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = slot < p.n_slots;
  float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0;
  TileRect tr = {0, 0, 0, 0};
  int c = 0;
  uint64_t key = 0;
  if (active) {
    const float4* rp = reinterpret_cast<const float4*>(p.splats + (size_t)slot * 12);
    r0 = __ldg(rp); r1 = __ldg(rp + 1); r2 = __ldg(rp + 2);
    const int64_t idx = (int64_t)__float_as_int(r2.z);
    c = (int)(idx / N);
    int ty0, ty1;
    band_rows2(p.d, p.tile_h, c, ty0, ty1);
    // same candidate rectangle + same hit test (both non-inlined bodies) as the counting pass
    tr = candidate_rect(r0.x, r0.y, (float)p.radii[idx], r0.z, r0.w, r1.x, r2.w + kLog2_255, p.tile_w, p.tile_h, ty0, ty1);
    // positive floats order like their bit patterns; the low word carries the splat slot (exact depth ties are
    // re-ordered by Gaussian id after the sort, tile_sort_gather_kernel)
    key = ((uint64_t)(uint32_t)__float_as_int(r2.y) << 32) | (uint64_t)(uint32_t)slot;
  }
  // Same flat warp-cooperative enumeration as the counting pass (projection.cu).  The capacity guard only
  // makes a count / emission disagreement memory-safe should a toolchain ever break the shared-body
  // contract; positions a tile's cursor never reached are read as sentinels by the sort and become null records.
  int ncand = active ? (tr.x1 - tr.x0) * (tr.y1 - tr.y0) : 0;
  if (ncand > kBigCand) {   // very large footprint: queued for big_splat_kernel (one CTA per splat)
    const int q = atomicAdd(p.big_q, 1);
    if (q < kBigQueueCap) {
      p.big_q[1 + q] = slot;
      ncand = 0;
    }
  }
  const int incl = warp_inclusive_scan_i32(ncand);
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  const int excl = incl - ncand;
  const int rw = tr.x1 - tr.x0;
  const unsigned key_lo = (unsigned)key, key_hi = (unsigned)(key >> 32);
  for (int base = 0; base < total; base += 32) {
    const int wi = min(base + lane, total - 1);
    const int owner = warp_find_owner(excl, wi);
    const int local = wi - __shfl_sync(0xffffffffu, excl, owner);
    const int ow = __shfl_sync(0xffffffffu, rw, owner);
    const int ox0 = __shfl_sync(0xffffffffu, tr.x0, owner), oy0 = __shfl_sync(0xffffffffu, tr.y0, owner);
    const float gx = __shfl_sync(0xffffffffu, r0.x, owner), gy = __shfl_sync(0xffffffffu, r0.y, owner);
    const float ga = __shfl_sync(0xffffffffu, r0.z, owner), gb = __shfl_sync(0xffffffffu, r0.w, owner);
    const float gc = __shfl_sync(0xffffffffu, r1.x, owner), gcut = __shfl_sync(0xffffffffu, r2.w, owner) + kLog2_255;
    const int gcam = __shfl_sync(0xffffffffu, c, owner);
    const unsigned glo = __shfl_sync(0xffffffffu, key_lo, owner), ghi = __shfl_sync(0xffffffffu, key_hi, owner);
    if (base + lane < total) {
      const int ry = local / ow;
      const int tx = ox0 + local - ry * ow, ty = oy0 + ry;
      if (tile_hit(gx, gy, ga, gb, gc, gcut, tx, ty, p.d.width, p.d.height)) {
        const int tile = (gcam * p.tile_h + ty - p.d.row_begin) * p.tile_w + tx;
        const int seg0 = p.tile_offsets[tile], cap = p.tile_offsets[tile + 1] - seg0;
        const int pos = atomicAdd(p.cursors + tile, 1);
        if (pos < cap) p.keys[seg0 + pos] = ((uint64_t)ghi << 32) | glo;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// per-tile sort + gather.  Bitonic network in its "all merges ascending" form (first step of a merge
// pairs i with its mirror inside the block, the rest pair i with i + j): every compare-exchange puts the
// smaller key at the lower index, so positions >= n behave as +inf padding that never moves and pairs
// reaching beyond n are simply skipped - any n, no power-of-two padding in memory.
// ---------------------------------------------------------------------------------------------
constexpr int kTileSortCap = 4096;    // 64 KB instance: 2 x 4096 x 8 B (merge sort ping-pong)
constexpr int kTileSortSmall = 2048;  // segments up to here go to the 32 KB instance
constexpr uint32_t kNullSlot = 0xffffffffu;

template <typename KeyPtr>
BDS_D void block_bitonic_sort(KeyPtr keys, int n) {
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  const int half = np2 >> 1;
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j >= 1; j >>= 1) {
      const bool mirror = (j == (k >> 1));
      for (int t = threadIdx.x; t < half; t += blockDim.x) {
        const int lo = t & (j - 1);
        const int i = ((t - lo) << 1) + lo;              // 2 j (t / j) + t % j
        const int q = mirror ? (i - lo) + (k - 1 - lo) : i + j;
        if (q < n) {
          const uint64_t a = keys[i], b = keys[q];
          if (a > b) { keys[i] = b; keys[q] = a; }
        }
      }
      __syncthreads();
    }
  }
}

// Block merge sort for a segment that fits shared memory: every warp sorts 32-key chunks in registers (bitonic over
// the lanes, shuffles only), then log2(n / 32) merge levels in which each key finds its output position by a
// binary search in the partner run (keys are unique - the slot is part of the word - so the position is
// unambiguous; the left run counts strictly smaller partners, the right run smaller-or-equal ones, which also
// keeps equal sentinels apart).  Less than half the instructions of the bitonic network at the segment lengths
// that hold most records (500-1400) and 5-7 block barriers instead of 55-66.
BDS_D uint64_t warp_sort32(uint64_t key) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j >= 1; j >>= 1) {
      const uint64_t other = __shfl_xor_sync(0xffffffffu, key, j);
      const bool take_min = ((lane & j) == 0) == ((lane & k) == 0);
      key = (take_min == (key < other)) ? key : other;
    }
  }
  return key;
}

// a: the n keys (shared), b: scratch of the same size; returns the buffer that holds the sorted keys
BDS_D uint64_t* block_merge_sort(uint64_t* a, uint64_t* b, int n) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c0 = warp * 32; c0 < n; c0 += 256) {
    const int i = c0 + lane;
    uint64_t k = i < n ? a[i] : ~0ull;   // padding sorts to the end of the chunk and is not stored
    k = warp_sort32(k);
    if (i < n) a[i] = k;
  }
  __syncthreads();
  uint64_t *src = a, *dst = b;
  for (int m = 32; m < n; m <<= 1) {
    for (int i = threadIdx.x; i < n; i += 256) {
      const int run = i / m;
      const bool right = run & 1;
      const int pair0 = (run & ~1) * m;
      const int other0 = right ? pair0 : pair0 + m;   // start of the partner run
      const int olen = max(0, min(m, n - other0));
      const uint64_t x = src[i];
      int lo = 0, hi = olen;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const uint64_t y = src[other0 + mid];
        if (right ? (y <= x) : (y < x)) lo = mid + 1; else hi = mid;
      }
      dst[i - (right ? m : 0) + lo] = x;   // pair0 + (index inside the own run) + (partners in front)
    }
    __syncthreads();
    uint64_t* t = src; src = dst; dst = t;
  }
  return src;
}

// gsplat orders equal depths by Gaussian id (stable radix sort over Gaussian-major intersections).  The sort above
// orders them by splat slot; slots are handed out by warp-aggregated atomics, so exact fp32 depth ties (a few
// hundred tiles per step at the bench size) are put into id order here: one thread, insertion sort inside each run
// of equal depth.  Uniform early-out when the tile has no tie.
template <typename KeyPtr>
BDS_D void fix_depth_ties(KeyPtr keys, int n, const float4* __restrict__ splats) {
  int tie = 0;
  for (int i = threadIdx.x; i + 1 < n; i += blockDim.x) {
    const uint64_t a = keys[i], b = keys[i + 1];
    tie |= (uint32_t)(a >> 32) == (uint32_t)(b >> 32) && (uint32_t)b != kNullSlot;
  }
  if (!__syncthreads_or(tie)) return;
  if (threadIdx.x == 0) {
    int i = 0;
    while (i + 1 < n) {
      int e = i + 1;
      const uint32_t depth = (uint32_t)(keys[i] >> 32);
      while (e < n && (uint32_t)(keys[e] >> 32) == depth && (uint32_t)keys[e] != kNullSlot) ++e;
      for (int a = i + 1; a < e; ++a) {          // insertion sort of [i, e) by Gaussian id
        const uint64_t ka = keys[a];
        const int ida = __float_as_int(__ldg(splats + (size_t)(uint32_t)ka * 3 + 2).z);
        int b = a - 1;
        while (b >= i && __float_as_int(__ldg(splats + (size_t)(uint32_t)keys[b] * 3 + 2).z) > ida) {
          keys[b + 1] = keys[b];
          --b;
        }
        keys[b + 1] = ka;
      }
      i = e;
    }
  }
  __syncthreads();
}

template <int CAP, bool BIG>
__global__ void __launch_bounds__(256) tile_sort_gather_kernel(const int32_t* __restrict__ tile_offsets,
                                                               uint64_t* __restrict__ keys,
                                                               const float4* __restrict__ splats,
                                                               float4* __restrict__ sorted,
                                                               int32_t* __restrict__ sorted_slots,
                                                               int32_t* __restrict__ big_list,
                                                               const int32_t* __restrict__ cursors) {
  extern __shared__ __align__(16) uint64_t s_keys[];   // 2 x CAP keys (merge sort ping-pong)
  // the 32 KB instance (one CTA per tile, 7 CTAs per SM) sorts every segment up to kTileSortSmall records and
  // queues longer ones (big_list[0] = count, then tile ids) for the 64 KB instance, a fixed grid over that queue
  // that costs a few microseconds when the queue is empty
  for (int item = blockIdx.x; BIG ? item < big_list[0] : item == (int)blockIdx.x; item += gridDim.x) {
    const int tile = BIG ? big_list[1 + item] : item;
    const int seg0 = tile_offsets[tile];
    const int n = tile_offsets[tile + 1] - seg0;
    if (n <= 0) continue;
    if (!BIG && n > CAP) {
      if (threadIdx.x == 0) big_list[1 + atomicAdd(big_list, 1)] = tile;
      return;
    }
    if (BIG) __syncthreads();   // the previous item's gather is done with the shared keys
    // the emission filled the first cursors[tile] positions of the segment; should a toolchain ever make the
    // counting and the emission pass disagree, the rest are sentinels (they sort last and gather null records)
    const int filled = min(cursors[tile], n);
    const uint64_t* order;
    if (n <= CAP) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) s_keys[i] = i < filled ? keys[seg0 + i] : ~0ull;
      __syncthreads();
      uint64_t* sorted_keys = s_keys;
      if (n > 1) {
        sorted_keys = block_merge_sort(s_keys, s_keys + CAP, n);
        fix_depth_ties(sorted_keys, n, splats);
      }
      order = sorted_keys;
    } else {
      for (int i = filled + threadIdx.x; i < n; i += blockDim.x) keys[seg0 + i] = ~0ull;
      __syncthreads();
      block_bitonic_sort(keys + seg0, n);  // rare: oversized segment, in place in global memory
      fix_depth_ties(keys + seg0, n, splats);
      order = keys + seg0;
    }
    // sorted[seg0 + i] = splats[slot of order[i]] with the id field replaced by the slot; one thread per float4
    for (int t = threadIdx.x; t < n * 3; t += blockDim.x) {
      const int i = t / 3, part = t - i * 3;
      const uint32_t slot = (uint32_t)order[i];
      float4 v;
      if (slot == kNullSlot) {  // never-filled position (see emit_pairs_kernel): a record that contributes nothing
        v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (part == 2) v = make_float4(0.f, 0.f, __int_as_float(0), -1.0e30f);
      } else {
        v = __ldg(splats + (size_t)slot * 3 + part);
        if (part == 2) v.z = __int_as_float((int)slot);
      }
      sorted[(size_t)seg0 * 3 + t] = v;
    }
    if (sorted_slots)
      for (int i = threadIdx.x; i < n; i += blockDim.x) sorted_slots[seg0 + i] = (int32_t)(uint32_t)order[i];
  }
}

struct SortWorkspace {
  size_t keys, cursors, big_list, big_q, total;
};
static SortWorkspace carve_sort(int64_t n_isect, int n_tiles) {
  SortWorkspace w;
  size_t off = 0;
  size_t nk = (size_t)(n_isect > 0 ? n_isect : 1);
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
  w.keys = take(nk * 8);
  // cursors | big_list | big_q are contiguous: ONE memset zeroes the cursors and the two queue counters
  w.cursors = off; off += (size_t)(n_tiles + 1) * 4;
  w.big_list = off; off += (size_t)(n_tiles + 1) * 4;
  w.big_q = off; off += (size_t)(kBigQueueCap + 1) * 4;
  off = align_up(off, 256);
  w.total = off;
  return w;
}

int check_render_desc(const bds_render_desc* d);  // projection.cu

}  // namespace bds

using namespace bds;

static int band_tiles(const bds_render_desc* d) {
  const int tile_w = (d->width + kTile - 1) / kTile;
  return (d->row_end - d->row_begin) * tile_w;
}

extern "C" size_t bds_bin_count_workspace_bytes(const bds_render_desc* d) {
  return scan_workspace_bytes(d ? (int64_t)band_tiles(d) + 1 : 1) + 256;
}

extern "C" int bds_bin_count(const bds_render_desc* d, const int32_t* tile_counts, int32_t* tile_offsets,
                             int64_t* n_isect_dev, void* workspace, bds_stream_t stream_) {
  if (int rc = check_render_desc(d)) return rc;
  BDS_REQUIRE(tile_counts && tile_offsets && n_isect_dev && workspace, "bin_count: null pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  // tile_counts holds n_tiles + 1 entries, the last one zero: the exclusive scan then also yields the end offset
  return exclusive_scan<int32_t, int32_t>(tile_counts, tile_offsets, (int64_t)band_tiles(d) + 1, n_isect_dev, workspace,
                                          stream);
}

extern "C" size_t bds_bin_sort_workspace_bytes(const bds_render_desc* d, int64_t n_isect) {
  return carve_sort(n_isect, d ? band_tiles(d) : 1).total + 256;
}

extern "C" int bds_bin_sort(const bds_render_desc* d, int64_t n_isect, int32_t n_slots, const int32_t* radii,
                            const float* splats, const int32_t* tile_offsets, float* sorted_splats,
                            int32_t* sorted_slots, void* workspace, bds_stream_t stream_) {
  if (int rc = check_render_desc(d)) return rc;
  BDS_REQUIRE(n_isect >= 0 && n_isect < ((int64_t)1 << 31), "bin_sort: n_isect must fit int32 (got %lld)", (long long)n_isect);
  if (n_isect == 0) return 0;
  // n_slots may exceed n_isect: a deferred very large splat gets a record before its tiles are counted and a thin
  // diagonal ellipse can end up touching none (a slot with 0 records is harmless downstream)
  BDS_REQUIRE(radii && splats && tile_offsets && sorted_splats && workspace && n_slots > 0,
              "bin_sort: null pointer or non-positive n_slots");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int tile_w = (d->width + kTile - 1) / kTile, tile_h = (d->height + kTile - 1) / kTile;
  const int n_tiles = band_tiles(d);
  char* ws = static_cast<char*>(workspace);
  SortWorkspace w = carve_sort(n_isect, n_tiles);
  uint64_t* keys = reinterpret_cast<uint64_t*>(ws + w.keys);
  int32_t* cursors = reinterpret_cast<int32_t*>(ws + w.cursors);
  int32_t* big_list = reinterpret_cast<int32_t*>(ws + w.big_list);
  BDS_CHECK_CUDA(cudaMemsetAsync(cursors, 0, w.big_q + (size_t)(kBigQueueCap + 1) * 4 - w.cursors, stream));
  EmitParams ep;
  ep.d = *d; ep.tile_w = tile_w; ep.tile_h = tile_h; ep.radii = radii; ep.tile_offsets = tile_offsets;
  ep.cursors = cursors; ep.n_slots = n_slots; ep.splats = splats; ep.keys = keys;
  ep.big_q = reinterpret_cast<int32_t*>(ws + w.big_q);
  emit_pairs_kernel<<<ceil_div(n_slots, 256), 256, 0, stream>>>(ep);
  BDS_CHECK_LAUNCH();
  {
    BigSplatParams b{};
    b.d = *d; b.tile_w = tile_w; b.tile_h = tile_h; b.splats = splats; b.radii = radii;
    b.n_queue = ep.big_q; b.queue = ep.big_q + 1; b.tile_offsets = tile_offsets; b.cursors = cursors; b.keys = keys;
    big_splat_kernel<true><<<4 * 148, 256, 0, stream>>>(b);
    BDS_CHECK_LAUNCH();
  }
  static bool attr_set[64] = {};   // per device; idempotent: two racing first calls set the same value twice
  int dev = 0;
  BDS_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    BDS_CHECK_CUDA(cudaFuncSetAttribute(tile_sort_gather_kernel<kTileSortCap, true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kTileSortCap * 8));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  tile_sort_gather_kernel<kTileSortSmall, false><<<n_tiles, 256, 2 * kTileSortSmall * 8, stream>>>(
      tile_offsets, keys, reinterpret_cast<const float4*>(splats), reinterpret_cast<float4*>(sorted_splats),
      sorted_slots, big_list, cursors);
  BDS_CHECK_LAUNCH();
  const int big_grid = n_tiles < 3 * 148 ? n_tiles : 3 * 148;   // 64 KB of shared memory: 3 CTAs per SM
  tile_sort_gather_kernel<kTileSortCap, true><<<big_grid, 256, 2 * kTileSortCap * 8, stream>>>(
      tile_offsets, keys, reinterpret_cast<const float4*>(splats), reinterpret_cast<float4*>(sorted_splats),
      sorted_slots, big_list, cursors);
  BDS_CHECK_LAUNCH();
  return 0;
}
