// Tile binning: one scan of the per-(camera, Gaussian) tile counts (record total + id-ordered rank of
// every emitting splat), a stable LSD radix sort written for this library (no CUB / thrust) used twice
// - visible splats by fp32 depth bits, then the emitted (band tile, slot) records by tile -, per-tile
// start offsets and the gather of the packed 48-byte splat records into sorted order so that the
// composite kernels read contiguous chunks (TMA bulk copies).
//
// Replaces gsplat's isect_tiles + torch.cumsum + cub::DeviceRadixSort + isect_offset_encode for the
// reference call at models/trainers/base.py:393-408.  Ordering contract = gsplat's: ascending
// (camera, tile, depth bits), ties in Gaussian-id order.
//
// All kernels are HBM-bound integer / copy work: coalesced loads, grid sized to the data.
#include "projection_math.cuh"

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
// LSD radix sort, 8-bit digits, (KeyT key, uint32 value), stable.  Per pass:
//   rs_hist_kernel     per-block digit histogram -> hist[digit][block]
//   exclusive_scan     over the digit-major array -> global base of every (digit, block)
//   rs_scatter_kernel  stable in-block ranking (warp match_any + per-warp counters) and scatter
// ---------------------------------------------------------------------------------------------
constexpr int kRsThreads = 256;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems;  // 4096 keys per block
constexpr int kRsWarps = kRsThreads / 32;

template <typename KeyT>
__global__ void __launch_bounds__(kRsThreads) rs_hist_kernel(const KeyT* __restrict__ keys, int64_t n, int shift,
                                                             int nblocks, uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  int64_t base = (int64_t)blockIdx.x * kRsTile;
#pragma unroll
  for (int k = 0; k < kRsItems; ++k) {
    int64_t i = base + (int64_t)k * kRsThreads + threadIdx.x;
    if (i < n) atomicAdd(&sh[(uint32_t)(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = sh[threadIdx.x];
}

template <typename KeyT>
__global__ void __launch_bounds__(kRsThreads) rs_scatter_kernel(const KeyT* __restrict__ keys_in,
                                                                const uint32_t* __restrict__ vals_in, int64_t n,
                                                                int shift, int nblocks,
                                                                const uint32_t* __restrict__ base,
                                                                KeyT* __restrict__ keys_out,
                                                                uint32_t* __restrict__ vals_out) {
  __shared__ uint32_t warp_hist[kRsWarps][256];
  __shared__ uint32_t gbase[256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kRsWarps * 256; i += kRsThreads) (&warp_hist[0][0])[i] = 0;
  gbase[threadIdx.x] = base[(size_t)threadIdx.x * nblocks + blockIdx.x];
  __syncthreads();
  // warp w owns the contiguous segment [w*512, (w+1)*512) of the block tile, in rounds of 32
  int64_t seg = (int64_t)blockIdx.x * kRsTile + (int64_t)warp * (32 * kRsItems);
  KeyT key[kRsItems];
  uint32_t rank[kRsItems];
#pragma unroll
  for (int r = 0; r < kRsItems; ++r) {
    int64_t i = seg + r * 32 + lane;
    bool valid = i < n;
    key[r] = valid ? keys_in[i] : (KeyT)0;
    uint32_t digit = valid ? ((uint32_t)(key[r] >> shift) & 255u) : 256u;  // 256 = "no element"
    unsigned peers = __match_any_sync(0xffffffffu, digit);
    uint32_t before = __popc(peers & ((1u << lane) - 1u));
    int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (valid && lane == leader) {
      old = warp_hist[warp][digit];
      warp_hist[warp][digit] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[r] = old + before;
    __syncwarp();
  }
  __syncthreads();
  {  // exclusive scan over warps, per digit
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w) {
      uint32_t cnt = warp_hist[w][threadIdx.x];
      warp_hist[w][threadIdx.x] = run;
      run += cnt;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kRsItems; ++r) {
    int64_t i = seg + r * 32 + lane;
    if (i < n) {
      uint32_t digit = (uint32_t)(key[r] >> shift) & 255u;
      uint32_t pos = gbase[digit] + warp_hist[warp][digit] + rank[r];
      keys_out[pos] = key[r];
      vals_out[pos] = vals_in[i];
    }
  }
}

// sorts (keys[0], vals[0]) over bits [0, end_bit); returns the index (0/1) of the buffer holding the result
template <typename KeyT>
static int radix_sort_pairs(KeyT* keys[2], uint32_t* vals[2], int64_t n, int end_bit, uint32_t* hist, void* scan_ws,
                            cudaStream_t stream, int* result_buf) {
  const int nblocks = ceil_div(n, kRsTile);
  int cur = 0;
  for (int shift = 0; shift < end_bit; shift += 8) {
    rs_hist_kernel<KeyT><<<nblocks, kRsThreads, 0, stream>>>(keys[cur], n, shift, nblocks, hist);
    BDS_CHECK_LAUNCH();
    if (int rc = exclusive_scan<uint32_t, uint32_t>(hist, hist, (int64_t)256 * nblocks, nullptr, scan_ws, stream)) return rc;
    rs_scatter_kernel<KeyT><<<nblocks, kRsThreads, 0, stream>>>(keys[cur], vals[cur], n, shift, nblocks, hist,
                                                                keys[cur ^ 1], vals[cur ^ 1]);
    BDS_CHECK_LAUNCH();
    cur ^= 1;
  }
  *result_buf = cur;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Two-stage ordering.  gsplat sorts every (tile, splat) record by a 64-bit (tile | depth) key; here the
// visible splats (n_slots, ~7x fewer than records) are first sorted by (depth bits, Gaussian id), the
// records are then emitted in that order and only need a STABLE sort on the tile index
// (17 bits -> 3 passes of 20 B/record instead of 6 passes of 32 B/record).  Result: per tile ascending
// (depth bits, id) - exactly the order of gsplat's stable 64-bit sort over id-ordered emission.
// ---------------------------------------------------------------------------------------------
constexpr int kRankShift = 40;  // bds_bin_count packs (rank of emitting splats << 40) | tile prefix

// stage-1 input in Gaussian-id order: key = depth bits, value = slot (one thread per compact slot)
__global__ void __launch_bounds__(256) stage1_fill_kernel(const float* __restrict__ splats, int n_slots,
                                                          const int64_t* __restrict__ packed_prefix,
                                                          uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n_slots) return;
  const float* rec = splats + (size_t)slot * 12;
  int64_t idx = (int64_t)__float_as_int(__ldg(rec + 10));
  int64_t rank = packed_prefix[idx] >> kRankShift;
  keys[rank] = (uint32_t)__float_as_int(__ldg(rec + 9));  // positive floats order like their bit patterns
  vals[rank] = (uint32_t)slot;
}

// tiles_touched in depth-sorted order (input of the second scan)
__global__ void __launch_bounds__(256) gather_tiles_kernel(const float* __restrict__ splats,
                                                           const uint32_t* __restrict__ sorted_slots, int n_slots,
                                                           const int32_t* __restrict__ tiles_touched,
                                                           int32_t* __restrict__ out) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_slots) return;
  int64_t idx = (int64_t)__float_as_int(__ldg(splats + (size_t)sorted_slots[r] * 12 + 10));
  out[r] = tiles_touched[idx];
}

struct EmitParams {
  bds_render_desc d;
  int tile_w, tile_h;
  const int32_t* radii;
  const int32_t* tiles_sorted;   // [n_slots] tile counts in depth order
  const int64_t* offsets;        // [n_slots] exclusive prefix of tiles_sorted
  const uint32_t* sorted_slots;  // [n_slots] slots in (depth, id) order
  int n_tiles;
  int n_slots;
  uint32_t perm_mul;             // multiplier coprime with n_slots
  const float* splats;
  uint32_t* keys;
  uint32_t* vals;
};

// one thread per splat in (depth, id) order: emits (band tile, slot) for every tile it reaches
__global__ void __launch_bounds__(256) emit_keys_kernel(EmitParams p) {
  __shared__ int s_run[8][32];
  const int N = p.d.n_gauss;
  const int lane = threadIdx.x & 31;
  // Depth order puts the largest (nearest) splats side by side; a multiplicative permutation of the
  // thread -> rank map spreads them over the grid (output positions come from offsets[r], so the
  // emitted order is unchanged).
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = t < p.n_slots;
  const int r = active ? (int)(((uint64_t)t * p.perm_mul) % (uint64_t)p.n_slots) : 0;
  const int slot = active ? (int)p.sorted_slots[r] : 0;
  float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0;
  TileRect tr = {0, 0, 0, 0};
  long long o = 0, o_end = 0;
  int c = 0;
  if (active) {
    const float4* rp = reinterpret_cast<const float4*>(p.splats + (size_t)slot * 12);
    r0 = __ldg(rp); r1 = __ldg(rp + 1); r2 = __ldg(rp + 2);
    const int64_t idx = (int64_t)__float_as_int(r2.z);
    c = (int)(idx / N);
    int ty0, ty1;
    band_rows2(p.d, p.tile_h, c, ty0, ty1);
    // same candidate rectangle + same hit test (both non-inlined bodies) as the counting pass
    tr = candidate_rect(r0.x, r0.y, (float)p.radii[idx], r0.z, r0.w, r1.x, r2.w + kLog2_255, p.tile_w, p.tile_h, ty0, ty1);
    o = p.offsets[r];
    o_end = o + p.tiles_sorted[r];
  }
  // Same flat warp-cooperative enumeration as the counting pass.  The o_end guard and the sentinel
  // padding (a key that sorts behind every real tile) only make a count / emission disagreement
  // memory-safe should a toolchain ever break the shared-body contract.
  {
    const int ncand = active ? (tr.x1 - tr.x0) * (tr.y1 - tr.y0) : 0;
    const int incl = warp_inclusive_scan_i32(ncand);
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const int excl = incl - ncand;
    const int rw = tr.x1 - tr.x0;
    int* run = s_run[threadIdx.x >> 5];
    run[lane] = 0;
    __syncwarp();
    for (int base = 0; base < total; base += 32) {
      const int wi = min(base + lane, total - 1);
      const int owner = warp_find_owner(excl, wi);
      const int local = wi - __shfl_sync(0xffffffffu, excl, owner);
      const int ow = __shfl_sync(0xffffffffu, rw, owner);
      const int ox0 = __shfl_sync(0xffffffffu, tr.x0, owner), oy0 = __shfl_sync(0xffffffffu, tr.y0, owner);
      const float gx = __shfl_sync(0xffffffffu, r0.x, owner), gy = __shfl_sync(0xffffffffu, r0.y, owner);
      const float ga = __shfl_sync(0xffffffffu, r0.z, owner), gb = __shfl_sync(0xffffffffu, r0.w, owner);
      const float gc = __shfl_sync(0xffffffffu, r1.x, owner), gcut = __shfl_sync(0xffffffffu, r2.w, owner) + kLog2_255;
      const int gcam = __shfl_sync(0xffffffffu, c, owner), gslot = __shfl_sync(0xffffffffu, slot, owner);
      const long long go = __shfl_sync(0xffffffffu, o, owner), gend = __shfl_sync(0xffffffffu, o_end, owner);
      bool hit = false;
      int tx = 0, ty = 0;
      if (base + lane < total) {
        const int ry = local / ow;
        tx = ox0 + local - ry * ow;
        ty = oy0 + ry;
        hit = tile_hit(gx, gy, ga, gb, gc, gcut, tx, ty, p.d.width, p.d.height);
      }
      const unsigned hm = __ballot_sync(0xffffffffu, hit);
      const unsigned grp = __match_any_sync(0xffffffffu, owner);
      const int before = run[owner];
      __syncwarp();
      const long long pos = go + before + __popc(hm & grp & ((1u << lane) - 1u));
      if (hit && pos < gend) {
        p.keys[pos] = (uint32_t)((gcam * p.tile_h + ty - p.d.row_begin) * p.tile_w + tx);
        p.vals[pos] = (uint32_t)gslot;
      }
      if (lane == __ffs(grp) - 1) run[owner] = before + __popc(hm & grp);
      __syncwarp();
    }
    o += run[lane];
    if (o > o_end) o = o_end;
  }
  for (; o < o_end; ++o) {
    p.keys[o] = (uint32_t)p.n_tiles;
    p.vals[o] = (uint32_t)slot;
  }
}

// per-tile start offsets from the sorted tile keys (the role of isect_offset_encode)
__global__ void __launch_bounds__(256) tile_offsets_kernel(const uint32_t* __restrict__ keys, int64_t n, int n_tiles,
                                                           int32_t* __restrict__ offsets) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n == 0) {
    if (i <= n_tiles) offsets[i] = 0;
    return;
  }
  if (i >= n) return;
  int cur = (int)keys[i];
  if (i == 0) {
    for (int t = 0; t <= cur && t <= n_tiles; ++t) offsets[t] = 0;
  } else {
    int prev = (int)keys[i - 1];
    for (int t = prev + 1; t <= cur && t <= n_tiles; ++t) offsets[t] = (int32_t)i;
  }
  if (i == n - 1) {
    for (int t = cur + 1; t <= n_tiles; ++t) offsets[t] = (int32_t)n;
  }
}

// sorted_splats[i] = splats[vals[i]] with the id field replaced by the slot; one thread per float4
__global__ void __launch_bounds__(256) gather_records_kernel(const uint32_t* __restrict__ vals, int64_t n,
                                                             const float4* __restrict__ splats,
                                                             float4* __restrict__ sorted) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * 3) return;
  int64_t i = t / 3;
  int part = (int)(t - i * 3);
  uint32_t slot = vals[i];
  float4 v = __ldg(splats + (size_t)slot * 3 + part);
  if (part == 2) v.z = __int_as_float((int)slot);
  sorted[t] = v;
}

struct SortWorkspace {
  size_t keys_a, keys_b, vals_a, vals_b, s1k_a, s1k_b, s1v_a, s1v_b, tiles_sorted, offs2, hist, scan, total;
};
static SortWorkspace carve_sort(int64_t n_isect, int64_t n_slots) {
  SortWorkspace w;
  size_t off = 0;
  size_t nk = (size_t)(n_isect > 0 ? n_isect : 1), ns = (size_t)(n_slots > 0 ? n_slots : 1);
  size_t nmax = nk > ns ? nk : ns;
  int nblocks = ceil_div((int64_t)nmax, kRsTile);
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
  w.keys_a = take(nk * 4); w.keys_b = take(nk * 4); w.vals_a = take(nk * 4); w.vals_b = take(nk * 4);
  w.s1k_a = take(ns * 4); w.s1k_b = take(ns * 4); w.s1v_a = take(ns * 4); w.s1v_b = take(ns * 4);
  w.tiles_sorted = take(ns * 4); w.offs2 = take(ns * 8);
  w.hist = take((size_t)256 * nblocks * 4);
  size_t scan_n = (size_t)256 * nblocks;
  w.scan = take(scan_workspace_bytes((int64_t)(scan_n > ns ? scan_n : ns)));
  w.total = off;
  return w;
}

static int tile_bits_for(int n_tiles) {
  int bits = 1;
  while ((1 << bits) < n_tiles) ++bits;
  return bits;
}

// packs (tiles > 0) << kRankShift | tiles so that ONE scan yields both the record count and the id-ordered
// rank of every emitting splat
__global__ void __launch_bounds__(256) pack_counts_kernel(const int32_t* __restrict__ tiles, int64_t n,
                                                          int64_t* __restrict__ packed) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t t = tiles[i];
  packed[i] = t > 0 ? (((int64_t)1 << kRankShift) | t) : 0;
}
__global__ void unpack_total_kernel(int64_t* total) { *total &= (((int64_t)1 << kRankShift) - 1); }

int check_render_desc(const bds_render_desc* d);  // projection.cu

}  // namespace bds

using namespace bds;

extern "C" size_t bds_bin_count_workspace_bytes(int64_t n_elems) { return scan_workspace_bytes(n_elems) + 256; }

extern "C" int bds_bin_count(const bds_render_desc* d, const int32_t* tiles_touched, int64_t* isect_offsets,
                             int64_t* n_isect_dev, void* workspace, bds_stream_t stream_) {
  if (int rc = check_render_desc(d)) return rc;
  int64_t n = (int64_t)d->n_gauss * d->n_cams;
  BDS_REQUIRE(n_isect_dev, "bin_count: null n_isect pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n == 0) {
    BDS_CHECK_CUDA(cudaMemsetAsync(n_isect_dev, 0, sizeof(int64_t), stream));
    return 0;
  }
  BDS_REQUIRE(tiles_touched && isect_offsets && workspace, "bin_count: null pointer");
  pack_counts_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(tiles_touched, n, isect_offsets);
  BDS_CHECK_LAUNCH();
  if (int rc = exclusive_scan<int64_t, int64_t>(isect_offsets, isect_offsets, n, n_isect_dev, workspace, stream)) return rc;
  unpack_total_kernel<<<1, 1, 0, stream>>>(n_isect_dev);
  BDS_CHECK_LAUNCH();
  return 0;
}

extern "C" size_t bds_bin_sort_workspace_bytes(const bds_render_desc* d, int64_t n_isect) {
  int64_t ns = d ? (int64_t)d->n_gauss * d->n_cams : 1;  // upper bound of the visible splats
  if (ns > n_isect && n_isect > 0) ns = n_isect;         // every slot emits at least one record
  return carve_sort(n_isect, ns).total + 256;
}

extern "C" int bds_bin_sort(const bds_render_desc* d, int64_t n_isect, int32_t n_slots, const int32_t* radii,
                            const int32_t* tiles_touched, const int64_t* isect_offsets, const float* splats,
                            float* sorted_splats, int32_t* sorted_slots, int32_t* tile_offsets, void* workspace,
                            bds_stream_t stream_) {
  if (int rc = check_render_desc(d)) return rc;
  BDS_REQUIRE(n_isect >= 0 && n_isect < ((int64_t)1 << 31), "bin_sort: n_isect must fit int32 (got %lld)", (long long)n_isect);
  BDS_REQUIRE(tile_offsets, "bin_sort: null tile_offsets");
  if (n_isect > 0)
    BDS_REQUIRE(radii && tiles_touched && isect_offsets && splats && workspace && n_slots > 0 && n_slots <= n_isect,
                "bin_sort: null pointer or inconsistent n_slots");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int tile_w = (d->width + kTile - 1) / kTile, tile_h = (d->height + kTile - 1) / kTile;
  const int n_tiles = (d->row_end - d->row_begin) * tile_w;
  if (n_isect == 0) {
    tile_offsets_kernel<<<ceil_div(n_tiles + 1, 256), 256, 0, stream>>>(nullptr, 0, n_tiles, tile_offsets);
    BDS_CHECK_LAUNCH();
    return 0;
  }
  BDS_REQUIRE(sorted_splats, "bin_sort: null sorted_splats");
  char* ws = static_cast<char*>(workspace);
  int64_t ns_cap = (int64_t)d->n_gauss * d->n_cams;
  if (ns_cap > n_isect) ns_cap = n_isect;
  SortWorkspace w = carve_sort(n_isect, ns_cap);
  uint32_t* keys[2] = {reinterpret_cast<uint32_t*>(ws + w.keys_a), reinterpret_cast<uint32_t*>(ws + w.keys_b)};
  uint32_t* vals[2] = {reinterpret_cast<uint32_t*>(ws + w.vals_a), reinterpret_cast<uint32_t*>(ws + w.vals_b)};
  uint32_t* s1k[2] = {reinterpret_cast<uint32_t*>(ws + w.s1k_a), reinterpret_cast<uint32_t*>(ws + w.s1k_b)};
  uint32_t* s1v[2] = {reinterpret_cast<uint32_t*>(ws + w.s1v_a), reinterpret_cast<uint32_t*>(ws + w.s1v_b)};
  int32_t* tiles_sorted = reinterpret_cast<int32_t*>(ws + w.tiles_sorted);
  int64_t* offs2 = reinterpret_cast<int64_t*>(ws + w.offs2);
  uint32_t* hist = reinterpret_cast<uint32_t*>(ws + w.hist);
  void* scan_ws = ws + w.scan;

  // stage 1: visible splats by (depth bits, Gaussian id)
  stage1_fill_kernel<<<ceil_div(n_slots, 256), 256, 0, stream>>>(splats, n_slots, isect_offsets, s1k[0], s1v[0]);
  BDS_CHECK_LAUNCH();
  int b1 = 0;
  if (int rc = radix_sort_pairs<uint32_t>(s1k, s1v, n_slots, 32, hist, scan_ws, stream, &b1)) return rc;
  // stage 2: emit in that order, stable sort on the band tile index
  gather_tiles_kernel<<<ceil_div(n_slots, 256), 256, 0, stream>>>(splats, s1v[b1], n_slots, tiles_touched, tiles_sorted);
  BDS_CHECK_LAUNCH();
  if (int rc = exclusive_scan<int32_t, int64_t>(tiles_sorted, offs2, n_slots, nullptr, scan_ws, stream)) return rc;
  EmitParams ep;
  ep.d = *d; ep.tile_w = tile_w; ep.tile_h = tile_h; ep.radii = radii; ep.tiles_sorted = tiles_sorted; ep.offsets = offs2;
  ep.sorted_slots = s1v[b1]; ep.n_tiles = n_tiles; ep.n_slots = n_slots; ep.splats = splats; ep.keys = keys[0];
  ep.vals = vals[0];
  {
    static const uint32_t primes[] = {2654435761u, 2246822519u, 3266489917u, 668265263u, 374761393u};
    ep.perm_mul = 1;
    for (uint32_t pr : primes)
      if ((uint64_t)n_slots % pr != 0) { ep.perm_mul = pr; break; }  // prime not dividing n => bijection mod n
  }
  emit_keys_kernel<<<ceil_div(n_slots, 256), 256, 0, stream>>>(ep);
  BDS_CHECK_LAUNCH();
  int b2 = 0;
  if (int rc = radix_sort_pairs<uint32_t>(keys, vals, n_isect, tile_bits_for(n_tiles + 1), hist, scan_ws, stream, &b2)) return rc;
  tile_offsets_kernel<<<ceil_div(n_isect, 256), 256, 0, stream>>>(keys[b2], n_isect, n_tiles, tile_offsets);
  BDS_CHECK_LAUNCH();
  gather_records_kernel<<<ceil_div(n_isect * 3, 256), 256, 0, stream>>>(vals[b2], n_isect,
                                                                       reinterpret_cast<const float4*>(splats),
                                                                       reinterpret_cast<float4*>(sorted_splats));
  BDS_CHECK_LAUNCH();
  if (sorted_slots)
    BDS_CHECK_CUDA(cudaMemcpyAsync(sorted_slots, vals[b2], (size_t)n_isect * 4, cudaMemcpyDeviceToDevice, stream));
  return 0;
}
