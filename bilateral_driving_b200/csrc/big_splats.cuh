// Splats whose footprint covers hundreds of tiles (a Gaussian a metre in front of a camera covers the whole
// image: 8160 tiles) do not belong in the warp-cooperative tile enumeration of the projection / emission passes:
// the warp that owns one runs 255 rounds while the others run two, and the kernel waits for it.  Both passes push
// such splats (more than kBigCand candidate tiles; a few hundred of the 2.2 M at the bench size) onto a small
// device queue and this kernel handles them afterwards, one CTA per splat, 256 candidate tiles at a time.
#pragma once
#include "projection_math.cuh"

namespace bds {

#ifndef BDS_BIG_CAND
#define BDS_BIG_CAND 256
#endif
constexpr int kBigCand = BDS_BIG_CAND;        // candidate tiles above which a splat leaves the warp-cooperative loop
constexpr int kBigQueueCap = BDS_COUNTERS_LEN - 4;   // queue entries; splats beyond it stay in the warp loop

struct BigSplatParams {
  bds_render_desc d;
  int tile_w, tile_h;
  const float* splats;         // packed records (bds_project_fwd)
  const int32_t* radii;        // [C,N]
  const int32_t* n_queue;      // device count (may exceed the capacity: clamped)
  const int32_t* queue;        // splat slots (-1: dropped)
  // count mode
  int32_t* tile_counts;
  int32_t* tiles_touched;
  // emit mode
  const int32_t* tile_offsets;
  int32_t* cursors;
  uint64_t* keys;
};

template <bool EMIT>
__global__ void __launch_bounds__(256) big_splat_kernel(BigSplatParams p) {
  __shared__ int s_hits;
  const int n = min(*p.n_queue, kBigQueueCap);
  const int N = p.d.n_gauss;
  for (int e = blockIdx.x; e < n; e += gridDim.x) {
    const int slot = p.queue[e];
    if (slot < 0) continue;  // uniform over the block
    const float4* rp = reinterpret_cast<const float4*>(p.splats + (size_t)slot * 12);
    const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1), r2 = __ldg(rp + 2);
    const int64_t idx = (int64_t)__float_as_int(r2.z);
    const int c = (int)(idx / N);
    int g0 = c * p.tile_h, g1 = g0 + p.tile_h;
    int lo = p.d.row_begin > g0 ? p.d.row_begin : g0, hi = p.d.row_end < g1 ? p.d.row_end : g1;
    const int ty0 = hi > lo ? lo - g0 : 0, ty1 = hi > lo ? hi - g0 : 0;
    const float cut = r2.w + kLog2_255;
    const TileRect tr = candidate_rect(r0.x, r0.y, (float)p.radii[idx], r0.z, r0.w, r1.x, cut, p.tile_w, p.tile_h, ty0, ty1);
    const int rw = tr.x1 - tr.x0, total = rw * (tr.y1 - tr.y0);
    if (!EMIT) {
      if (threadIdx.x == 0) s_hits = 0;
      __syncthreads();
    }
    const uint64_t key = ((uint64_t)(uint32_t)__float_as_int(r2.y) << 32) | (uint64_t)(uint32_t)slot;
    int mine = 0;
    for (int i = threadIdx.x; i < total; i += 256) {
      const int ry = i / rw;
      const int tx = tr.x0 + i - ry * rw, ty = tr.y0 + ry;
      if (tile_hit(r0.x, r0.y, r0.z, r0.w, r1.x, cut, tx, ty, p.d.width, p.d.height)) {
        const int tile = (c * p.tile_h + ty - p.d.row_begin) * p.tile_w + tx;
        if (EMIT) {
          const int seg0 = p.tile_offsets[tile], cap = p.tile_offsets[tile + 1] - seg0;
          const int pos = atomicAdd(p.cursors + tile, 1);
          if (pos < cap) p.keys[seg0 + pos] = key;
        } else {
          if (p.tile_counts) atomicAdd(p.tile_counts + tile, 1);
          ++mine;
        }
      }
    }
    if (!EMIT) {
      mine = (int)warp_sum((float)mine);   // <= 8160 per lane: exact in fp32
      if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_hits, mine);
      __syncthreads();
      if (threadIdx.x == 0) p.tiles_touched[idx] = s_hits;
      __syncthreads();
    }
  }
}

}  // namespace bds
