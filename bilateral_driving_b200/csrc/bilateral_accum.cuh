// Block-cooperative accumulation of bilateral-grid node gradients, shared by the fused composite
// backward (composite.cu) and the tiled stand-alone bilateral backward (bilateral.cu).
#pragma once
#include "bilateral_math.cuh"

namespace bds {

constexpr unsigned kFull = 0xffffffffu;

// Grid-node gradients of one tile land in a handful of lattice nodes (a 16x16 tile spans a fraction
// of a grid cell at 1080p).  They are accumulated in shared memory and leave as one global reduction
// per touched (node, channel) per tile instead of 96 per pixel per level:
//   1. every thread stages its pixel (vA[12], 4 xy-corner weights, 2 z weights) at a position given by
//      a counting sort on its cell key (z0, oy, ox) inside a 3x3x(L+1)-node window (integer shared
//      atomics are native; pixels whose cell falls outside the window - tiny images / huge grids -
//      scatter straight to global memory instead);
//   2. the 16 half-warps take equal slices of the sorted list; 12 lanes (4 xy-corners x 3 channel
//      quads) accumulate runs of equal key in REGISTERS.  A run that ends inside a slice goes to the
//      window with shared atomic adds (fp32 shared atomics are CAS loops: kept rare); the LAST run of
//      each slice - all slices end together, mostly on the same cells - is parked in a per-half-warp
//      slot with plain stores and 192 threads merge equal-key neighbours into the window afterwards;
//   3. the window is flushed to global memory with one reduction per touched (node, channel).
constexpr int kWinNodes = 3;                   // window is kWinNodes x kWinNodes lattice nodes in xy
constexpr int kWinMaxL = 16;
constexpr int kStageFloats = 20;               // per pixel: vA[12] | wxy[4] | wz0, wz1, base, pad
constexpr int kWinFloats = kWinNodes * kWinNodes * (kWinMaxL + 1) * 12;
constexpr int kWinKeys = kWinNodes * kWinNodes * kWinMaxL;         // 144 cell keys
constexpr int kSlotFloats = 16 * 96;  // last-run register sums of the 16 half-warps: 12 lanes x (a0 | a1)
constexpr size_t kBwdSmemBil =
    (size_t)(256 * kStageFloats + kWinFloats + kSlotFloats) * sizeof(float) + 2 * 160 * sizeof(int);

// tile_x01, tile_y01: lin01() of the tile's first pixel column / row (hoisted by the caller: one IEEE
// division per axis per thread instead of one per level)
BDS_D void level_grad_accumulate(float* smem, const Tri& t, const float vAff[12], bool valid, float tile_x01,
                                 float tile_y01, int L, int GY, int GX, float* __restrict__ v_grid) {
  float* stage = smem;
  float* win = smem + 256 * kStageFloats;
  int* hist = reinterpret_cast<int*>(win + kWinFloats);  // [160] counts -> start offsets
  int* misc = hist + 160;                                // [0] = number of staged pixels
  int* slot_key = misc + 8;                              // [16] window offset of each half-warp's last run (-1: none)
  float* slot_val = reinterpret_cast<float*>(hist + 320);  // [16][12 lanes][a0 | a1]
  // window origin = cell of the tile's first pixel (uniform over the block)
  const float fx0 = fminf(fmaxf(unit_coord(tile_x01, GX), 0.f), (float)(GX - 1));
  const float fy0 = fminf(fmaxf(unit_coord(tile_y01, GY), 0.f), (float)(GY - 1));
  const int nx0 = (int)floorf(fx0), ny0 = (int)floorf(fy0);
  const bool use_win = L <= kWinMaxL;
  const int slab = kWinNodes * kWinNodes * 12;           // floats per z slab
  const int per_win = slab * (L + 1);                    // + one dummy slab so z0 + 1 is always in range
  const int ox = t.x0 - nx0, oy = t.y0 - ny0;
  const bool in_win = use_win && ox >= 0 && ox + 1 < kWinNodes && oy >= 0 && oy + 1 < kWinNodes;
  if (valid && !in_win) tri_scatter(v_grid, t, vAff);    // rare: straight to global memory
  if (!use_win) return;                                  // uniform over the block
  {
    float4* w4 = reinterpret_cast<float4*>(win);
    for (int i = threadIdx.x; i < per_win / 4; i += 256) w4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x < 160) hist[threadIdx.x] = 0;
  }
  __syncthreads();
  // ---- counting sort by cell key
  const bool staged = valid && in_win;
  const int key = staged ? (t.z0 * kWinNodes + oy) * kWinNodes + ox : 0;
  int rank = 0;
  if (staged) rank = atomicAdd(&hist[key], 1);
  __syncthreads();
  if (threadIdx.x < 32) {  // exclusive scan of 160 counters: 5 per lane
    int v[5], s = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) { v[k] = hist[threadIdx.x * 5 + k]; s += v[k]; }
    int inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int u = __shfl_up_sync(kFull, inc, o);
      if ((int)threadIdx.x >= o) inc += u;
    }
    int ex = inc - s;
#pragma unroll
    for (int k = 0; k < 5; ++k) { hist[threadIdx.x * 5 + k] = ex; ex += v[k]; }
    if (threadIdx.x == 31) misc[0] = inc;
  }
  __syncthreads();
  if (staged) {
    const float wx0 = 1.f - t.wx1, wy0 = 1.f - t.wy1;
    float4* sp = reinterpret_cast<float4*>(stage + (hist[key] + rank) * kStageFloats);
    sp[0] = make_float4(vAff[0], vAff[1], vAff[2], vAff[3]);
    sp[1] = make_float4(vAff[4], vAff[5], vAff[6], vAff[7]);
    sp[2] = make_float4(vAff[8], vAff[9], vAff[10], vAff[11]);
    sp[3] = make_float4(wx0 * wy0, t.wx1 * wy0, wx0 * t.wy1, t.wx1 * t.wy1);
    sp[4] = make_float4(1.f - t.wz1, t.dz != 0 ? t.wz1 : 0.f, __int_as_float(key * 12), 0.f);
  }
  __syncthreads();
  {
    // half-warp h takes the h-th sixteenth of the sorted list; lanes 0..11 = 4 corners x 3 channel quads
    const int n_staged = misc[0];
    const int hw = threadIdx.x >> 4, l16 = threadIdx.x & 15;
    const int per = (n_staged + 15) >> 4;
    const int p0 = hw * per, p1 = min(n_staged, p0 + per);
    const bool worker = l16 < 12;
    const int corner = l16 / 3, quad = l16 - corner * 3;
    const int coff = ((corner >> 1) * kWinNodes + (corner & 1)) * 12 + quad * 4;  // (dy, dx) node + channel quad
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    int cur = -1;
    if (worker) {
      if (p0 < p1) {
        // outer loop = runs of equal cell key, inner loop = the pixels of one run (register accumulation
        // with nothing else live across it); the key of the next pixel rides on the float4 the next
        // iteration needs anyway
        const float* sp = stage + p0 * kStageFloats;
        float4 m = *reinterpret_cast<const float4*>(sp + 16);
        int px = p0;
        for (;;) {
          cur = __float_as_int(m.z);
          f32x2 a0l = pk2(0.f, 0.f), a0h = a0l, a1l = a0l, a1h = a0l;   // packed fp32x2 accumulators (FFMA2)
          int key;
          do {
            const float4 va = *reinterpret_cast<const float4*>(sp + quad * 4);
            const float wc = sp[12 + corner];
            const float w0 = wc * m.x, w1 = wc * m.y;
            const f32x2 ww0 = pk2(w0, w0), ww1 = pk2(w1, w1), vl = pk2(va.x, va.y), vh = pk2(va.z, va.w);
            fma2_acc(a0l, ww0, vl); fma2_acc(a0h, ww0, vh);
            fma2_acc(a1l, ww1, vl); fma2_acc(a1h, ww1, vh);
            ++px;
            sp += kStageFloats;
            key = -1;
            if (px < p1) {
              m = *reinterpret_cast<const float4*>(sp + 16);
              key = __float_as_int(m.z);
            }
          } while (key == cur);
          upk2(a0l, a0.x, a0.y); upk2(a0h, a0.z, a0.w);
          upk2(a1l, a1.x, a1.y); upk2(a1h, a1.z, a1.w);
          if (px >= p1) break;  // the slice's last run stays in registers (parked below)
          float* c0 = win + cur + coff;   // a run that ends inside the slice: shared atomics (rare)
          if (a0.x != 0.f) atomicAdd(c0, a0.x);
          if (a0.y != 0.f) atomicAdd(c0 + 1, a0.y);
          if (a0.z != 0.f) atomicAdd(c0 + 2, a0.z);
          if (a0.w != 0.f) atomicAdd(c0 + 3, a0.w);
          if (a1.x != 0.f) atomicAdd(c0 + slab, a1.x);
          if (a1.y != 0.f) atomicAdd(c0 + slab + 1, a1.y);
          if (a1.z != 0.f) atomicAdd(c0 + slab + 2, a1.z);
          if (a1.w != 0.f) atomicAdd(c0 + slab + 3, a1.w);
        }
      }
      // every slice ends at about the same time and mostly on the same few cells: instead of 16 x 96
      // contended shared atomics the last runs are parked in per-half-warp slots ...
      float4* sv = reinterpret_cast<float4*>(slot_val + (hw * 12 + l16) * 8);
      sv[0] = a0;
      sv[1] = a1;
    }
    if (l16 == 0) slot_key[hw] = cur;
  }
  __syncthreads();
  if (threadIdx.x < 192) {
    // ... and merged here: slices are contiguous pieces of a sorted list, so equal last keys sit in
    // consecutive slots; thread (lane, value) sums each group and adds it to the window once
    const int j = threadIdx.x % 96, h0 = (threadIdx.x / 96) * 8;
    const int lane12 = j >> 3, v = j & 7;
    const int corner = lane12 / 3, quad = lane12 - corner * 3;
    const int off = ((corner >> 1) * kWinNodes + (corner & 1)) * 12 + quad * 4 + (v >> 2) * slab + (v & 3);
    float sum = 0.f;
    int kcur = slot_key[h0];
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      sum += slot_val[(h0 + h) * 96 + j];
      const int knext = (h < 7) ? slot_key[h0 + h + 1] : -2;
      if (knext != kcur) {
        if (kcur >= 0 && sum != 0.f) atomicAdd(win + kcur + off, sum);
        sum = 0.f;
        kcur = knext;
      }
    }
  }
  __syncthreads();
  {
    const int n_out4 = slab * L / 4;  // float4 entries; a float4 never straddles a node (12 floats per node)
    const float4* win4 = reinterpret_cast<const float4*>(win);
    for (int e = threadIdx.x; e < n_out4; e += 256) {
      const float4 v = win4[e];
      if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) {
        const int node = e / 3, ch0 = (e - node * 3) * 4;
        const int nx = node % kWinNodes, ny = (node / kWinNodes) % kWinNodes, z = node / (kWinNodes * kWinNodes);
        const int gx = nx0 + nx, gy = ny0 + ny;
        if (gx < GX && gy < GY) {
          float* dst = v_grid + (size_t)bil_node(gx, gy, z, L, GX) * 12 + ch0;
          if (v.x != 0.f) red_add(dst, v.x);
          if (v.y != 0.f) red_add(dst + 1, v.y);
          if (v.z != 0.f) red_add(dst + 2, v.z);
          if (v.w != 0.f) red_add(dst + 3, v.w);
        }
      }
    }
  }
  __syncthreads();
}



// =================================================================================================
// Warp-level variant for RANK-ONE cotangents (fused composite backward): vAff = g (x) [x; 1].
//
// The block-wide counting sort above costs seven block barriers per level, and measured on the benchmark workload
// the barriers - not the instructions - were what the grid-gradient accumulation cost inside the composite backward
// (every warp of the tile waits for the slowest one at each of them).  Here a warp reduces its own 32 pixels and
// sends the result to global memory by itself - no block barrier, no shared window:
//   1. each lane stages its pixel - [x; 1], g, the eight corner weights - in a 2 KB per-warp panel;
//   2. __match_any_sync groups the lanes by lattice cell (x0, y0, z0).  Rendered images are smooth: nearly every
//      warp of the benchmark meets ONE cell per level.  For one cell at a time the warp TRANSPOSES the work: lane =
//      (pixel row of the 8x4 rectangle, corner).  The eight lanes of a row own the eight corner nodes of the cell and
//      sum w_corner * g (x) [x; 1] - all twelve channels, six packed fp32x2 FMAs per pixel - over the eight pixels of
//      their row (fully unrolled, broadcast shared-memory reads; a pixel of another cell takes part with weight
//      zero); two xor-shuffles fold the four rows;
//   3. the 3x4 sum of every corner node leaves as three 128-bit vector reductions (red.global.add.v4.f32): 72 per
//      warp and level in the common case instead of 96 scalar ones per PIXEL and level.
// Entry of pixel (row q, column m) sits at index 9 q + m: the four rows are read at the same m by the four lane groups,
// and a row stride of 9 entries puts them in different shared-memory banks.
struct WarpPanel {
  float4 x1[36];      // {x_r, x_g, x_b, 1}: level input of the pixel
  float4 g[36];       // {g_r, g_g, g_b, -}: cotangent of the level output
  float w[36][8];     // corner weights, index = (slab << 2) | (y << 1) | x
};
constexpr size_t kPanelBytes = 8 * sizeof(WarpPanel);                                  // 8 warps

BDS_D void warp_level_accumulate(WarpPanel* pn, const Tri& t, bool valid, float g0, float g1, float g2, float x0,
                                 float x1, float x2, int L, int GY, int GX, float* __restrict__ v_grid) {
  const int lane = threadIdx.x & 31;
  const int key = valid ? (t.z0 * GY + t.y0) * GX + t.x0 : -1;     // lattice cell of the pixel
  __syncwarp();                 // the previous level's panel has been consumed
  {   // every lane writes its entry (zeros without a pixel: the unrolled loop below reads all of them)
    const float sv = valid ? 1.f : 0.f;
    const float wx0 = 1.f - t.wx1, wy0 = 1.f - t.wy1;
    const float wz0 = sv * (1.f - t.wz1), wz1 = t.dz != 0 ? sv * t.wz1 : 0.f;
    const float w00 = wx0 * wy0, w01 = t.wx1 * wy0, w10 = wx0 * t.wy1, w11 = t.wx1 * t.wy1;
    const int slot = lane + (lane >> 3);
    pn->x1[slot] = valid ? make_float4(x0, x1, x2, 1.f) : make_float4(0.f, 0.f, 0.f, 0.f);
    pn->g[slot] = valid ? make_float4(g0, g1, g2, 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4* wp = reinterpret_cast<float4*>(pn->w[slot]);
    wp[0] = make_float4(w00 * wz0, w01 * wz0, w10 * wz0, w11 * wz0);
    wp[1] = make_float4(w00 * wz1, w01 * wz1, w10 * wz1, w11 * wz1);
  }
  const unsigned grp = __match_any_sync(kFull, key);
  unsigned leaders = __ballot_sync(kFull, valid && lane == __ffs(grp) - 1);
  __syncwarp();
  const int c = lane & 7, q = lane >> 3;               // corner node of the cell, pixel row of the 8x4 rectangle
  const float* wf = &pn->w[9 * q][0] + c;
  const float4* gq = pn->g + 9 * q;
  const float4* x1q = pn->x1 + 9 * q;
  while (leaders) {             // uniform: one cell per round
    const int l = __ffs(leaders) - 1;
    leaders &= leaders - 1;
    const unsigned gmask = __shfl_sync(kFull, grp, l);
    const int k = __shfl_sync(kFull, key, l);
    const unsigned rowmask = (gmask >> (8 * q)) & 0xffu;   // the cell's pixels in this lane group's row
    f32x2 a0 = pk2(0.f, 0.f), a1 = a0, a2 = a0, a3 = a0, a4 = a0, a5 = a0;   // rows r, g, b of the 3x4 sum
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const float wc = (rowmask >> m) & 1u ? wf[m * 8] : 0.f;
      const float4 gg = gq[m];
      const float4 x = x1q[m];
      const f32x2 xa = pk2(x.x, x.y), xb = pk2(x.z, x.w);
      const float s0 = wc * gg.x, s1 = wc * gg.y, s2 = wc * gg.z;
      const f32x2 p0 = pk2(s0, s0), p1 = pk2(s1, s1), p2 = pk2(s2, s2);
      fma2_acc(a0, p0, xa); fma2_acc(a1, p0, xb);
      fma2_acc(a2, p1, xa); fma2_acc(a3, p1, xb);
      fma2_acc(a4, p2, xa); fma2_acc(a5, p2, xb);
    }
    // fold the four rows (lanes c, c + 8, c + 16, c + 24 hold the same corner)
    float v[12];
    upk2(a0, v[0], v[1]); upk2(a1, v[2], v[3]); upk2(a2, v[4], v[5]);
    upk2(a3, v[6], v[7]); upk2(a4, v[8], v[9]); upk2(a5, v[10], v[11]);
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      v[i] += __shfl_xor_sync(kFull, v[i], 8);
      v[i] += __shfl_xor_sync(kFull, v[i], 16);
    }
    // a zero weight (clamped slab / lattice edge) leaves exact zeros in every channel: nothing to add
    if (q == 0 && (v[3] != 0.f || v[7] != 0.f || v[11] != 0.f || v[0] != 0.f || v[5] != 0.f || v[10] != 0.f)) {
      const int cx = k % GX, cy = (k / GX) % GY, cz = k / (GX * GY);
      const int gx = min(cx + (c & 1), GX - 1), gy = min(cy + ((c >> 1) & 1), GY - 1), gz = min(cz + (c >> 2), L - 1);
      float* dst = v_grid + (size_t)bil_node(gx, gy, gz, L, GX) * 12;
      red_add_v4(dst, v[0], v[1], v[2], v[3]);
      red_add_v4(dst + 4, v[4], v[5], v[6], v[7]);
      red_add_v4(dst + 8, v[8], v[9], v[10], v[11]);
    }
  }
}

}  // namespace bds
