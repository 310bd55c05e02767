// Per-warp reduction of bilateral-grid node gradients, shared by the fused composite backward (composite.cu, rank-one
// cotangents) and the tiled stand-alone bilateral backward (bilateral.cu, general 12-channel cotangents).
#pragma once
#include "bilateral_math.cuh"

namespace bds {

constexpr unsigned kFull = 0xffffffffu;

// Round 1 reduced a tile's grid-node gradients block-wide: a counting sort of the 256 pixels by lattice cell, register
// accumulation over the sorted runs, a shared window of nodes - seven block barriers per level.  Measured inside the
// composite backward, the BARRIERS (every warp of the tile waits for the slowest at each) - not the instructions -
// were what that cost: four designs with very different instruction counts ran within 2 % of each other until the
// barriers went.  Both variants below therefore work per warp and send their sums straight to global memory.

// RANK-ONE cotangents (fused composite backward): vAff = g (x) [x; 1].  A warp reduces its own 32 pixels and sends the
// result to global memory by itself - no block barrier, no shared window:
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

// The same reduction for a GENERAL 12-channel cotangent per pixel (stand-alone bilateral backward: the cotangent of a
// low-resolution affine field is a gathered sum, not an outer product).  Lanes may map to any pixels of the tile.
struct WarpPanel12 {
  float4 v[36][3];    // the pixel's 3x4 cotangent, rows r, g, b (entry index 9 q + m, see WarpPanel)
  float w[36][8];     // corner weights, index = (slab << 2) | (y << 1) | x
};

BDS_D void warp_level_accumulate12(WarpPanel12* pn, const Tri& t, bool valid, const float vA[12], int L, int GY, int GX,
                                   float* __restrict__ v_grid) {
  const int lane = threadIdx.x & 31;
  const int key = valid ? (t.z0 * GY + t.y0) * GX + t.x0 : -1;     // lattice cell of the pixel
  __syncwarp();                 // the previous use of the panel has been consumed
  {
    const float sv = valid ? 1.f : 0.f;
    const float wx0 = 1.f - t.wx1, wy0 = 1.f - t.wy1;
    const float wz0 = sv * (1.f - t.wz1), wz1 = t.dz != 0 ? sv * t.wz1 : 0.f;
    const float w00 = wx0 * wy0, w01 = t.wx1 * wy0, w10 = wx0 * t.wy1, w11 = t.wx1 * t.wy1;
    const int slot = lane + (lane >> 3);
    pn->v[slot][0] = make_float4(vA[0] * sv, vA[1] * sv, vA[2] * sv, vA[3] * sv);
    pn->v[slot][1] = make_float4(vA[4] * sv, vA[5] * sv, vA[6] * sv, vA[7] * sv);
    pn->v[slot][2] = make_float4(vA[8] * sv, vA[9] * sv, vA[10] * sv, vA[11] * sv);
    float4* wp = reinterpret_cast<float4*>(pn->w[slot]);
    wp[0] = make_float4(w00 * wz0, w01 * wz0, w10 * wz0, w11 * wz0);
    wp[1] = make_float4(w00 * wz1, w01 * wz1, w10 * wz1, w11 * wz1);
  }
  const unsigned grp = __match_any_sync(kFull, key);
  unsigned leaders = __ballot_sync(kFull, valid && lane == __ffs(grp) - 1);
  __syncwarp();
  const int c = lane & 7, q = lane >> 3;               // corner node of the cell, group of eight lanes
  const float* wf = &pn->w[9 * q][0] + c;
  const float4(*vq)[3] = pn->v + 9 * q;
  while (leaders) {             // uniform: one cell per round
    const int l = __ffs(leaders) - 1;
    leaders &= leaders - 1;
    const unsigned gmask = __shfl_sync(kFull, grp, l);
    const int k = __shfl_sync(kFull, key, l);
    const unsigned rowmask = (gmask >> (8 * q)) & 0xffu;
    f32x2 a0 = pk2(0.f, 0.f), a1 = a0, a2 = a0, a3 = a0, a4 = a0, a5 = a0;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const float wc = (rowmask >> m) & 1u ? wf[m * 8] : 0.f;
      const float4 r0 = vq[m][0], r1 = vq[m][1], r2 = vq[m][2];
      const f32x2 w2 = pk2(wc, wc);
      fma2_acc(a0, w2, pk2(r0.x, r0.y)); fma2_acc(a1, w2, pk2(r0.z, r0.w));
      fma2_acc(a2, w2, pk2(r1.x, r1.y)); fma2_acc(a3, w2, pk2(r1.z, r1.w));
      fma2_acc(a4, w2, pk2(r2.x, r2.y)); fma2_acc(a5, w2, pk2(r2.z, r2.w));
    }
    float v[12];
    upk2(a0, v[0], v[1]); upk2(a1, v[2], v[3]); upk2(a2, v[4], v[5]);
    upk2(a3, v[6], v[7]); upk2(a4, v[8], v[9]); upk2(a5, v[10], v[11]);
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      v[i] += __shfl_xor_sync(kFull, v[i], 8);
      v[i] += __shfl_xor_sync(kFull, v[i], 16);
    }
    bool any = false;
#pragma unroll
    for (int i = 0; i < 12; ++i) any |= v[i] != 0.f;
    if (q == 0 && any) {
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
