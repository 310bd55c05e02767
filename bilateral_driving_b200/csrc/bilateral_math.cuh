// Per-pixel arithmetic of the multi-scale bilateral-grid chain, shared by the stand-alone
// bilateral kernels (bilateral.cu) and the fused composite epilogue (composite.cu).
//
// Semantics follow the reference (paths relative to /root/reference/project):
//   bilateral/lib_bilagrid.py:317-368   xy in [0,1] -> [-1,1], z = luma*2-1, F.grid_sample 5-D
//                                       trilinear, align_corners=True, padding_mode="border",
//                                       channel c = 4*row + col of the 3x4 affine
//   models/modules.py:494-504, 409-420  bilinear resize, align_corners=False (low-res guidance)
//   models/trainers/scene_graph.py:112-117   x <- A[:, :3] x + A[:, 3], level after level
//
// Grid VALUES are read from a guidance-contiguous, quad-major repack [GY][GX][3][L][4] that the host
// entry points build from the reference's channel-first [12][L][GY][GX] parameter slot: the four floats
// of one affine row (one 128-bit load) of the L slabs of an xy node are 16*L contiguous bytes.  The lanes
// of a warp share the xy cell and differ in their guidance slab, so ONE load instruction (fixed corner,
// fixed row) touches at most 16*L bytes = 2 cache lines at L = 16 - the slice is L1-wavefront bound, and
// the node-major [GY][GX][L][12] order it replaced spread the same instruction over 48*L bytes (6 lines).
// Grid GRADIENTS keep the node-major [GY][GX][L][12] order (bil_node), which is what the window flush of
// bilateral_accum.cuh writes.
#pragma once
#include "bds_common.cuh"

namespace bds {

constexpr float kLumaR = 0.299f, kLumaG = 0.587f, kLumaB = 0.114f;

// torch bilinear (align_corners=False) source taps: models/modules.py:497, :414-419
struct LinTap {
  int i0, i1;
  float t;
};
// IEEE division even under --use_fast_math: the resize taps must round like torch's fp32 ones
BDS_HD float div_rn(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}
BDS_HD LinTap lin_src(int d, int in_size, int out_size) {
  LinTap r;
  float scale = div_rn((float)in_size, (float)out_size);
  float s = scale * ((float)d + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  int i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  r.i0 = i0;
  r.i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  r.t = s - (float)i0;
  return r;
}

// torch.linspace(0, 1, n)[j] in fp32 (symmetric evaluation, as ATen does)
BDS_HD float lin01(int j, int n) {
  if (n <= 1) return 0.f;
  float step = div_rn(1.0f, (float)(n - 1));
  return (j < n / 2) ? step * (float)j : 1.0f - step * (float)(n - 1 - j);
}

// lattice coordinate -> voxel units, as BilateralGrid.forward + grid_sampler_unnormalize do
BDS_HD float lattice_coord(int j, int n, int gsize) {
  float x = lin01(j, n);
  float xn = (x - 0.5f) * 2.0f;
  return ((xn + 1.0f) * 0.5f) * (float)(gsize - 1);
}
BDS_HD float unit_coord(float x01, int gsize) {
  float xn = (x01 - 0.5f) * 2.0f;
  return ((xn + 1.0f) * 0.5f) * (float)(gsize - 1);
}
BDS_HD float luma_of(float r, float g, float b) { return r * kLumaR + g * kLumaG + b * kLumaB; }
BDS_HD float luma_coord(float luma, int L) {
  float z = luma * 2.0f - 1.0f;
  return ((z + 1.0f) * 0.5f) * (float)(L - 1);
}

// node index of (x, y, z) in the [GY][GX][L] repack, and the inverse map used by the repack kernels
BDS_HD int bil_node(int x, int y, int z, int L, int GX) { return (y * GX + x) * L + z; }
// element i of the repack ([node][12]) <-> element of the parameter layout [12][L][GY][GX]
BDS_HD size_t bil_param_index(int node, int ch, int L, int GY, int GX) {
  int z = node % L, xy = node / L;
  int x = xy % GX, y = xy / GX;
  return (((size_t)ch * L + z) * GY + y) * GX + x;
}
// value repack [GY][GX][3][L][4]: float offset of (node = xy * L + z, affine row k), and the inverse map
// (element i of the value repack -> element of the parameter layout) used by the repack kernels
BDS_HD int bil_value_offset(int node, int z, int k, int L) { return 12 * node - 8 * z + 4 * k * L; }
BDS_HD size_t bil_value_param_index(int i, int L, int GY, int GX) {
  int xy = i / (12 * L), rem = i - xy * 12 * L;
  int k = rem / (4 * L), rem2 = rem - k * 4 * L;
  int z = rem2 >> 2, ch = 4 * k + (rem2 & 3);
  int x = xy % GX, y = xy / GX;
  return (((size_t)ch * L + z) * GY + y) * GX + x;
}

// Trilinear set-up with border clamping.  Offsets are in lattice NODES of the [GY][GX][L] repack.
struct Tri {
  int n00, n01, n10, n11;  // (y0,x0) (y0,x1) (y1,x0) (y1,x1) node offsets inside slab z0
  int dz;                  // node offset from slab z0 to slab z1 (0 when clamped)
  int x0, y0, z0;          // lower lattice node of the cell
  float wx1, wy1, wz1;     // upper weights; lower = 1 - upper
  bool z_inside;           // 0 < fz < L-1 (strict): guidance gradient flows
  int L;                   // slabs of the level (value-repack addressing)
};
BDS_HD Tri tri_setup(float fx, float fy, float fz, int L, int GY, int GX) {
  Tri t;
  t.z_inside = (fz > 0.f) && (fz < (float)(L - 1));
  fx = fminf(fmaxf(fx, 0.f), (float)(GX - 1));
  fy = fminf(fmaxf(fy, 0.f), (float)(GY - 1));
  fz = fminf(fmaxf(fz, 0.f), (float)(L - 1));
  int x0 = (int)floorf(fx), y0 = (int)floorf(fy), z0 = (int)floorf(fz);
  t.wx1 = fx - (float)x0;
  t.wy1 = fy - (float)y0;
  t.wz1 = fz - (float)z0;
  int x1 = x0 + 1 < GX ? x0 + 1 : x0;  // weight of the clamped corner is exactly 0
  int y1 = y0 + 1 < GY ? y0 + 1 : y0;
  int z1 = z0 + 1 < L ? z0 + 1 : z0;
  t.x0 = x0; t.y0 = y0; t.z0 = z0;
  t.n00 = bil_node(x0, y0, z0, L, GX);
  t.n01 = bil_node(x1, y0, z0, L, GX);
  t.n10 = bil_node(x0, y1, z0, L, GX);
  t.n11 = bil_node(x1, y1, z0, L, GX);
  t.dz = z1 - z0;
  t.L = L;
  return t;
}

#ifdef __CUDACC__
BDS_D void load12(const float* p, float v[12]) {
  const float4* q = reinterpret_cast<const float4*>(p);
  float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w;
}

// the three affine rows of one node of the value repack (rows are 4 L floats apart) as six fp32 pairs
BDS_D void load12v(const float* p, int L, f32x2 v[6]) {
  const float4* q = reinterpret_cast<const float4*>(p);
  float4 a = __ldg(q), b = __ldg(q + L), c = __ldg(q + 2 * L);
  v[0] = pk2(a.x, a.y); v[1] = pk2(a.z, a.w);
  v[2] = pk2(b.x, b.y); v[3] = pk2(b.z, b.w);
  v[4] = pk2(c.x, c.y); v[5] = pk2(c.z, c.w);
}

// A = trilerp(grid)(t); optionally dA/dfz (slab difference, xy-interpolated).  g = value repack.
// 8 corners x 12 channels as packed fp32x2 FMAs (48 FFMA2 instead of 96 FFMA): the slice is issue-bound.
template <bool WITH_DZ>
BDS_D void tri_fetch(const float* __restrict__ g, const Tri& t, float A[12], float dAdz[12]) {
#ifdef BDS_DIAG_NO_FETCH   // timing experiments only: no grid loads (results are wrong)
#pragma unroll
  for (int k = 0; k < 12; ++k) { A[k] = t.wx1 + (float)k; if (WITH_DZ) dAdz[k] = t.wy1 * (float)k; }
  return;
#endif
  const float w00 = (1.f - t.wx1) * (1.f - t.wy1), w01 = t.wx1 * (1.f - t.wy1);
  const float w10 = (1.f - t.wx1) * t.wy1, w11 = t.wx1 * t.wy1;
  const f32x2 p00 = pk2(w00, w00), p01 = pk2(w01, w01), p10 = pk2(w10, w10), p11 = pk2(w11, w11);
  f32x2 c0[6], c1[6], v[6];
  g -= 8 * t.z0;   // value repack: row k of node n lives at 12 n - 8 z + 4 k L (bil_value_offset)
  load12v(g + 12 * t.n00, t.L, v);
#pragma unroll
  for (int k = 0; k < 6; ++k) c0[k] = mul2(p00, v[k]);
  load12v(g + 12 * t.n01, t.L, v);
#pragma unroll
  for (int k = 0; k < 6; ++k) fma2_acc(c0[k], p01, v[k]);
  load12v(g + 12 * t.n10, t.L, v);
#pragma unroll
  for (int k = 0; k < 6; ++k) fma2_acc(c0[k], p10, v[k]);
  load12v(g + 12 * t.n11, t.L, v);
#pragma unroll
  for (int k = 0; k < 6; ++k) fma2_acc(c0[k], p11, v[k]);
  g += 4 * t.dz;   // upper slab: node + dz, z + dz  ->  12 dz - 8 dz
  load12v(g + 12 * t.n00, t.L, v);
#pragma unroll
  for (int k = 0; k < 6; ++k) c1[k] = mul2(p00, v[k]);
  load12v(g + 12 * t.n01, t.L, v);
#pragma unroll
  for (int k = 0; k < 6; ++k) fma2_acc(c1[k], p01, v[k]);
  load12v(g + 12 * t.n10, t.L, v);
#pragma unroll
  for (int k = 0; k < 6; ++k) fma2_acc(c1[k], p10, v[k]);
  load12v(g + 12 * t.n11, t.L, v);
#pragma unroll
  for (int k = 0; k < 6; ++k) fma2_acc(c1[k], p11, v[k]);
  const f32x2 pz = pk2(t.wz1, t.wz1);
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const f32x2 d = sub2(c1[k], c0[k]);
    upk2(fma2(pz, d, c0[k]), A[2 * k], A[2 * k + 1]);
    if (WITH_DZ) upk2(d, dAdz[2 * k], dAdz[2 * k + 1]);
  }
}

// scatter w * vA into the 8 nodes (global fp32 reductions)
BDS_D void tri_scatter(float* __restrict__ vg, const Tri& t, const float vA[12]) {
  float wx0 = 1.f - t.wx1, wy0 = 1.f - t.wy1, wz0 = 1.f - t.wz1;
  const int nodes[4] = {t.n00, t.n01, t.n10, t.n11};
  const float wxy[4] = {wx0 * wy0, t.wx1 * wy0, wx0 * t.wy1, t.wx1 * t.wy1};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float w0 = wxy[c] * wz0, w1 = wxy[c] * t.wz1;
    float* p0 = vg + 12 * nodes[c];
    float* p1 = vg + 12 * (nodes[c] + t.dz);
    if (w0 != 0.f) {
#pragma unroll
      for (int k = 0; k < 12; ++k) red_add(p0 + k, w0 * vA[k]);
    }
    if (w1 != 0.f) {
#pragma unroll
      for (int k = 0; k < 12; ++k) red_add(p1 + k, w1 * vA[k]);
    }
  }
}
#endif  // __CUDACC__

// x <- A[:, :3] x + A[:, 3]
BDS_HD void affine_apply(const float A[12], float& r, float& g, float& b) {
  float nr = A[0] * r + A[1] * g + A[2] * b + A[3];
  float ng = A[4] * r + A[5] * g + A[6] * b + A[7];
  float nb = A[8] * r + A[9] * g + A[10] * b + A[11];
  r = nr; g = ng; b = nb;
}
// cotangents: vA += gout (x) [x;1];  gin = A[:, :3]^T gout
BDS_HD void affine_apply_bwd(const float A[12], float xr, float xg, float xb, float gr, float gg,
                             float gb, float vA[12], float& or_, float& og, float& ob) {
  vA[0] += gr * xr; vA[1] += gr * xg; vA[2] += gr * xb; vA[3] += gr;
  vA[4] += gg * xr; vA[5] += gg * xg; vA[6] += gg * xb; vA[7] += gg;
  vA[8] += gb * xr; vA[9] += gb * xg; vA[10] += gb * xb; vA[11] += gb;
  or_ = A[0] * gr + A[4] * gg + A[8] * gb;
  og = A[1] * gr + A[5] * gg + A[9] * gb;
  ob = A[2] * gr + A[6] * gg + A[10] * gb;
}

// Device-side description of one level / the whole chain ------------------------------------------
struct BilLevel {
  const float* grid_cl;  // [L][GY][GX][12] repack
  float* v_grid_cl;      // same shape, accumulated
  const float* a_low;    // [Hd][Wd][12] low-res affine field (factor > 1)
  float* v_a_low;        // same shape, accumulated
  float* affine_out;     // optional full-res [H][W][12]
  const float* v_affine; // optional cotangent on affine_out
  int L, GY, GX, factor, Hd, Wd;
};
struct BilChain {
  int n_levels;
  int H, W;
  BilLevel lv[BDS_MAX_LEVELS];
};

}  // namespace bds
