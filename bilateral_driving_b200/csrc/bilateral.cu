// Stand-alone multi-scale bilateral-grid slice + sequential 3x4 apply, forward and backward.
// This is the drop-in for the reference's MultiScaleBilateralAffineTransform.forward
// (models/modules.py:505-584) + scene_graph.py:112-117 when it is NOT fused into the composite
// kernel: default low-resolution guidance ([4,4,2]) and the guidance_factor=None branch.
//
// Kernels (all HBM/L2-bound gather / pointwise work; no tensor cores):
//   repack_grid_kernel       [12,L,GY,GX] -> value repack [GY,GX,3,L,4]          (tiny)
//   lowres_slice_fwd_kernel  per low-res pixel: bilinear-down guidance -> luma -> trilerp -> A_low
//   apply_fwd_kernel         per pixel: A_l = up(A_low) or direct trilerp; x <- A_l x
//   apply_bwd_tiled_kernel   per 16x16 tile: recompute chain; g_l; vA_l -> v_A_low by a shared-memory
//                            gather over the tile's low-res footprint, or (full-res level) the
//                            block-cooperative grid-node accumulation + guidance gradient
//   lowres_slice_bwd_tiled_kernel  per 16x16 low-res tile: v_A_low -> v_grid (block-cooperative),
//                            guidance gradient -> transposed down-sample into v_rgb_in
//   unpack_add_grid_kernel   v_grid_cl [L,GY,GX,12] += into the parameter-layout gradient slot
#include "bilateral_accum.cuh"

namespace bds {

__global__ void repack_grid_kernel(const float* __restrict__ cf, float* __restrict__ cl, int L, int GY, int GX) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;  // over nodes*12, channel fastest
  if (i >= L * GY * GX * 12) return;
  cl[i] = cf[bil_value_param_index(i, L, GY, GX)];   // value repack [GY][GX][3][L][4]
}

__global__ void unpack_add_grid_kernel(const float* __restrict__ cl, float* __restrict__ cf, int L, int GY, int GX) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;  // over the parameter layout [12][L][GY][GX] (coalesced writes)
  int nodes = L * GY * GX;
  if (i >= nodes * 12) return;
  int ch = i / nodes, rem = i - ch * nodes;
  int x = rem % GX, y = (rem / GX) % GY, z = rem / (GX * GY);
  cf[i] += cl[(size_t)bil_node(x, y, z, L, GX) * 12 + ch];
}

// bilinear gather of the guidance RGB at low-res pixel (h, w)
BDS_D void downsample_rgb(const float* __restrict__ rgb, int H, int W, const LinTap& ty, const LinTap& tx,
                          float& r, float& g, float& b) {
  const float* p00 = rgb + ((size_t)ty.i0 * W + tx.i0) * 3;
  const float* p01 = rgb + ((size_t)ty.i0 * W + tx.i1) * 3;
  const float* p10 = rgb + ((size_t)ty.i1 * W + tx.i0) * 3;
  const float* p11 = rgb + ((size_t)ty.i1 * W + tx.i1) * 3;
  float w0 = 1.f - tx.t, w1 = tx.t, h0 = 1.f - ty.t, h1 = ty.t;
  r = h0 * (w0 * __ldg(p00) + w1 * __ldg(p01)) + h1 * (w0 * __ldg(p10) + w1 * __ldg(p11));
  g = h0 * (w0 * __ldg(p00 + 1) + w1 * __ldg(p01 + 1)) + h1 * (w0 * __ldg(p10 + 1) + w1 * __ldg(p11 + 1));
  b = h0 * (w0 * __ldg(p00 + 2) + w1 * __ldg(p01 + 2)) + h1 * (w0 * __ldg(p10 + 2) + w1 * __ldg(p11 + 2));
}

__global__ void __launch_bounds__(256) lowres_slice_fwd_kernel(const float* __restrict__ rgb, int H, int W,
                                                               BilLevel lv, float* __restrict__ a_low) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= lv.Hd * lv.Wd) return;
  int h = idx / lv.Wd, w = idx - h * lv.Wd;
  LinTap ty = lin_src(h, H, lv.Hd), tx = lin_src(w, W, lv.Wd);
  float r, g, b;
  downsample_rgb(rgb, H, W, ty, tx, r, g, b);
  float fx = lattice_coord(w, lv.Wd, lv.GX), fy = lattice_coord(h, lv.Hd, lv.GY);
  float fz = luma_coord(luma_of(r, g, b), lv.L);
  Tri t = tri_setup(fx, fy, fz, lv.L, lv.GY, lv.GX);
  float A[12];
  tri_fetch<false>(lv.grid_cl, t, A, nullptr);
  float4* o = reinterpret_cast<float4*>(a_low + (size_t)idx * 12);
  o[0] = make_float4(A[0], A[1], A[2], A[3]);
  o[1] = make_float4(A[4], A[5], A[6], A[7]);
  o[2] = make_float4(A[8], A[9], A[10], A[11]);
}

// full-res affine of one level at pixel (i, j)
BDS_D void level_affine_fwd(const BilLevel& lv, int H, int W, int i, int j, float lum, float A[12]) {
  if (lv.factor > 1) {
    LinTap ty = lin_src(i, lv.Hd, H), tx = lin_src(j, lv.Wd, W);
    float v00[12], v01[12], v10[12], v11[12];
    load12(lv.a_low + ((size_t)ty.i0 * lv.Wd + tx.i0) * 12, v00);
    load12(lv.a_low + ((size_t)ty.i0 * lv.Wd + tx.i1) * 12, v01);
    load12(lv.a_low + ((size_t)ty.i1 * lv.Wd + tx.i0) * 12, v10);
    load12(lv.a_low + ((size_t)ty.i1 * lv.Wd + tx.i1) * 12, v11);
    float w0 = 1.f - tx.t, w1 = tx.t, h0 = 1.f - ty.t, h1 = ty.t;
#pragma unroll
    for (int k = 0; k < 12; ++k) A[k] = h0 * (w0 * v00[k] + w1 * v01[k]) + h1 * (w0 * v10[k] + w1 * v11[k]);
  } else {
    Tri t = tri_setup(lattice_coord(j, W, lv.GX), lattice_coord(i, H, lv.GY), luma_coord(lum, lv.L), lv.L,
                      lv.GY, lv.GX);
    tri_fetch<false>(lv.grid_cl, t, A, nullptr);
  }
}

__global__ void __launch_bounds__(256) apply_fwd_kernel(const float* __restrict__ rgb_in, float* __restrict__ rgb_out,
                                                        BilChain ch) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ch.H * ch.W) return;
  int i = idx / ch.W, j = idx - i * ch.W;
  float r = rgb_in[(size_t)idx * 3], g = rgb_in[(size_t)idx * 3 + 1], b = rgb_in[(size_t)idx * 3 + 2];
  float lum = luma_of(r, g, b);
  for (int l = 0; l < ch.n_levels; ++l) {
    float A[12];
    level_affine_fwd(ch.lv[l], ch.H, ch.W, i, j, lum, A);
    if (ch.lv[l].affine_out) {
      float4* o = reinterpret_cast<float4*>(ch.lv[l].affine_out + (size_t)idx * 12);
      o[0] = make_float4(A[0], A[1], A[2], A[3]);
      o[1] = make_float4(A[4], A[5], A[6], A[7]);
      o[2] = make_float4(A[8], A[9], A[10], A[11]);
    }
    affine_apply(A, r, g, b);
  }
  rgb_out[(size_t)idx * 3] = r;
  rgb_out[(size_t)idx * 3 + 1] = g;
  rgb_out[(size_t)idx * 3 + 2] = b;
}

// ---------------------------------------------------------------------------------------------
// Tiled backward (16x16 pixel tiles, 256 threads): no per-pixel global atomics.
//   apply_bwd_tiled_kernel        full-res tile: chain backward per pixel; for a low-res level the 12
//                                 cotangents of the tile are staged in shared memory and every low-res
//                                 pixel of the tile's footprint GATHERS its transposed-bilinear sum
//                                 (one global reduction per (low-res pixel, channel) per tile instead of
//                                 48 per pixel); full-res levels use warp_level_accumulate12.
//   lowres_slice_bwd_tiled_kernel low-res tile: grid-node gradients through warp_level_accumulate12,
//                                 guidance gradient through the transposed down-sample (12 reductions
//                                 per low-res pixel).
// ---------------------------------------------------------------------------------------------
constexpr int kMaxFoot = 18;  // low-res rows/cols a 16-pixel tile edge can touch (factor >= 2 -> <= 10)
// shared memory of a tile: the cotangent stage + row sums of the transposed up-sample, or the eight per-warp panels
constexpr size_t kBwdSmemBil = (256 * 12 + 16 * kMaxFoot * 12) * sizeof(float) > 8 * sizeof(WarpPanel12)
                                   ? (256 * 12 + 16 * kMaxFoot * 12) * sizeof(float) : 8 * sizeof(WarpPanel12);
static_assert((256 * 12 + 16 * kMaxFoot * 12) * sizeof(float) <= kBwdSmemBil, "cotangent stage + row sums must fit the tile's shared memory");
static_assert(8 * sizeof(WarpPanel12) <= kBwdSmemBil, "per-warp panels must fit the tile's shared memory");

__global__ void __launch_bounds__(256) apply_bwd_tiled_kernel(const float* __restrict__ rgb_in,
                                                              const float* __restrict__ v_rgb_out,
                                                              float* __restrict__ v_rgb_in, BilChain ch) {
  __shared__ __align__(16) float smem[kBwdSmemBil / sizeof(float)];
  __shared__ float wy_tab[kMaxFoot][16], wx_tab[kMaxFoot][16];
  __shared__ int foot[4];  // qy0, nly, qx0, nlx
  const int tiles_x = (ch.W + 15) / 16;
  const int tile_x0 = (blockIdx.x % tiles_x) * 16, tile_y0 = (blockIdx.x / tiles_x) * 16;
  const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
  const int j = tile_x0 + lx, i = tile_y0 + ly;
  const bool inside = j < ch.W && i < ch.H;
  const int jc = min(j, ch.W - 1), ic = min(i, ch.H - 1);
  const size_t idx = (size_t)ic * ch.W + jc;
  float r0 = 0.f, g0 = 0.f, b0 = 0.f, gr = 0.f, gg = 0.f, gb = 0.f;
  if (inside) {
    r0 = rgb_in[idx * 3]; g0 = rgb_in[idx * 3 + 1]; b0 = rgb_in[idx * 3 + 2];
    gr = v_rgb_out[idx * 3]; gg = v_rgb_out[idx * 3 + 1]; gb = v_rgb_out[idx * 3 + 2];
  }
  const float lum = luma_of(r0, g0, b0);
  float xs[BDS_MAX_LEVELS][3];
  {
    float r = r0, g = g0, b = b0;
#pragma unroll
    for (int l = 0; l < BDS_MAX_LEVELS; ++l) {
      if (l < ch.n_levels) {
        xs[l][0] = r; xs[l][1] = g; xs[l][2] = b;
        float A[12];
        level_affine_fwd(ch.lv[l], ch.H, ch.W, ic, jc, lum, A);
        affine_apply(A, r, g, b);
      }
    }
  }
  float v_lum = 0.f;
#pragma unroll
  for (int l = BDS_MAX_LEVELS - 1; l >= 0; --l) {
    if (l < ch.n_levels) {
      const BilLevel& lv = ch.lv[l];
      float A[12], vA[12];
      level_affine_fwd(lv, ch.H, ch.W, ic, jc, lum, A);
      if (lv.v_affine && inside) {
        load12(lv.v_affine + idx * 12, vA);
      } else {
#pragma unroll
        for (int k = 0; k < 12; ++k) vA[k] = 0.f;
      }
      float nr, ng, nb;
      affine_apply_bwd(A, xs[l][0], xs[l][1], xs[l][2], gr, gg, gb, vA, nr, ng, nb);
      gr = nr; gg = ng; gb = nb;
      if (lv.factor > 1) {
        // ---- transposed bilinear up-sample as a gather over the tile's low-res footprint
        __syncthreads();    // a full-resolution level before this one may still be reading its per-warp panels
        float* sva = smem;  // [256][12]
        {
          float4* sp = reinterpret_cast<float4*>(sva + threadIdx.x * 12);
          const float m = inside ? 1.f : 0.f;
          sp[0] = make_float4(vA[0] * m, vA[1] * m, vA[2] * m, vA[3] * m);
          sp[1] = make_float4(vA[4] * m, vA[5] * m, vA[6] * m, vA[7] * m);
          sp[2] = make_float4(vA[8] * m, vA[9] * m, vA[10] * m, vA[11] * m);
        }
        if (threadIdx.x == 0) {
          LinTap a = lin_src(tile_y0, lv.Hd, ch.H), b = lin_src(min(tile_y0 + 15, ch.H - 1), lv.Hd, ch.H);
          LinTap c = lin_src(tile_x0, lv.Wd, ch.W), d = lin_src(min(tile_x0 + 15, ch.W - 1), lv.Wd, ch.W);
          foot[0] = a.i0; foot[1] = min(b.i1 - a.i0 + 1, kMaxFoot);
          foot[2] = c.i0; foot[3] = min(d.i1 - c.i0 + 1, kMaxFoot);
        }
        __syncthreads();
        const int qy0 = foot[0], nly = foot[1], qx0 = foot[2], nlx = foot[3];
        // weight tables: contribution of tile row/col t onto low-res row/col q
        for (int e = threadIdx.x; e < kMaxFoot * 16; e += 256) {
          const int q = e >> 4, t = e & 15;
          float wy = 0.f, wx = 0.f;
          if (q < nly && tile_y0 + t < ch.H) {
            LinTap tp = lin_src(tile_y0 + t, lv.Hd, ch.H);
            if (tp.i0 == qy0 + q) wy += 1.f - tp.t;
            if (tp.i1 == qy0 + q) wy += tp.t;
          }
          if (q < nlx && tile_x0 + t < ch.W) {
            LinTap tp = lin_src(tile_x0 + t, lv.Wd, ch.W);
            if (tp.i0 == qx0 + q) wx += 1.f - tp.t;
            if (tp.i1 == qx0 + q) wx += tp.t;
          }
          wy_tab[q][t] = wy;
          wx_tab[q][t] = wx;
        }
        __syncthreads();
        // separable gather: along x first (tile row ty, low-res column qx), then along y - 16 x fewer FMAs than
        // the direct double sum; one thread per (row, column, channel quad)
        float4* rows = reinterpret_cast<float4*>(smem + 256 * 12);   // [16][nlx][3]
        const float4* sva4 = reinterpret_cast<const float4*>(sva);
        for (int item = threadIdx.x; item < 16 * nlx * 3; item += 256) {
          const int q4 = item % 3, r = item / 3;
          const int qx = r % nlx, ty = r / nlx;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int tx = 0; tx < 16; ++tx) {
            const float w = wx_tab[qx][tx];
            const float4 v = sva4[(ty * 16 + tx) * 3 + q4];
            acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
          }
          rows[(ty * nlx + qx) * 3 + q4] = acc;
        }
        __syncthreads();
        for (int item = threadIdx.x; item < nly * nlx * 3; item += 256) {
          const int q4 = item % 3, q = item / 3;
          const int qx = q % nlx, qy = q / nlx;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int ty = 0; ty < 16; ++ty) {
            const float w = wy_tab[qy][ty];
            const float4 v = rows[(ty * nlx + qx) * 3 + q4];
            acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
          }
          float* dst = lv.v_a_low + ((size_t)(qy0 + qy) * lv.Wd + qx0 + qx) * 12 + q4 * 4;
          if (acc.x != 0.f) red_add(dst, acc.x);
          if (acc.y != 0.f) red_add(dst + 1, acc.y);
          if (acc.z != 0.f) red_add(dst + 2, acc.z);
          if (acc.w != 0.f) red_add(dst + 3, acc.w);
        }
        __syncthreads();
        // a tile wider than kMaxFoot low-res pixels cannot happen for factor >= 2 (<= 10)
      } else {
        Tri t = tri_setup(lattice_coord(jc, ch.W, lv.GX), lattice_coord(ic, ch.H, lv.GY), luma_coord(lum, lv.L), lv.L,
                          lv.GY, lv.GX);
        if (t.z_inside && inside) {
          float Ad[12], dAdz[12];
          tri_fetch<true>(lv.grid_cl, t, Ad, dAdz);
          float s = 0.f;
#pragma unroll
          for (int k = 0; k < 12; ++k) s = fmaf(vA[k], dAdz[k], s);
          v_lum += s * (float)(lv.L - 1);
        }
        // per-warp reduction straight to global memory (no block barrier, bilateral_accum.cuh)
        warp_level_accumulate12(reinterpret_cast<WarpPanel12*>(smem) + (threadIdx.x >> 5), t, inside, vA, lv.L, lv.GY, lv.GX,
                                lv.v_grid_cl);
      }
    }
  }
  if (inside) {
    v_rgb_in[idx * 3] = gr + v_lum * kLumaR;
    v_rgb_in[idx * 3 + 1] = gg + v_lum * kLumaG;
    v_rgb_in[idx * 3 + 2] = gb + v_lum * kLumaB;
  }
}

__global__ void __launch_bounds__(256) lowres_slice_bwd_tiled_kernel(const float* __restrict__ rgb, int H, int W,
                                                                     BilLevel lv, float* __restrict__ v_rgb_in) {
  __shared__ __align__(16) float smem[kBwdSmemBil / sizeof(float)];
  const int tiles_x = (lv.Wd + 15) / 16;
  const int tile_x0 = (blockIdx.x % tiles_x) * 16, tile_y0 = (blockIdx.x / tiles_x) * 16;
  const int w = tile_x0 + (threadIdx.x & 15), h = tile_y0 + (threadIdx.x >> 4);
  const bool inside = w < lv.Wd && h < lv.Hd;
  const int wc = min(w, lv.Wd - 1), hc = min(h, lv.Hd - 1);
  LinTap ty = lin_src(hc, H, lv.Hd), tx = lin_src(wc, W, lv.Wd);
  float r, g, b;
  downsample_rgb(rgb, H, W, ty, tx, r, g, b);
  Tri t = tri_setup(lattice_coord(wc, lv.Wd, lv.GX), lattice_coord(hc, lv.Hd, lv.GY),
                    luma_coord(luma_of(r, g, b), lv.L), lv.L, lv.GY, lv.GX);
  float vA[12];
  if (inside) {
    load12(lv.v_a_low + ((size_t)hc * lv.Wd + wc) * 12, vA);
  } else {
#pragma unroll
    for (int k = 0; k < 12; ++k) vA[k] = 0.f;
  }
  if (inside && t.z_inside) {
    float Ad[12], dAdz[12];
    tri_fetch<true>(lv.grid_cl, t, Ad, dAdz);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 12; ++k) s = fmaf(vA[k], dAdz[k], s);
    const float v_lum = s * (float)(lv.L - 1);
    const float w0 = 1.f - tx.t, w1 = tx.t, h0 = 1.f - ty.t, h1 = ty.t;
    const float lw[3] = {kLumaR, kLumaG, kLumaB};
    float* p00 = v_rgb_in + ((size_t)ty.i0 * W + tx.i0) * 3;
    float* p01 = v_rgb_in + ((size_t)ty.i0 * W + tx.i1) * 3;
    float* p10 = v_rgb_in + ((size_t)ty.i1 * W + tx.i0) * 3;
    float* p11 = v_rgb_in + ((size_t)ty.i1 * W + tx.i1) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = v_lum * lw[c];
      red_add(p00 + c, h0 * w0 * v);
      red_add(p01 + c, h0 * w1 * v);
      red_add(p10 + c, h1 * w0 * v);
      red_add(p11 + c, h1 * w1 * v);
    }
  }
  warp_level_accumulate12(reinterpret_cast<WarpPanel12*>(smem) + (threadIdx.x >> 5), t, inside, vA, lv.L, lv.GY, lv.GX,
                          lv.v_grid_cl);
}

// generic per-sample slice on the channel-first parameter layout (BilateralGrid.forward) ----------
__global__ void slice_generic_fwd_kernel(const float* __restrict__ grid, int L, int GY, int GX, int n,
                                         const float* __restrict__ xy, const float* __restrict__ rgb,
                                         float* __restrict__ affine) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  float fx = unit_coord(xy[2 * (size_t)idx], GX), fy = unit_coord(xy[2 * (size_t)idx + 1], GY);
  float fz = luma_coord(luma_of(rgb[3 * (size_t)idx], rgb[3 * (size_t)idx + 1], rgb[3 * (size_t)idx + 2]), L);
  Tri t = tri_setup(fx, fy, fz, L, GY, GX);
  int nodes = L * GY * GX;
  float wx0 = 1.f - t.wx1, wy0 = 1.f - t.wy1, wz0 = 1.f - t.wz1;
  // offsets in the PARAMETER layout [L][GY][GX] (this entry reads the grid in place)
  const int x1 = t.n01 != t.n00 ? t.x0 + 1 : t.x0, y1 = t.n10 != t.n00 ? t.y0 + 1 : t.y0;
  const int q00 = (t.z0 * GY + t.y0) * GX + t.x0, q01 = (t.z0 * GY + t.y0) * GX + x1;
  const int q10 = (t.z0 * GY + y1) * GX + t.x0, q11 = (t.z0 * GY + y1) * GX + x1, qz = t.dz * GY * GX;
  for (int k = 0; k < 12; ++k) {
    const float* g = grid + (size_t)k * nodes;
    float c0 = wy0 * (wx0 * __ldg(g + q00) + t.wx1 * __ldg(g + q01)) +
               t.wy1 * (wx0 * __ldg(g + q10) + t.wx1 * __ldg(g + q11));
    float c1 = wy0 * (wx0 * __ldg(g + q00 + qz) + t.wx1 * __ldg(g + q01 + qz)) +
               t.wy1 * (wx0 * __ldg(g + q10 + qz) + t.wx1 * __ldg(g + q11 + qz));
    affine[(size_t)idx * 12 + k] = wz0 * c0 + t.wz1 * c1;
  }
}

__global__ void slice_generic_bwd_kernel(const float* __restrict__ grid, int L, int GY, int GX, int n,
                                         const float* __restrict__ xy, const float* __restrict__ rgb,
                                         const float* __restrict__ v_affine, float* __restrict__ v_grid,
                                         float* __restrict__ v_rgb) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  float fx = unit_coord(xy[2 * (size_t)idx], GX), fy = unit_coord(xy[2 * (size_t)idx + 1], GY);
  float fz = luma_coord(luma_of(rgb[3 * (size_t)idx], rgb[3 * (size_t)idx + 1], rgb[3 * (size_t)idx + 2]), L);
  Tri t = tri_setup(fx, fy, fz, L, GY, GX);
  int nodes = L * GY * GX;
  float wx0 = 1.f - t.wx1, wy0 = 1.f - t.wy1, wz0 = 1.f - t.wz1;
  const int x1 = t.n01 != t.n00 ? t.x0 + 1 : t.x0, y1 = t.n10 != t.n00 ? t.y0 + 1 : t.y0;
  const int q00 = (t.z0 * GY + t.y0) * GX + t.x0, q01 = (t.z0 * GY + t.y0) * GX + x1;
  const int q10 = (t.z0 * GY + y1) * GX + t.x0, q11 = (t.z0 * GY + y1) * GX + x1, qz = t.dz * GY * GX;
  float s = 0.f;
  for (int k = 0; k < 12; ++k) {
    float va = v_affine[(size_t)idx * 12 + k];
    const float* g = grid + (size_t)k * nodes;
    float* vg = v_grid + (size_t)k * nodes;
    float c0 = wy0 * (wx0 * __ldg(g + q00) + t.wx1 * __ldg(g + q01)) +
               t.wy1 * (wx0 * __ldg(g + q10) + t.wx1 * __ldg(g + q11));
    float c1 = wy0 * (wx0 * __ldg(g + q00 + qz) + t.wx1 * __ldg(g + q01 + qz)) +
               t.wy1 * (wx0 * __ldg(g + q10 + qz) + t.wx1 * __ldg(g + q11 + qz));
    s = fmaf(va, c1 - c0, s);
    red_add(vg + q00, va * wz0 * wy0 * wx0);
    red_add(vg + q01, va * wz0 * wy0 * t.wx1);
    red_add(vg + q10, va * wz0 * t.wy1 * wx0);
    red_add(vg + q11, va * wz0 * t.wy1 * t.wx1);
    red_add(vg + q00 + qz, va * t.wz1 * wy0 * wx0);
    red_add(vg + q01 + qz, va * t.wz1 * wy0 * t.wx1);
    red_add(vg + q10 + qz, va * t.wz1 * t.wy1 * wx0);
    red_add(vg + q11 + qz, va * t.wz1 * t.wy1 * t.wx1);
  }
  float v_lum = t.z_inside ? s * (float)(L - 1) : 0.f;
  v_rgb[(size_t)idx * 3] = v_lum * kLumaR;
  v_rgb[(size_t)idx * 3 + 1] = v_lum * kLumaG;
  v_rgb[(size_t)idx * 3 + 2] = v_lum * kLumaB;
}

// TV loss (lib_bilagrid.py:152-168): sum over the 3 spatial axes of mean squared forward
// differences, divided by the batch size N.  One thread per element; block reduction; one atomic.
// All levels of a multi-scale module (modules.py:466-472) go in ONE launch: blockIdx.y = level.
struct TvLevels {
  const float* g[BDS_MAX_LEVELS];
  float* v_g[BDS_MAX_LEVELS];
  int N[BDS_MAX_LEVELS], L[BDS_MAX_LEVELS], GY[BDS_MAX_LEVELS], GX[BDS_MAX_LEVELS];
  float weight[BDS_MAX_LEVELS];
};
__global__ void __launch_bounds__(256) tv_kernel(TvLevels lv, float v_loss, float* __restrict__ loss) {
  const int l = blockIdx.y;
  const float* __restrict__ g = lv.g[l];
  float* __restrict__ v_g = lv.v_g[l];
  const int N = lv.N[l], L = lv.L[l], GY = lv.GY[l], GX = lv.GX[l];
  const float weight = lv.weight[l];
  size_t total = (size_t)N * 12 * L * GY * GX;
  float acc = 0.f;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    int x = idx % GX;
    size_t r = idx / GX;
    int y = r % GY;
    r /= GY;
    int z = r % L;
    float v = g[idx];
    // counts = elements of one batch item of the differenced tensor (x1.size()[1:])
    float cz = fmaxf(12.f * (L - 1) * GY * GX, 1.f), cy = fmaxf(12.f * L * (GY - 1) * GX, 1.f),
          cx = fmaxf(12.f * L * GY * (GX - 1), 1.f);
    float sN = weight / (float)N;
    float grad = 0.f, a = 0.f;
    if (x + 1 < GX) { float d = g[idx + 1] - v; a += d * d / cx; grad -= 2.f * d / cx; }
    if (x > 0) { float d = v - g[idx - 1]; grad += 2.f * d / cx; }
    if (y + 1 < GY) { float d = g[idx + GX] - v; a += d * d / cy; grad -= 2.f * d / cy; }
    if (y > 0) { float d = v - g[idx - GX]; grad += 2.f * d / cy; }
    size_t sz = (size_t)GY * GX;
    if (z + 1 < L) { float d = g[idx + sz] - v; a += d * d / cz; grad -= 2.f * d / cz; }
    if (z > 0) { float d = v - g[idx - sz]; grad += 2.f * d / cz; }
    acc += a * sN;
    if (v_g) v_g[idx] += v_loss * sN * grad;
  }
  __shared__ float red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = red[threadIdx.x];
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffu, v, o);
    if (threadIdx.x == 0 && v != 0.f) red_add(loss, v);
  }
}

// workspace carving ---------------------------------------------------------------------------------
struct BilWorkspace {
  size_t grid_cl[BDS_MAX_LEVELS], v_grid_cl[BDS_MAX_LEVELS], a_low[BDS_MAX_LEVELS], v_a_low[BDS_MAX_LEVELS];
  size_t total;
};
static BilWorkspace carve(const bds_bilateral_desc* d, int H, int W) {
  BilWorkspace w;
  size_t off = 0;
  for (int l = 0; l < d->n_levels; ++l) {
    size_t gbytes = align_up((size_t)d->L[l] * d->GY[l] * d->GX[l] * 12 * sizeof(float), 256);
    w.grid_cl[l] = off; off += gbytes;
    w.v_grid_cl[l] = off; off += gbytes;
    size_t abytes = 0;
    if (d->factor[l] > 1) abytes = align_up((size_t)(H / d->factor[l]) * (W / d->factor[l]) * 12 * sizeof(float), 256);
    w.a_low[l] = off; off += abytes;
    w.v_a_low[l] = off; off += abytes;
  }
  w.total = off;
  return w;
}

static int check_desc(const bds_bilateral_desc* d, int H, int W) {
  BDS_REQUIRE(d && d->n_levels >= 1 && d->n_levels <= BDS_MAX_LEVELS, "bilateral: n_levels out of range");
  BDS_REQUIRE(H > 0 && W > 0, "bilateral: empty image");
  for (int l = 0; l < d->n_levels; ++l) {
    BDS_REQUIRE(d->L[l] >= 1 && d->GY[l] >= 1 && d->GX[l] >= 1, "bilateral: bad grid size at level %d", l);
    BDS_REQUIRE(d->factor[l] >= 0, "bilateral: negative guidance factor");
    if (d->factor[l] > 1)
      BDS_REQUIRE(H / d->factor[l] >= 1 && W / d->factor[l] >= 1, "bilateral: image smaller than guidance factor");
  }
  return 0;
}

static void fill_chain(BilChain& ch, const bds_bilateral_desc* d, int H, int W, char* ws, const BilWorkspace& w) {
  ch.n_levels = d->n_levels;
  ch.H = H;
  ch.W = W;
  for (int l = 0; l < d->n_levels; ++l) {
    BilLevel& lv = ch.lv[l];
    lv.grid_cl = reinterpret_cast<const float*>(ws + w.grid_cl[l]);
    lv.v_grid_cl = reinterpret_cast<float*>(ws + w.v_grid_cl[l]);
    lv.a_low = reinterpret_cast<const float*>(ws + w.a_low[l]);
    lv.v_a_low = reinterpret_cast<float*>(ws + w.v_a_low[l]);
    lv.affine_out = nullptr;
    lv.v_affine = nullptr;
    lv.L = d->L[l]; lv.GY = d->GY[l]; lv.GX = d->GX[l];
    lv.factor = d->factor[l] > 1 ? d->factor[l] : 1;
    lv.Hd = H / lv.factor; lv.Wd = W / lv.factor;
  }
}

}  // namespace bds

using namespace bds;

extern "C" size_t bds_bilateral_workspace_bytes(const bds_bilateral_desc* d, int H, int W) {
  if (!d || d->n_levels < 1 || d->n_levels > BDS_MAX_LEVELS) return 0;
  return carve(d, H, W).total + 256;
}

extern "C" int bds_bilateral_fwd(const bds_bilateral_desc* d, int H, int W, const float* rgb_in,
                                 const float* const* host_grids, float* rgb_out,
                                 float* const* host_affine_out, void* workspace, bds_stream_t stream_) {
  if (int rc = check_desc(d, H, W)) return rc;
  BDS_REQUIRE(rgb_in && rgb_out && host_grids && workspace, "bilateral_fwd: null pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BilWorkspace w = carve(d, H, W);
  BilChain ch;
  fill_chain(ch, d, H, W, static_cast<char*>(workspace), w);
  for (int l = 0; l < d->n_levels; ++l) {
    BDS_REQUIRE(host_grids[l], "bilateral_fwd: null grid at level %d", l);
    int nodes = d->L[l] * d->GY[l] * d->GX[l];
    repack_grid_kernel<<<ceil_div(nodes * 12, 256), 256, 0, stream>>>(host_grids[l], const_cast<float*>(ch.lv[l].grid_cl), d->L[l], d->GY[l], d->GX[l]);
    BDS_CHECK_LAUNCH();
    if (host_affine_out) ch.lv[l].affine_out = host_affine_out[l];
  }
  for (int l = 0; l < d->n_levels; ++l) {
    if (ch.lv[l].factor > 1) {
      lowres_slice_fwd_kernel<<<ceil_div((int64_t)ch.lv[l].Hd * ch.lv[l].Wd, 256), 256, 0, stream>>>(
          rgb_in, H, W, ch.lv[l], const_cast<float*>(ch.lv[l].a_low));
      BDS_CHECK_LAUNCH();
    }
  }
  apply_fwd_kernel<<<ceil_div((int64_t)H * W, 256), 256, 0, stream>>>(rgb_in, rgb_out, ch);
  BDS_CHECK_LAUNCH();
  return 0;
}

extern "C" int bds_bilateral_bwd(const bds_bilateral_desc* d, int H, int W, const float* rgb_in,
                                 const float* const* host_grids, const float* v_rgb_out,
                                 const float* const* host_v_affine, float* v_rgb_in,
                                 float* const* host_v_grids, void* workspace, bds_stream_t stream_) {
  if (int rc = check_desc(d, H, W)) return rc;
  BDS_REQUIRE(rgb_in && v_rgb_out && v_rgb_in && host_grids && host_v_grids && workspace,
              "bilateral_bwd: null pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BilWorkspace w = carve(d, H, W);
  BilChain ch;
  fill_chain(ch, d, H, W, static_cast<char*>(workspace), w);
  // the forward left grid_cl and a_low in the workspace; rebuild them anyway so that backward is
  // self-contained (the workspace may have been reused) - both are tiny next to the image passes
  for (int l = 0; l < d->n_levels; ++l) {
    BDS_REQUIRE(host_grids[l] && host_v_grids[l], "bilateral_bwd: null grid at level %d", l);
    int nodes = d->L[l] * d->GY[l] * d->GX[l];
    repack_grid_kernel<<<ceil_div(nodes * 12, 256), 256, 0, stream>>>(host_grids[l], const_cast<float*>(ch.lv[l].grid_cl), d->L[l], d->GY[l], d->GX[l]);
    BDS_CHECK_LAUNCH();
    BDS_CHECK_CUDA(cudaMemsetAsync(ch.lv[l].v_grid_cl, 0, (size_t)nodes * 12 * sizeof(float), stream));
    if (ch.lv[l].factor > 1) {
      size_t n_low = (size_t)ch.lv[l].Hd * ch.lv[l].Wd;
      lowres_slice_fwd_kernel<<<ceil_div((int64_t)n_low, 256), 256, 0, stream>>>(rgb_in, H, W, ch.lv[l],
                                                                                const_cast<float*>(ch.lv[l].a_low));
      BDS_CHECK_LAUNCH();
      BDS_CHECK_CUDA(cudaMemsetAsync(ch.lv[l].v_a_low, 0, n_low * 12 * sizeof(float), stream));
    }
    if (host_v_affine) ch.lv[l].v_affine = host_v_affine[l];
  }
  apply_bwd_tiled_kernel<<<((W + 15) / 16) * ((H + 15) / 16), 256, 0, stream>>>(rgb_in, v_rgb_out, v_rgb_in, ch);
  BDS_CHECK_LAUNCH();
  for (int l = 0; l < d->n_levels; ++l) {
    if (ch.lv[l].factor > 1) {
      lowres_slice_bwd_tiled_kernel<<<((ch.lv[l].Wd + 15) / 16) * ((ch.lv[l].Hd + 15) / 16), 256, 0, stream>>>(
          rgb_in, H, W, ch.lv[l], v_rgb_in);
      BDS_CHECK_LAUNCH();
    }
    int nodes = d->L[l] * d->GY[l] * d->GX[l];
    unpack_add_grid_kernel<<<ceil_div(nodes * 12, 256), 256, 0, stream>>>(ch.lv[l].v_grid_cl, host_v_grids[l], d->L[l], d->GY[l], d->GX[l]);
    BDS_CHECK_LAUNCH();
  }
  return 0;
}

extern "C" int bds_bilagrid_slice_fwd(const float* grid, int L, int GY, int GX, int n, const float* xy,
                                      const float* rgb, float* affine, bds_stream_t stream) {
  BDS_REQUIRE(L >= 1 && GY >= 1 && GX >= 1 && n >= 0, "slice_fwd: bad sizes");
  if (n == 0) return 0;
  BDS_REQUIRE(grid && xy && rgb && affine, "slice_fwd: null pointer");
  slice_generic_fwd_kernel<<<ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(grid, L, GY, GX, n, xy, rgb, affine);
  BDS_CHECK_LAUNCH();
  return 0;
}

extern "C" int bds_bilagrid_slice_bwd(const float* grid, int L, int GY, int GX, int n, const float* xy,
                                      const float* rgb, const float* v_affine, float* v_grid, float* v_rgb,
                                      bds_stream_t stream) {
  BDS_REQUIRE(L >= 1 && GY >= 1 && GX >= 1 && n >= 0, "slice_bwd: bad sizes");
  if (n == 0) return 0;
  BDS_REQUIRE(grid && xy && rgb && v_affine && v_grid && v_rgb, "slice_bwd: null pointer");
  slice_generic_bwd_kernel<<<ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(grid, L, GY, GX, n, xy, rgb,
                                                                                          v_affine, v_grid, v_rgb);
  BDS_CHECK_LAUNCH();
  return 0;
}

extern "C" int bds_tv_fwd_bwd(const float* grids, int N, int L, int GY, int GX, float weight, float v_loss,
                              float* loss, float* v_grids, bds_stream_t stream) {
  BDS_REQUIRE(N >= 1 && L >= 1 && GY >= 1 && GX >= 1, "tv: bad sizes");
  BDS_REQUIRE(grids && loss, "tv: null pointer");
  size_t total = (size_t)N * 12 * L * GY * GX;
  TvLevels lv{};
  lv.g[0] = grids; lv.v_g[0] = v_grids; lv.N[0] = N; lv.L[0] = L; lv.GY[0] = GY; lv.GX[0] = GX; lv.weight[0] = weight;
  tv_kernel<<<dim3(ceil_div((int64_t)total, 256), 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(lv, v_loss, loss);
  BDS_CHECK_LAUNCH();
  return 0;
}

extern "C" int bds_tv_levels_fwd_bwd(int n_levels, const float* const* host_grids, const int* N, const int* L,
                                     const int* GY, const int* GX, const float* weights, float v_loss, float* loss,
                                     float* const* host_v_grids, bds_stream_t stream) {
  BDS_REQUIRE(n_levels >= 1 && n_levels <= BDS_MAX_LEVELS, "tv_levels: 1..%d levels", BDS_MAX_LEVELS);
  BDS_REQUIRE(host_grids && N && L && GY && GX && weights && loss, "tv_levels: null pointer");
  TvLevels lv{};
  size_t most = 0;
  for (int l = 0; l < n_levels; ++l) {
    BDS_REQUIRE(host_grids[l] && N[l] >= 1 && L[l] >= 1 && GY[l] >= 1 && GX[l] >= 1, "tv_levels: bad level %d", l);
    lv.g[l] = host_grids[l]; lv.v_g[l] = host_v_grids ? host_v_grids[l] : nullptr;
    lv.N[l] = N[l]; lv.L[l] = L[l]; lv.GY[l] = GY[l]; lv.GX[l] = GX[l]; lv.weight[l] = weights[l];
    size_t total = (size_t)N[l] * 12 * L[l] * GY[l] * GX[l];
    most = total > most ? total : most;
  }
  // blockIdx.y = level; the x extent covers the largest level (grid-stride loop, smaller levels leave early)
  int bx = ceil_div((int64_t)most, 256);
  if (bx > 148 * 16) bx = 148 * 16;
  tv_kernel<<<dim3(bx, n_levels), 256, 0, static_cast<cudaStream_t>(stream)>>>(lv, v_loss, loss);
  BDS_CHECK_LAUNCH();
  return 0;
}
