// Fused front-to-back alpha compositing + reference glue + multi-scale bilateral chain, forward and
// backward.  One CTA per 16x16 tile of the band: 128 threads = 4 warps x 8x8 pixels (two pixels per lane) in the
// forward, 256 threads = 8 warps x 8x4 pixels in the backward.
//
// Replaces, in one launch each way (paths relative to /root/reference/project):
//   gsplat rasterize_to_pixels_fwd/bwd          called via models/trainers/base.py:393-408
//   clamp(rgb, max=1), RGB+ED depth normalise   base.py:414-417 (+ gsplat rendering ED branch)
//   rgb_gaussians + rgb_sky * (1 - opacity)     models/trainers/scene_graph.py:287-294
//   MultiScaleBilateralAffineTransform (guidance_factor=None branch, modules.py:548-559) and the
//   sequential 3x4 apply                         scene_graph.py:112-117
//
// B200 mapping:
//   * the tile's depth-sorted 48-byte splat records are contiguous in HBM (binning.cu) and are staged
//     into shared memory with 1-D TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx), 3 stages
//     of 128 records, issued by one elected thread;
//   * each warp tests 32 records at a time (one per lane) against its own pixel rectangle with an
//     exact ellipse-vs-rectangle bound, ballots, and only walks the survivors - the blend itself reads
//     records from shared memory as 128-bit broadcasts; transmittance lives in a register per pixel and
//     a warp vote retires the warp when all its pixels are saturated; stages are handed back with
//     consumer-release mbarriers (no block barrier in either walk);
//   * the backward re-walks the same records back to front (2 stages of 128 records) with one running scalar
//     per pixel; the per-record reduction across the warp is DEFERRED: the walk stores two scalars per (pixel,
//     record) into a per-warp shared panel and every 16 records the warp transposes the work (lane = record)
//     and sums the panel into pixel moments in registers - no shuffle tree - that leave as three 128-bit vector
//     reductions into a 48-byte-per-splat gradient record (see flush_batch); the walk's bookkeeping is in
//     registers (lane i remembers where the record of panel row i lies in shared memory);
//   * FMA-dense parts (trilinear slice, colour sums) use packed fp32x2 FMAs (FFMA2);
//   * no tensor cores: the work is gather / pointwise / scatter.
#include <stdlib.h>

#include "bilateral_accum.cuh"
#include "tma.cuh"
#include "projection_math.cuh"

namespace bds {

#ifndef BDS_BWD_CHUNK
#define BDS_BWD_CHUNK 128
#endif
constexpr int kChunk = BDS_BWD_CHUNK;       // backward: records per TMA stage
#ifndef BDS_FWD_CHUNK
#define BDS_FWD_CHUNK 128
#endif
#ifndef BDS_FWD_STAGES
#define BDS_FWD_STAGES 3
#endif
constexpr int kFChunk = BDS_FWD_CHUNK;      // forward: records per TMA stage
constexpr int kFStages = BDS_FWD_STAGES;    // forward: stages in flight
constexpr int kRecBytes = 48;

// ---- parameters ----------------------------------------------------------------------------------
struct FusedBil {       // per-camera bilateral chain, full-resolution guidance
  int n_levels;
  const float* grid_cl[BDS_MAX_LEVELS];  // base of [C][L][GY][GX][12]
  float* v_grid_cl[BDS_MAX_LEVELS];
  int L[BDS_MAX_LEVELS], GY[BDS_MAX_LEVELS], GX[BDS_MAX_LEVELS];
};

struct CompParams {
  const float4* recs;
  const int32_t* tile_offsets;
  int W, H, tile_w, tile_h, row_begin;
  int64_t pix_row0;     // first stacked pixel row (c*H + y) of the band
  int channels, expected_depth;
  const float* backgrounds;  // [C, channels] or null
  const float* sky;          // band pixels x 3 or null
  float *out_rgb, *out_rgbg, *out_depth, *out_alpha;
  int32_t* last_ids;
  const uint8_t* slot_keep;  // optional per-splat keep flags (masked re-render over the same sorted lists)
  FusedBil bil;
  // backward only
  const float *v_rgb, *v_rgbg, *v_depth, *v_alpha;
  float *v_splats, *v_sky, *v_backgrounds;
};

struct TileGeom {
  int cam, px, py;      // pixel of this thread
  bool inside;
  int64_t pix;          // index into band-pixel arrays
  float wx0, wy0;       // warp rectangle origin (pixel index)
  int start, end;       // record range of the tile
};

BDS_D TileGeom tile_geom(const CompParams& p) {
  TileGeom g;
  int t = blockIdx.x;
  int grow = p.row_begin + t / p.tile_w;
  int tx = t - (t / p.tile_w) * p.tile_w;
  g.cam = grow / p.tile_h;
  int ty = grow - g.cam * p.tile_h;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int sx = (warp & 1) * 8, sy = (warp >> 1) * 4;
  g.px = tx * kTile + sx + (lane & 7);
  g.py = ty * kTile + sy + (lane >> 3);
  g.wx0 = (float)(tx * kTile + sx);
  g.wy0 = (float)(ty * kTile + sy);
  g.inside = g.px < p.W && g.py < p.H;
  g.pix = ((int64_t)g.cam * p.H + g.py - p.pix_row0) * p.W + g.px;
  g.start = p.tile_offsets[t];
  g.end = p.tile_offsets[t + 1];
  return g;
}

// full-resolution-guidance chain on one pixel; fills A (all levels) and the level inputs xs
BDS_D void fused_chain_fwd(const FusedBil& b, int cam, int H, int W, int i, int j, float& r, float& g, float& bl,
                           float lum) {
  // the lattice coordinate of (pixel, level) is unit_coord(lin01(pixel), grid size): the IEEE division inside
  // lin01 is paid once per axis instead of once per level
  const float x01 = lin01(j, W), y01 = lin01(i, H);
  for (int l = 0; l < b.n_levels; ++l) {
    const float* grid = b.grid_cl[l] + (size_t)cam * b.L[l] * b.GY[l] * b.GX[l] * 12;
    Tri t = tri_setup(unit_coord(x01, b.GX[l]), unit_coord(y01, b.GY[l]), luma_coord(lum, b.L[l]), b.L[l],
                      b.GY[l], b.GX[l]);
    float A[12];
    tri_fetch<false>(grid, t, A, nullptr);
    affine_apply(A, r, g, bl);
  }
}

// =================================================================================================
// forward
// =================================================================================================
// One CTA of 128 threads per 16x16 tile: 4 warps x (8x8 pixels), TWO pixels per lane - (x, y) and (x, y + 4).  What a
// (warp, record) evaluation pays once - the bit iteration over the rectangle test's survivors, three 128-bit
// broadcast loads of the record, the loop control - is shared by 64 pixels instead of 32, and the per-pixel arithmetic
// of the two pixels runs as packed fp32x2 instructions (FFMA2 / FMUL2 / FADD2) with the record operand shared.
constexpr int kFwdThreads = 128;
constexpr int kFwdWarps = kFwdThreads / 32;
#ifndef BDS_FWD_MINB
#define BDS_FWD_MINB 8     // resident CTAs per SM the forward is compiled for (register cap 65536 / (128 * MINB) = 64)
#endif

struct TileGeom2 {
  int cam, px, py0, py1;    // the lane's two pixels: (px, py0) and (px, py1 = py0 + 4)
  bool in0, in1;
  int64_t pix0, pix1;       // indices into band-pixel arrays
  float wx0, wy0;           // origin (pixel index) of the warp's 8x8 rectangle
  int start, end;           // record range of the tile
};

BDS_D TileGeom2 tile_geom2(const CompParams& p) {
  TileGeom2 g;
  const int t = blockIdx.x;
  const int grow = p.row_begin + t / p.tile_w;
  const int tx = t - (t / p.tile_w) * p.tile_w;
  g.cam = grow / p.tile_h;
  const int ty = grow - g.cam * p.tile_h;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sx = (warp & 1) * 8, sy = (warp >> 1) * 8;
  g.px = tx * kTile + sx + (lane & 7);
  g.py0 = ty * kTile + sy + (lane >> 3);
  g.py1 = g.py0 + 4;
  g.wx0 = (float)(tx * kTile + sx);
  g.wy0 = (float)(ty * kTile + sy);
  g.in0 = g.px < p.W && g.py0 < p.H;
  g.in1 = g.px < p.W && g.py1 < p.H;
  g.pix0 = ((int64_t)g.cam * p.H + g.py0 - p.pix_row0) * p.W + g.px;
  g.pix1 = g.pix0 + 4 * (int64_t)p.W;
  g.start = p.tile_offsets[t];
  g.end = p.tile_offsets[t + 1];
  return g;
}

// per-pixel epilogue: gsplat outputs (mode 0) or the reference glue (+ the bilateral chain)
template <int MODE>
BDS_D void fwd_epilogue(const CompParams& p, int cam, int64_t pix, int px, int py, float T, float cr, float cg, float cb,
                        float cd, int last) {
  const float A = 1.f - T;
  p.last_ids[pix] = last;
  p.out_alpha[pix] = A;
  if (MODE == 0) {
    float bgv[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.backgrounds)
      for (int c = 0; c < p.channels; ++c) bgv[c] = p.backgrounds[cam * p.channels + c];
    float o[4] = {cr + T * bgv[0], cg + T * bgv[1], cb + T * bgv[2], cd + T * bgv[3]};
    if (p.channels == 4 && p.expected_depth) o[3] = o[3] / fmaxf(A, 1e-10f);
    for (int c = 0; c < p.channels; ++c) p.out_rgb[pix * p.channels + c] = o[c];
  } else {
    float rg = fminf(cr, 1.f), gg = fminf(cg, 1.f), bg = fminf(cb, 1.f);  // base.py:417
    p.out_rgbg[pix * 3] = rg; p.out_rgbg[pix * 3 + 1] = gg; p.out_rgbg[pix * 3 + 2] = bg;
    p.out_depth[pix] = cd / fmaxf(A, 1e-10f);                             // RGB+ED
    float r = rg, gr = gg, b = bg;
    if (p.sky) {                                                          // scene_graph.py:293
      r = fmaf(p.sky[pix * 3], T, r);
      gr = fmaf(p.sky[pix * 3 + 1], T, gr);
      b = fmaf(p.sky[pix * 3 + 2], T, b);
    }
    if (MODE == 2) fused_chain_fwd(p.bil, cam, p.H, p.W, py, px, r, gr, b, luma_of(r, gr, b));
    p.out_rgb[pix * 3] = r; p.out_rgb[pix * 3 + 1] = gr; p.out_rgb[pix * 3 + 2] = b;
  }
}

template <int MODE>
__global__ void __launch_bounds__(kFwdThreads, BDS_FWD_MINB) composite_fwd_kernel(CompParams p) {
  __shared__ __align__(128) float4 srec[kFStages][kFChunk * 3];
  __shared__ __align__(8) uint64_t bars[kFStages];    // stage filled (TMA complete_tx)
  __shared__ __align__(8) uint64_t freed[kFStages];   // stage consumed by all warps
  __shared__ int s_done_warps;                        // warps whose 64 pixels are all saturated
  __shared__ int s_stop;                              // first chunk that will NOT be loaded (block-wide early exit)
  __shared__ int s_decided;                           // refill decisions are taken in chunk order (no holes)

  const TileGeom2 g = tile_geom2(p);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = g.end - g.start;
  const int nchunks = (n + kFChunk - 1) / kFChunk;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kFStages; ++s) { mbar_init(&bars[s], 1); mbar_init(&freed[s], kFwdWarps); }
    s_done_warps = 0;
    s_stop = nchunks;
    s_decided = 0;
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 0; k < nchunks && k < kFStages; ++k) {
      int cnt = min(kFChunk, n - k * kFChunk);
      mbar_expect_tx(&bars[k], cnt * kRecBytes);
      bulk_g2s(&srec[k][0], p.recs + (size_t)(g.start + k * kFChunk) * 3, cnt * kRecBytes, &bars[k]);
    }
  }

  const float pxf = (float)g.px + 0.5f;
  const f32x2 pyf2 = pk2((float)g.py0 + 0.5f, (float)g.py1 + 0.5f);
  const float rxmin = g.wx0 + 0.5f, rxmax = g.wx0 + 7.5f, rymin = g.wy0 + 0.5f, rymax = g.wy0 + 7.5f;
  const f32x2 one2 = pk2(1.f, 1.f);
  f32x2 T2 = one2;                                         // transmittance of the two pixels
  f32x2 cr2 = pk2(0.f, 0.f), cg2 = cr2, cb2 = cr2, cd2 = cr2;   // (pixel 0, pixel 1) per channel
  int last0 = -1, last1 = -1;
  // alpha >= 1/255  <=>  e >= -log2(255); a saturated (or outside) pixel raises its threshold to +inf, so
  // "done" costs no extra test in the blend loop
  float emin0 = g.in0 ? -kLog2_255 : INFINITY, emin1 = g.in1 ? -kLog2_255 : INFINITY;
  bool warp_done = __all_sync(kFull, !g.in0 && !g.in1);

  // No block barrier in the loop: a warp releases a stage when it is done with it (mbarrier `freed`) and runs ahead by
  // up to kFStages - 1 chunks; lane 0 of warp k % 4 refills the stage with chunk k + kFStages once all warps have
  // released it - unless every warp has reported its pixels saturated, in which case it lowers s_stop instead and the
  // block leaves after the chunks already in flight (the block-wide early exit).
  bool counted_done = false;
  for (int k = 0; k < nchunks; ++k) {
    const int st = k % kFStages;
    {   // wait for chunk k, or learn that it will never be loaded
      bool stop = false;
      for (unsigned it = 0;; ++it) {
        if (mbar_try_wait(&bars[st], (k / kFStages) & 1)) break;
        if (k >= *reinterpret_cast<volatile int*>(&s_stop)) { stop = true; break; }
        if (it > (1u << 24)) __trap();   // a byte-count mismatch would otherwise hang the GPU
      }
      if (__any_sync(kFull, stop)) break;
    }
    if (!warp_done) {
      const int cnt = min(kFChunk, n - k * kFChunk);
      const float4* sr = &srec[st][0];
      const int idx0 = g.start + k * kFChunk;
      for (int base = 0; base < cnt && !warp_done; base += 32) {
        // lane j tests record base+j against the warp's 8x8 pixel rectangle
        const int j = base + lane;
        bool hit = false;
        if (j < cnt) {
          const float4 r0 = sr[j * 3], r1 = sr[j * 3 + 1], r2 = sr[j * 3 + 2];
          const float s = min_sigma_rect(r0.x, r0.y, r0.z, r0.w, r1.x, rxmin, rxmax, rymin, rymax);
          hit = !(s > r2.w + (kLog2_255 + kCullMargin));
          // a masked-out splat is a splat of opacity 0: it fails the alpha >= 1/255 test on every pixel
          if (p.slot_keep && hit) hit = p.slot_keep[__float_as_int(r2.z)] != 0;
        }
        // walked front to back: the mask is bit-reversed so that "highest set bit" (one FLO) is the first record;
        // bit fb <-> record base + 31 - fb
        unsigned m = __brev(__ballot_sync(kFull, hit));
        const uint32_t top_addr = smem_addr(sr + (base + 31) * 3);
        const int top_idx = idx0 + base + 31;
        while (m) {
          const int fb = bfind_u32(m);
          m ^= bit_mask(fb);
          const float4* rp = smem_ptr<float4>(top_addr - (uint32_t)fb * kRecBytes);
          const float4 r0 = rp[0], r1 = rp[1], r2 = rp[2];
          // e = log2(opacity) - sigma', sigma' = a' dx^2 + b' dx dy + c' dy^2  (alpha = 2^e); dx is shared by the
          // lane's two pixels, everything in dy runs packed.  ndx = -dx exactly, so the bits are those of the
          // backward's scalar form
          const float ndx = pxf - r0.x;
          const float adx = r0.z * -ndx;
          const f32x2 dy2 = sub2(pk2(r0.y, r0.y), pyf2);
          const f32x2 t2 = fma2(pk2(r0.w, r0.w), dy2, pk2(adx, adx));
          f32x2 e2 = fma2(pk2(ndx, ndx), t2, pk2(r2.w, r2.w));
          e2 = fma2(mul2(pk2(-r1.x, -r1.x), dy2), dy2, e2);
          float e0, e1;
          upk2(e2, e0, e1);
          const float a0 = fminf(kAlphaMax, exp2f(e0)), a1 = fminf(kAlphaMax, exp2f(e1));
          const bool ok0 = e0 >= emin0 && e0 <= r2.w;   // alpha >= 1/255 (and not saturated) and sigma' >= 0
          const bool ok1 = e1 >= emin1 && e1 <= r2.w;
          const f32x2 a2 = pk2(a0, a1);
          const f32x2 nT2 = mul2(T2, sub2(one2, a2));
          float nT0, nT1, T0, T1, v0, v1;
          upk2(nT2, nT0, nT1);
          upk2(T2, T0, T1);
          const bool go0 = ok0 && nT0 > kTStop;         // gsplat stops BEFORE adding the saturating record
          const bool go1 = ok1 && nT1 > kTStop;
          upk2(mul2(a2, T2), v0, v1);
          const f32x2 vis2 = pk2(go0 ? v0 : 0.f, go1 ? v1 : 0.f);
          fma2_acc(cr2, vis2, pk2(r1.z, r1.z));
          fma2_acc(cg2, vis2, pk2(r1.w, r1.w));
          fma2_acc(cb2, vis2, pk2(r2.x, r2.x));
          fma2_acc(cd2, vis2, pk2(r2.y, r2.y));
          T2 = pk2(go0 ? nT0 : T0, go1 ? nT1 : T1);
          last0 = go0 ? top_idx - fb : last0;
          last1 = go1 ? top_idx - fb : last1;
          emin0 = (ok0 && !go0) ? INFINITY : emin0;
          emin1 = (ok1 && !go1) ? INFINITY : emin1;
        }
        warp_done = __all_sync(kFull, emin0 > 0.f && emin1 > 0.f);
      }
    }
    if (warp_done && !counted_done) {
      counted_done = true;
      if (lane == 0) atomicAdd(&s_done_warps, 1);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&freed[st]);   // release: orders the counter update above before the refill decision
    if (k + kFStages < nchunks && warp == (k & (kFwdWarps - 1)) && lane == 0) {
      mbar_wait(&freed[st], (k / kFStages) & 1);
      for (unsigned it = 0; *reinterpret_cast<volatile int*>(&s_decided) != k; ++it)   // the decision for chunk k - 1 first
        if (it > (1u << 24)) __trap();
      if (*reinterpret_cast<volatile int*>(&s_stop) == nchunks) {   // nobody has stopped the loads yet
        if (*reinterpret_cast<volatile int*>(&s_done_warps) == kFwdWarps) {
          *reinterpret_cast<volatile int*>(&s_stop) = k + kFStages;
        } else {
          const int kk = k + kFStages;
          const int cnt = min(kFChunk, n - kk * kFChunk);
          mbar_expect_tx(&bars[st], cnt * kRecBytes);
          bulk_g2s(&srec[st][0], p.recs + (size_t)(g.start + kk * kFChunk) * 3, cnt * kRecBytes, &bars[st]);
        }
      }
      __threadfence_block();
      *reinterpret_cast<volatile int*>(&s_decided) = k + 1;
    }
  }
  // every chunk that was loaded has been waited for by every warp (a warp only leaves at a chunk >= s_stop): no bulk
  // copy is in flight into this CTA's shared memory when it exits

  float T0, T1, r0_, r1_, g0_, g1_, b0_, b1_, d0_, d1_;
  upk2(T2, T0, T1);
  upk2(cr2, r0_, r1_); upk2(cg2, g0_, g1_); upk2(cb2, b0_, b1_); upk2(cd2, d0_, d1_);
  if (g.in0) fwd_epilogue<MODE>(p, g.cam, g.pix0, g.px, g.py0, T0, r0_, g0_, b0_, d0_, last0);
  if (g.in1) fwd_epilogue<MODE>(p, g.cam, g.pix1, g.px, g.py1, T1, r1_, g1_, b1_, d1_, last1);
}

// =================================================================================================
// backward
// =================================================================================================
// Deferred per-record reduction.  Summing a record's per-pixel gradient terms across the 32 lanes right
// away costs a 12-value transposing shuffle tree per (warp, record).  Instead the walk only stores the two
// per-(pixel, record) scalars every gradient is linear in,
//     w   = [alpha unclamped] * araw * v_alpha     (d loss / d log-ish opacity term; v_sigma' = -ln2 * w)
//     fac = alpha * T                              (blend weight: v_colour = fac * v_C)
// into a per-warp [kBatch records][32 pixels] shared-memory panel.  When kBatch records are pending the warp
// TRANSPOSES the work: lane = (record, half of the 8x4 rectangle) walks its 16 pixels, forming the six
// pixel-local moments of w (1, u, v, u^2, uv, v^2), the four colour sums and the two absgrad sums in
// registers - no shuffles except one 16-lane fold - re-centres the moments on the record's mean and leaves
// as three 128-bit vector reductions (red.global.add.v4.f32) into the record's 48-byte gradient line:
//     {m_x, m_y, m_xx, m_xy, m_yy, m_0, v_r, v_g, v_b, v_depth, sum|w g_x|, sum|w g_y|}
// with m_ab = sum_p w_p dx_p^a dy_p^b, (dx, dy) = mean2d - pixel centre, g = (2a' dx + b' dy, b' dx + 2c' dy).
// project_bwd_kernel turns the moments into v_mean2d / v_conic / v_opacity (all linear in them).
#ifndef BDS_BWD_STAGES
#define BDS_BWD_STAGES 2
#endif
constexpr int kBStages = BDS_BWD_STAGES;   // backward: TMA stages of kChunk records
#ifndef BDS_BATCH
#define BDS_BATCH 16
#endif
constexpr int kBatch = BDS_BATCH;      // records per deferred-reduction batch (8 or 16)
constexpr int kParts = 32 / kBatch;    // lanes per record in the transposed phase
constexpr int kPartPix = 32 / kParts;  // pixels per lane: 8 = one row of the 8x4 rectangle, 16 = two rows
constexpr int kPanelStride = 33;       // float2 per record row: 32 pixels + 1 pad (conflict-free transposed reads)

struct BatchSmem {                     // per warp
  float2 panel[kBatch * kPanelStride]; // {w, fac}
  float4 carry[kBatch * 3];            // records still pending when their TMA stage was released (whole 48-byte records)
  float4 vc[32];                       // per-pixel cotangent of the raw accumulators (C_r, C_g, C_b, D)
};

BDS_D float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// rx0, ry0: centre of the warp rectangle's first pixel; nb: pending records (warp-uniform); rec_addr: lane i < nb
// holds the shared-window address of the record pending in panel row i - inside a TMA stage that is still held,
// or inside the warp's carry area
BDS_D void flush_batch(const BatchSmem* bs, int nb, float rx0, float ry0, float* __restrict__ v_splats,
                       uint32_t rec_addr) {
  __syncwarp();
  const int lane = threadIdx.x & 31;
  const int rec = lane & (kBatch - 1), part = lane / kBatch;
  const float2* row = bs->panel + rec * kPanelStride + part * kPartPix;
  const float4* vcp = bs->vc + part * kPartPix;
  const float4* rp = smem_ptr<float4>(__shfl_sync(kFull, rec_addr, rec));
  const float4 q0 = rp[0];                                         // record {x, y, a', b'}
  const float q1x = reinterpret_cast<const float*>(rp)[4];         // c'
  const int q1slot = reinterpret_cast<const int*>(rp)[10];         // bits(slot)
  const float X = q0.x - rx0, Y = q0.y - ry0;          // mean relative to pixel (u, v) = (0, 0)
  const float A2 = 2.f * q0.z, B = q0.w, C2 = 2.f * q1x;
  float m0 = 0.f, mu = 0.f, mv = 0.f, muu = 0.f, muv = 0.f, mvv = 0.f;
  float ax = 0.f, ay = 0.f;
  f32x2 c01 = pk2(0.f, 0.f), c23 = c01;   // colour sums as packed fp32x2 (FFMA2)
  const f32x2 gstep = pk2(-A2, -B);       // d(g_x, g_y) / du
#pragma unroll
  for (int r = 0; r < kPartPix / 8; ++r) {
    const float v = (float)(part * (kPartPix / 8) + r);
    const float dy = Y - v;
    f32x2 g2 = pk2(fmaf(A2, X, B * dy), fmaf(B, X, C2 * dy));   // (g_x, g_y) at u = 0, stepped along the row
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float2 d = row[r * 8 + u];
      const float4 c = vcp[r * 8 + u];
      s0 += d.x;
      s1 = fmaf(d.x, (float)u, s1);
      s2 = fmaf(d.x, (float)(u * u), s2);
      const f32x2 fac2 = pk2(d.y, d.y);
      fma2_acc(c01, fac2, pk2(c.x, c.y));
      fma2_acc(c23, fac2, pk2(c.z, c.w));
      float wgx, wgy;
      upk2(mul2(pk2(d.x, d.x), g2), wgx, wgy);
      ax += fabsf(wgx);
      ay += fabsf(wgy);
      if (u < 7) g2 = add2(g2, gstep);
    }
    m0 += s0; mu += s1; muu += s2;
    mv = fmaf(v, s0, mv); muv = fmaf(v, s1, muv); mvv = fmaf(v * v, s0, mvv);
  }
  float c0, c1, c2, c3;
  upk2(c01, c0, c1);
  upk2(c23, c2, c3);
#pragma unroll
  for (int o = 16; o >= kBatch; o >>= 1) {   // fold the kParts partial sums of each record
    m0 += __shfl_xor_sync(kFull, m0, o); mu += __shfl_xor_sync(kFull, mu, o); mv += __shfl_xor_sync(kFull, mv, o);
    muu += __shfl_xor_sync(kFull, muu, o); muv += __shfl_xor_sync(kFull, muv, o); mvv += __shfl_xor_sync(kFull, mvv, o);
    c0 += __shfl_xor_sync(kFull, c0, o); c1 += __shfl_xor_sync(kFull, c1, o); c2 += __shfl_xor_sync(kFull, c2, o);
    c3 += __shfl_xor_sync(kFull, c3, o); ax += __shfl_xor_sync(kFull, ax, o); ay += __shfl_xor_sync(kFull, ay, o);
  }
  if (rec < nb && part < 3) {
    // pixel-local -> mean-centred moments: dx = X - u, dy = Y - v
    const float mx = fmaf(X, m0, -mu), my = fmaf(Y, m0, -mv);
    const float mxx = fmaf(X, mx - mu, muu);
    const float mxy = fmaf(X, my, fmaf(-Y, mu, muv));
    const float myy = fmaf(Y, my - mv, mvv);
    float* dst = v_splats + (size_t)q1slot * 12;
    if (kParts >= 3) {          // one 128-bit reduction per lane
      if (part == 0) red_add_v4(dst, mx, my, mxx, mxy);
      else if (part == 1) red_add_v4(dst + 4, myy, m0, c0, c1);
      else red_add_v4(dst + 8, c2, c3, ax, ay);
    } else if (part == 0) {
      red_add_v4(dst, mx, my, mxx, mxy);
      red_add_v4(dst + 4, myy, m0, c0, c1);
    } else {
      red_add_v4(dst + 8, c2, c3, ax, ay);
    }
  }
  __syncwarp();
}

constexpr size_t kBwdSmemRec = (size_t)kBStages * kChunk * kRecBytes;
constexpr size_t kBwdSmemWalk = kBwdSmemRec + 8 * sizeof(BatchSmem);
constexpr size_t kBwdSmemBilPark = kPanelBytes + (size_t)(BDS_MAX_LEVELS - 1) * 3 * 256 * sizeof(float4);   // panels + parked levels
constexpr size_t kBwdSmem = kBwdSmemBilPark > kBwdSmemWalk ? kBwdSmemBilPark : kBwdSmemWalk;

#ifndef BDS_BWD_MINB
#define BDS_BWD_MINB 4     // resident CTAs per SM the backward is compiled for
#endif
template <int MODE>
__global__ void __launch_bounds__(256, BDS_BWD_MINB) composite_bwd_kernel(CompParams p) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  float4 (*srec)[kChunk * 3] = reinterpret_cast<float4 (*)[kChunk * 3]>(dyn_smem);
  __shared__ __align__(8) uint64_t bars[kBStages];    // stage filled (TMA complete_tx)
  __shared__ __align__(8) uint64_t freed[kBStages];   // stage consumed by all eight warps
  __shared__ int s_last[8];

  const TileGeom g = tile_geom(p);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // ---- per-pixel epilogue backward: cotangents of the raw accumulators C (3), D and of A = 1 - T
  float vC[4] = {0.f, 0.f, 0.f, 0.f};
  float vA = 0.f, Tfin = 1.f;
  int last = -1;
  float bgv[4] = {0.f, 0.f, 0.f, 0.f};
  float A = 0.f;
  float gr = 0.f, gg2 = 0.f, gb = 0.f;           // cotangent of rgb_in (modes 1/2)
  float rg = 0.f, gg = 0.f, bg = 0.f, sk[3] = {0.f, 0.f, 0.f};
  if (g.inside) {
    last = p.last_ids[g.pix];
    A = p.out_alpha[g.pix];
    Tfin = 1.f - A;
    if (p.v_alpha) vA = p.v_alpha[g.pix];
    if (MODE == 0) {
      float vo[4] = {0.f, 0.f, 0.f, 0.f};
      for (int c = 0; c < p.channels; ++c) vo[c] = p.v_rgb[g.pix * p.channels + c];
      if (p.backgrounds)  // T_final * <bg, v> enters through the walk (bg_dot below)
        for (int c = 0; c < p.channels; ++c) bgv[c] = p.backgrounds[g.cam * p.channels + c];
      vC[0] = vo[0]; vC[1] = vo[1]; vC[2] = vo[2];
      if (p.channels == 4) {
        if (p.expected_depth) {
          float Ac = fmaxf(A, 1e-10f);
          vC[3] = vo[3] / Ac;
          if (A >= 1e-10f) vA -= vo[3] * p.out_depth[g.pix] / Ac;  // out_depth holds the normalised depth
        } else {
          vC[3] = vo[3];
        }
      }
    } else {
      rg = p.out_rgbg[g.pix * 3]; gg = p.out_rgbg[g.pix * 3 + 1]; bg = p.out_rgbg[g.pix * 3 + 2];
      if (p.sky) { sk[0] = p.sky[g.pix * 3]; sk[1] = p.sky[g.pix * 3 + 1]; sk[2] = p.sky[g.pix * 3 + 2]; }
      gr = p.v_rgb[g.pix * 3]; gg2 = p.v_rgb[g.pix * 3 + 1]; gb = p.v_rgb[g.pix * 3 + 2];
    }
  }
  if (MODE == 2) {
    // chain backward (appendix A.3 of SURVEY.md); every warp reduces the grid-node gradients of its own 8x4 pixels
    // (warp_level_accumulate, bilateral_accum.cuh: no block barrier), pixels outside the image contribute nothing
    WarpPanel* const pn = reinterpret_cast<WarpPanel*>(dyn_smem) + warp;
    const float x0r = fmaf(sk[0], Tfin, rg), x0g = fmaf(sk[1], Tfin, gg), x0b = fmaf(sk[2], Tfin, bg);
    const float lum = luma_of(x0r, x0g, x0b);
    const int pxc = min(g.px, p.W - 1), pyc = min(g.py, p.H - 1);
    const float x01 = lin01(pxc, p.W), y01 = lin01(pyc, p.H);   // one IEEE division per axis, not per level
    // Forward order first: levels 0 .. n-2 are fetched ONCE (values and slab differences); what the backward order
    // needs of them - the 3x3 part of A_l (for A_l^T g) and d_l = dA_l/dfz [x_l; 1] (for the guidance gradient) - is
    // parked in shared memory (12 floats per level and thread) instead of being fetched a second time.
    float4* const park = reinterpret_cast<float4*>(dyn_smem + kPanelBytes) + threadIdx.x;   // [level][3][256]
    float xs[BDS_MAX_LEVELS][3];
    {
      float r = x0r, gq = x0g, b = x0b;
#pragma unroll
      for (int l = 0; l < BDS_MAX_LEVELS; ++l) {
        if (l < p.bil.n_levels) {
          xs[l][0] = r; xs[l][1] = gq; xs[l][2] = b;
          if (l + 1 < p.bil.n_levels) {   // the last level is fetched in the backward loop
            const float* grid = p.bil.grid_cl[l] + (size_t)g.cam * p.bil.L[l] * p.bil.GY[l] * p.bil.GX[l] * 12;
            Tri t = tri_setup(unit_coord(x01, p.bil.GX[l]), unit_coord(y01, p.bil.GY[l]),
                              luma_coord(lum, p.bil.L[l]), p.bil.L[l], p.bil.GY[l], p.bil.GX[l]);
            float Al[12], dAdz[12];
            tri_fetch<true>(grid, t, Al, dAdz);
            const float d0 = fmaf(dAdz[0], r, fmaf(dAdz[1], gq, fmaf(dAdz[2], b, dAdz[3])));
            const float d1 = fmaf(dAdz[4], r, fmaf(dAdz[5], gq, fmaf(dAdz[6], b, dAdz[7])));
            const float d2 = fmaf(dAdz[8], r, fmaf(dAdz[9], gq, fmaf(dAdz[10], b, dAdz[11])));
            park[(l * 3 + 0) * 256] = make_float4(Al[0], Al[1], Al[2], d0);
            park[(l * 3 + 1) * 256] = make_float4(Al[4], Al[5], Al[6], d1);
            park[(l * 3 + 2) * 256] = make_float4(Al[8], Al[9], Al[10], d2);
            affine_apply(Al, r, gq, b);
          }
        }
      }
    }
    float v_lum = 0.f;
#pragma unroll
    for (int l = BDS_MAX_LEVELS - 1; l >= 0; --l) {
      if (l < p.bil.n_levels) {
        const int Ll = p.bil.L[l], GYl = p.bil.GY[l], GXl = p.bil.GX[l];
        const size_t goff = (size_t)g.cam * Ll * GYl * GXl * 12;
        Tri t = tri_setup(unit_coord(x01, GXl), unit_coord(y01, GYl), luma_coord(lum, Ll), Ll, GYl, GXl);
        const float xr = xs[l][0], xg = xs[l][1], xb = xs[l][2];
        float R[9], d0, d1, d2;   // 3x3 part of A_l, dA_l/dfz [x_l; 1]
        if (l + 1 == p.bil.n_levels) {
          float Al[12], dAdz[12];
          tri_fetch<true>(p.bil.grid_cl[l] + goff, t, Al, dAdz);
          d0 = fmaf(dAdz[0], xr, fmaf(dAdz[1], xg, fmaf(dAdz[2], xb, dAdz[3])));
          d1 = fmaf(dAdz[4], xr, fmaf(dAdz[5], xg, fmaf(dAdz[6], xb, dAdz[7])));
          d2 = fmaf(dAdz[8], xr, fmaf(dAdz[9], xg, fmaf(dAdz[10], xb, dAdz[11])));
          R[0] = Al[0]; R[1] = Al[1]; R[2] = Al[2]; R[3] = Al[4]; R[4] = Al[5]; R[5] = Al[6];
          R[6] = Al[8]; R[7] = Al[9]; R[8] = Al[10];
        } else {
          const float4 q0 = park[(l * 3 + 0) * 256], q1 = park[(l * 3 + 1) * 256], q2 = park[(l * 3 + 2) * 256];
          R[0] = q0.x; R[1] = q0.y; R[2] = q0.z; d0 = q0.w;
          R[3] = q1.x; R[4] = q1.y; R[5] = q1.z; d1 = q1.w;
          R[6] = q2.x; R[7] = q2.y; R[8] = q2.z; d2 = q2.w;
        }
        if (t.z_inside)   // guidance gradient: <g (x) [x; 1], dA/dfz> (L - 1)
          v_lum = fmaf(fmaf(gr, d0, fmaf(gg2, d1, gb * d2)), (float)(Ll - 1), v_lum);
#ifndef BDS_DIAG_NO_ACCUM   // timing experiments only (scripts/gpu_variants.sh): results are wrong without it
        warp_level_accumulate(pn, t, g.inside, gr, gg2, gb, xr, xg, xb, Ll, GYl, GXl, p.bil.v_grid_cl[l] + goff);
#else
        if (gr == 123.456f) p.bil.v_grid_cl[l][goff] = gr + xr;
#endif
        // cotangent of the level input: A[:, :3]^T g
        const float nr = R[0] * gr + R[3] * gg2 + R[6] * gb;
        const float ng = R[1] * gr + R[4] * gg2 + R[7] * gb;
        const float nb = R[2] * gr + R[5] * gg2 + R[8] * gb;
        gr = nr; gg2 = ng; gb = nb;
      }
    }
    gr += v_lum * kLumaR; gg2 += v_lum * kLumaG; gb += v_lum * kLumaB;
    // the panel bytes are about to be overwritten by TMA (async proxy): order the
    // generic-proxy accesses above before it (the __syncthreads below publishes it block-wide)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (MODE != 0 && g.inside) {
    // (gr, gg2, gb) = cotangent of rgb_in = rgb_gauss + sky * (1 - A)
    if (p.v_sky) { p.v_sky[g.pix * 3] = gr * Tfin; p.v_sky[g.pix * 3 + 1] = gg2 * Tfin; p.v_sky[g.pix * 3 + 2] = gb * Tfin; }
    vA -= sk[0] * gr + sk[1] * gg2 + sk[2] * gb;
    float vr = gr, vg = gg2, vb = gb;
    if (p.v_rgbg) { vr += p.v_rgbg[g.pix * 3]; vg += p.v_rgbg[g.pix * 3 + 1]; vb += p.v_rgbg[g.pix * 3 + 2]; }
    vC[0] = rg < 1.f ? vr : 0.f;   // clamp(max=1) passes gradient below the bound
    vC[1] = gg < 1.f ? vg : 0.f;
    vC[2] = bg < 1.f ? vb : 0.f;
    if (p.v_depth) {
      float Ac = fmaxf(A, 1e-10f);
      float vd = p.v_depth[g.pix];
      vC[3] = vd / Ac;
      if (A >= 1e-10f) vA -= vd * p.out_depth[g.pix] / Ac;
    }
  }

  // ---- block-wide last contributing record
  int wl = last;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wl = max(wl, __shfl_xor_sync(kFull, wl, o));
  if (lane == 0) s_last[warp] = wl;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kBStages; ++s) { mbar_init(&bars[s], 1); mbar_init(&freed[s], 8); }
    mbar_fence_init();
  }
  // this warp's deferred-reduction panel (beyond the record stages; the bilateral prologue is done with
  // the bytes it shared with them)
  BatchSmem* bs = reinterpret_cast<BatchSmem*>(dyn_smem + kBwdSmemRec) + warp;
  __syncthreads();   // (mode 2) the prologue's shared memory is free: panels / record stages may be written
  bs->vc[lane] = make_float4(vC[0], vC[1], vC[2], vC[3]);   // read by this warp only (flush_batch syncs the warp)
  int block_last = -1;
#pragma unroll
  for (int w = 0; w < 8; ++w) block_last = max(block_last, s_last[w]);
  const int warp_last = wl;

  const int n = (block_last < 0) ? 0 : block_last - g.start + 1;
  const int nchunks = (n + kChunk - 1) / kChunk;
  // chunks are walked from the back: walk index q = 0.. corresponds to chunk k = nchunks-1-q
  int issued = 0;
  if (threadIdx.x == 0) {
    for (; issued < nchunks && issued < kBStages; ++issued) {
      int k = nchunks - 1 - issued;
      int cnt = min(kChunk, n - k * kChunk);
      mbar_expect_tx(&bars[issued], cnt * kRecBytes);
      bulk_g2s(&srec[issued][0], p.recs + (size_t)(g.start + k * kChunk) * 3, cnt * kRecBytes, &bars[issued]);
    }
  }

  const float pxf = (float)g.px + 0.5f, pyf = (float)g.py + 0.5f;
  const float rxmin = g.wx0 + 0.5f, rxmax = g.wx0 + 7.5f, rymin = g.wy0 + 0.5f, rymax = g.wy0 + 3.5f;
  // background: render = C + T_final * bg  ->  d/dalpha_i carries -T_final/(1-alpha_i) <bg, v>
  float bg_dot = 0.f;
  if (MODE == 0 && p.backgrounds) {
    bg_dot = bgv[0] * vC[0] + bgv[1] * vC[1] + bgv[2] * vC[2];
    if (p.channels == 4) bg_dot += bgv[3] * vC[3];
  }
  // per-pixel walk state.  With S_i = sum over the records behind i of fac_j <c_j, v_C>:
  //   v_alpha_i = T_i <c_i, v_C> - (S_i - T_final (v_A - <bg, v_C>)) / (1 - alpha_i)
  // so ONE running scalar (bufdot) replaces the four colour buffers of the textbook form.
  float T = Tfin;
  float bufdot = -Tfin * (vA - bg_dot);
  if (!g.inside) last = -1;   // pixels outside the image never match a record
  // deferred-reduction bookkeeping, all in registers: lane i holds in rec_addr the shared-window address of the
  // record pending in panel row i (no per-record metadata stores)
  float2* const pw0 = bs->panel + lane;
  float2* pw = pw0;           // this lane's cell of the next free panel row
  int nb = 0;                 // records pending in this warp's panel (warp-uniform)
  uint32_t rec_addr = smem_addr(dyn_smem);   // always a readable address, also in lanes >= nb
  const uint32_t stages_end = smem_addr(dyn_smem) + (uint32_t)kBwdSmemRec;
  const f32x2 pxy2 = pk2(pxf, pyf);
  const f32x2 vC01 = pk2(vC[0], vC[1]), vC23 = pk2(vC[2], vC[3]);

  for (int q = 0; q < nchunks; ++q) {
    const int st = q % kBStages;
    const int k = nchunks - 1 - q;
    mbar_wait(&bars[st], (q / kBStages) & 1);
    const int cnt = min(kChunk, n - k * kChunk);
    const int chunk0 = g.start + k * kChunk;
    if (warp_last >= chunk0) {
      const float4* sr = &srec[st][0];
      for (int base = ((cnt - 1) / 32) * 32; base >= 0; base -= 32) {
        if (chunk0 + base > warp_last) continue;
        int j = base + lane;
        bool hit = false;
        if (j < cnt) {
          float4 r0 = sr[j * 3], r1 = sr[j * 3 + 1], r2 = sr[j * 3 + 2];
          float s = min_sigma_rect(r0.x, r0.y, r0.z, r0.w, r1.x, rxmin, rxmax, rymin, rymax);
          hit = !(s > r2.w + (kLog2_255 + kCullMargin));
        }
        unsigned m = __ballot_sync(kFull, hit);
        const uint32_t group_addr = smem_addr(sr + base * 3);
        const int last_rel = last - chunk0 - base;   // this pixel takes bit <= last_rel of the group
        while (m) {
          const int bit = bfind_u32(m);              // from the back: highest record first
          m ^= bit_mask(bit);
          const uint32_t raddr = group_addr + (uint32_t)bit * kRecBytes;
          const float4* rp = smem_ptr<float4>(raddr);
          const float4 r0 = rp[0], r1 = rp[1], r2 = rp[2];
          float dx, dy;
          upk2(sub2(pk2(r0.x, r0.y), pxy2), dx, dy);
          // e = log2(opacity) - sigma', sigma' = a' dx^2 + b' dx dy + c' dy^2  (alpha = 2^e)
          float e = fmaf(-dx, fmaf(r0.w, dy, r0.z * dx), r2.w);
          e = fmaf(-(r1.x * dy), dy, e);
          const bool valid = (bit <= last_rel) & (e >= -kLog2_255) & (r2.w >= e);
          if (!__any_sync(kFull, valid)) continue;
          // branch-free: a lane that does not take the record runs the same arithmetic with alpha = 0, which
          // leaves T and bufdot unchanged (rcp(1) = 1, 0 * finite = 0) and stores w = fac = 0
          const float araw = valid ? exp2f(e) : 0.f;   // opacity * exp(-sigma)
          const float alpha = fminf(kAlphaMax, araw);
          const float ra = rcp_approx(1.f - alpha);
          T *= ra;                                     // transmittance in front of this record
          const float fac = alpha * T;
          float cd0, cd1;                              // <colour, v_C>, two channels per lane of the packed ops
          upk2(fma2(pk2(r2.x, r2.y), vC23, mul2(pk2(r1.z, r1.w), vC01)), cd0, cd1);
          const float cdot = cd0 + cd1;
          const float v_alpha = fmaf(T, cdot, -ra * bufdot);
          bufdot = fmaf(fac, cdot, bufdot);
          const float w = araw <= kAlphaMax ? araw * v_alpha : 0.f;  // the clamp at 0.999 blocks the gradient
          *pw = make_float2(w, fac);
          rec_addr = nb == (int)lane_id() ? raddr : rec_addr;
          pw += kPanelStride;
          if (++nb == kBatch) {
            flush_batch(bs, kBatch, rxmin, rymin, p.v_splats, rec_addr);
            nb = 0;
            pw = pw0;
          }
        }
      }
      // the stage is about to be released: records of it that are still pending move to the warp's carry area
      if (lane < nb && rec_addr < stages_end) {
        const float4* src = smem_ptr<float4>(rec_addr);
        float4* dst = bs->carry + lane * 3;
        const float4 c0 = src[0], c1 = src[1], c2 = src[2];
        dst[0] = c0; dst[1] = c1; dst[2] = c2;
        rec_addr = smem_addr(dst);
      }
    }
    // consumer release: the warp is done with stage st.  No block barrier - the other warps run ahead by up to
    // kBStages - 1 chunks; the refill of the stage (walk index q + kBStages) is issued by lane 0 of warp q % 8 once
    // all eight warps have released it
    __syncwarp();
    if (lane == 0) mbar_arrive(&freed[st]);
    if (q + kBStages < nchunks && warp == (q & 7) && lane == 0) {
      mbar_wait(&freed[st], (q / kBStages) & 1);
      const int kk = nchunks - 1 - (q + kBStages);
      const int c2 = min(kChunk, n - kk * kChunk);
      mbar_expect_tx(&bars[st], c2 * kRecBytes);
      bulk_g2s(&srec[st][0], p.recs + (size_t)(g.start + kk * kChunk) * 3, c2 * kRecBytes, &bars[st]);
    }
  }
  if (nb > 0) flush_batch(bs, nb, rxmin, rymin, p.v_splats, rec_addr);
  // v_backgrounds: sum over pixels of T_final * v (MODE 0)
  if (MODE == 0 && p.v_backgrounds) {
    for (int c = 0; c < p.channels; ++c) {
      float contrib = g.inside ? Tfin * vC[c] : 0.f;
      float sum = warp_sum(contrib);
      if (lane == 0 && sum != 0.f) red_add(p.v_backgrounds + g.cam * p.channels + c, sum);
    }
  }
}

// channel-first parameter slots <-> channel-last workspace, all (camera, level) pairs in ONE launch
constexpr int kMaxRepackJobs = 48;
struct RepackJobs {
  const float* src[kMaxRepackJobs];
  float* dst[kMaxRepackJobs];
  int L[kMaxRepackJobs], GY[kMaxRepackJobs], GX[kMaxRepackJobs];
  int n;
};
template <bool UNPACK_ADD>
__global__ void repack_jobs_kernel(RepackJobs jobs) {
  const int job = blockIdx.y;
  const int L = jobs.L[job], GY = jobs.GY[job], GX = jobs.GX[job];
  const int nodes = L * GY * GX;
  const float* __restrict__ src = jobs.src[job];
  float* __restrict__ dst = jobs.dst[job];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nodes * 12; i += gridDim.x * blockDim.x) {
    if (UNPACK_ADD) {  // dst = parameter layout [12][L][GY][GX] += src repack [GY][GX][L][12]
      int ch = i / nodes, rem = i - ch * nodes;
      int x = rem % GX, y = (rem / GX) % GY, z = rem / (GX * GY);
      dst[i] += src[(size_t)bil_node(x, y, z, L, GX) * 12 + ch];
    } else {           // dst = value repack [GY][GX][3][L][4] = src parameter layout
      dst[i] = src[bil_value_param_index(i, L, GY, GX)];
    }
  }
}
template <bool UNPACK_ADD>
static int launch_repack(RepackJobs& jobs, cudaStream_t stream) {
  if (jobs.n == 0) return 0;
  int max_nodes = 0;
  for (int k = 0; k < jobs.n; ++k) {
    int nk = jobs.L[k] * jobs.GY[k] * jobs.GX[k];
    max_nodes = nk > max_nodes ? nk : max_nodes;
  }
  dim3 grid(ceil_div((int64_t)max_nodes * 12, 256), jobs.n);
  repack_jobs_kernel<UNPACK_ADD><<<grid, 256, 0, stream>>>(jobs);
  BDS_CHECK_LAUNCH();
  jobs.n = 0;
  return 0;
}

struct CompWorkspace {
  size_t grid_cl[BDS_MAX_LEVELS], v_grid_cl[BDS_MAX_LEVELS];
  size_t total;
};
static CompWorkspace carve_comp(const bds_render_desc* d, const bds_epilogue_desc* e) {
  CompWorkspace w;
  size_t off = 0;
  for (int l = 0; l < BDS_MAX_LEVELS; ++l) w.grid_cl[l] = w.v_grid_cl[l] = 0;
  if (e->mode == 2) {
    for (int l = 0; l < e->bil.n_levels; ++l) {
      size_t b = align_up((size_t)d->n_cams * e->bil.L[l] * e->bil.GY[l] * e->bil.GX[l] * 12 * sizeof(float), 256);
      w.grid_cl[l] = off; off += b;
      w.v_grid_cl[l] = off; off += b;
    }
  }
  w.total = off;
  return w;
}

int check_render_desc(const bds_render_desc* d);

// does camera c own a tile row of the band [row_begin, row_end) (global tile rows, tile_h per camera)?
static bool cam_in_band(const bds_render_desc* d, int tile_h, int c) {
  return d->row_begin < (c + 1) * tile_h && d->row_end > c * tile_h;
}

static int fill_common(CompParams& p, const bds_render_desc* d, const bds_epilogue_desc* e) {
  BDS_REQUIRE(e && e->mode >= 0 && e->mode <= 2, "composite: epilogue mode must be 0, 1 or 2");
  if (e->mode == 0) BDS_REQUIRE(e->channels == 3 || e->channels == 4, "composite: channels must be 3 or 4");
  if (e->mode == 2) {
    BDS_REQUIRE(e->bil.n_levels >= 1 && e->bil.n_levels <= BDS_MAX_LEVELS, "composite: bad n_levels");
    for (int l = 0; l < e->bil.n_levels; ++l)
      BDS_REQUIRE(e->bil.factor[l] <= 1, "composite: the fused epilogue implements full-resolution guidance "
                                          "(guidance_factor=None); run mode 1 + bds_bilateral_* for low-res guidance");
  }
  p.W = d->width; p.H = d->height;
  p.tile_w = (d->width + kTile - 1) / kTile;
  p.tile_h = (d->height + kTile - 1) / kTile;
  p.row_begin = d->row_begin;
  // first stacked pixel row of the band
  int cam0 = d->row_begin / p.tile_h, ty0 = d->row_begin - cam0 * p.tile_h;
  p.pix_row0 = (int64_t)cam0 * d->height + (int64_t)ty0 * kTile;
  p.channels = e->mode == 0 ? e->channels : 4;
  p.expected_depth = e->expected_depth;
  return 0;
}

}  // namespace bds

using namespace bds;

extern "C" size_t bds_composite_workspace_bytes(const bds_render_desc* d, const bds_epilogue_desc* e) {
  if (!d || !e) return 0;
  return carve_comp(d, e).total + 256;
}

static int composite_fwd_impl(const bds_render_desc* d, const bds_epilogue_desc* e, const float* sorted_splats,
                              const int32_t* tile_offsets, const float* backgrounds, const float* sky,
                              const float* const* host_grids, float* out_rgb, float* out_rgb_gauss,
                              float* out_depth, float* out_alpha, int32_t* last_ids, void* workspace,
                              const uint8_t* slot_keep, bds_stream_t stream_);

extern "C" int bds_composite_fwd(const bds_render_desc* d, const bds_epilogue_desc* e, const float* sorted_splats,
                                 const int32_t* tile_offsets, const float* backgrounds, const float* sky,
                                 const float* const* host_grids, float* out_rgb, float* out_rgb_gauss,
                                 float* out_depth, float* out_alpha, int32_t* last_ids, void* workspace,
                                 bds_stream_t stream_) {
  return composite_fwd_impl(d, e, sorted_splats, tile_offsets, backgrounds, sky, host_grids, out_rgb, out_rgb_gauss,
                            out_depth, out_alpha, last_ids, workspace, nullptr, stream_);
}

extern "C" int bds_composite_fwd_masked(const bds_render_desc* d, const bds_epilogue_desc* e,
                                        const float* sorted_splats, const int32_t* tile_offsets,
                                        const uint8_t* slot_keep, const float* backgrounds, const float* sky,
                                        const float* const* host_grids, float* out_rgb, float* out_rgb_gauss,
                                        float* out_depth, float* out_alpha, int32_t* last_ids, void* workspace,
                                        bds_stream_t stream_) {
  BDS_REQUIRE(slot_keep, "composite_fwd_masked: slot_keep is null");
  return composite_fwd_impl(d, e, sorted_splats, tile_offsets, backgrounds, sky, host_grids, out_rgb, out_rgb_gauss,
                            out_depth, out_alpha, last_ids, workspace, slot_keep, stream_);
}

// slot_keep[slot] = gaussian_keep[Gaussian of the splat]
__global__ void slot_keep_kernel(const float* __restrict__ splats, const int32_t* __restrict__ counters, int n_gauss,
                                 const uint8_t* __restrict__ gaussian_keep, uint8_t* __restrict__ slot_keep) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= counters[0]) return;
  const int64_t idx = (int64_t)__float_as_int(splats[(size_t)slot * 12 + 10]);   // camera * N + Gaussian
  slot_keep[slot] = gaussian_keep[idx % n_gauss];
}

extern "C" int bds_slot_keep(const bds_render_desc* d, const float* splats, const int32_t* counters, int32_t n_slots,
                             const uint8_t* gaussian_keep, uint8_t* slot_keep, bds_stream_t stream_) {
  if (int rc = check_render_desc(d)) return rc;
  BDS_REQUIRE(splats && counters && gaussian_keep && slot_keep && n_slots >= 0, "slot_keep: bad arguments");
  if (n_slots == 0) return 0;
  slot_keep_kernel<<<ceil_div(n_slots, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(splats, counters, d->n_gauss,
                                                                                        gaussian_keep, slot_keep);
  BDS_CHECK_LAUNCH();
  return 0;
}

static int composite_fwd_impl(const bds_render_desc* d, const bds_epilogue_desc* e, const float* sorted_splats,
                              const int32_t* tile_offsets, const float* backgrounds, const float* sky,
                              const float* const* host_grids, float* out_rgb, float* out_rgb_gauss,
                              float* out_depth, float* out_alpha, int32_t* last_ids, void* workspace,
                              const uint8_t* slot_keep, bds_stream_t stream_) {
  if (int rc = check_render_desc(d)) return rc;
  CompParams p{};
  if (int rc = fill_common(p, d, e)) return rc;
  const int n_tiles = (d->row_end - d->row_begin) * p.tile_w;
  if (n_tiles == 0) return 0;
  BDS_REQUIRE(tile_offsets && out_rgb && out_alpha && last_ids, "composite_fwd: null pointer");
  if (e->mode != 0) BDS_REQUIRE(out_rgb_gauss && out_depth, "composite_fwd: modes 1/2 need rgb_gauss and depth outputs");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  p.recs = reinterpret_cast<const float4*>(sorted_splats);
  p.tile_offsets = tile_offsets;
  p.backgrounds = backgrounds; p.sky = sky;
  p.out_rgb = out_rgb; p.out_rgbg = out_rgb_gauss; p.out_depth = out_depth; p.out_alpha = out_alpha; p.last_ids = last_ids;
  p.slot_keep = slot_keep;
  if (e->mode == 2) {
    BDS_REQUIRE(host_grids && workspace, "composite_fwd: mode 2 needs grids and workspace");
    CompWorkspace w = carve_comp(d, e);
    p.bil.n_levels = e->bil.n_levels;
    RepackJobs jobs;
    jobs.n = 0;
    for (int l = 0; l < e->bil.n_levels; ++l) {
      int nodes = e->bil.L[l] * e->bil.GY[l] * e->bil.GX[l];
      float* base = reinterpret_cast<float*>(static_cast<char*>(workspace) + w.grid_cl[l]);
      p.bil.grid_cl[l] = base;
      p.bil.L[l] = e->bil.L[l]; p.bil.GY[l] = e->bil.GY[l]; p.bil.GX[l] = e->bil.GX[l];
      for (int c = 0; c < d->n_cams; ++c) {
        const float* src = host_grids[c * e->bil.n_levels + l];
        // a camera whose tile rows meet the band runs the chain on every pixel: its repack must not be skipped
        BDS_REQUIRE(src || !cam_in_band(d, p.tile_h, c), "composite_fwd: null grid slot for a camera inside the band");
        if (!src) continue;  // camera outside the band
        jobs.src[jobs.n] = src; jobs.dst[jobs.n] = base + (size_t)c * nodes * 12;
        jobs.L[jobs.n] = e->bil.L[l]; jobs.GY[jobs.n] = e->bil.GY[l]; jobs.GX[jobs.n] = e->bil.GX[l];
        if (++jobs.n == kMaxRepackJobs)
          if (int rc = launch_repack<false>(jobs, stream)) return rc;
      }
    }
    if (int rc = launch_repack<false>(jobs, stream)) return rc;
  }
  switch (e->mode) {
    case 0: composite_fwd_kernel<0><<<n_tiles, kFwdThreads, 0, stream>>>(p); break;
    case 1: composite_fwd_kernel<1><<<n_tiles, kFwdThreads, 0, stream>>>(p); break;
    default: composite_fwd_kernel<2><<<n_tiles, kFwdThreads, 0, stream>>>(p); break;
  }
  BDS_CHECK_LAUNCH();
  return 0;
}

extern "C" int bds_composite_bwd(const bds_render_desc* d, const bds_epilogue_desc* e, const float* sorted_splats,
                                 const int32_t* sorted_slots, const int32_t* tile_offsets, const float* backgrounds,
                                 const float* sky, const float* const* host_grids, const float* out_rgb_gauss,
                                 const float* out_depth, const float* out_alpha, const int32_t* last_ids,
                                 const float* v_rgb, const float* v_rgb_gauss, const float* v_depth,
                                 const float* v_alpha, float* v_splats, float* v_sky, float* const* host_v_grids,
                                 float* v_backgrounds, void* workspace, bds_stream_t stream_) {
  (void)sorted_slots;
  if (int rc = check_render_desc(d)) return rc;
  CompParams p{};
  if (int rc = fill_common(p, d, e)) return rc;
  const int n_tiles = (d->row_end - d->row_begin) * p.tile_w;
  if (n_tiles == 0) return 0;
  BDS_REQUIRE(tile_offsets && out_alpha && last_ids && v_rgb && v_splats, "composite_bwd: null pointer");
  BDS_REQUIRE(((uintptr_t)v_splats & 15u) == 0, "composite_bwd: v_splats must be 16-byte aligned (128-bit reductions)");
  if (e->mode != 0) BDS_REQUIRE(out_rgb_gauss && out_depth, "composite_bwd: modes 1/2 need rgb_gauss and depth");
  if (e->mode == 0 && e->channels == 4 && e->expected_depth)
    BDS_REQUIRE(out_depth, "composite_bwd: ED mode needs the normalised depth output (out_depth)");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  p.recs = reinterpret_cast<const float4*>(sorted_splats);
  p.tile_offsets = tile_offsets;
  p.backgrounds = backgrounds; p.sky = sky;
  p.out_rgbg = const_cast<float*>(out_rgb_gauss); p.out_depth = const_cast<float*>(out_depth);
  p.out_alpha = const_cast<float*>(out_alpha); p.last_ids = const_cast<int32_t*>(last_ids);
  p.v_rgb = v_rgb; p.v_rgbg = v_rgb_gauss; p.v_depth = v_depth; p.v_alpha = v_alpha;
  p.v_splats = v_splats; p.v_sky = v_sky; p.v_backgrounds = v_backgrounds;
  CompWorkspace w = carve_comp(d, e);
  if (e->mode == 2) {
    BDS_REQUIRE(host_grids && host_v_grids && workspace, "composite_bwd: mode 2 needs grids, v_grids and workspace");
    p.bil.n_levels = e->bil.n_levels;
    RepackJobs jobs;
    jobs.n = 0;
    for (int l = 0; l < e->bil.n_levels; ++l) {
      int nodes = e->bil.L[l] * e->bil.GY[l] * e->bil.GX[l];
      float* base = reinterpret_cast<float*>(static_cast<char*>(workspace) + w.grid_cl[l]);
      float* vbase = reinterpret_cast<float*>(static_cast<char*>(workspace) + w.v_grid_cl[l]);
      p.bil.grid_cl[l] = base; p.bil.v_grid_cl[l] = vbase;
      p.bil.L[l] = e->bil.L[l]; p.bil.GY[l] = e->bil.GY[l]; p.bil.GX[l] = e->bil.GX[l];
      for (int c = 0; c < d->n_cams; ++c) {
        const float* src = host_grids[c * e->bil.n_levels + l];
        BDS_REQUIRE(src || !cam_in_band(d, p.tile_h, c), "composite_bwd: null grid slot for a camera inside the band");
        if (!src) continue;
        jobs.src[jobs.n] = src; jobs.dst[jobs.n] = base + (size_t)c * nodes * 12;
        jobs.L[jobs.n] = e->bil.L[l]; jobs.GY[jobs.n] = e->bil.GY[l]; jobs.GX[jobs.n] = e->bil.GX[l];
        if (++jobs.n == kMaxRepackJobs)
          if (int rc = launch_repack<false>(jobs, stream)) return rc;
      }
    }
    if (int rc = launch_repack<false>(jobs, stream)) return rc;
    // the gradient halves of the workspace are contiguous per level pair (grid_cl | v_grid_cl): zero them
    for (int l = 0; l < e->bil.n_levels; ++l) {
      int nodes = e->bil.L[l] * e->bil.GY[l] * e->bil.GX[l];
      BDS_CHECK_CUDA(cudaMemsetAsync(p.bil.v_grid_cl[l], 0, (size_t)d->n_cams * nodes * 12 * sizeof(float), stream));
    }
  }
  switch (e->mode) {
    case 0:
      BDS_CHECK_CUDA(cudaFuncSetAttribute(composite_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)kBwdSmemWalk));
      composite_bwd_kernel<0><<<n_tiles, 256, kBwdSmemWalk, stream>>>(p);
      break;
    case 1:
      BDS_CHECK_CUDA(cudaFuncSetAttribute(composite_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)kBwdSmemWalk));
      composite_bwd_kernel<1><<<n_tiles, 256, kBwdSmemWalk, stream>>>(p);
      break;
    default:
      BDS_CHECK_CUDA(cudaFuncSetAttribute(composite_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)kBwdSmem));
      composite_bwd_kernel<2><<<n_tiles, 256, kBwdSmem, stream>>>(p);
      break;
  }
  BDS_CHECK_LAUNCH();
  if (e->mode == 2) {
    RepackJobs jobs;
    jobs.n = 0;
    for (int l = 0; l < e->bil.n_levels; ++l) {
      int nodes = e->bil.L[l] * e->bil.GY[l] * e->bil.GX[l];
      for (int c = 0; c < d->n_cams; ++c) {
        float* dst = host_v_grids[c * e->bil.n_levels + l];
        if (!dst) continue;
        jobs.src[jobs.n] = p.bil.v_grid_cl[l] + (size_t)c * nodes * 12; jobs.dst[jobs.n] = dst;
        jobs.L[jobs.n] = e->bil.L[l]; jobs.GY[jobs.n] = e->bil.GY[l]; jobs.GX[jobs.n] = e->bil.GX[l];
        if (++jobs.n == kMaxRepackJobs)
          if (int rc = launch_repack<true>(jobs, stream)) return rc;
      }
    }
    if (int rc = launch_repack<true>(jobs, stream)) return rc;
  }
  return 0;
}

