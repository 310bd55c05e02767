// Projection (+ optional fused activations and SH colour), tile counting with exact
// ellipse-vs-tile culling, compaction of visible splats into packed 48-byte records; and its VJP.
// Also the stand-alone spherical-harmonics entry points.
//
// One thread per (camera, Gaussian): 44-56 B read, <= 48 B written; HBM-bound pointwise work.
// Replaces gsplat's fully_fused_projection_fwd/bwd + compute_sh_fwd/bwd for the reference call
// sites models/trainers/base.py:393-408 and models/gaussians/vanilla.py:383-395.
#include "big_splats.cuh"
#include "sh_math.cuh"
#include "tma.cuh"

namespace bds {

struct ProjParams {
  bds_render_desc d;
  int tile_w, tile_h;
  const float *means, *quats, *scales, *opacities, *colors, *fdc, *frest, *viewmats, *Ks;
  int colors_per_cam;
  int32_t* radii;
  float *means2d, *depths, *conics, *compensations;
  int32_t* tiles_touched;
  int32_t* tile_counts;  // [band tiles + 1] per-tile record counts (atomics), or NULL
  float* splats;
  int32_t splat_cap;
  int32_t* slot_of;
  int32_t* counters;
};

BDS_D void load_cam(const float* __restrict__ viewmats, const float* __restrict__ Ks, int c, CamIntr& cam) {
  const float* V = viewmats + 16 * c;
  const float* K = Ks + 9 * c;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) cam.R[i * 3 + j] = __ldg(V + i * 4 + j);
    cam.t[i] = __ldg(V + i * 4 + 3);
  }
  cam.fx = __ldg(K); cam.fy = __ldg(K + 4); cam.cx = __ldg(K + 2); cam.cy = __ldg(K + 5);
}

BDS_D float sigmoidf(float x) { return 1.0f / (1.0f + __expf(-x)); }

// rows of camera c covered by the band, as tile rows [ty0, ty1)
BDS_HD void band_rows(const bds_render_desc& d, int tile_h, int c, int& ty0, int& ty1) {
  int g0 = c * tile_h, g1 = g0 + tile_h;
  int lo = d.row_begin > g0 ? d.row_begin : g0;
  int hi = d.row_end < g1 ? d.row_end : g1;
  if (hi <= lo) { ty0 = ty1 = 0; return; }
  ty0 = lo - g0;
  ty1 = hi - g0;
}

// Two phases per block of 256 Gaussians.
//  A. one thread per GAUSSIAN: parameters, activations and the world covariance are paid once; the
//     cameras of the band are walked with only the cheap conservative cull (depth range + screen bound).
//     A Gaussian is visible in ~1 of the 6 rig cameras, so the (Gaussian, camera) pairs that survive are
//     pushed to a shared-memory queue instead of being processed by the few lanes that own them.
//  B. one thread per QUEUE ENTRY: full EWA projection, exact tile count, SH colour, packed record.  Every
//     lane has work; the SH coefficients of the block were staged in shared memory by coalesced loads.
constexpr int kProjBlock = 256;
constexpr int kProjCamsPerRound = 8;
constexpr int kProjGaussFloats = 12;   // mu[3] cov[6] smax opacity pad
constexpr int kProjCamFloats = 24;     // R[9] t[3] fx fy cx cy | camera position[3] | pad
constexpr size_t kProjSmemBase = (size_t)kProjBlock * kProjGaussFloats * sizeof(float) +
                                 (size_t)kProjCamsPerRound * kProjCamFloats * sizeof(float) +
                                 (size_t)kProjBlock * kProjCamsPerRound * sizeof(uint16_t) + 8 * 32 * sizeof(int) + 32;
BDS_HD size_t proj_smem_bytes(int sh_floats) { return kProjSmemBase + (size_t)kProjBlock * sh_floats * sizeof(float); }

__global__ void __launch_bounds__(kProjBlock) project_fwd_kernel(ProjParams p, int sh_floats, int sh_bulk_ok) {
  extern __shared__ __align__(128) unsigned char proj_smem[];
  const int K3 = sh_floats;  // 3 * sh_K on the SH path, else 0
  float* s_dc = reinterpret_cast<float*>(proj_smem);                              // [256][3]
  float* s_rest = s_dc + (K3 > 0 ? kProjBlock * 3 : 0);                           // [256][3 (K - 1)]
  float* s_gauss = s_rest + (K3 > 0 ? kProjBlock * (K3 - 3) : 0);                 // [256][12]
  float* s_cam = s_gauss + kProjBlock * kProjGaussFloats;                         // [8][24]
  int* s_run = reinterpret_cast<int*>(s_cam + kProjCamsPerRound * kProjCamFloats);  // [8 warps][32]
  int* s_count = s_run + 8 * 32;                                                  // [4]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_count + 4);                     // SH staging barrier
  uint16_t* s_queue = reinterpret_cast<uint16_t*>(s_bar + 2);                     // [256 * 8]

  const int N = p.d.n_gauss;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * kProjBlock;
  const int n = n0 + threadIdx.x;
  const bool in_range = n < N;
  const int per = K3 - 3;

  // ---- stage the block's SH coefficients (contiguous in global memory) with two bulk copies that
  // complete on an mbarrier while phase A runs; odd tails / unaligned bases use plain loads
  const int cnt = min(kProjBlock, N - n0);
  const bool bulk = K3 > 0 && sh_bulk_ok && (cnt & 3) == 0;
  if (K3 > 0 && bulk) {
    if (threadIdx.x == 0) {
      mbar_init(s_bar, 1);
      mbar_fence_init();
      mbar_expect_tx(s_bar, (unsigned)(cnt * K3 * sizeof(float)));
      bulk_g2s(s_dc, p.fdc + (size_t)n0 * 3, (unsigned)(cnt * 3 * sizeof(float)), s_bar);
      if (per > 0) bulk_g2s(s_rest, p.frest + (size_t)n0 * per, (unsigned)(cnt * per * sizeof(float)), s_bar);
    }
  } else if (K3 > 0) {
    const float* fdc = p.fdc + (size_t)n0 * 3;
    for (int i = threadIdx.x; i < cnt * 3; i += kProjBlock) s_dc[i] = __ldg(fdc + i);
    if (per > 0) {
      const float* fr = p.frest + (size_t)n0 * per;
      for (int i = threadIdx.x; i < cnt * per; i += kProjBlock) s_rest[i] = __ldg(fr + i);
    }
  }
  // ---- per-Gaussian parameters
  float mu[3] = {0.f, 0.f, 0.f}, smax = 0.f;
  if (in_range) {
    mu[0] = p.means[3 * (size_t)n]; mu[1] = p.means[3 * (size_t)n + 1]; mu[2] = p.means[3 * (size_t)n + 2];
    float q[4] = {p.quats[4 * (size_t)n], p.quats[4 * (size_t)n + 1], p.quats[4 * (size_t)n + 2], p.quats[4 * (size_t)n + 3]};
    float s[3] = {p.scales[3 * (size_t)n], p.scales[3 * (size_t)n + 1], p.scales[3 * (size_t)n + 2]};
    if (p.d.raw_params) { s[0] = __expf(s[0]); s[1] = __expf(s[1]); s[2] = __expf(s[2]); }
    float Rq[9], cov[6];
    quat_to_rotmat(q, Rq);
    covar_world(Rq, s, cov);
    smax = fmaxf(s[0], fmaxf(s[1], s[2]));
    float op_base = p.opacities[n];
    if (p.d.raw_params) op_base = sigmoidf(op_base);
    float4* sg = reinterpret_cast<float4*>(s_gauss + threadIdx.x * kProjGaussFloats);
    sg[0] = make_float4(mu[0], mu[1], mu[2], cov[0]);
    sg[1] = make_float4(cov[1], cov[2], cov[3], cov[4]);
    sg[2] = make_float4(cov[5], smax, op_base, 0.f);
  }

  for (int c0 = 0; c0 < p.d.n_cams; c0 += kProjCamsPerRound) {
    const int nc = min(kProjCamsPerRound, p.d.n_cams - c0);
    __syncthreads();  // previous round's queue and cameras are consumed; (first round) staging is visible
    if (threadIdx.x < nc) {
      CamIntr cam;
      load_cam(p.viewmats, p.Ks, c0 + threadIdx.x, cam);
      float* sc = s_cam + threadIdx.x * kProjCamFloats;
#pragma unroll
      for (int i = 0; i < 9; ++i) sc[i] = cam.R[i];
      sc[9] = cam.t[0]; sc[10] = cam.t[1]; sc[11] = cam.t[2];
      sc[12] = cam.fx; sc[13] = cam.fy; sc[14] = cam.cx; sc[15] = cam.cy;
      // camera position = -R^T t  (vanilla.py:384-385)
      sc[16] = -(cam.R[0] * cam.t[0] + cam.R[3] * cam.t[1] + cam.R[6] * cam.t[2]);
      sc[17] = -(cam.R[1] * cam.t[0] + cam.R[4] * cam.t[1] + cam.R[7] * cam.t[2]);
      sc[18] = -(cam.R[2] * cam.t[0] + cam.R[5] * cam.t[1] + cam.R[8] * cam.t[2]);
    }
    if (threadIdx.x == 0) s_count[0] = 0;
    __syncthreads();
    // ---- phase A: cheap cull, defaults for the culled pairs, queue for the rest
    for (int ci = 0; ci < nc; ++ci) {
      const int c = c0 + ci;
      int ty0, ty1;
      band_rows(p.d, p.tile_h, c, ty0, ty1);
      if (ty1 <= ty0) continue;  // camera outside the band (uniform)
      const float* sc = s_cam + ci * kProjCamFloats;
      bool keep = false;
      if (in_range) {
        const float x = sc[0] * mu[0] + sc[1] * mu[1] + sc[2] * mu[2] + sc[9];
        const float y = sc[3] * mu[0] + sc[4] * mu[1] + sc[5] * mu[2] + sc[10];
        const float z = sc[6] * mu[0] + sc[7] * mu[1] + sc[8] * mu[2] + sc[11];
        keep = !(z < p.d.near_plane || z > p.d.far_plane) &&
               !screen_cull(x, y, z, smax, sc[12], sc[13], sc[14], sc[15], p.d.width, p.d.height, p.d.eps2d);
        if (!keep) {
          const int64_t idx = (int64_t)c * N + n;
          if (p.radii) p.radii[idx] = 0;
          p.tiles_touched[idx] = 0;
          if (p.slot_of) p.slot_of[idx] = -1;
        }
      }
      const unsigned km = __ballot_sync(0xffffffffu, keep);
      if (km) {
        int base = 0;
        const int leader = __ffs(km) - 1;
        if (lane == leader) base = atomicAdd(&s_count[0], __popc(km));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (keep) s_queue[base + __popc(km & ((1u << lane) - 1u))] = (uint16_t)((ci << 8) | threadIdx.x);
      }
    }
    __syncthreads();
    if (bulk && c0 == 0) mbar_wait(s_bar, 0);  // SH coefficients have landed (waiting again is harmless)
    // ---- phase B: one queue entry per thread
    const int n_queue = s_count[0];
    for (int q0 = 0; q0 < n_queue; q0 += kProjBlock) {
      const int qi = q0 + threadIdx.x;
      const bool active = qi < n_queue;
      const int ent = active ? (int)s_queue[qi] : 0;
      const int ci = ent >> 8, ln = ent & 255;
      const int c = c0 + ci, gn = n0 + ln;
      const int64_t idx = (int64_t)c * N + gn;
      int ty0, ty1;
      band_rows(p.d, p.tile_h, c, ty0, ty1);
      int n_tiles = 0, radius_i = 0;
      bool cand = false;
      float mx = 0.f, my = 0.f, qa = 0.f, qb = 0.f, qc = 0.f, op = 0.f, sigma_cut = 0.f, lop = 0.f, depth = 0.f;
      TileRect tr = {0, 0, 0, 0};
      if (active) {
        const float4* sg = reinterpret_cast<const float4*>(s_gauss + ln * kProjGaussFloats);
        const float4 g0 = sg[0], g1 = sg[1], g2 = sg[2];
        const float gmu[3] = {g0.x, g0.y, g0.z};
        const float cov[6] = {g0.w, g1.x, g1.y, g1.z, g1.w, g2.x};
        const float* sc = s_cam + ci * kProjCamFloats;
        CamIntr cam;
#pragma unroll
        for (int i = 0; i < 9; ++i) cam.R[i] = sc[i];
        cam.t[0] = sc[9]; cam.t[1] = sc[10]; cam.t[2] = sc[11];
        cam.fx = sc[12]; cam.fy = sc[13]; cam.cx = sc[14]; cam.cy = sc[15];
        Proj o;
        if (project_gaussian_cov(gmu, cov, g2.y, cam, p.d.width, p.d.height, p.d.eps2d, p.d.near_plane, p.d.far_plane,
                                 p.d.radius_clip, o)) {
          radius_i = (int)o.radius;
          if (p.means2d) { p.means2d[2 * idx] = o.mx; p.means2d[2 * idx + 1] = o.my; }
          if (p.depths) p.depths[idx] = o.z;
          if (p.conics) { p.conics[3 * idx] = o.a; p.conics[3 * idx + 1] = o.b; p.conics[3 * idx + 2] = o.c; }
          if (p.compensations) p.compensations[idx] = o.comp;
          op = g2.z;
          if (p.d.antialiased) op *= o.comp;
          mx = o.mx; my = o.my; depth = o.z;
          qa = 0.5f * kLog2e * o.a; qb = kLog2e * o.b; qc = 0.5f * kLog2e * o.c;
          lop = __log2f(op);
          sigma_cut = lop + kLog2_255;  // alpha >= 1/255  <=>  sigma' <= log2(255 o)
          if (op >= kAlphaMin) {
            tr = candidate_rect(mx, my, o.radius, qa, qb, qc, sigma_cut, p.tile_w, p.tile_h, ty0, ty1);
            cand = (tr.x1 > tr.x0) && (tr.y1 > tr.y0);
          }
        }
      }
      // ---- exact tile count.  The candidates of the warp's splats form one flat list that the 32 lanes
      // test 32 at a time (splat footprints differ by orders of magnitude: per-lane loops would idle).
      // A splat with more than kBigCand candidate tiles goes to the big-splat queue (big_splats.cuh) instead:
      // its tiles are counted by a follow-up launch, one CTA per splat.
      int qpos = -1;
      bool deferred = false;
      {
        int ncand = cand ? (tr.x1 - tr.x0) * (tr.y1 - tr.y0) : 0;
        if (ncand > kBigCand && p.radii) {   // the follow-up launch rebuilds the candidate rectangle from radii
          qpos = atomicAdd(p.counters + 2, 1);
          deferred = qpos < kBigQueueCap;
          if (deferred) ncand = 0;
        }
        const int incl = warp_inclusive_scan_i32(ncand);
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const int excl = incl - ncand;
        const int rw = tr.x1 - tr.x0;
        int* run = s_run + (threadIdx.x >> 5) * 32;
        run[lane] = 0;
        __syncwarp();
        for (int base = 0; base < total; base += 32) {
          const int wi = min(base + lane, total - 1);
          const int owner = warp_find_owner(excl, wi);
          const int local = wi - __shfl_sync(0xffffffffu, excl, owner);
          const int ow = __shfl_sync(0xffffffffu, rw, owner);
          const int ox0 = __shfl_sync(0xffffffffu, tr.x0, owner), oy0 = __shfl_sync(0xffffffffu, tr.y0, owner);
          const float gx = __shfl_sync(0xffffffffu, mx, owner), gy = __shfl_sync(0xffffffffu, my, owner);
          const float ga = __shfl_sync(0xffffffffu, qa, owner), gb = __shfl_sync(0xffffffffu, qb, owner);
          const float gc = __shfl_sync(0xffffffffu, qc, owner), gcut = __shfl_sync(0xffffffffu, sigma_cut, owner);
          const int gcam = __shfl_sync(0xffffffffu, c, owner);
          bool hit = false;
          if (base + lane < total) {
            const int ry = local / ow;
            const int tx = ox0 + local - ry * ow, ty = oy0 + ry;
            hit = tile_hit(gx, gy, ga, gb, gc, gcut, tx, ty, p.d.width, p.d.height);
            if (hit && p.tile_counts) atomicAdd(p.tile_counts + ((gcam * p.tile_h + ty - p.d.row_begin) * p.tile_w + tx), 1);
          }
          const unsigned hm = __ballot_sync(0xffffffffu, hit);
          const unsigned grp = __match_any_sync(0xffffffffu, owner);
          if (lane == __ffs(grp) - 1) run[owner] += __popc(hm & grp);  // one leader per owner: no atomics
          __syncwarp();
        }
        n_tiles = run[lane];
        __syncwarp();
      }
      // ---- packed record (SH colour only for splats that reach some tile)
      const bool emit = n_tiles > 0 || deferred;   // a footprint that large reaches some tile
      float rec[12];
      if (emit) {
        rec[0] = mx; rec[1] = my; rec[2] = qa; rec[3] = qb; rec[4] = qc; rec[5] = op;
        if (p.d.sh_degree >= 0) {
          const float* sg = s_gauss + ln * kProjGaussFloats;
          const float* sc = s_cam + ci * kProjCamFloats;
          float dx = sg[0] - sc[16], dy = sg[1] - sc[17], dz = sg[2] - sc[18];  // mean - camera position
          float inv = rsqrtf(dx * dx + dy * dy + dz * dz);
          float b[16];
          sh_basis(p.d.sh_degree, dx * inv, dy * inv, dz * inv, b);
          int nb = (p.d.sh_degree + 1) * (p.d.sh_degree + 1);
          if (nb > p.d.sh_K) nb = p.d.sh_K;
          const float* dc = s_dc + ln * 3;
          const float* rest = s_rest + ln * per - 3;  // rest[k * 3 + ch], k >= 1
          float col[3];
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) col[ch] = b[0] * dc[ch];
          for (int k = 1; k < nb; ++k) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) col[ch] = fmaf(b[k], rest[k * 3 + ch], col[ch]);
          }
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) rec[6 + ch] = fminf(fmaxf(col[ch] + 0.5f, 0.f), 1.f);
        } else {
          const float* cp = p.colors + (p.colors_per_cam ? 3 * (size_t)idx : 3 * (size_t)gn);
          rec[6] = cp[0]; rec[7] = cp[1]; rec[8] = cp[2];
        }
        rec[9] = depth;
        rec[10] = __int_as_float((int)idx);
        rec[11] = lop;  // log2(opacity): alpha = exp2(rec[11] - sigma')
      }
      // warp-aggregated compaction: one atomic per warp
      unsigned ballot = __ballot_sync(0xffffffffu, emit);
      int slot = -1;
      if (ballot) {
        int leader = __ffs(ballot) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(p.counters, __popc(ballot));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (emit) {
          slot = base + __popc(ballot & ((1u << lane) - 1u));
          if (slot >= p.splat_cap) {  // capacity overflow: flag, drop the record
            p.counters[1] = 1;
            slot = -1;
            n_tiles = 0;
          }
        }
      }
      if (deferred) p.counters[4 + qpos] = slot;   // -1 when the record was dropped
      if (active) {
        if (p.radii) p.radii[idx] = radius_i;
        p.tiles_touched[idx] = n_tiles;             // deferred: 0 here, written by big_splat_kernel
        if (p.slot_of) p.slot_of[idx] = slot;
        if (slot >= 0) {
          float4* dst = reinterpret_cast<float4*>(p.splats + (size_t)slot * 12);
          dst[0] = make_float4(rec[0], rec[1], rec[2], rec[3]);
          dst[1] = make_float4(rec[4], rec[5], rec[6], rec[7]);
          dst[2] = make_float4(rec[8], rec[9], rec[10], rec[11]);
        }
      }
    }
  }
}

struct ProjBwdParams {
  bds_render_desc d;
  const float *means, *quats, *scales, *opacities, *colors, *fdc, *frest, *viewmats, *Ks;
  int colors_per_cam;
  const float* splats;
  const int32_t* counters;
  const float* v_splats;
  const float *v_means2d_extra, *v_depths_extra, *v_conics_extra;
  float *v_means, *v_quats, *v_scales, *v_opacities, *v_colors, *v_fdc, *v_frest, *v_viewmats, *v_means2d, *absgrad;
  float* v_sh_color;   // optional [C,N,3]: the SH colour cotangent per (camera, Gaussian) instead of v_fdc / v_frest
};

// Gradient row of the 15 higher SH coefficients of one Gaussian (45 floats, value j = b[1 + j/3] * vcol[j % 3]):
// HEAD scalar reductions up to 16-byte alignment, then 128-bit vector reductions, then the scalar tail -
// 13-14 requests instead of 45.
template <int HEAD>
BDS_D void sh_rest_red(float* fr, const float* b, const float* vcol) {
#define BDS_SHV(j) (b[1 + (j) / 3] * vcol[(j) % 3])
#pragma unroll
  for (int j = 0; j < HEAD; ++j) red_add(fr + j, BDS_SHV(j));
  constexpr int NV = (45 - HEAD) / 4;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    const int j = HEAD + 4 * q;
    red_add_v4(fr + j, BDS_SHV(j), BDS_SHV(j + 1), BDS_SHV(j + 2), BDS_SHV(j + 3));
  }
#pragma unroll
  for (int j = HEAD + 4 * NV; j < 45; ++j) red_add(fr + j, BDS_SHV(j));
#undef BDS_SHV
}

__global__ void __launch_bounds__(256) project_bwd_kernel(ProjBwdParams p) {
  int n_slots = p.counters[0];
  int slot = blockIdx.x * blockDim.x + threadIdx.x;
  bool active = slot < n_slots;
  float vview[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) vview[k] = 0.f;
  int c = 0;
  if (active) {
    const float4* rp = reinterpret_cast<const float4*>(p.splats + (size_t)slot * 12);
    float4 r0 = rp[0], r1 = rp[1], r2 = rp[2];
    const float4* vp = reinterpret_cast<const float4*>(p.v_splats + (size_t)slot * 12);
    float4 g0 = vp[0], g1 = vp[1], g2 = vp[2];
    int64_t idx = (int64_t)__float_as_int(r2.z);
    const int N = p.d.n_gauss;
    c = (int)(idx / N);
    int n = (int)(idx - (int64_t)c * N);
    // v_splats = {m_x, m_y, m_xx, m_xy, m_yy, m_0, v_r, v_g, v_b, v_depth, sum|w g_x|, sum|w g_y|}: moments of
    // w = araw * v_alpha over the pixels (composite.cu, flush_batch).  With sigma' = a' dx^2 + b' dx dy + c' dy^2
    // and v_sigma' = -ln2 * w per pixel:  v_mean2d = -ln2 (2a' m_x + b' m_y, b' m_x + 2c' m_y),
    // v_(a',b',c') = -ln2 (m_xx, m_xy, m_yy) and (a', b', c') = log2e (a/2, b, c/2)  =>  v_conic = (-m_xx/2, -m_xy, -m_yy/2);
    // v_opacity = m_0 / opacity (alpha = opacity * vis).
    float vmx = -kLn2 * fmaf(2.f * r0.z, g0.x, r0.w * g0.y);
    float vmy = -kLn2 * fmaf(r0.w, g0.x, 2.f * r1.x * g0.y);
    float va = -0.5f * g0.z, vb = -g0.w, vc = -0.5f * g1.x;
    float vop = r1.y > 0.f ? g1.y / r1.y : 0.f;
    float vr = g1.z, vg = g1.w, vbl = g2.x, vz = g2.y;
    if (p.v_means2d) { p.v_means2d[2 * idx] = vmx; p.v_means2d[2 * idx + 1] = vmy; }
    if (p.absgrad) { p.absgrad[2 * idx] = kLn2 * g2.z; p.absgrad[2 * idx + 1] = kLn2 * g2.w; }
    if (p.v_means2d_extra) { vmx += p.v_means2d_extra[2 * idx]; vmy += p.v_means2d_extra[2 * idx + 1]; }
    if (p.v_depths_extra) vz += p.v_depths_extra[idx];
    if (p.v_conics_extra) { va += p.v_conics_extra[3 * idx]; vb += p.v_conics_extra[3 * idx + 1]; vc += p.v_conics_extra[3 * idx + 2]; }
    CamIntr cam;
    load_cam(p.viewmats, p.Ks, c, cam);
    float mu[3] = {p.means[3 * (size_t)n], p.means[3 * (size_t)n + 1], p.means[3 * (size_t)n + 2]};
    float q[4] = {p.quats[4 * (size_t)n], p.quats[4 * (size_t)n + 1], p.quats[4 * (size_t)n + 2], p.quats[4 * (size_t)n + 3]};
    float sraw[3] = {p.scales[3 * (size_t)n], p.scales[3 * (size_t)n + 1], p.scales[3 * (size_t)n + 2]};
    float s[3] = {sraw[0], sraw[1], sraw[2]};
    if (p.d.raw_params) { s[0] = __expf(s[0]); s[1] = __expf(s[1]); s[2] = __expf(s[2]); }
    // opacity (+ antialias compensation)
    float op_in = p.opacities[n];
    float op = p.d.raw_params ? sigmoidf(op_in) : op_in;
    float vcomp = 0.f;
    if (p.d.antialiased) {
      float comp = r1.y / fmaxf(op, 1e-30f);  // record opacity = op * comp
      vcomp = vop * op;
      vop = vop * comp;
    }
    if (p.d.raw_params) vop *= op * (1.f - op);
    red_add(p.v_opacities + n, vop);
    // colours
    if (p.d.sh_degree >= 0 && p.v_sh_color) {
      // compact form (multi-GPU exchange): the SH gradient of a (camera, Gaussian) pair is the outer product of the
      // basis values of its view direction and this 3-vector; bds_sh_expand_bwd rebuilds the 48 coefficients
      float* dst = p.v_sh_color + 3 * (size_t)idx;
      dst[0] = (r1.z > 0.f && r1.z < 1.f) ? vr : 0.f;
      dst[1] = (r1.w > 0.f && r1.w < 1.f) ? vg : 0.f;
      dst[2] = (r2.x > 0.f && r2.x < 1.f) ? vbl : 0.f;
    } else if (p.d.sh_degree >= 0) {
      float cp[3] = {-(cam.R[0] * cam.t[0] + cam.R[3] * cam.t[1] + cam.R[6] * cam.t[2]),
                     -(cam.R[1] * cam.t[0] + cam.R[4] * cam.t[1] + cam.R[7] * cam.t[2]),
                     -(cam.R[2] * cam.t[0] + cam.R[5] * cam.t[1] + cam.R[8] * cam.t[2])};
      float dx = mu[0] - cp[0], dy = mu[1] - cp[1], dz = mu[2] - cp[2];
      float inv = rsqrtf(dx * dx + dy * dy + dz * dz);
      float b[16];
      sh_basis(p.d.sh_degree, dx * inv, dy * inv, dz * inv, b);
      int nb = (p.d.sh_degree + 1) * (p.d.sh_degree + 1);
      if (nb > p.d.sh_K) nb = p.d.sh_K;
      // clamp(+0.5, 0, 1) passes gradient strictly inside (torch.clamp: also at the bounds; measure zero)
      float vcol[3] = {(r1.z > 0.f && r1.z < 1.f) ? vr : 0.f, (r1.w > 0.f && r1.w < 1.f) ? vg : 0.f,
                       (r2.x > 0.f && r2.x < 1.f) ? vbl : 0.f};
      if (vcol[0] != 0.f || vcol[1] != 0.f || vcol[2] != 0.f) {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) red_add(p.v_fdc + 3 * (size_t)n + ch, b[0] * vcol[ch]);
        float* fr = p.v_frest + (size_t)n * (p.d.sh_K - 1) * 3;
        if (nb == 16 && p.d.sh_K == 16) {  // degree 3: the 45-float row leaves as 128-bit reductions
          switch ((int)(((16u - (unsigned)((uintptr_t)fr & 15u)) & 15u) >> 2)) {
            case 0: sh_rest_red<0>(fr, b, vcol); break;
            case 1: sh_rest_red<1>(fr, b, vcol); break;
            case 2: sh_rest_red<2>(fr, b, vcol); break;
            default: sh_rest_red<3>(fr, b, vcol); break;
          }
        } else
        for (int k = 1; k < nb; ++k) {
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) red_add(fr + (k - 1) * 3 + ch, b[k] * vcol[ch]);
        }
      }
    } else if (p.v_colors) {
      float* cp = p.v_colors + (p.colors_per_cam ? 3 * (size_t)idx : 3 * (size_t)n);
      red_add(cp, vr); red_add(cp + 1, vg); red_add(cp + 2, vbl);
    }
    // geometry
    float v_mu[3] = {0.f, 0.f, 0.f}, v_q[4] = {0.f, 0.f, 0.f, 0.f}, v_s[3] = {0.f, 0.f, 0.f}, v_R[9], v_t[3];
    project_gaussian_vjp(mu, q, s, cam, p.d.width, p.d.height, p.d.eps2d, vmx, vmy, vz, va, vb, vc, vcomp, v_mu, v_q,
                         v_s, v_R, v_t);
    if (p.d.raw_params) { v_s[0] *= s[0]; v_s[1] *= s[1]; v_s[2] *= s[2]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      red_add(p.v_means + 3 * (size_t)n + k, v_mu[k]);
      red_add(p.v_scales + 3 * (size_t)n + k, v_s[k]);
    }
    if (((uintptr_t)p.v_quats & 15u) == 0) {
      red_add_v4(p.v_quats + 4 * (size_t)n, v_q[0], v_q[1], v_q[2], v_q[3]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) red_add(p.v_quats + 4 * (size_t)n + k, v_q[k]);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      vview[i * 4] = v_R[i * 3]; vview[i * 4 + 1] = v_R[i * 3 + 1]; vview[i * 4 + 2] = v_R[i * 3 + 2];
      vview[i * 4 + 3] = v_t[i];
    }
  }
  if (p.v_viewmats) {
    // slots are not grouped by camera: reduce per camera over the warp, then one atomic per value
    for (int cam_i = 0; cam_i < p.d.n_cams; ++cam_i) {
      bool mine = active && c == cam_i;
      if (!__any_sync(0xffffffffu, mine)) continue;
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        float v = warp_sum(mine ? vview[k] : 0.f);
        if ((threadIdx.x & 31) == 0 && v != 0.f) red_add(p.v_viewmats + 16 * cam_i + k, v);
      }
    }
  }
}

// Cotangents a user of the gsplat info tensors puts on means2d / depths / conics (v_*_extra) for Gaussians that gsplat
// calls visible (radii > 0) but that own NO splat record here (their alpha >= 1/255 footprint misses every tile of the
// band, so the exact binning never emitted them): project_bwd_kernel walks the records and cannot see them.
__global__ void __launch_bounds__(256) project_bwd_extras_kernel(ProjBwdParams p, const int32_t* __restrict__ radii,
                                                                 const int32_t* __restrict__ slot_of) {
  const int64_t total = (int64_t)p.d.n_gauss * p.d.n_cams;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = idx < total && radii[idx] > 0 && slot_of[idx] < 0;
  float vview[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) vview[k] = 0.f;
  const int N = p.d.n_gauss;
  const int c = idx < total ? (int)(idx / N) : 0;
  if (active) {
    const int n = (int)(idx - (int64_t)c * N);
    float vmx = 0.f, vmy = 0.f, vz = 0.f, va = 0.f, vb = 0.f, vc = 0.f;
    if (p.v_means2d_extra) { vmx = p.v_means2d_extra[2 * idx]; vmy = p.v_means2d_extra[2 * idx + 1]; }
    if (p.v_depths_extra) vz = p.v_depths_extra[idx];
    if (p.v_conics_extra) { va = p.v_conics_extra[3 * idx]; vb = p.v_conics_extra[3 * idx + 1]; vc = p.v_conics_extra[3 * idx + 2]; }
    CamIntr cam;
    load_cam(p.viewmats, p.Ks, c, cam);
    float mu[3] = {p.means[3 * (size_t)n], p.means[3 * (size_t)n + 1], p.means[3 * (size_t)n + 2]};
    float q[4] = {p.quats[4 * (size_t)n], p.quats[4 * (size_t)n + 1], p.quats[4 * (size_t)n + 2], p.quats[4 * (size_t)n + 3]};
    float s[3] = {p.scales[3 * (size_t)n], p.scales[3 * (size_t)n + 1], p.scales[3 * (size_t)n + 2]};
    if (p.d.raw_params) { s[0] = __expf(s[0]); s[1] = __expf(s[1]); s[2] = __expf(s[2]); }
    float v_mu[3] = {0.f, 0.f, 0.f}, v_q[4] = {0.f, 0.f, 0.f, 0.f}, v_s[3] = {0.f, 0.f, 0.f}, v_R[9], v_t[3];
    project_gaussian_vjp(mu, q, s, cam, p.d.width, p.d.height, p.d.eps2d, vmx, vmy, vz, va, vb, vc, 0.f, v_mu, v_q, v_s,
                         v_R, v_t);
    if (p.d.raw_params) { v_s[0] *= s[0]; v_s[1] *= s[1]; v_s[2] *= s[2]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      red_add(p.v_means + 3 * (size_t)n + k, v_mu[k]);
      red_add(p.v_scales + 3 * (size_t)n + k, v_s[k]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) red_add(p.v_quats + 4 * (size_t)n + k, v_q[k]);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      vview[i * 4] = v_R[i * 3]; vview[i * 4 + 1] = v_R[i * 3 + 1]; vview[i * 4 + 2] = v_R[i * 3 + 2];
      vview[i * 4 + 3] = v_t[i];
    }
  }
  if (p.v_viewmats && __any_sync(0xffffffffu, active)) {
    // consecutive idx share the camera except where a warp straddles a camera boundary
    for (int cam_i = __shfl_sync(0xffffffffu, c, 0); cam_i <= __shfl_sync(0xffffffffu, c, 31); ++cam_i) {
      const bool mine = active && c == cam_i;
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        float v = warp_sum(mine ? vview[k] : 0.f);
        if ((threadIdx.x & 31) == 0 && v != 0.f) red_add(p.v_viewmats + 16 * cam_i + k, v);
      }
    }
  }
}

// stand-alone SH ---------------------------------------------------------------------------------
__global__ void sh_fwd_kernel(int n, int degree, int K, const float* __restrict__ dirs,
                              const float* __restrict__ coeffs, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x = dirs[3 * (size_t)i], y = dirs[3 * (size_t)i + 1], z = dirs[3 * (size_t)i + 2];
  float inv = rsqrtf(x * x + y * y + z * z);
  float b[16];
  sh_basis(degree, x * inv, y * inv, z * inv, b);
  int nb = (degree + 1) * (degree + 1);
  if (nb > K) nb = K;
  const float* cf = coeffs + (size_t)i * K * 3;
  float c0 = 0.f, c1 = 0.f, c2 = 0.f;
  for (int k = 0; k < nb; ++k) {
    c0 = fmaf(b[k], cf[3 * k], c0);
    c1 = fmaf(b[k], cf[3 * k + 1], c1);
    c2 = fmaf(b[k], cf[3 * k + 2], c2);
  }
  out[3 * (size_t)i] = c0; out[3 * (size_t)i + 1] = c1; out[3 * (size_t)i + 2] = c2;
}

__global__ void sh_bwd_kernel(int n, int degree, int K, const float* __restrict__ dirs,
                              const float* __restrict__ coeffs, const float* __restrict__ v_out,
                              float* __restrict__ v_coeffs, float* __restrict__ v_dirs) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x = dirs[3 * (size_t)i], y = dirs[3 * (size_t)i + 1], z = dirs[3 * (size_t)i + 2];
  float inv = rsqrtf(x * x + y * y + z * z);
  float nx = x * inv, ny = y * inv, nz = z * inv;
  float b[16];
  sh_basis(degree, nx, ny, nz, b);
  int nb = (degree + 1) * (degree + 1);
  if (nb > K) nb = K;
  float g0 = v_out[3 * (size_t)i], g1 = v_out[3 * (size_t)i + 1], g2 = v_out[3 * (size_t)i + 2];
  float* vc = v_coeffs + (size_t)i * K * 3;
  for (int k = 0; k < K; ++k) {
    float bk = k < nb ? b[k] : 0.f;
    vc[3 * k] = bk * g0; vc[3 * k + 1] = bk * g1; vc[3 * k + 2] = bk * g2;
  }
  if (v_dirs) {
    float dbx[16], dby[16], dbz[16];
    sh_basis_grad(degree, nx, ny, nz, dbx, dby, dbz);
    const float* cf = coeffs + (size_t)i * K * 3;
    float vx = 0.f, vy = 0.f, vz = 0.f;
    for (int k = 0; k < nb; ++k) {
      float w = cf[3 * k] * g0 + cf[3 * k + 1] * g1 + cf[3 * k + 2] * g2;
      vx = fmaf(dbx[k], w, vx); vy = fmaf(dby[k], w, vy); vz = fmaf(dbz[k], w, vz);
    }
    float dot = nx * vx + ny * vy + nz * vz;  // through the normalisation
    v_dirs[3 * (size_t)i] = (vx - nx * dot) * inv;
    v_dirs[3 * (size_t)i + 1] = (vy - ny * dot) * inv;
    v_dirs[3 * (size_t)i + 2] = (vz - nz * dot) * inv;
  }
}

int check_render_desc(const bds_render_desc* d) {
  BDS_REQUIRE(d, "render desc is null");
  BDS_REQUIRE(d->n_gauss >= 0 && d->n_cams >= 1, "render desc: bad n_gauss/n_cams");
  BDS_REQUIRE(d->width >= 1 && d->height >= 1, "render desc: empty image");
  BDS_REQUIRE((int64_t)d->n_gauss * d->n_cams < (int64_t)1 << 31, "render desc: n_gauss*n_cams must fit int32");
  int tile_h = (d->height + kTile - 1) / kTile;
  BDS_REQUIRE(d->row_begin >= 0 && d->row_end >= d->row_begin && d->row_end <= d->n_cams * tile_h,
              "render desc: band [%d,%d) outside [0,%d]", d->row_begin, d->row_end, d->n_cams * tile_h);
  BDS_REQUIRE(d->sh_degree <= 3, "render desc: sh_degree <= 3 supported");
  if (d->sh_degree >= 0) BDS_REQUIRE(d->sh_K >= 1 && d->sh_K <= 16, "render desc: sh_K out of range");
  return 0;
}

}  // namespace bds

using namespace bds;

extern "C" int bds_project_fwd(const bds_render_desc* d, const float* means, const float* quats, const float* scales,
                               const float* opacities, const float* colors, int colors_per_cam,
                               const float* features_dc, const float* features_rest, const float* viewmats,
                               const float* Ks, int32_t* radii, float* means2d, float* depths, float* conics,
                               float* compensations, int32_t* tiles_touched, int32_t* tile_counts, float* splats,
                               int32_t splat_cap, int32_t* slot_of, int32_t* counters, bds_stream_t stream) {
  if (int rc = check_render_desc(d)) return rc;
  int64_t total = (int64_t)d->n_gauss * d->n_cams;
  if (tile_counts)
    BDS_CHECK_CUDA(cudaMemsetAsync(tile_counts, 0,
                                   ((size_t)(d->row_end - d->row_begin) * ((d->width + kTile - 1) / kTile) + 1) * sizeof(int32_t),
                                   static_cast<cudaStream_t>(stream)));
  if (total == 0) return 0;
  BDS_REQUIRE(means && quats && scales && opacities && viewmats && Ks && tiles_touched && splats && counters,
              "project_fwd: null pointer");
  if (d->sh_degree >= 0)
    BDS_REQUIRE(features_dc && (features_rest || d->sh_K == 1), "project_fwd: SH path needs features_dc/rest");
  else
    BDS_REQUIRE(colors, "project_fwd: colors required when sh_degree < 0");
  ProjParams p;
  p.d = *d;
  p.tile_w = (d->width + kTile - 1) / kTile;
  p.tile_h = (d->height + kTile - 1) / kTile;
  p.means = means; p.quats = quats; p.scales = scales; p.opacities = opacities; p.colors = colors;
  p.fdc = features_dc; p.frest = features_rest; p.viewmats = viewmats; p.Ks = Ks; p.colors_per_cam = colors_per_cam;
  p.radii = radii; p.means2d = means2d; p.depths = depths; p.conics = conics; p.compensations = compensations;
  p.tiles_touched = tiles_touched; p.tile_counts = tile_counts; p.splats = splats; p.splat_cap = splat_cap; p.slot_of = slot_of; p.counters = counters;
  // cameras outside the band are never visited: their tiles_touched must read 0 (the caller's buffer may be fresh)
  BDS_CHECK_CUDA(cudaMemsetAsync(tiles_touched, 0, (size_t)total * sizeof(int32_t), static_cast<cudaStream_t>(stream)));
  const int sh_floats = d->sh_degree >= 0 ? 3 * d->sh_K : 0;
  const size_t smem = proj_smem_bytes(sh_floats);
  BDS_CHECK_CUDA(cudaFuncSetAttribute(project_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int sh_bulk_ok = sh_floats > 0 && ((uintptr_t)features_dc % 16 == 0) &&
                         (sh_floats == 3 || (uintptr_t)features_rest % 16 == 0);
  project_fwd_kernel<<<ceil_div(d->n_gauss, kProjBlock), kProjBlock, smem, static_cast<cudaStream_t>(stream)>>>(
      p, sh_floats, sh_bulk_ok);
  BDS_CHECK_LAUNCH();
  if (radii) {  // tiles of the queued very large splats (a few microseconds when the queue is empty)
    BigSplatParams b{};
    b.d = *d; b.tile_w = p.tile_w; b.tile_h = p.tile_h; b.splats = splats; b.radii = radii;
    b.n_queue = counters + 2; b.queue = counters + 4; b.tile_counts = tile_counts; b.tiles_touched = tiles_touched;
    big_splat_kernel<false><<<4 * 148, 256, 0, static_cast<cudaStream_t>(stream)>>>(b);
    BDS_CHECK_LAUNCH();
  }
  return 0;
}

static int project_bwd_impl(const bds_render_desc* d, const float* means, const float* quats, const float* scales,
                            const float* opacities, const float* colors, int colors_per_cam,
                            const float* features_dc, const float* features_rest, const float* viewmats,
                            const float* Ks, const float* splats, const int32_t* counters, const float* v_splats,
                            const float* v_means2d_extra, const float* v_depths_extra, const float* v_conics_extra,
                            float* v_means, float* v_quats, float* v_scales, float* v_opacities, float* v_colors,
                            float* v_features_dc, float* v_features_rest, float* v_viewmats, float* v_means2d,
                            float* absgrad, float* v_sh_color, bds_stream_t stream) {
  if (int rc = check_render_desc(d)) return rc;
  int64_t total = (int64_t)d->n_gauss * d->n_cams;
  if (total == 0) return 0;
  BDS_REQUIRE(means && quats && scales && opacities && viewmats && Ks && splats && counters && v_splats,
              "project_bwd: null input pointer");
  BDS_REQUIRE(v_means && v_quats && v_scales && v_opacities, "project_bwd: null output pointer");
  if (d->sh_degree >= 0 && !v_sh_color)
    BDS_REQUIRE(v_features_dc && (v_features_rest || d->sh_K == 1), "project_bwd: SH grads missing");
  ProjBwdParams p;
  p.v_sh_color = d->sh_degree >= 0 ? v_sh_color : nullptr;
  p.d = *d;
  p.means = means; p.quats = quats; p.scales = scales; p.opacities = opacities; p.colors = colors;
  p.fdc = features_dc; p.frest = features_rest; p.viewmats = viewmats; p.Ks = Ks; p.colors_per_cam = colors_per_cam;
  p.splats = splats; p.counters = counters; p.v_splats = v_splats;
  p.v_means2d_extra = v_means2d_extra; p.v_depths_extra = v_depths_extra; p.v_conics_extra = v_conics_extra;
  p.v_means = v_means; p.v_quats = v_quats; p.v_scales = v_scales; p.v_opacities = v_opacities; p.v_colors = v_colors;
  p.v_fdc = v_features_dc; p.v_frest = v_features_rest; p.v_viewmats = v_viewmats; p.v_means2d = v_means2d;
  p.absgrad = absgrad;
  // the slot count lives on the device; launch over the worst case (every (cam, gauss) visible) and
  // let threads beyond counters[0] exit - the grid is cheap next to an extra host sync
  project_bwd_kernel<<<ceil_div(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  BDS_CHECK_LAUNCH();
  return 0;
}

extern "C" int bds_project_bwd(const bds_render_desc* d, const float* means, const float* quats, const float* scales,
                               const float* opacities, const float* colors, int colors_per_cam,
                               const float* features_dc, const float* features_rest, const float* viewmats,
                               const float* Ks, const float* splats, const int32_t* counters, const float* v_splats,
                               const float* v_means2d_extra, const float* v_depths_extra, const float* v_conics_extra,
                               float* v_means, float* v_quats, float* v_scales, float* v_opacities, float* v_colors,
                               float* v_features_dc, float* v_features_rest, float* v_viewmats, float* v_means2d,
                               float* absgrad, bds_stream_t stream) {
  return project_bwd_impl(d, means, quats, scales, opacities, colors, colors_per_cam, features_dc, features_rest, viewmats,
                          Ks, splats, counters, v_splats, v_means2d_extra, v_depths_extra, v_conics_extra, v_means, v_quats,
                          v_scales, v_opacities, v_colors, v_features_dc, v_features_rest, v_viewmats, v_means2d, absgrad,
                          nullptr, stream);
}

extern "C" int bds_project_bwd_compact_sh(const bds_render_desc* d, const float* means, const float* quats,
                                          const float* scales, const float* opacities, const float* viewmats,
                                          const float* Ks, const float* splats, const int32_t* counters,
                                          const float* v_splats, float* v_means, float* v_quats, float* v_scales,
                                          float* v_opacities, float* v_sh_color, float* v_viewmats, bds_stream_t stream) {
  BDS_REQUIRE(d && d->sh_degree >= 0 && v_sh_color, "project_bwd_compact_sh: needs the SH fast path and v_sh_color");
  return project_bwd_impl(d, means, quats, scales, opacities, nullptr, 0, nullptr, nullptr, viewmats, Ks, splats, counters,
                          v_splats, nullptr, nullptr, nullptr, v_means, v_quats, v_scales, v_opacities, nullptr, nullptr,
                          nullptr, v_viewmats, nullptr, nullptr, v_sh_color, stream);
}

// v_features_dc / v_features_rest of every Gaussian from the compact per-(camera, Gaussian) colour cotangents:
// sum over cameras of basis(view direction) (x) v_sh_color.  One thread per Gaussian builds the 3 K sums in registers;
// the block's rows are contiguous in v_features_rest, so they go through shared memory and leave as coalesced stores.
constexpr int kShExpBlock = 128;
__global__ void __launch_bounds__(kShExpBlock) sh_expand_bwd_kernel(bds_render_desc d, const float* __restrict__ means,
                                                                   const float* __restrict__ viewmats,
                                                                   const float* __restrict__ v_sh_color,
                                                                   float* __restrict__ v_fdc, float* __restrict__ v_frest) {
  __shared__ float s_out[kShExpBlock * 49];   // [thread][48] padded to 49 (conflict-free row writes)
  const int n0 = blockIdx.x * kShExpBlock;
  const int n = n0 + threadIdx.x;
  float acc[48];
#pragma unroll
  for (int k = 0; k < 48; ++k) acc[k] = 0.f;
  if (n < d.n_gauss) {
    const float mu[3] = {means[3 * (size_t)n], means[3 * (size_t)n + 1], means[3 * (size_t)n + 2]};
    for (int c = 0; c < d.n_cams; ++c) {
      const float* vc = v_sh_color + 3 * ((size_t)c * d.n_gauss + n);
      const float v0 = vc[0], v1 = vc[1], v2 = vc[2];
      if (v0 == 0.f && v1 == 0.f && v2 == 0.f) continue;
      const float* V = viewmats + 16 * c;   // camera position = -R^T t
      const float cx = -(V[0] * V[3] + V[4] * V[7] + V[8] * V[11]);
      const float cy = -(V[1] * V[3] + V[5] * V[7] + V[9] * V[11]);
      const float cz = -(V[2] * V[3] + V[6] * V[7] + V[10] * V[11]);
      const float dx = mu[0] - cx, dy = mu[1] - cy, dz = mu[2] - cz;
      const float inv = rsqrtf(dx * dx + dy * dy + dz * dz);
      float b[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) b[k] = 0.f;     // bands above the active degree contribute nothing
      sh_basis(d.sh_degree, dx * inv, dy * inv, dz * inv, b);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        acc[3 * k] = fmaf(b[k], v0, acc[3 * k]);
        acc[3 * k + 1] = fmaf(b[k], v1, acc[3 * k + 1]);
        acc[3 * k + 2] = fmaf(b[k], v2, acc[3 * k + 2]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 48; ++k) s_out[threadIdx.x * 49 + k] = acc[k];
  __syncthreads();
  const int rows = min(kShExpBlock, d.n_gauss - n0);
  const int K = d.sh_K;
  for (int i = threadIdx.x; i < rows * 3; i += kShExpBlock) v_fdc[(size_t)n0 * 3 + i] = s_out[(i / 3) * 49 + (i % 3)];
  const int per = (K - 1) * 3;   // floats of one row of v_features_rest
  for (int i = threadIdx.x; i < rows * per; i += kShExpBlock) {
    const int r = i / per, k = i - r * per;
    v_frest[(size_t)n0 * per + i] = s_out[r * 49 + 3 + k];
  }
}

extern "C" int bds_sh_expand_bwd(const bds_render_desc* d, const float* means, const float* viewmats,
                                 const float* v_sh_color, float* v_features_dc, float* v_features_rest,
                                 bds_stream_t stream) {
  if (int rc = check_render_desc(d)) return rc;
  if (d->n_gauss == 0) return 0;
  BDS_REQUIRE(d->sh_degree >= 0 && d->sh_K >= 1 && d->sh_K <= 16, "sh_expand_bwd: needs the SH fast path, K <= 16");
  BDS_REQUIRE(means && viewmats && v_sh_color && v_features_dc && (v_features_rest || d->sh_K == 1),
              "sh_expand_bwd: null pointer");
  sh_expand_bwd_kernel<<<ceil_div(d->n_gauss, kShExpBlock), kShExpBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      *d, means, viewmats, v_sh_color, v_features_dc, v_features_rest);
  BDS_CHECK_LAUNCH();
  return 0;
}

extern "C" int bds_project_bwd_extras(const bds_render_desc* d, const float* means, const float* quats,
                                      const float* scales, const float* viewmats, const float* Ks, const int32_t* radii,
                                      const int32_t* slot_of, const float* v_means2d_extra, const float* v_depths_extra,
                                      const float* v_conics_extra, float* v_means, float* v_quats, float* v_scales,
                                      float* v_viewmats, bds_stream_t stream) {
  if (int rc = check_render_desc(d)) return rc;
  int64_t total = (int64_t)d->n_gauss * d->n_cams;
  if (total == 0 || !(v_means2d_extra || v_depths_extra || v_conics_extra)) return 0;
  BDS_REQUIRE(means && quats && scales && viewmats && Ks && radii && slot_of, "project_bwd_extras: null input pointer");
  BDS_REQUIRE(v_means && v_quats && v_scales, "project_bwd_extras: null output pointer");
  ProjBwdParams p{};
  p.d = *d;
  p.means = means; p.quats = quats; p.scales = scales; p.viewmats = viewmats; p.Ks = Ks;
  p.v_means2d_extra = v_means2d_extra; p.v_depths_extra = v_depths_extra; p.v_conics_extra = v_conics_extra;
  p.v_means = v_means; p.v_quats = v_quats; p.v_scales = v_scales; p.v_viewmats = v_viewmats;
  project_bwd_extras_kernel<<<ceil_div(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, radii, slot_of);
  BDS_CHECK_LAUNCH();
  return 0;
}

extern "C" int bds_sh_fwd(int n, int degree, int K, const float* dirs, const float* coeffs, float* out,
                          bds_stream_t stream) {
  BDS_REQUIRE(n >= 0 && degree >= 0 && degree <= 3 && K >= 1, "sh_fwd: degree in [0,3], K >= 1");
  if (n == 0) return 0;
  BDS_REQUIRE(dirs && coeffs && out, "sh_fwd: null pointer");
  sh_fwd_kernel<<<ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(n, degree, K, dirs, coeffs, out);
  BDS_CHECK_LAUNCH();
  return 0;
}

extern "C" int bds_sh_bwd(int n, int degree, int K, const float* dirs, const float* coeffs, const float* v_out,
                          float* v_coeffs, float* v_dirs, bds_stream_t stream) {
  BDS_REQUIRE(n >= 0 && degree >= 0 && degree <= 3 && K >= 1, "sh_bwd: degree in [0,3], K >= 1");
  if (n == 0) return 0;
  BDS_REQUIRE(dirs && coeffs && v_out && v_coeffs, "sh_bwd: null pointer");
  sh_bwd_kernel<<<ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(n, degree, K, dirs, coeffs, v_out, v_coeffs,
                                                                                v_dirs);
  BDS_CHECK_LAUNCH();
  return 0;
}
