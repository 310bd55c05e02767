// TEST INFRASTRUCTURE: compiles the host/device math headers of the product with g++ so that the
// per-element arithmetic (projection + VJP, SH bases, exact ellipse-vs-rectangle bound, resize taps)
// is checked on the CPU against the oracle before the kernels run on a GPU.  Never loaded by the
// product package.
#include "../../bilateral_driving_b200/csrc/bilateral_math.cuh"
#include "../../bilateral_driving_b200/csrc/projection_math.cuh"
#include "../../bilateral_driving_b200/csrc/sh_math.cuh"

using namespace bds;

static CamIntr make_cam(const float* viewmat, const float* K) {
  CamIntr cam;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) cam.R[i * 3 + j] = viewmat[i * 4 + j];
    cam.t[i] = viewmat[i * 4 + 3];
  }
  cam.fx = K[0]; cam.fy = K[4]; cam.cx = K[2]; cam.cy = K[5];
  return cam;
}

extern "C" {

// out per gaussian: [mx, my, z, a, b, c, radius, comp]
void hc_project(int n, const float* means, const float* quats, const float* scales, const float* viewmat,
                const float* K, int width, int height, float eps2d, float near_plane, float far_plane,
                float radius_clip, float* out) {
  CamIntr cam = make_cam(viewmat, K);
  for (int i = 0; i < n; ++i) {
    Proj o;
    bool vis = project_gaussian(means + 3 * i, quats + 4 * i, scales + 3 * i, cam, width, height, eps2d, near_plane,
                                far_plane, radius_clip, o);
    float* q = out + 8 * i;
    q[0] = o.mx; q[1] = o.my; q[2] = o.z; q[3] = o.a; q[4] = o.b; q[5] = o.c; q[6] = vis ? o.radius : 0.f; q[7] = o.comp;
  }
}

// cot per gaussian: [vmx, vmy, vz, va, vb, vc, vcomp]; outputs v_means[n,3], v_quats[n,4], v_scales[n,3],
// v_view[12] (R row-major 9 + t 3), accumulated over gaussians
void hc_project_vjp(int n, const float* means, const float* quats, const float* scales, const float* viewmat,
                    const float* K, int width, int height, float eps2d, const float* cot, float* v_means,
                    float* v_quats, float* v_scales, float* v_view) {
  CamIntr cam = make_cam(viewmat, K);
  for (int k = 0; k < 12; ++k) v_view[k] = 0.f;
  for (int i = 0; i < n; ++i) {
    const float* c = cot + 7 * i;
    float v_mu[3] = {0, 0, 0}, v_q[4] = {0, 0, 0, 0}, v_s[3] = {0, 0, 0}, v_R[9], v_t[3];
    project_gaussian_vjp(means + 3 * i, quats + 4 * i, scales + 3 * i, cam, width, height, eps2d, c[0], c[1], c[2],
                         c[3], c[4], c[5], c[6], v_mu, v_q, v_s, v_R, v_t);
    for (int k = 0; k < 3; ++k) { v_means[3 * i + k] = v_mu[k]; v_scales[3 * i + k] = v_s[k]; }
    for (int k = 0; k < 4; ++k) v_quats[4 * i + k] = v_q[k];
    for (int k = 0; k < 9; ++k) v_view[k] += v_R[k];
    for (int k = 0; k < 3; ++k) v_view[9 + k] += v_t[k];
  }
}

float hc_min_sigma_rect(float gx, float gy, float qa, float qb, float qc, float xmin, float xmax, float ymin,
                        float ymax) {
  return min_sigma_rect(gx, gy, qa, qb, qc, xmin, xmax, ymin, ymax);
}

void hc_sh_basis(int degree, float x, float y, float z, float* b, float* dbx, float* dby, float* dbz) {
  sh_basis(degree, x, y, z, b);
  sh_basis_grad(degree, x, y, z, dbx, dby, dbz);
}

void hc_lin_src(int d, int in_size, int out_size, int* i0, int* i1, float* t) {
  LinTap tp = lin_src(d, in_size, out_size);
  *i0 = tp.i0; *i1 = tp.i1; *t = tp.t;
}
float hc_lattice_coord(int j, int n, int g) { return lattice_coord(j, n, g); }
float hc_unit_lin01(int j, int n, int g) { return unit_coord(lin01(j, n), g); }  // the hoisted form the kernels use
// value repack [GY][GX][3][L][4]: element -> parameter-layout index, and (node, z, row) -> float offset
long hc_value_param_index(int i, int L, int GY, int GX) { return (long)bil_value_param_index(i, L, GY, GX); }
int hc_value_offset(int node, int z, int k, int L) { return bil_value_offset(node, z, k, L); }
int hc_node(int x, int y, int z, int L, int GX) { return bil_node(x, y, z, L, GX); }

// tri set-up: returns nodes[8] and weights[8] of the trilinear stencil, and z_inside
int hc_tri(float fx, float fy, float fz, int L, int GY, int GX, int* nodes, float* w) {
  Tri t = tri_setup(fx, fy, fz, L, GY, GX);
  int nn[4] = {t.n00, t.n01, t.n10, t.n11};
  float wxy[4] = {(1 - t.wx1) * (1 - t.wy1), t.wx1 * (1 - t.wy1), (1 - t.wx1) * t.wy1, t.wx1 * t.wy1};
  for (int c = 0; c < 4; ++c) {
    nodes[c] = nn[c]; w[c] = wxy[c] * (1 - t.wz1);
    nodes[4 + c] = nn[c] + t.dz; w[4 + c] = wxy[c] * t.wz1;
  }
  return t.z_inside ? 1 : 0;
}
}

// candidate_rect must contain every tile that tile_hit accepts inside the gsplat square bound.
// Returns the number of hit tiles outside the candidate rectangle (must be 0) and writes
// stats[0] = tiles in the square bound, stats[1] = candidate tiles, stats[2] = hit tiles.
extern "C" int hc_candidate_rect_check(float mx, float my, float radius, float qa, float qb, float qc,
                                       float sigma_cut, int width, int height, int* stats) {
  int tile_w = (width + kTile - 1) / kTile, tile_h = (height + kTile - 1) / kTile;
  TileRect c = candidate_rect(mx, my, radius, qa, qb, qc, sigma_cut, tile_w, tile_h, 0, tile_h);
  float tr = radius / kTile, tx = mx / kTile, ty = my / kTile;
  int x0 = (int)fminf(fmaxf(floorf(tx - tr), 0.f), (float)tile_w), x1 = (int)fminf(fmaxf(ceilf(tx + tr), 0.f), (float)tile_w);
  int y0 = (int)fminf(fmaxf(floorf(ty - tr), 0.f), (float)tile_h), y1 = (int)fminf(fmaxf(ceilf(ty + tr), 0.f), (float)tile_h);
  int outside = 0, hits = 0;
  for (int y = y0; y < y1; ++y)
    for (int x = x0; x < x1; ++x)
      if (tile_hit(mx, my, qa, qb, qc, sigma_cut, x, y, width, height)) {
        ++hits;
        if (!(x >= c.x0 && x < c.x1 && y >= c.y0 && y < c.y1)) ++outside;
      }
  stats[0] = (x1 - x0) * (y1 - y0);
  stats[1] = (c.x1 - c.x0) * (c.y1 - c.y0);
  stats[2] = hits;
  return outside;
}
