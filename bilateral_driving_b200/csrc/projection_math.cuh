// Per-Gaussian projection arithmetic (forward + VJP), host/device so the same code is unit-tested
// on the CPU against the oracle's autograd (tests/hostcheck) before it ever runs on the GPU.
//
// Restates what gsplat v1.3.0 fully_fused_projection computes for the call at
// models/trainers/base.py:393-408 of the reference:
//   p_c = R mu + t;  Sigma = Rq diag(s^2) Rq^T (q normalised);  Sigma_c = R Sigma R^T
//   J with the symmetric 1.3*tan(fov) clamp;  Sigma_2 = J Sigma_c J^T + eps2d I
//   conic = Sigma_2^-1;  radius = ceil(3 sqrt(lambda_max));  culls: near/far, det<=0, radius_clip,
//   off-screen.
#pragma once
#include <math.h>

#include "bds_common.cuh"

namespace bds {

struct CamIntr {
  float fx, fy, cx, cy;
  float R[9];  // row-major world->camera rotation
  float t[3];
};

BDS_HD void quat_to_rotmat(const float q[4], float Rm[9]) {
  float w = q[0], x = q[1], y = q[2], z = q[3];
  float inv = 1.0f / sqrtf(w * w + x * x + y * y + z * z);
  w *= inv; x *= inv; y *= inv; z *= inv;
  float x2 = x * x, y2 = y * y, z2 = z * z, xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
  Rm[0] = 1.f - 2.f * (y2 + z2); Rm[1] = 2.f * (xy - wz);       Rm[2] = 2.f * (xz + wy);
  Rm[3] = 2.f * (xy + wz);       Rm[4] = 1.f - 2.f * (x2 + z2); Rm[5] = 2.f * (yz - wx);
  Rm[6] = 2.f * (xz - wy);       Rm[7] = 2.f * (yz + wx);       Rm[8] = 1.f - 2.f * (x2 + y2);
}

// cotangent of the UN-normalised quaternion given the cotangent of the rotation matrix
BDS_HD void quat_to_rotmat_vjp(const float q[4], const float vR[9], float vq[4]) {
  float w = q[0], x = q[1], y = q[2], z = q[3];
  float inv = 1.0f / sqrtf(w * w + x * x + y * y + z * z);
  w *= inv; x *= inv; y *= inv; z *= inv;
  // d/d(normalised q)
  float vw = 2.f * (x * (vR[7] - vR[5]) + y * (vR[2] - vR[6]) + z * (vR[3] - vR[1]));
  float vx = 2.f * (-2.f * x * (vR[4] + vR[8]) + y * (vR[1] + vR[3]) + z * (vR[2] + vR[6]) + w * (vR[7] - vR[5]));
  float vy = 2.f * (x * (vR[1] + vR[3]) - 2.f * y * (vR[0] + vR[8]) + z * (vR[5] + vR[7]) + w * (vR[2] - vR[6]));
  float vz = 2.f * (x * (vR[2] + vR[6]) + y * (vR[5] + vR[7]) - 2.f * z * (vR[0] + vR[4]) + w * (vR[3] - vR[1]));
  // through the normalisation: v_q = (v_qn - qn <qn, v_qn>) / |q|
  float dot = w * vw + x * vx + y * vy + z * vz;
  vq[0] = (vw - w * dot) * inv;
  vq[1] = (vx - x * dot) * inv;
  vq[2] = (vy - y * dot) * inv;
  vq[3] = (vz - z * dot) * inv;
}

struct Proj {
  float x, y, z;        // camera-space mean
  float mx, my;         // means2d
  float a, b, c;        // conic (inverse of the blurred 2-D covariance)
  float comp;           // antialias compensation sqrt(max(0, det0/det))
  float radius;         // ceil(3 sqrt(lambda_max)), 0 if culled
  // kept for the backward
  float cov2a, cov2b, cov2c;  // blurred 2-D covariance
  float det0, det;
};

// cov3 (world) symmetric 6: xx xy xz yy yz zz
BDS_HD void covar_world(const float Rq[9], const float s[3], float cov[6]) {
  float M[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) M[i * 3 + j] = Rq[i * 3 + j] * s[j];
  cov[0] = M[0] * M[0] + M[1] * M[1] + M[2] * M[2];
  cov[1] = M[0] * M[3] + M[1] * M[4] + M[2] * M[5];
  cov[2] = M[0] * M[6] + M[1] * M[7] + M[2] * M[8];
  cov[3] = M[3] * M[3] + M[4] * M[4] + M[5] * M[5];
  cov[4] = M[3] * M[6] + M[4] * M[7] + M[5] * M[8];
  cov[5] = M[6] * M[6] + M[7] * M[7] + M[8] * M[8];
}

// Sc = R S R^T for symmetric S (6) -> full 3x3 row-major (symmetric)
BDS_HD void covar_cam(const float R[9], const float S[6], float Sc[9]) {
  float Sf[9] = {S[0], S[1], S[2], S[1], S[3], S[4], S[2], S[4], S[5]};
  float RS[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) RS[i * 3 + j] = R[i * 3] * Sf[j] + R[i * 3 + 1] * Sf[3 + j] + R[i * 3 + 2] * Sf[6 + j];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) Sc[i * 3 + j] = RS[i * 3] * R[j * 3] + RS[i * 3 + 1] * R[j * 3 + 1] + RS[i * 3 + 2] * R[j * 3 + 2];
}

// Cheap conservative screen cull before the covariance arithmetic (5/6 of a rig's Gaussians are outside
// any one camera): radius <= 3 sqrt(lambda_max) + 1 and lambda_max <= ||J||_F^2 smax^2 + 0.7 (trace bound
// incl. the eps2d blur and the 0.01 floor), so a splat whose mean (camera coordinates x, y, z) is further
// than that bound outside the image is culled by gsplat's own test as well.  true = culled.
BDS_HD bool screen_cull(float x, float y, float z, float smax, float fx, float fy, float cx, float cy, int width,
                        int height, float eps2d) {
  float lx = 1.3f * (0.5f * (float)width / fx), ly = 1.3f * (0.5f * (float)height / fy);
  float rzq = 1.0f / z;
  float jf2 = (fx * fx * (1.f + lx * lx) + fy * fy * (1.f + ly * ly)) * rzq * rzq;
  float rb = 3.f * sqrtf(jf2 * smax * smax + 0.7f + eps2d) * 1.001f + 2.f;
  float px = fx * x * rzq + cx, py = fy * y * rzq + cy;
  return px + rb <= 0.f || px - rb >= (float)width || py + rb <= 0.f || py - rb >= (float)height;
}

// Projection of one Gaussian whose world covariance (symmetric 6) is already known; smax = largest
// scale.  Returns false when culled (radius = 0).
BDS_HD bool project_gaussian_cov(const float mu[3], const float cov[6], float smax, const CamIntr& cam, int width,
                                 int height, float eps2d, float near_plane, float far_plane, float radius_clip,
                                 Proj& o) {
  o.radius = 0.f;
  o.x = cam.R[0] * mu[0] + cam.R[1] * mu[1] + cam.R[2] * mu[2] + cam.t[0];
  o.y = cam.R[3] * mu[0] + cam.R[4] * mu[1] + cam.R[5] * mu[2] + cam.t[1];
  o.z = cam.R[6] * mu[0] + cam.R[7] * mu[1] + cam.R[8] * mu[2] + cam.t[2];
  if (o.z < near_plane || o.z > far_plane) return false;
  if (screen_cull(o.x, o.y, o.z, smax, cam.fx, cam.fy, cam.cx, cam.cy, width, height, eps2d)) return false;
  float Sc[9];
  covar_cam(cam.R, cov, Sc);
  float limx = 1.3f * (0.5f * (float)width / cam.fx), limy = 1.3f * (0.5f * (float)height / cam.fy);
  float rz = 1.0f / o.z, rz2 = rz * rz;
  float tx = o.z * fminf(limx, fmaxf(-limx, o.x * rz));
  float ty = o.z * fminf(limy, fmaxf(-limy, o.y * rz));
  float j00 = cam.fx * rz, j02 = -cam.fx * tx * rz2, j11 = cam.fy * rz, j12 = -cam.fy * ty * rz2;
  // cov2 = J Sc J^T with J = [[j00, 0, j02], [0, j11, j12]]
  float c00 = j00 * j00 * Sc[0] + 2.f * j00 * j02 * Sc[2] + j02 * j02 * Sc[8];
  float c01 = j00 * j11 * Sc[1] + j00 * j12 * Sc[2] + j02 * j11 * Sc[5] + j02 * j12 * Sc[8];
  float c11 = j11 * j11 * Sc[4] + 2.f * j11 * j12 * Sc[5] + j12 * j12 * Sc[8];
  o.mx = cam.fx * o.x * rz + cam.cx;
  o.my = cam.fy * o.y * rz + cam.cy;
  o.det0 = c00 * c11 - c01 * c01;
  c00 += eps2d;
  c11 += eps2d;
  float det = c00 * c11 - c01 * c01;
  o.det = det;
  if (det <= 0.f) return false;
  float inv = 1.0f / det;
  o.a = c11 * inv;
  o.b = -c01 * inv;
  o.c = c00 * inv;
  o.cov2a = c00; o.cov2b = c01; o.cov2c = c11;
  o.comp = sqrtf(fmaxf(0.f, o.det0 * inv));
  float bh = 0.5f * (c00 + c11);
  float lam = bh + sqrtf(fmaxf(0.01f, bh * bh - det));
  float radius = ceilf(3.f * sqrtf(lam));
  if (radius <= radius_clip) return false;
  if (o.mx + radius <= 0.f || o.mx - radius >= (float)width || o.my + radius <= 0.f || o.my - radius >= (float)height)
    return false;
  o.radius = radius;
  return true;
}

// Returns false when culled (radius = 0).
BDS_HD bool project_gaussian(const float mu[3], const float q[4], const float s[3], const CamIntr& cam, int width,
                             int height, float eps2d, float near_plane, float far_plane, float radius_clip,
                             Proj& o) {
  float Rq[9], cov[6];
  quat_to_rotmat(q, Rq);
  covar_world(Rq, s, cov);
  return project_gaussian_cov(mu, cov, fmaxf(s[0], fmaxf(s[1], s[2])), cam, width, height, eps2d, near_plane,
                              far_plane, radius_clip, o);
}

// VJP.  Inputs: cotangents of means2d (vmx, vmy), depth (vz), conic (va, vb, vc) and (antialiased
// only) of the compensation factor.  Outputs: v_mu[3], v_q[4], v_s[3] (ACCUMULATED by the caller),
// and v_R[9] / v_t[3] of the view matrix (written).
BDS_HD void project_gaussian_vjp(const float mu[3], const float q[4], const float s[3], const CamIntr& cam, int width,
                                 int height, float eps2d, float vmx, float vmy, float vz, float va, float vb,
                                 float vc, float vcomp, float v_mu[3], float v_q[4], float v_s[3], float v_R[9],
                                 float v_t[3]) {
  // ---- recompute forward intermediates
  float x = cam.R[0] * mu[0] + cam.R[1] * mu[1] + cam.R[2] * mu[2] + cam.t[0];
  float y = cam.R[3] * mu[0] + cam.R[4] * mu[1] + cam.R[5] * mu[2] + cam.t[1];
  float z = cam.R[6] * mu[0] + cam.R[7] * mu[1] + cam.R[8] * mu[2] + cam.t[2];
  float Rq[9], cov[6], Sc[9];
  quat_to_rotmat(q, Rq);
  covar_world(Rq, s, cov);
  covar_cam(cam.R, cov, Sc);
  float limx = 1.3f * (0.5f * (float)width / cam.fx), limy = 1.3f * (0.5f * (float)height / cam.fy);
  float rz = 1.0f / z, rz2 = rz * rz, rz3 = rz2 * rz;
  float xr = x * rz, yr = y * rz;
  bool in_x = (xr >= -limx) && (xr <= limx), in_y = (yr >= -limy) && (yr <= limy);
  float tx = z * fminf(limx, fmaxf(-limx, xr));
  float ty = z * fminf(limy, fmaxf(-limy, yr));
  float j00 = cam.fx * rz, j02 = -cam.fx * tx * rz2, j11 = cam.fy * rz, j12 = -cam.fy * ty * rz2;
  float c00 = j00 * j00 * Sc[0] + 2.f * j00 * j02 * Sc[2] + j02 * j02 * Sc[8];
  float c01 = j00 * j11 * Sc[1] + j00 * j12 * Sc[2] + j02 * j11 * Sc[5] + j02 * j12 * Sc[8];
  float c11 = j11 * j11 * Sc[4] + 2.f * j11 * j12 * Sc[5] + j12 * j12 * Sc[8];
  float det0 = c00 * c11 - c01 * c01;
  float b00 = c00 + eps2d, b11 = c11 + eps2d;
  float det = b00 * b11 - c01 * c01;
  float inv = 1.0f / det;
  float ca = b11 * inv, cb = -c01 * inv, cc = b00 * inv;  // conic
  // ---- conic -> blurred covariance:  v_Sigma2 = -conic V conic, V = [[va, vb/2],[vb/2, vc]]
  float V00 = va, V01 = 0.5f * vb, V11 = vc;
  // X = conic * V
  float X00 = ca * V00 + cb * V01, X01 = ca * V01 + cb * V11, X10 = cb * V00 + cc * V01, X11 = cb * V01 + cc * V11;
  float g00 = -(X00 * ca + X01 * cb), g01 = -(X00 * cb + X01 * cc), g11 = -(X10 * cb + X11 * cc);
  // g01 is the symmetric off-diagonal entry of v_Sigma2 (full matrix [[g00,g01],[g01,g11]])
  if (vcomp != 0.f) {
    // comp = sqrt(max(0, det0/det)); det0 = c00 c11 - c01^2 (unblurred), det = b00 b11 - c01^2
    float ratio = det0 * inv;
    if (ratio > 0.f) {
      float comp = sqrtf(ratio);
      float v_ratio = vcomp * 0.5f / comp;
      float v_det0 = v_ratio * inv, v_det = -v_ratio * det0 * inv * inv;
      g00 += v_det0 * c11 + v_det * b11;
      g11 += v_det0 * c00 + v_det * b00;
      g01 += -(v_det0 + v_det) * c01;  // d/dc01 of both dets is -2 c01, split over the two symmetric entries
    }
  }
  // ---- cov2 = J Sc J^T:  v_Sc = J^T G J,  v_J = 2 G J Sc   (G symmetric 2x2)
  float J[6] = {j00, 0.f, j02, 0.f, j11, j12};
  float G[4] = {g00, g01, g01, g11};
  float GJ[6];  // 2x3
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) GJ[i * 3 + k] = G[i * 2] * J[k] + G[i * 2 + 1] * J[3 + k];
  float vSc[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) vSc[i * 3 + k] = J[i] * GJ[k] + J[3 + i] * GJ[3 + k];
  float vJ[6];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k)
      vJ[i * 3 + k] = 2.f * (GJ[i * 3] * Sc[k] + GJ[i * 3 + 1] * Sc[3 + k] + GJ[i * 3 + 2] * Sc[6 + k]);
  // ---- mean: means2d and depth
  float vpx = cam.fx * rz * vmx, vpy = cam.fy * rz * vmy;
  float vpz = -(cam.fx * x * vmx + cam.fy * y * vmy) * rz2 + vz;
  // ---- J entries
  // j00 = fx/z, j11 = fy/z
  vpz += -cam.fx * rz2 * vJ[0] - cam.fy * rz2 * vJ[4];
  // j02 = -fx tx / z^2 with tx = x (inside) or z*lim (outside)
  if (in_x) {
    vpx += -cam.fx * rz2 * vJ[2];
    vpz += 2.f * cam.fx * tx * rz3 * vJ[2];
  } else {
    vpz += cam.fx * tx * rz3 * vJ[2];  // j02 = -fx * (+-lim) / z
  }
  if (in_y) {
    vpy += -cam.fy * rz2 * vJ[5];
    vpz += 2.f * cam.fy * ty * rz3 * vJ[5];
  } else {
    vpz += cam.fy * ty * rz3 * vJ[5];
  }
  // ---- p = R mu + t
  const float* R = cam.R;
  v_mu[0] += R[0] * vpx + R[3] * vpy + R[6] * vpz;
  v_mu[1] += R[1] * vpx + R[4] * vpy + R[7] * vpz;
  v_mu[2] += R[2] * vpx + R[5] * vpy + R[8] * vpz;
  v_t[0] = vpx; v_t[1] = vpy; v_t[2] = vpz;
  float vp[3] = {vpx, vpy, vpz};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) v_R[i * 3 + k] = vp[i] * mu[k];
  // ---- Sc = R S R^T:  v_S = R^T vSc R,  v_R += (vSc + vSc^T) R S  (vSc symmetric here -> 2 vSc R S)
  float Sf[9] = {cov[0], cov[1], cov[2], cov[1], cov[3], cov[4], cov[2], cov[4], cov[5]};
  float vScR[9], RS[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      vScR[i * 3 + k] = vSc[i * 3] * R[k] + vSc[i * 3 + 1] * R[3 + k] + vSc[i * 3 + 2] * R[6 + k];
      RS[i * 3 + k] = R[i * 3] * Sf[k] + R[i * 3 + 1] * Sf[3 + k] + R[i * 3 + 2] * Sf[6 + k];
    }
  float vS[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      vS[i * 3 + k] = R[i] * vScR[k] + R[3 + i] * vScR[3 + k] + R[6 + i] * vScR[6 + k];
      // (vSc + vSc^T) R S, vSc symmetric
      v_R[i * 3 + k] += 2.f * (vSc[i * 3] * RS[k] + vSc[i * 3 + 1] * RS[3 + k] + vSc[i * 3 + 2] * RS[6 + k]);
    }
  // ---- S = M M^T, M = Rq diag(s):  v_M = (vS + vS^T) M
  float M[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) M[i * 3 + k] = Rq[i * 3 + k] * s[k];
  float vM[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k)
      vM[i * 3 + k] = (vS[i * 3] + vS[i]) * M[k] + (vS[i * 3 + 1] + vS[3 + i]) * M[3 + k] +
                      (vS[i * 3 + 2] + vS[6 + i]) * M[6 + k];
  float vRq[9];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    v_s[k] += Rq[k] * vM[k] + Rq[3 + k] * vM[3 + k] + Rq[6 + k] * vM[6 + k];
#pragma unroll
    for (int i = 0; i < 3; ++i) vRq[i * 3 + k] = vM[i * 3 + k] * s[k];
  }
  float vq[4];
  quat_to_rotmat_vjp(q, vRq, vq);
#pragma unroll
  for (int k = 0; k < 4; ++k) v_q[k] += vq[k];
}

// Conservative lower bound of sigma' = a' dx^2 + b' dx dy + c' dy^2 (d = mu - p) over the pixel-centre
// rectangle [xmin,xmax] x [ymin,ymax]: exact minimum of the convex quadratic on the box.
BDS_HD float min_sigma_rect(float gx, float gy, float qa, float qb, float qc, float xmin, float xmax, float ymin,
                            float ymax) {
  float dxlo = gx - xmax, dxhi = gx - xmin, dylo = gy - ymax, dyhi = gy - ymin;
  float dxn = dxlo > 0.f ? dxlo : (dxhi < 0.f ? dxhi : 0.f);
  float dyn = dylo > 0.f ? dylo : (dyhi < 0.f ? dyhi : 0.f);
  if (dxn == 0.f && dyn == 0.f) return 0.f;
  float s = 3.0e38f;
  if (dxn != 0.f) {
    float dy = fminf(fmaxf(-qb * dxn * (0.5f / qc), dylo), dyhi);
    s = qa * dxn * dxn + qb * dxn * dy + qc * dy * dy;
  }
  if (dyn != 0.f) {
    float dx = fminf(fmaxf(-qb * dyn * (0.5f / qa), dxlo), dxhi);
    s = fminf(s, qa * dx * dx + qb * dx * dyn + qc * dyn * dyn);
  }
  return s;
}
constexpr float kCullMargin = 0.02f;  // slack on sigma' (log2 units) so rounding never culls a contributor

// Does the splat reach alpha >= 1/255 on some pixel centre of tile (tx, ty)?  NOT inlined on the
// device: the counting pass (projection.cu) and the emission pass (binning.cu) must take the very
// same decision, so both call this one body compiled from this one source.
#ifdef __CUDACC__
static __host__ __device__ __noinline__
#else
static inline
#endif
bool tile_hit(float gx, float gy, float qa, float qb, float qc, float sigma_cut, int tx, int ty, int width,
              int height) {
  float xmin = (float)(tx * kTile) + 0.5f, ymin = (float)(ty * kTile) + 0.5f;
  float xmax = fminf((float)(tx * kTile + kTile) - 0.5f, (float)width - 0.5f);
  float ymax = fminf((float)(ty * kTile + kTile) - 0.5f, (float)height - 0.5f);
  float s = min_sigma_rect(gx, gy, qa, qb, qc, xmin, xmax, ymin, ymax);
  return !(s > sigma_cut + kCullMargin);
}


// Candidate tile rectangle of one splat inside the band rows [ty0, ty1) of its camera: gsplat's
// isect_tiles square (3-sigma radius) bound intersected with the bounding box of the alpha >= 1/255
// ellipse {sigma' <= sigma_cut}.  Only a candidate set: tile_hit() decides.  NOT inlined, for the same
// reason as tile_hit (count and emission must agree).
struct TileRect { int x0, x1, y0, y1; };
#ifdef __CUDACC__
static __host__ __device__ __noinline__
#else
static inline
#endif
TileRect candidate_rect(float mx, float my, float radius, float qa, float qb, float qc, float sigma_cut, int tile_w,
                        int tile_h, int ty0, int ty1) {
  const float inv = 1.0f / (float)kTile;
  float tr = radius * inv, tx = mx * inv, ty = my * inv;
  TileRect r;
  r.x0 = (int)fminf(fmaxf(floorf(tx - tr), 0.f), (float)tile_w);
  r.x1 = (int)fminf(fmaxf(ceilf(tx + tr), 0.f), (float)tile_w);
  r.y0 = (int)fminf(fmaxf(floorf(ty - tr), 0.f), (float)tile_h);
  r.y1 = (int)fminf(fmaxf(ceilf(ty + tr), 0.f), (float)tile_h);
  // ellipse bounding box (pixel-centre coordinates), with slack
  float det = qa * qc - 0.25f * qb * qb;
  float cut = sigma_cut + 2.f * kCullMargin;
  if (det > 0.f && cut > 0.f) {
    float hx = sqrtf(cut * qc / det) + 0.05f, hy = sqrtf(cut * qa / det) + 0.05f;
    // tile t covers pixel centres [16 t + 0.5, 16 t + 15.5]
    float lx = ceilf((mx - hx - 15.5f) * inv), ux = floorf((mx + hx - 0.5f) * inv) + 1.f;
    float ly = ceilf((my - hy - 15.5f) * inv), uy = floorf((my + hy - 0.5f) * inv) + 1.f;
    if (lx > (float)r.x0) r.x0 = (int)fminf(lx, (float)tile_w);
    if (ux < (float)r.x1) r.x1 = (int)fmaxf(ux, 0.f);
    if (ly > (float)r.y0) r.y0 = (int)fminf(ly, (float)tile_h);
    if (uy < (float)r.y1) r.y1 = (int)fmaxf(uy, 0.f);
  }
  if (r.y0 < ty0) r.y0 = ty0;
  if (r.y1 > ty1) r.y1 = ty1;
  if (r.x1 < r.x0) r.x1 = r.x0;
  if (r.y1 < r.y0) r.y1 = r.y0;
  return r;
}


#ifdef __CUDACC__
// Flat warp-cooperative enumeration of candidate tiles: lane j owns ncand_j candidates; work item w of the
// concatenated list belongs to the largest lane whose exclusive prefix is <= w (zero-count lanes share
// their successor's prefix, so they are never picked for w below the total).
BDS_D int warp_inclusive_scan_i32(int v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
BDS_D int warp_find_owner(int excl, int w) {
  int owner = 0;
#pragma unroll
  for (int step = 16; step > 0; step >>= 1) {
    const int cl = owner + step;
    const int e = __shfl_sync(0xffffffffu, excl, cl & 31);
    if (cl < 32 && e <= w) owner = cl;
  }
  return owner;
}
#endif

}  // namespace bds
