// Real spherical harmonics (degree <= 3... 4 via bases 0..15), Sloan's fast evaluation, as
// gsplat.cuda._wrapper.spherical_harmonics evaluates them for the reference call at
// models/gaussians/vanilla.py:383-389 (direction normalised inside the kernel).
#pragma once
#include "bds_common.cuh"

namespace bds {

// basis values for a NORMALISED direction; nb = (degree+1)^2 <= 16
BDS_HD void sh_basis(int degree, float x, float y, float z, float b[16]) {
  b[0] = 0.2820947917738781f;
  if (degree < 1) return;
  b[1] = -0.48860251190292f * y;
  b[2] = 0.48860251190292f * z;
  b[3] = -0.48860251190292f * x;
  if (degree < 2) return;
  float z2 = z * z;
  float fTmp0B = -1.092548430592079f * z;
  float fC1 = x * x - y * y;
  float fS1 = 2.f * x * y;
  b[4] = 0.5462742152960395f * fS1;
  b[5] = fTmp0B * y;
  b[6] = 0.9461746957575601f * z2 - 0.3153915652525201f;
  b[7] = fTmp0B * x;
  b[8] = 0.5462742152960395f * fC1;
  if (degree < 3) return;
  float fTmp0C = -2.285228997322329f * z2 + 0.4570457994644658f;
  float fTmp1B = 1.445305721320277f * z;
  float fC2 = x * fC1 - y * fS1;
  float fS2 = x * fS1 + y * fC1;
  b[9] = -0.5900435899266435f * fS2;
  b[10] = fTmp1B * fS1;
  b[11] = fTmp0C * y;
  b[12] = z * (1.865881662950577f * z2 - 1.119528997770346f);
  b[13] = fTmp0C * x;
  b[14] = fTmp1B * fC1;
  b[15] = -0.5900435899266435f * fC2;
}

// d(basis)/d(x,y,z) for a normalised direction (used only when dirs require grad)
BDS_HD void sh_basis_grad(int degree, float x, float y, float z, float dbx[16], float dby[16], float dbz[16]) {
#pragma unroll
  for (int k = 0; k < 16; ++k) dbx[k] = dby[k] = dbz[k] = 0.f;
  if (degree < 1) return;
  dby[1] = -0.48860251190292f;
  dbz[2] = 0.48860251190292f;
  dbx[3] = -0.48860251190292f;
  if (degree < 2) return;
  const float k1 = 1.092548430592079f;
  dbx[4] = k1 * y;              dby[4] = k1 * x;
  dby[5] = -k1 * z;             dbz[5] = -k1 * y;
  dbz[6] = 2.f * 0.9461746957575601f * z;
  dbx[7] = -k1 * z;             dbz[7] = -k1 * x;
  dbx[8] = k1 * x;              dby[8] = -k1 * y;
  if (degree < 3) return;
  float z2 = z * z;
  float fC1 = x * x - y * y, fS1 = 2.f * x * y;
  float fTmp0C = -2.285228997322329f * z2 + 0.4570457994644658f;
  float fTmp1B = 1.445305721320277f * z;
  const float k3 = 0.5900435899266435f;
  // fS2 = 3x^2 y - y^3 ; fC2 = x^3 - 3 x y^2
  dbx[9] = -k3 * 6.f * x * y;           dby[9] = -k3 * 3.f * fC1;
  dbx[10] = fTmp1B * 2.f * y;           dby[10] = fTmp1B * 2.f * x;      dbz[10] = 1.445305721320277f * fS1;
  dby[11] = fTmp0C;                     dbz[11] = -2.f * 2.285228997322329f * z * y;
  dbz[12] = 3.f * 1.865881662950577f * z2 - 1.119528997770346f;
  dbx[13] = fTmp0C;                     dbz[13] = -2.f * 2.285228997322329f * z * x;
  dbx[14] = fTmp1B * 2.f * x;           dby[14] = -fTmp1B * 2.f * y;     dbz[14] = 1.445305721320277f * fC1;
  dbx[15] = -k3 * 3.f * fC1;            dby[15] = k3 * 6.f * x * y;
}

}  // namespace bds
