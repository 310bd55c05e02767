// Fused photometric loss of the benchmark step:
//   loss = mean((rgb - gt)^2) + lambda_d * mean(depth) + lambda_a * mean(alpha)
// One pass: reads rgb/gt/depth/alpha, writes the three cotangents, block-reduces the loss.
// Pointwise, HBM-bound: 32 B read + 20 B written per pixel.
#include "bds_common.cuh"

namespace bds {
__global__ void __launch_bounds__(256) loss_kernel(int64_t n_pix, const float* __restrict__ rgb,
                                                   const float* __restrict__ gt, const float* __restrict__ depth,
                                                   const float* __restrict__ alpha, float lambda_d, float lambda_a,
                                                   float inv_count, float* __restrict__ loss, float* __restrict__ v_rgb,
                                                   float* __restrict__ v_depth, float* __restrict__ v_alpha) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
  if (i < n_pix) {
    const float k = inv_count * (1.0f / 3.0f);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float d = rgb[i * 3 + c] - gt[i * 3 + c];
      acc = fmaf(d * d, k, acc);
      v_rgb[i * 3 + c] = 2.f * d * k;
    }
    if (depth) { acc = fmaf(depth[i], lambda_d * inv_count, acc); v_depth[i] = lambda_d * inv_count; }
    if (alpha) { acc = fmaf(alpha[i], lambda_a * inv_count, acc); v_alpha[i] = lambda_a * inv_count; }
  }
  __shared__ float red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = red[threadIdx.x];
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffu, v, o);
    if (threadIdx.x == 0) red_add(loss, v);
  }
}
}  // namespace bds

extern "C" int bds_loss_fwd_bwd(int64_t n_pix, const float* rgb, const float* gt, const float* depth,
                                const float* alpha, float lambda_d, float lambda_a, float inv_count, float* loss,
                                float* v_rgb, float* v_depth, float* v_alpha, bds_stream_t stream) {
  if (n_pix == 0) return 0;
  BDS_REQUIRE(rgb && gt && loss && v_rgb, "loss: null pointer");
  BDS_REQUIRE((!depth || v_depth) && (!alpha || v_alpha), "loss: cotangent buffer missing");
  bds::loss_kernel<<<bds::ceil_div(n_pix, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      n_pix, rgb, gt, depth, alpha, lambda_d, lambda_a, inv_count, loss, v_rgb, v_depth, v_alpha);
  BDS_CHECK_LAUNCH();
  return 0;
}
