// Densification statistics of one training step, one launch.
//
// Replaces the dozen ATen kernels (boolean-mask gathers / scatters, norm, maximum) behind
// VanillaGaussians.after_train (models/gaussians/vanilla.py:163-191, paths relative to /root/reference/project), which
// BasicTrainer.postprocess_per_train_step (models/trainers/base.py:279-297) calls every step with the taps this
// library emits: radii [N] (info["radii"]), xys_grad [N,2] (info["means2d"].absgrad, already scaled by W/2, H/2).
//
//   first step after a reset (the statistics were None):   xys_grad_norm = |xys_grad| for EVERY Gaussian, vis_counts = 1
//   otherwise, for visible Gaussians (radii > 0):           xys_grad_norm += |xys_grad|, vis_counts += 1
//   always, for visible Gaussians:                          max_2Dsize = max(max_2Dsize, radii / last_size)
//
// Pointwise, HBM-bound: 12 B read + up to 12 B read-modify-written per Gaussian.
#include "bds_common.cuh"

namespace bds {
__global__ void __launch_bounds__(256) densify_stats_kernel(int64_t n, const int32_t* __restrict__ radii,
                                                            const float* __restrict__ xys_grad, float scale_x,
                                                            float scale_y, float inv_last_size, int first,
                                                            float* __restrict__ xys_grad_norm,
                                                            float* __restrict__ vis_counts,
                                                            float* __restrict__ max_2dsize) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = radii[i];
  const bool vis = r > 0;
  const float gx = xys_grad[2 * i] * scale_x, gy = xys_grad[2 * i + 1] * scale_y;
  const float g = sqrtf(gx * gx + gy * gy);
  if (first) {
    xys_grad_norm[i] = g;
    vis_counts[i] = 1.f;
  } else if (vis) {
    xys_grad_norm[i] += g;
    vis_counts[i] += 1.f;
  }
  if (vis) max_2dsize[i] = fmaxf(max_2dsize[i], (float)r * inv_last_size);
}
}  // namespace bds

extern "C" int bds_densify_stats(int64_t n, const int32_t* radii, const float* xys_grad, float scale_x, float scale_y,
                                 float last_size, int first, float* xys_grad_norm, float* vis_counts,
                                 float* max_2dsize, bds_stream_t stream) {
  if (n == 0) return 0;
  BDS_REQUIRE(radii && xys_grad && xys_grad_norm && vis_counts && max_2dsize, "densify_stats: null pointer");
  BDS_REQUIRE(last_size > 0.f, "densify_stats: last_size must be positive");
  bds::densify_stats_kernel<<<bds::ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      n, radii, xys_grad, scale_x, scale_y, 1.0f / last_size, first, xys_grad_norm, vis_counts, max_2dsize);
  BDS_CHECK_LAUNCH();
  return 0;
}
