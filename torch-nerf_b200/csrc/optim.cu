// Adam update of the flat parameter buffer (torch.optim.Adam as configured by runners/runner_utils.py:691-695:
// lr and eps given, betas (0.9, 0.999), no weight decay, no amsgrad), one launch over all 1 191 688 parameters of the
// two networks.  HBM-bound: 16 B read + 12 B written per parameter.
#include "common.cuh"

namespace nerf {

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, float grad_scale, float beta1,
                                                   float beta2, float step_size, float inv_sqrt_bc2, float eps) {
  const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 + 3 < n) {
    const float4 gv = *reinterpret_cast<const float4*>(g + i4);
    float4 pv = *reinterpret_cast<float4*>(p + i4), mv = *reinterpret_cast<float4*>(m + i4), vv = *reinterpret_cast<float4*>(v + i4);
    const float gs[4] = {gv.x * grad_scale, gv.y * grad_scale, gv.z * grad_scale, gv.w * grad_scale};
    float* pe = &pv.x;
    float* me = &mv.x;
    float* ve = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      me[k] = me[k] + (1.f - beta1) * (gs[k] - me[k]);            // lerp(exp_avg, grad, 1 - beta1)
      ve[k] = beta2 * ve[k] + (1.f - beta2) * gs[k] * gs[k];
      const float denom = sqrtf(ve[k]) * inv_sqrt_bc2 + eps;
      pe[k] -= step_size * (me[k] / denom);
    }
    *reinterpret_cast<float4*>(p + i4) = pv;
    *reinterpret_cast<float4*>(m + i4) = mv;
    *reinterpret_cast<float4*>(v + i4) = vv;
  } else {
    for (int64_t i = i4; i < n; ++i) {
      const float gg = g[i] * grad_scale;
      const float mm = m[i] + (1.f - beta1) * (gg - m[i]);
      const float vv = beta2 * v[i] + (1.f - beta2) * gg * gg;
      m[i] = mm;
      v[i] = vv;
      p[i] -= step_size * (mm / (sqrtf(vv) * inv_sqrt_bc2 + eps));
    }
  }
}

}  // namespace nerf

extern "C" int nerf_adam_step(float* param_dev, const float* grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev, int64_t n,
                              double lr, double beta1, double beta2, double eps, int64_t step, double grad_scale,
                              nerf_stream_t stream) {
  NERF_CHECK_ARG(n >= 0 && step >= 1, "nerf_adam_step: n must be >= 0 and step >= 1");
  if (n == 0) return NERF_OK;
  NERF_CHECK_ARG(param_dev && grad_dev && exp_avg_dev && exp_avg_sq_dev, "nerf_adam_step: null pointer");
  NERF_CHECK_ARG((reinterpret_cast<uintptr_t>(param_dev) | reinterpret_cast<uintptr_t>(grad_dev) |
                  reinterpret_cast<uintptr_t>(exp_avg_dev) | reinterpret_cast<uintptr_t>(exp_avg_sq_dev)) % 16 == 0,
                 "nerf_adam_step: buffers must be 16-byte aligned");
  // bias corrections in double on the host, like torch's single-tensor path
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  const int64_t quads = (n + 3) / 4;
  nerf::adam_kernel<<<(unsigned)nerf::ceil_div64(quads, 256), 256, 0, nerf::as_stream(stream)>>>(
      param_dev, grad_dev, exp_avg_dev, exp_avg_sq_dev, n, (float)grad_scale, (float)beta1, (float)beta2, (float)(lr / bc1),
      (float)(1.0 / sqrt(bc2)), (float)eps);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}
