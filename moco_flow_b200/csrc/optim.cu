// Adam step over flat fp32 parameter / gradient / moment buffers (the optimizer step that follows the
// backward + all-reduce in the timed training step).  One launch per contiguous segment instead of torch's
// multi-tensor-apply chain; step count and learning rate live on the device so that the launch is CUDA-graph
// replayable while MultiStepLR keeps changing the rate between replays.
// Reference semantics: torch.optim.Adam as configured by trainer/base.py:122-133 (eps 1e-8, L2 weight decay,
// no amsgrad); op order follows torch/optim/adam.py::_single_tensor_adam.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/moco_flow_b200.h"

namespace mcf {

__global__ void k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                       float* __restrict__ v, long long n, const float* __restrict__ lr_dev,
                       const long long* __restrict__ step_dev, double beta1, double beta2, float eps, float wd,
                       float gscale) {
  // scalar arithmetic in double, then rounded to fp32, like the Python-side arithmetic of torch.optim.Adam
  const double t = (double)(step_dev[0] + 1);
  const double bc1 = 1.0 - pow(beta1, t);
  const double bc2 = 1.0 - pow(beta2, t);
  const float step_size = (float)((double)lr_dev[0] / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  const float w1 = (float)(1.0 - beta1), w2 = (float)(1.0 - beta2), b2f = (float)beta2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    float pi = p[i];
    float gi = g[i] * gscale;
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    float mi = m[i], vi = v[i];
    mi = fmaf(w1, gi - mi, mi);                 // exp_avg.lerp_(grad, 1 - beta1)
    vi = fmaf(w2 * gi, gi, vi * b2f);           // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi = fmaf(-step_size, mi / denom, pi);      // param.addcdiv_(exp_avg, denom, value=-step_size)
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
  }
}

__global__ void k_adam_tick(long long* step_dev) { step_dev[0] += 1; }

}  // namespace mcf

extern "C" int mcf_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                             const float* lr_dev, long long* step_dev, double beta1, double beta2, float eps,
                             float weight_decay, float grad_scale, int advance_step, cudaStream_t stream) {
  if (n < 0 || !lr_dev || !step_dev) return MCF_ERR_BAD_ARG;
  if (n > 0) {
    if (!params || !grads || !exp_avg || !exp_avg_sq) return MCF_ERR_BAD_ARG;
    long long blocks = (n + 256 * 4 - 1) / (256 * 4);
    if (blocks > 148 * 16) blocks = 148 * 16;
    mcf::k_adam<<<(unsigned)blocks, 256, 0, stream>>>(params, grads, exp_avg, exp_avg_sq, n, lr_dev, step_dev, beta1,
                                                      beta2, eps, weight_decay, grad_scale);
  }
  if (advance_step) mcf::k_adam_tick<<<1, 1, 0, stream>>>(step_dev);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}
