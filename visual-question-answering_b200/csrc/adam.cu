// Fused Adam over the flat parameter / gradient buffers (replaces torch.optim.Adam(model.parameters(), lr), reference
// main.py:180,222: default betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad).
//
//   m = b1 m + (1 - b1) g ;  v = b2 v + (1 - b2) g^2 ;  p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
//
// One pass over four flat fp32 arrays (read p, g, m, v; write p, m, v: 28 bytes per parameter) instead of the eight
// multi-tensor passes of the foreach implementation.  The step count lives on the device so the update can be captured in
// a CUDA graph; a one-thread kernel advances it and derives the two bias-correction factors in double precision.
#include "common.cuh"

namespace hca {
namespace {

__global__ void adam_prep_kernel(long long* __restrict__ step, float* __restrict__ coef, float lr, float b1, float b2) {
  pdl_enter();
  const long long t = step[0] + 1;
  step[0] = t;
  coef[0] = (float)((double)lr / (1.0 - pow((double)b1, (double)t)));   // step size
  coef[1] = (float)(1.0 / sqrt(1.0 - pow((double)b2, (double)t)));      // 1 / sqrt(bias correction 2)
}

__global__ void __launch_bounds__(256) adam_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                                                   float4* __restrict__ v, int64_t n4, const float* __restrict__ coef, float b1, float b2,
                                                   float eps) {
  pdl_enter();
  const float step_size = coef[0], inv_bc2 = coef[1];
  const float c1 = 1.f - b1, c2 = 1.f - b2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = p[i], mm = m[i], vv = v[i];
    const float4 gg = g[i];
#define HCA_ADAM1(c)                                         \
  mm.c = fmaf(b1, mm.c, c1 * gg.c);                          \
  vv.c = fmaf(b2, vv.c, c2 * gg.c * gg.c);                   \
  pp.c -= step_size * (mm.c / fmaf(sqrtf(vv.c), inv_bc2, eps));
    HCA_ADAM1(x) HCA_ADAM1(y) HCA_ADAM1(z) HCA_ADAM1(w)
#undef HCA_ADAM1
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}

}  // namespace
}  // namespace hca

extern "C" int hca_adam_step(float* p, const float* g, float* m, float* v, int64_t n, long long* step, float* coef, float lr,
                             float beta1, float beta2, float eps, void* stream) {
  using namespace hca;
  cudaStream_t s = (cudaStream_t)stream;
  HCA_CHECK_ARG(p && g && m && v && step && coef, "adam_step: null pointer");
  HCA_CHECK_ARG(n > 0 && n % 4 == 0, "adam_step: the flat buffers must hold a multiple of 4 elements (got %lld)", (long long)n);
  HCA_CHECK_ARG(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                  reinterpret_cast<uintptr_t>(v)) & 15) == 0, "adam_step: buffers must be 16-byte aligned");
  HCA_LAUNCH_K((adam_prep_kernel), 1, 1, 0, s, step, coef, lr, beta1, beta2);
  HCA_LAUNCHED();
  HCA_LAUNCH_K((adam_kernel), ew_grid(n / 4), 256, 0, s, (float4*)p, (const float4*)g, (float4*)m, (float4*)v, n / 4, coef, beta1, beta2, eps);
  HCA_LAUNCHED();
  return 0;
}
