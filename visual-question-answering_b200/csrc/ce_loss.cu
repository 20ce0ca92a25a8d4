// Mean softmax cross-entropy over the K answer classes, forward and gradient in ONE launch (replaces
// nn.CrossEntropyLoss()(logits, labels) of the reference training loop, main.py:179,214, and its autograd backward).
//
//   loss = scale / B * sum_b ( logsumexp(logits[b,:]) - logits[b, label_b] )
//   dlogits[b,k] = scale / B * ( softmax(logits[b,:])[k] - [k == label_b] )        (gradient for d loss = 1)
//
// One block per row: the row (K <= a few thousand fp32) is read once into registers / L1, max and sum-exp are block
// reductions, the gradient row is written in the same pass.  The row losses are summed in a fixed order by the last block
// to finish (deterministic; a completion counter in the workspace orders it): no atomics on fp32.
#include "common.cuh"
#include "util_kernels.cuh"

namespace hca {
namespace {

constexpr int CE_THREADS = 256;

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();                       // red may still be read from the previous reduction
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < CE_THREADS / 32; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
  return r;
}

__global__ void __launch_bounds__(CE_THREADS) ce_loss_kernel(const float* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels,
                                                             int B, int K, float scale, float* __restrict__ row_loss,
                                                             float* __restrict__ loss, float* __restrict__ dlogits, int64_t ldd,
                                                             unsigned int* __restrict__ counter) {
  pdl_enter();
  __shared__ float red[CE_THREADS / 32];
  __shared__ bool last;
  const int b = blockIdx.x;
  const float* x = logits + (int64_t)b * ld;
  float m = -INFINITY;
  for (int k = threadIdx.x; k < K; k += CE_THREADS) m = fmaxf(m, x[k]);
  m = block_reduce(m, true, red);
  float s = 0.f;
  for (int k = threadIdx.x; k < K; k += CE_THREADS) s += expf(x[k] - m);
  s = block_reduce(s, false, red);
  const float lse = m + logf(s);
  const int64_t y = labels[b];
  const bool y_ok = y >= 0 && y < K;
  const float g = scale / (float)B;
  if (dlogits) {
    float* d = dlogits + (int64_t)b * ldd;
    for (int k = threadIdx.x; k < K; k += CE_THREADS) d[k] = g * (expf(x[k] - lse) - ((int64_t)k == y ? 1.f : 0.f));
  }
  if (threadIdx.x == 0) {
    // an out-of-range label poisons the loss (torch raises a device-side assert there)
    row_loss[b] = y_ok ? lse - x[y] : __int_as_float(0x7fc00000);
    __threadfence();
    last = (atomicAdd(counter, 1u) == (unsigned)(B - 1));
  }
  __syncthreads();
  if (last) {
    __threadfence();
    float t = 0.f;
    for (int i = threadIdx.x; i < B; i += CE_THREADS) t += __ldcg(row_loss + i);
    t = block_reduce(t, false, red);
    if (threadIdx.x == 0) {
      loss[0] = g * t;
    }
  }
}

}  // namespace
}  // namespace hca

extern "C" size_t hca_ce_loss_workspace(int B) { return hca::align_up((size_t)B * 4) + 256; }

extern "C" int hca_ce_loss(const float* logits, int64_t ld, const int64_t* labels, int B, int K, float scale, float* loss,
                           float* dlogits, int64_t ldd, void* ws, size_t ws_bytes, void* stream) {
  using namespace hca;
  HCA_CHECK_ARG(logits && labels && loss && ws, "ce_loss: null pointer");
  HCA_CHECK_ARG(B > 0 && K > 0 && ld >= K && (!dlogits || ldd >= K), "ce_loss: bad sizes B=%d K=%d", B, K);
  HCA_CHECK_ARG(ws_bytes >= hca_ce_loss_workspace(B), "ce_loss: workspace too small");
  // ws = [completion counter (256 B)][row losses]
  unsigned int* counter = (unsigned int*)ws;
  HCA_TRY(zero_async(counter, 4, (cudaStream_t)stream));
  float* row_loss = (float*)((char*)ws + 256);
  HCA_LAUNCH_K((ce_loss_kernel), B, CE_THREADS, 0, (cudaStream_t)stream, logits, ld, labels, B, K, scale, row_loss, loss, dlogits, ldd, counter);
  HCA_LAUNCHED();
  return 0;
}
