// Word-embedding gather / scatter-add (replaces nn.Embedding fwd+bwd, reference model.py:263,282).
#include "common.cuh"
#include "util_kernels.cuh"

namespace hca {
namespace {

__global__ void __launch_bounds__(256) embedding_fwd_kernel(const int64_t* __restrict__ tokens, const float4* __restrict__ table,
                                                            float4* __restrict__ out, int64_t rows, int E4, int64_t vocab) {
  pdl_enter();
  const int64_t total = rows * E4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / E4;
    const int c = (int)(i - r * E4);
    int64_t tok = tokens[r];
    tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);   // clamp: never read out of bounds
    out[i] = __ldg(table + tok * E4 + c);
  }
}

__global__ void __launch_bounds__(256) embedding_bwd_kernel(const int64_t* __restrict__ tokens, const float* __restrict__ dout,
                                                            float* __restrict__ dtable, int64_t rows, int E, int64_t vocab) {
  pdl_enter();
  const int64_t total = rows * E;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / E;
    const int c = (int)(i - r * E);
    const int64_t tok = tokens[r];
    if (tok <= 0 || tok >= vocab) continue;                 // padding_idx = 0 receives no gradient
    atomicAdd(dtable + tok * E + c, dout[i]);
  }
}

}  // namespace
}  // namespace hca

extern "C" int hca_embedding_fwd(const int64_t* tokens, const float* table, float* out, int64_t rows, int E, int64_t vocab,
                                 void* stream) {
  using namespace hca;
  HCA_CHECK_ARG(tokens && table && out, "embedding_fwd: null pointer");
  HCA_CHECK_ARG(rows >= 0 && E > 0 && E % 4 == 0 && vocab > 0, "embedding_fwd: bad sizes rows=%lld E=%d vocab=%lld", (long long)rows, E, (long long)vocab);
  if (rows == 0) return 0;
  const int64_t total = rows * (E / 4);
  const int grid = ew_grid(total);
  HCA_LAUNCH_K((embedding_fwd_kernel), grid, 256, 0, (cudaStream_t)stream, tokens, (const float4*)table, (float4*)out, rows, E / 4, vocab);
  HCA_LAUNCHED();
  return 0;
}

extern "C" int hca_embedding_bwd(const int64_t* tokens, const float* dout, float* dtable, int64_t rows, int E, int64_t vocab,
                                 void* stream) {
  using namespace hca;
  HCA_CHECK_ARG(tokens && dout && dtable, "embedding_bwd: null pointer");
  HCA_CHECK_ARG(rows >= 0 && E > 0 && vocab > 0, "embedding_bwd: bad sizes");
  HCA_TRY(zero_async(dtable, (size_t)vocab * E * sizeof(float), (cudaStream_t)stream));
  if (rows == 0) return 0;
  const int64_t total = rows * E;
  const int grid = ew_grid(total);
  HCA_LAUNCH_K((embedding_bwd_kernel), grid, 256, 0, (cudaStream_t)stream, tokens, dout, dtable, rows, E, vocab);
  HCA_LAUNCHED();
  return 0;
}
