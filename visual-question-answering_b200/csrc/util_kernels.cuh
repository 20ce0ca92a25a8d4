// Small shared kernels: column sums (bias gradients) and fills.
#pragma once
#include "common.cuh"

namespace hca {
// out[c] += sum_r X[r*ld + c]   (out must be zeroed by the caller unless accumulating on purpose)
int launch_colsum(const float* X, int64_t ld, int64_t rows, int cols, float* out, cudaStream_t s);
int zero_async(void* p, size_t bytes, cudaStream_t s);

// Several buffers cleared by ONE kernel launch instead of one memset node each (the step had 29 memsets, ~1.9 us apiece, most of them
// a few hundred bytes: bias gradients, counters, split-K accumulators).  Buffers and sizes must be 4-byte aligned.
struct ZeroJobs {                     // kernel argument: buffers to clear (4-byte words)
  static constexpr int MAX = 12;
  void* ptr[MAX];
  unsigned long long words[MAX];
  int count;
};
struct ZeroBatch {
  static constexpr int MAX = ZeroJobs::MAX;
  void* ptr[MAX];
  unsigned long long words[MAX];      // 4-byte words
  int count = 0;
  cudaStream_t stream;
  explicit ZeroBatch(cudaStream_t s) : stream(s) {}
  int add(void* p, size_t bytes);     // flushes first when full
  int flush();                        // launches (no-op when empty)
  // Hands the collected buffers to ANOTHER kernel of the same entry point instead of launching (that kernel calls zero_jobs_device at its
  // top; it must precede every consumer of the buffers in stream order, like the launch it replaces): one launch less per entry point.
  ZeroJobs take();
};
// grid-stride clear of the buffers of `jobs` by the calling kernel (all of its threads)
__device__ __forceinline__ void zero_jobs_device(const ZeroJobs& jobs) {
  const unsigned long long tid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + (unsigned long long)blockIdx.y * gridDim.x * blockDim.x,
                           nthr = (unsigned long long)gridDim.x * gridDim.y * blockDim.x;
  for (int j = 0; j < jobs.count; ++j) {
    uint32_t* p = (uint32_t*)jobs.ptr[j];
    const unsigned long long n = jobs.words[j];
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
      uint4* p4 = (uint4*)p;
      const unsigned long long n4 = n >> 2;
      for (unsigned long long i = tid; i < n4; i += nthr) p4[i] = make_uint4(0u, 0u, 0u, 0u);
      for (unsigned long long i = (n4 << 2) + tid; i < n; i += nthr) p[i] = 0u;
    } else {
      for (unsigned long long i = tid; i < n; i += nthr) p[i] = 0u;
    }
  }
}
}  // namespace hca
