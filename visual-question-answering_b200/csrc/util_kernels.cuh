// Small shared kernels: column sums (bias gradients) and fills.
#pragma once
#include "common.cuh"

namespace hca {
// out[c] += sum_r X[r*ld + c]   (out must be zeroed by the caller unless accumulating on purpose)
int launch_colsum(const float* X, int64_t ld, int64_t rows, int cols, float* out, cudaStream_t s);
int zero_async(void* p, size_t bytes, cudaStream_t s);

// Several buffers cleared by ONE kernel launch instead of one memset node each (the step had 29 memsets, ~1.9 us apiece, most of them
// a few hundred bytes: bias gradients, counters, split-K accumulators).  Buffers and sizes must be 4-byte aligned.
struct ZeroBatch {
  static constexpr int MAX = 12;
  void* ptr[MAX];
  unsigned long long words[MAX];      // 4-byte words
  int count = 0;
  cudaStream_t stream;
  explicit ZeroBatch(cudaStream_t s) : stream(s) {}
  int add(void* p, size_t bytes);     // flushes first when full
  int flush();                        // launches (no-op when empty)
};
}  // namespace hca
