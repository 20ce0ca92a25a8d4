// Small shared kernels: column sums (bias gradients) and fills.
#pragma once
#include "common.cuh"

namespace hca {
// out[c] += sum_r X[r*ld + c]   (out must be zeroed by the caller unless accumulating on purpose)
int launch_colsum(const float* X, int64_t ld, int64_t rows, int cols, float* out, cudaStream_t s);
int zero_async(void* p, size_t bytes, cudaStream_t s);
}  // namespace hca
