// Dense (un-batched) contractions in the three layouts the path needs, dispatched to the tcgen05
// tensor-core kernel (gemm_tc.cu) or to the exact-fp32 CUDA-core kernel (gemm_ffma.cu).
//
//   dense_nt : D[M,N] = A[M,K] . B[N,K]^T      forward of nn.Linear (x . W^T), B is a weight [out,in]
//   dense_nn : D[M,N] = A[M,K] . B[K,N]        data gradient (dY . W)
//   dense_tn : D[M,N] = A[K,M]^T . B[K,N]      weight gradient (dY^T . X), long K -> split-K
//
// All matrices fp32 row-major with explicit leading dimensions.
#pragma once
#include "common.cuh"

namespace hca {

struct DenseEpi {
  const float* bias = nullptr;   // + bias[n]
  int act_tanh = 0;              // tanh(.)
  const float* mulx = nullptr;   // * (1 - mulx[m][n]^2)   (tanh backward), leading dim mulx_ld
  int64_t mulx_ld = 0;
  int accumulate = 0;            // D += result
  int exact = 0;                 // force the exact-fp32 path (argmax-critical products)
};

int dense_nt(const float* A, int64_t lda, const float* B, int64_t ldb, float* D, int64_t ldd, int M, int N, int K,
             const DenseEpi& e, Workspace& ws, cudaStream_t s);
int dense_nn(const float* A, int64_t lda, const float* B, int64_t ldb, float* D, int64_t ldd, int M, int N, int K,
             const DenseEpi& e, Workspace& ws, cudaStream_t s);
// zero_first: D is cleared first; otherwise the product is accumulated into D.
int dense_tn(const float* A, int64_t lda, const float* B, int64_t ldb, float* D, int64_t ldd, int M, int N, int K,
             bool zero_first, Workspace& ws, cudaStream_t s);

// scratch the tensor-core path may carve from the workspace for one call of the given shape
size_t dense_scratch_bytes(int M, int N, int K);

}  // namespace hca
