// Generic strided, batched fp32 GEMM on CUDA cores with a fused epilogue.
//
//   acc[z][m][n] = sum_k A[z][m][k] * B[z][n][k]   (+ sum_k A2[z][m][k] * B2[z][n][k])
//
// Every operand is addressed through element strides, so transposed / permuted views (the reference's
// x_img.permute(0,2,1), C.transpose(2,1), ...) cost nothing.  This is the exact-fp32 path: it carries
// the phrase-conv forward (whose max-pool argmax must match the reference bit for bit) and all the
// small per-sample products; the tcgen05 path (gemm_tc.cu) takes over the large dense contractions.
#pragma once
#include "common.cuh"

namespace hca {

struct Operand {           // element (z, r, k) lives at p[(z % zmod) * sb + r * sr + k * sk]
  const float* p = nullptr;
  int64_t sb = 0, sr = 0, sk = 0;
  int zmod = 0;            // 0 = use z as is
};
struct MatRef {            // element (z, m, n)
  const float* p = nullptr;
  int64_t sb = 0, sm = 0, sn = 0;
  int zmod = 0;
};

enum Epi : int {
  EPI_STORE = 0,   // D = f(acc)                     (split-K / accumulate => atomicAdd / +=)
  EPI_ROWDOT = 1,  // red_row[z][m] += sum_n f(acc) * colv[n]            (nothing stored)
  EPI_DZ = 2,      // h = f(acc); D = rowv[z][m]*colv[n]*(1-h*h); red_col[n] += sum_m h*rowv[z][m]
};

struct GemmParams {
  Operand A, B, A2, B2;    // A2/B2 optional second product with its own K2
  int M = 0, N = 0, K = 0, K2 = 0, batch = 1, splitk = 1;
  float* D = nullptr;
  int64_t d_sb = 0, d_sm = 0, d_sn = 1;
  int epi = EPI_STORE;
  // f(acc) = act(acc + bias[n] + add[z][m][n])
  const float* bias = nullptr;
  MatRef add;
  int act_tanh = 0;
  // EPI_STORE extras, applied after f:  v += r1_row[z][m] * r1_col[z][n];  v *= (1 - mulx^2);  D (+)= v
  const float* r1_row = nullptr; int64_t r1r_sb = 0;
  const float* r1_col = nullptr; int64_t r1c_sb = 0;
  MatRef mulx;
  int accumulate = 0;
  // EPI_ROWDOT / EPI_DZ
  const float* rowv = nullptr; int64_t rowv_sb = 0;
  const float* colv = nullptr;
  float* red_row = nullptr; int64_t red_row_sb = 0;
  float* red_col = nullptr;
};

// big = 128x128 tiles (8x8 per thread) for the large dense products, else 64x64 (4x4 per thread)
int launch_gemm_ffma(const GemmParams& p, bool big, cudaStream_t stream);

}  // namespace hca
