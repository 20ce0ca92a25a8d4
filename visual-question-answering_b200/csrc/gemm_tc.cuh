// tcgen05 / TMEM / TMA GEMM on split-bf16 operand planes (sm_100a).
//
// fp32 operands are decomposed into P bf16 "planes"  x = x0 + x1 (+ x2),  x_{i+1} = bf16(x - x0 - .. - x_i),
// and the product is accumulated in fp32 TMEM from the plane pairs (i,j) with i + j < P:
//     P = 1 : 1 MMA   (plain bf16,            ~2^-8  operand precision)
//     P = 2 : 3 MMAs  (bf16x2 split,          ~2^-16)   <- default: meets the 1e-3 budget with margin
//     P = 3 : 6 MMAs  (bf16x3 split, fp32-grade ~2^-24) <- for argmax-critical products
// The smem tiles of one k-block are loaded once by TMA and reused by all plane pairs, so a P=2 k-block does
// 3 MMAs on 4 tiles (better smem/L2 reuse than a plain bf16 GEMM of the same tile shape).
//
// One CTA = one 128 x BN output tile (x one split-K slice): warp 0 = TMA producer, warp 1 = TMEM allocator +
// single-thread tcgen05.mma issuer, warps 2..5 = epilogue (tcgen05.ld -> bias / tanh / (1-x^2) -> global).
// Operands may be K-major ([rows, K], K contiguous) or MN-major ([K, rows], rows contiguous) -- both through
// 128-byte-swizzled TMA boxes and the matching UMMA shared-memory descriptors -- so NT (forward), NN (dgrad)
// and TN (wgrad) products need no transposes in HBM.
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace hca {

struct TcOperand {
  const __nv_bfloat16* planes = nullptr;  // [P][rows][ld]
  int64_t ld = 0;                         // elements, multiple of 8
  int64_t plane_stride = 0;               // elements between planes, multiple of 8
  int rows = 0, cols = 0;                 // K-major: rows = M|N, cols = K.  MN-major: rows = K, cols = M|N
  bool mn_major = false;
};

struct TcEpilogue {
  float* D = nullptr;
  int64_t ldd = 0;
  const float* bias = nullptr;
  int act_tanh = 0;
  const float* mulx = nullptr;
  int64_t mulx_ld = 0;
  int accumulate = 0;        // D += result (single split) ; split-K always accumulates atomically
};

// D[M,N] (+)= A . B^T with the layouts described by the operands.  splitk >= 1.
int launch_gemm_tc(const TcOperand& A, const TcOperand& B, int P, int M, int N, int K, const TcEpilogue& e, int splitk,
                   cudaStream_t s);

// fp32 [rows, cols] (leading dim ld) -> P bf16 planes [P][rows][ldp]
int launch_split_planes(const float* src, int64_t ld, int64_t rows, int cols, __nv_bfloat16* planes, int64_t ldp,
                        int64_t plane_stride, int P, cudaStream_t s);

bool tc_available();
// debug: record per-CTA clock64 stamps of the next gemm_tc launches into buf [nctas][8] (nullptr disables)
void tc_set_timeline(long long* buf, int nctas);   // TMA descriptor encoder resolved from the driver

}  // namespace hca
