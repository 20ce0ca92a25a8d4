// tcgen05 / TMEM / TMA GEMM on split-bf16 operand planes (sm_100a), optionally batched, with fused epilogues.
//
// fp32 operands are decomposed into P bf16 "planes"  x = x0 + x1 (+ x2),  x_{i+1} = bf16(x - x0 - .. - x_i),
// and the product is accumulated in fp32 TMEM from the plane pairs (i,j) with i + j < P:
//     P = 1 : 1 MMA   (plain bf16,            ~2^-8  operand precision)
//     P = 2 : 3 MMAs  (bf16x2 split,          ~2^-16)   <- default: meets the 1e-3 budget with margin
//     P = 3 : 6 MMAs  (bf16x3 split, ~fp32; the TMEM accumulation itself truncates at ~K * 2^-24)
// The smem tiles of one k-block are loaded once by TMA and reused by all plane pairs, so a P=2 k-block does
// 3 MMAs on 4 tiles (better smem/L2 reuse than a plain bf16 GEMM of the same tile shape).
//
// One CTA = one 128 x 128 output tile of one batch entry (x one split-K slice): warp 0 = TMA producer, warp 1 =
// TMEM allocator + tcgen05.mma issuer, warps 2..5 = epilogue (tcgen05.ld -> fused math -> swizzled smem -> TMA
// tile store / reduce-add).  Operands may be K-major ([rows, K]) or MN-major ([K, rows]) -- both through
// 128-byte-swizzled TMA boxes and the matching UMMA shared-memory descriptors -- so NT (forward), NN (dgrad) and
// TN (wgrad) products need no transposes in HBM.  A second operand pair may be chained along K
// (D = A.B^T + A2.B2^T in one accumulator).
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace hca {

struct TcOperand {
  const __nv_bfloat16* planes = nullptr;  // [P][batch][rows][ld] (any strides below)
  int64_t ld = 0;                         // elements, multiple of 8
  int64_t plane_stride = 0;               // elements between planes, multiple of 8
  int64_t batch_stride = 0;               // elements between batch entries, multiple of 8 (0 when nbatch == 1)
  int nbatch = 1;                         // batch entries present in memory; entry used = z % nbatch
  int rows = 0, cols = 0;                 // K-major: rows = M|N, cols = K.  MN-major: rows = K, cols = M|N
  bool mn_major = false;
};

enum TcEpiMode : int {
  TC_EPI_STORE = 0,   // D[z] (+)= f
  TC_EPI_ROWDOT = 1,  // red_row[z][m] += sum_n f[m][n] * colv[n]                     (nothing stored)
  TC_EPI_DZ = 2,      // D[z] = rowv[z][m] * colv[n] * (1 - f^2) ; red_col[n] += sum_m f[m][n] * rowv[z][m]
};
enum TcAuxMode : int { TC_AUX_NONE = 0, TC_AUX_ADD = 1 /* f = act(acc + aux) */, TC_AUX_MUL_1MX2 = 2 /* f *= 1 - aux^2 */ };

// f = act( acc + bias[z][n] + (aux if ADD) ) ; then (+ rowv[z][m] * r1col[z][n]) ; then (* (1 - aux^2) if MUL) ; then mode
struct TcEpilogue {
  float* D = nullptr;
  int64_t ldd = 0, d_batch_stride = 0;
  const float* bias = nullptr;  int64_t bias_batch_stride = 0;
  int act_tanh = 0;
  const float* mulx = nullptr;  int64_t mulx_ld = 0;     // legacy per-row (1 - x^2) factor, un-batched, read directly
  int accumulate = 0;                                     // D += result
  int mode = TC_EPI_STORE;
  // aux tile [M, N] fp32 per batch entry (entry = z % aux_nbatch), fetched by TMA: needs 16-byte aligned rows
  const float* aux = nullptr;  int64_t aux_ld = 0, aux_batch_stride = 0;  int aux_nbatch = 1;  int aux_mode = TC_AUX_NONE;
  const float* rowv = nullptr;  int64_t rowv_batch_stride = 0;      // [z][m]
  const float* colv = nullptr;                                      // [n]
  const float* r1col = nullptr; int64_t r1col_batch_stride = 0;     // rank-1 term rowv[z][m] * r1col[z][n]
  float* red_row = nullptr;     int64_t red_row_batch_stride = 0;
  float* red_col = nullptr;
};

// D[z][M,N] (+)= A[z] . B[z]^T (+ A2[z] . B2[z]^T with inner size K2) for z < batch.  splitk >= 1 (un-batched only).
int launch_gemm_tc(const TcOperand& A, const TcOperand& B, int P, int M, int N, int K, const TcEpilogue& e, int splitk,
                   cudaStream_t s, int batch = 1, const TcOperand* A2 = nullptr, const TcOperand* B2 = nullptr, int K2 = 0);

// fp32 [rows, cols] (leading dim ld) -> P bf16 planes [P][rows][ldp]
int launch_split_planes(const float* src, int64_t ld, int64_t rows, int cols, __nv_bfloat16* planes, int64_t ldp,
                        int64_t plane_stride, int P, cudaStream_t s);

bool tc_available();   // TMA descriptor encoder resolved from the driver
// debug: record per-CTA clock64 stamps of the next gemm_tc launches into buf [nctas][64] (nullptr disables)
void tc_set_timeline(long long* buf, int nctas);

}  // namespace hca
