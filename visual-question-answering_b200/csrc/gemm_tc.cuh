// tcgen05 / TMEM / TMA GEMM on split-bf16 operand planes (sm_100a): persistent, batched, with fused epilogues.
//
// fp32 operands are decomposed into P bf16 "planes"  x = x0 + x1 (+ x2),  x_{i+1} = bf16(x - x0 - .. - x_i),
// and the product is accumulated in fp32 TMEM from the plane pairs (i,j) with i + j < P:
//     P = 2 : 3 MMAs  (bf16x2 split, ~2^-16 operand precision)   <- default: meets the 1e-3 budget with margin
//     P = 3 : 6 MMAs  (bf16x3 split, ~fp32; the TMEM accumulation itself truncates at ~K * 2^-24)
// The smem tiles of one k-block are loaded once by TMA and reused by all plane pairs.
//
// One persistent CTA per SM walks a static round-robin schedule of 128 x BN output tiles (batch entry z, M tile, N tile,
// split-K slice): warp 0 = TMA producer, warp 1 = TMEM allocator + tcgen05.mma issuer, warps 2..5 = epilogue.  The fp32
// accumulator is double-buffered in TMEM (2 x BN columns), so the epilogue of tile i overlaps the mainloop of tile i+1.
// Operands may be K-major ([rows, K]) or MN-major ([K, rows]) -- both through 128-byte-swizzled TMA boxes and the matching
// UMMA shared-memory descriptors -- so NT (forward), NN (dgrad) and TN (wgrad) products and per-sample batches need no
// transposes in HBM.  A second operand pair may be chained along K (D = A.B^T + A2.B2^T in one accumulator).
//
// Epilogue (tcgen05.ld -> registers -> fused math -> swizzled smem -> TMA): the result can be written as fp32 (store or
// reduce-add), as bf16 hi/lo planes ready to be the operand of the next product (no fp32 round trip through HBM, no
// separate split pass), or both; or reduced on the fly (row dots, column sums).  BN = 32 instantiations run the
// "transposed" epilogue: the tile is written / its addend read as [n][m] with plain coalesced global accesses, which is
// how the products whose long dimension sits on M hand their result to consumers that want it K-major.
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"
#include "util_kernels.cuh"

namespace hca {

struct TcOperand {
  const __nv_bfloat16* planes = nullptr;  // [P][batch][rows][ld] (any strides below)
  int64_t ld = 0;                         // elements, multiple of 8
  int64_t plane_stride = 0;               // elements between planes, multiple of 8
  int64_t batch_stride = 0;               // elements between batch entries, multiple of 8 (0 when nbatch == 1)
  int nbatch = 1;                         // batch entries present in memory; entry used = (z / zdiv) % nbatch
  int zdiv = 1;
  int rows = 0, cols = 0;                 // K-major: rows = M|N, cols = K.  MN-major: rows = K, cols = M|N
  bool mn_major = false;
};

// bf16 hi/lo planes of a [batch][rows][ld] matrix (always 2 planes)
struct TcPlanes {
  __nv_bfloat16* p = nullptr;
  int64_t ld = 0, plane_stride = 0, batch_stride = 0;   // elements, multiples of 8
  int nbatch = 1, zdiv = 1;                              // entry = (z / zdiv) % nbatch  (inputs) ; z (outputs)
};

enum TcEpiMode : int {
  TC_EPI_STORE = 0,   // D[z] (+)= f  and / or  planes(f)
  TC_EPI_ROWDOT = 1,  // red_row[z][m] += sum_n f[m][n] * colv[n]                     (planes P of f may be kept; no fp32 store)
};
enum TcAuxMode : int { TC_AUX_NONE = 0, TC_AUX_ADD = 1 /* f = act(acc + aux) */, TC_AUX_MUL_1MX2 = 2 /* f *= 1 - aux^2 */ };

// f = act( acc + bias[z][n] + (aux if ADD) ) ; then (+ rowv[z][m] * r1col[z][group(m)][n]) ; then (* (1 - aux^2) if MUL) ; then mode
struct TcEpilogue {
  // fp32 output (optional).  Rows may be grouped: row m of the tile = group g = m / (M / d_groups), i = m % (M / d_groups),
  // stored at D + z * d_batch_stride + g * d_group_stride + i * ldd   (d_groups > 1 needs M <= 128).
  float* D = nullptr;
  int64_t ldd = 0, d_batch_stride = 0;
  int d_zdiv = 1;                                         // entry = z / d_zdiv (several z reduce-add into one entry)
  int d_groups = 1;  int64_t d_group_stride = 0;
  int accumulate = 0;                                     // D += result
  TcPlanes P;                                             // bf16 hi/lo planes output (optional)
  const float* bias = nullptr;  int64_t bias_batch_stride = 0;
  int act_tanh = 0;
  const float* mulx = nullptr;  int64_t mulx_ld = 0;     // per-element (1 - x^2) factor, un-batched, read directly
  int mode = TC_EPI_STORE;
  // addend / factor tile [M, N] per batch entry, fetched by TMA: either fp32 (aux) or bf16 hi/lo planes (auxp)
  const float* aux = nullptr;  int64_t aux_ld = 0, aux_batch_stride = 0;  int aux_nbatch = 1, aux_zdiv = 1;
  TcPlanes auxp;
  int aux_mode = TC_AUX_NONE;
  const float* rowv = nullptr;  int64_t rowv_batch_stride = 0;      // [z][m]
  const float* colv = nullptr;                                      // [n]
  const float* r1col = nullptr; int64_t r1col_batch_stride = 0;     // rank-1 term rowv[z][m] * r1col[z][g][n]
  int r1_rows_per_group = 0;    int64_t r1_group_stride = 0;        // g = m / r1_rows_per_group (0: one group)
  float* red_row = nullptr;     int64_t red_row_batch_stride = 0;
  float* red_col = nullptr;                                         // STORE: red_col[n] += sum_m f[m][n]
  // BN = 32 "transposed" epilogue: outputs and auxp are addressed [z][n][m] (ld = their leading dimension over m); fp32 D,
  // planes P, auxp (ADD / MUL_1MX2), act_tanh, red_row (un-batched: red_row[m] += sum_n f[m][n]) and a bias indexed by the
  // tile ROW (bias[z][m]: these products put the weight matrix on M) are supported, nothing else.
  int transposed = 0;
  // Block-structured contractions (un-batched, splitk = 1): output columns [g * kwin_ncol, (g + 1) * kwin_ncol) only have non-zero
  // operand data for k in [kwin_lo[g], kwin_hi[g]) (elements; up to 4 groups; 0 = off).  A tile contracts the union of the windows of
  // the groups it touches, rounded out to whole k-blocks -- the operand entries outside a group's own window must therefore BE zero in
  // memory (the windows only skip work, they never change the result).
  int kwin_ncol = 0;
  int kwin_lo[4] = {0, 0, 0, 0}, kwin_hi[4] = {0, 0, 0, 0};
  // Block-structured OUTPUTS (un-batched): of the output rows [g * nwin_nrow, (g + 1) * nwin_nrow) only the columns
  // [nwin_lo[g], nwin_hi[g]) are wanted (up to 4 groups; 0 = off).  Tiles entirely outside are not computed and their part of D is left
  // untouched; a tile that straddles is computed whole.
  int nwin_nrow = 0;
  int nwin_lo[4] = {0, 0, 0, 0}, nwin_hi[4] = {0, 0, 0, 0};
};

// D[z][M,N] (+)= A[z] . B[z]^T (+ A2[z] . B2[z]^T with inner size K2) for z < batch.  splitk >= 1 (un-batched only).
int launch_gemm_tc(const TcOperand& A, const TcOperand& B, int P, int M, int N, int K, const TcEpilogue& e, int splitk,
                   cudaStream_t s, int batch = 1, const TcOperand* A2 = nullptr, const TcOperand* B2 = nullptr, int K2 = 0);

// fp32 [rows, cols] (leading dim ld) -> P bf16 planes [P][rows][ldp]
int launch_split_planes(const float* src, int64_t ld, int64_t rows, int cols, __nv_bfloat16* planes, int64_t ldp,
                        int64_t plane_stride, int P, cudaStream_t s);
// three fp32 sources [B][T][cols] -> stacked bf16 hi/lo planes [2][B][3*T][ldp]  (rows of sample b: level-major, then t)
int launch_split_planes_stack3(const float* s0, const float* s1, const float* s2, int B, int T, int cols, __nv_bfloat16* planes,
                               int64_t ldp, int64_t plane_stride, cudaStream_t s);

// Several fp32 matrices -> bf16 hi/lo planes in ONE launch (the weight matrices of an entry point: each conversion alone is a
// ~3 us launch for ~1 us of work).  Sources 16-byte aligned with ld % 4 == 0 and cols % 4 == 0; always 2 planes.
struct SplitBatch {
  static constexpr int MAX = 6;
  struct Job { const float* src; __nv_bfloat16* dst; long long ld, rows, ldp, ps; int cols; } job[MAX];
  int count = 0;
  cudaStream_t stream;
  explicit SplitBatch(cudaStream_t s) : stream(s) {}
  int add(const float* src, int64_t ld, int64_t rows, int cols, __nv_bfloat16* planes, int64_t ldp, int64_t plane_stride);
  int flush();
  int flush(ZeroBatch& also_clear);    // the same launch also clears the buffers collected in `also_clear` (and empties it)
};

// Split-K factor for an un-batched product whose output has few tiles: as many K slices as give every SM ONE work item -- never more
// (tiles * slices > SMs means a second, nearly empty wave: measured as a 2x longer kernel on the weight gradients) -- with at least
// 4 k-blocks (256 of K) per slice.
int tc_splitk(int M, int N, int K);
int tc_splitk_tiles(int tiles, int K);                       // the same rule for a given number of output tiles
// 128 x 128 output tiles of an M x N product that the N windows of `e` (if any) leave to be computed
int tc_count_tiles(int M, int N, const TcEpilogue& e, int bm = 128, int bn = 128);

bool tc_available();   // TMA descriptor encoder resolved from the driver
// generic tiled TMA descriptor (rank 2..5) for the other hand-written kernels: `tm` points at a CUtensorMap; dims / box in
// elements (dim 0 innermost and contiguous), strides in BYTES for dims 1..rank-1; swizzle 0 none, 1 = 32 B, 2 = 64 B, 3 = 128 B
int tc_make_tmap(void* tm, bool bf16, int rank, const void* base, const uint64_t* dims, const uint64_t* strides_bytes,
                 const uint32_t* box, int swizzle);
// debug: record per-CTA clock64 stamps of the next gemm_tc launches into buf [nctas][64] (nullptr disables)
// launch_index >= 0: only that gemm_tc launch (counted from this call) records
void tc_set_timeline(long long* buf, int nctas, int launch_index = -1);
long long* tc_timeline_buffer();      // the buffer last handed to tc_set_timeline (the other kernels' debug stamps go there too)

}  // namespace hca
