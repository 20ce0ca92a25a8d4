// ParallelCoAttention, three levels with shared weights (replaces reference model.py:356-397).
//
// Per sample b and level l (V = image features [N,d], Q = question level [T,d]):
//   C  = tanh(Q V^T)                 PV = V Wv^T + bv  (level independent: once per step)
//   Hv = tanh(PV + C^T PQ)           PQ = Q Wq^T + bq  (once per level)
//   Hq = tanh(PQ + C PV)
//   av = softmax_N(Hv wv + cv)       aq = softmax_T(Hq wq + cq)   (all T positions, pads included)
//   vhat = av^T V                    qhat = aq^T Q
// W_b of the reference is dead code (model.py:347 vs :377) and does not appear.
//
// Backward follows SURVEY.md section 3.3; Hv / Hq are recomputed from the saved PV, PQ, C instead of being
// stored (3 x B x N x d floats otherwise).
//
// This file is the orchestration + the warp-level glue kernels; the contractions go through dense.cuh
// (large, tensor-core capable) and gemm_ffma.cuh (per-sample strided products with fused epilogues).
#include <algorithm>
#include "common.cuh"
#include "dense.cuh"
#include "gemm_ffma.cuh"
#include "gemm_tc.cuh"
#include "util_kernels.cuh"

namespace hca {
namespace {

__device__ __forceinline__ float block_sum(float v, float* red) {   // red: >= 33 floats of smem
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  const int nw = blockDim.x >> 5;
  float s = (l < nw) ? red[l] : 0.f;
  s = warp_sum(s);
  return s;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  const int nw = blockDim.x >> 5;
  float s = (l < nw) ? red[l] : -INFINITY;
  s = warp_max(s);
  return s;
}

// dense copy of a strided [B,N,d] tensor
__global__ void __launch_bounds__(256) gather_strided_kernel(const float* __restrict__ V, int64_t sb, int64_t sn, int64_t sd,
                                                             float* __restrict__ out, int B, int N, int d) {
  const int64_t total = (int64_t)B * N * d;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % d);
    const int n = (int)((i / d) % N);
    const int64_t b = i / ((int64_t)d * N);
    out[i] = V[b * sb + n * sn + c * sd];
  }
}

// softmax over L scores, then the attention-weighted sum of the L rows of X [L,d].
//   a[l] = softmax(s[l] + c) ; out[c] = sum_l a[l] X[l][c]
__device__ void softmax_wsum(const float* __restrict__ s, float cbias, const float* __restrict__ X, int L, int d,
                             float* __restrict__ a_out, float* __restrict__ out, float* a_sm, float* red) {
  float m = -INFINITY;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const float v = s[i] + cbias;
    a_sm[i] = v;
    m = fmaxf(m, v);
  }
  m = block_max(m, red);
  float sum = 0.f;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const float e = expf(a_sm[i] - m);
    a_sm[i] = e;
    sum += e;
  }
  sum = block_sum(sum, red);
  const float inv = 1.f / sum;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const float a = a_sm[i] * inv;
    a_sm[i] = a;
    a_out[i] = a;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float acc = 0.f;
    for (int l = 0; l < L; ++l) acc = fmaf(a_sm[l], X[(int64_t)l * d + c], acc);
    out[c] = acc;
  }
  __syncthreads();
}

// one block per (level, sample)
__global__ void __launch_bounds__(256) attn_finish_kernel(const float* __restrict__ sv, const float* __restrict__ sq,
                                                          const float* __restrict__ cv, const float* __restrict__ cq,
                                                          const float* __restrict__ V, const float* __restrict__ q0,
                                                          const float* __restrict__ q1, const float* __restrict__ q2,
                                                          float* __restrict__ av, float* __restrict__ aq,
                                                          float* __restrict__ vhat, float* __restrict__ qhat,
                                                          int B, int N, int T, int d) {
  extern __shared__ float sm[];
  float* red = sm;          // 64
  float* a_sm = sm + 64;    // max(N,T)
  const int z = blockIdx.x, l = z / B, b = z % B;
  const float* Q = (l == 0 ? q0 : (l == 1 ? q1 : q2)) + (int64_t)b * T * d;
  softmax_wsum(sv + (int64_t)z * N, cv[0], V + (int64_t)b * N * d, N, d, av + (int64_t)z * N, vhat + (int64_t)z * d, a_sm, red);
  softmax_wsum(sq + (int64_t)z * T, cq[0], Q, T, d, aq + (int64_t)z * T, qhat + (int64_t)z * d, a_sm, red);
}

// da[l] = X[l,:] . g ; ds = a * (da - <a,da>) ; dc += sum ds
__device__ void softmax_bwd(const float* __restrict__ a, const float* __restrict__ X, const float* __restrict__ g, int L, int d,
                            float* __restrict__ ds_out, float* __restrict__ dc, float* da_sm, float* red) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int l = w; l < L; l += nw) {
    const float* x = X + (int64_t)l * d;
    float acc = 0.f;
    for (int c = lane; c < d; c += 32) acc = fmaf(x[c], g[c], acc);
    acc = warp_sum(acc);
    if (lane == 0) da_sm[l] = acc;
  }
  __syncthreads();
  float dot = 0.f;
  for (int i = threadIdx.x; i < L; i += blockDim.x) dot = fmaf(a[i], da_sm[i], dot);
  dot = block_sum(dot, red);
  float tot = 0.f;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const float v = a[i] * (da_sm[i] - dot);
    ds_out[i] = v;
    tot += v;
  }
  tot = block_sum(tot, red);
  if (threadIdx.x == 0) atomicAdd(dc, tot);
  __syncthreads();
}

__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const float* __restrict__ av, const float* __restrict__ aq,
                                                            const float* __restrict__ V, const float* __restrict__ q0,
                                                            const float* __restrict__ q1, const float* __restrict__ q2,
                                                            const float* __restrict__ gv, const float* __restrict__ gq,
                                                            float* __restrict__ dsv, float* __restrict__ dsq,
                                                            float* __restrict__ dcv, float* __restrict__ dcq,
                                                            int B, int N, int T, int d) {
  extern __shared__ float sm[];
  float* red = sm;
  float* da_sm = sm + 64;
  const int z = blockIdx.x, l = z / B, b = z % B;
  const float* Q = (l == 0 ? q0 : (l == 1 ? q1 : q2)) + (int64_t)b * T * d;
  softmax_bwd(av + (int64_t)z * N, V + (int64_t)b * N * d, gv + (int64_t)z * d, N, d, dsv + (int64_t)z * N, dcv, da_sm, red);
  softmax_bwd(aq + (int64_t)z * T, Q, gq + (int64_t)z * d, T, d, dsq + (int64_t)z * T, dcq, da_sm, red);
}



// ---------------------------------------------------------------------------------------------------------------
// tensor-core path: every contraction of the module on tcgen05 with bf16x2 operand planes (gemm_tc.cuh).  Per-sample
// products are batched launches (grid.z = level * B + sample); image-side operands are shared across levels through
// the batch modulo of the TMA batch coordinate.
constexpr int TCP = 2;    // planes

struct Planes {
  __nv_bfloat16* p = nullptr;
  int64_t ld = 0, plane_stride = 0;
  int64_t rows = 0;
  int cols = 0;
};

inline int64_t round8(int64_t x) { return (x + 7) / 8 * 8; }

// allocate planes for an fp32 matrix [rows, cols] (rows may be filled by several split launches)
int alloc_planes(Planes& pl, int64_t rows, int cols, Workspace& w) {
  pl.ld = round8(cols);
  pl.rows = rows;
  pl.cols = cols;
  pl.plane_stride = rows * pl.ld;
  pl.p = w.take<__nv_bfloat16>((size_t)TCP * pl.plane_stride);
  if (!pl.p) return set_err(HCA_ERR_WORKSPACE, "coattn: workspace too small for bf16 operand planes (%lld x %d)", (long long)rows, cols);
  return 0;
}
int fill_planes(const Planes& pl, const float* src, int64_t ld, int64_t row0, int64_t rows, cudaStream_t s) {
  return launch_split_planes(src, ld, rows, pl.cols, pl.p + row0 * pl.ld, pl.ld, pl.plane_stride, TCP, s);
}
int make_planes(Planes& pl, const float* src, int64_t ld, int64_t rows, int cols, Workspace& w, cudaStream_t s) {
  HCA_TRY(alloc_planes(pl, rows, cols, w));
  return fill_planes(pl, src, ld, 0, rows, s);
}
// batched view: entry z = rows [row0 + z * rows_per_batch, +rows_per_batch) of the plane matrix
TcOperand view(const Planes& pl, int64_t row0, int rows_per_batch, int nbatch, bool mn_major) {
  TcOperand o;
  o.planes = pl.p + row0 * pl.ld;
  o.ld = pl.ld;
  o.plane_stride = pl.plane_stride;
  o.batch_stride = (int64_t)rows_per_batch * pl.ld;
  o.nbatch = nbatch;
  o.rows = rows_per_batch;
  o.cols = pl.cols;
  o.mn_major = mn_major;
  return o;
}

bool tc_path_ok(int N, int d) { return use_tc() && tc_available() && (d % 8 == 0) && (N % 4 == 0); }

int tc_splitk(int M, int N, int K) {
  const int tiles = ((M + 127) / 128) * ((N + 127) / 128);
  if (tiles >= 96) return 1;
  int sk = (148 + tiles - 1) / tiles;
  const int maxk = (K + 255) / 256;
  if (sk > maxk) sk = maxk;
  return sk < 1 ? 1 : sk;
}

struct CoattnPlanes {
  Planes V, Q, C, PV, PQ, Wv, Wq;
};

// planes every direction needs: V, the three question levels stacked [3*B*T, d], and the two projection weights
int common_planes(CoattnPlanes& cp, const float* Vd, const float* const q[3], const float* Wv, const float* Wq, int B, int N, int T,
                  int d, Workspace& w, cudaStream_t s) {
  const int64_t BT = (int64_t)B * T;
  HCA_TRY(make_planes(cp.V, Vd, d, (int64_t)B * N, d, w, s));
  HCA_TRY(alloc_planes(cp.Q, 3 * BT, d, w));
  for (int l = 0; l < 3; ++l) HCA_TRY(fill_planes(cp.Q, q[l], d, l * BT, BT, s));
  HCA_TRY(make_planes(cp.Wv, Wv, d, d, d, w, s));
  HCA_TRY(make_planes(cp.Wq, Wq, d, d, d, w, s));
  return 0;
}

int coattn_fwd_tc(const float* Vd, const float* const q[3], const float* Wv, const float* bv, const float* Wq, const float* bq,
                  const float* wv, const float* wq, float* PV, float* PQ, float* C, float* sv, float* sq, int B, int N, int T, int d,
                  Workspace& w, cudaStream_t s) {
  const int64_t BT = (int64_t)B * T, BN = (int64_t)B * N;
  CoattnPlanes cp;
  HCA_TRY(common_planes(cp, Vd, q, Wv, Wq, B, N, T, d, w, s));
  {  // PV = V Wv^T + bv   (once per step: level independent)
    TcEpilogue e; e.D = PV; e.ldd = d; e.bias = bv;
    HCA_TRY(launch_gemm_tc(view(cp.V, 0, (int)BN, 1, false), view(cp.Wv, 0, d, 1, false), TCP, (int)BN, d, d, e, 1, s));
  }
  {  // PQ = Q Wq^T + bq   (the three levels in one launch)
    TcEpilogue e; e.D = PQ; e.ldd = d; e.bias = bq;
    HCA_TRY(launch_gemm_tc(view(cp.Q, 0, (int)(3 * BT), 1, false), view(cp.Wq, 0, d, 1, false), TCP, (int)(3 * BT), d, d, e, 1, s));
  }
  {  // C[z] = tanh(Q[z] V[b]^T)
    TcEpilogue e; e.D = C; e.ldd = N; e.d_batch_stride = (int64_t)T * N; e.act_tanh = 1;
    HCA_TRY(launch_gemm_tc(view(cp.Q, 0, T, 3 * B, false), view(cp.V, 0, N, B, false), TCP, T, N, d, e, 1, s, 3 * B));
  }
  HCA_TRY(make_planes(cp.C, C, N, 3 * BT, N, w, s));
  HCA_TRY(make_planes(cp.PV, PV, d, BN, d, w, s));
  HCA_TRY(make_planes(cp.PQ, PQ, d, 3 * BT, d, w, s));
  {  // sq[z][t] = sum_j tanh(PQ + C PV)[t][j] wq[j]
    TcEpilogue e; e.mode = TC_EPI_ROWDOT; e.act_tanh = 1; e.colv = wq; e.red_row = sq; e.red_row_batch_stride = T;
    e.aux = PQ; e.aux_ld = d; e.aux_batch_stride = (int64_t)T * d; e.aux_nbatch = 3 * B; e.aux_mode = TC_AUX_ADD;
    HCA_TRY(launch_gemm_tc(view(cp.C, 0, T, 3 * B, false), view(cp.PV, 0, N, B, true), TCP, T, d, N, e, 1, s, 3 * B));
  }
  {  // sv[z][n] = sum_j tanh(PV + C^T PQ)[n][j] wv[j]
    TcEpilogue e; e.mode = TC_EPI_ROWDOT; e.act_tanh = 1; e.colv = wv; e.red_row = sv; e.red_row_batch_stride = N;
    e.aux = PV; e.aux_ld = d; e.aux_batch_stride = (int64_t)N * d; e.aux_nbatch = B; e.aux_mode = TC_AUX_ADD;
    HCA_TRY(launch_gemm_tc(view(cp.C, 0, T, 3 * B, true), view(cp.PQ, 0, T, 3 * B, true), TCP, N, d, T, e, 1, s, 3 * B));
  }
  return 0;
}

int coattn_bwd_tc(const float* Vd, const float* const q[3], const float* Wv, const float* Wq, const float* wv, const float* wq,
                  const float* PV, const float* PQ, const float* C, const float* av, const float* aq, const float* gvhat,
                  const float* gqhat, const float* dsv, const float* dsq, float* dZv, float* dZq, float* dPQ, float* dPV, float* dS,
                  float* dV, float* dQ, float* dWv, float* dbv, float* dWq, float* dbq, float* dwv, float* dwq, int B, int N, int T,
                  int d, Workspace& w, cudaStream_t s) {
  const int64_t BT = (int64_t)B * T, BN = (int64_t)B * N;
  CoattnPlanes cp;
  HCA_TRY(common_planes(cp, Vd, q, Wv, Wq, B, N, T, d, w, s));
  HCA_TRY(make_planes(cp.C, C, N, 3 * BT, N, w, s));
  HCA_TRY(make_planes(cp.PV, PV, d, BN, d, w, s));
  HCA_TRY(make_planes(cp.PQ, PQ, d, 3 * BT, d, w, s));
  {  // dZv = (dsv x wv) * (1 - Hv^2), Hv = tanh(PV + C^T PQ) recomputed ; dwv += Hv^T dsv
    TcEpilogue e; e.mode = TC_EPI_DZ; e.act_tanh = 1; e.colv = wv; e.rowv = dsv; e.rowv_batch_stride = N; e.red_col = dwv;
    e.aux = PV; e.aux_ld = d; e.aux_batch_stride = (int64_t)N * d; e.aux_nbatch = B; e.aux_mode = TC_AUX_ADD;
    e.D = dZv; e.ldd = d; e.d_batch_stride = (int64_t)N * d;
    HCA_TRY(launch_gemm_tc(view(cp.C, 0, T, 3 * B, true), view(cp.PQ, 0, T, 3 * B, true), TCP, N, d, T, e, 1, s, 3 * B));
  }
  {  // dZq likewise from Hq = tanh(PQ + C PV)
    TcEpilogue e; e.mode = TC_EPI_DZ; e.act_tanh = 1; e.colv = wq; e.rowv = dsq; e.rowv_batch_stride = T; e.red_col = dwq;
    e.aux = PQ; e.aux_ld = d; e.aux_batch_stride = (int64_t)T * d; e.aux_nbatch = 3 * B; e.aux_mode = TC_AUX_ADD;
    e.D = dZq; e.ldd = d; e.d_batch_stride = (int64_t)T * d;
    HCA_TRY(launch_gemm_tc(view(cp.C, 0, T, 3 * B, false), view(cp.PV, 0, N, B, true), TCP, T, d, N, e, 1, s, 3 * B));
  }
  Planes pZv, pZq, pS, pPQg, pPVg;
  HCA_TRY(make_planes(pZv, dZv, d, 3 * BN, d, w, s));
  HCA_TRY(make_planes(pZq, dZq, d, 3 * BT, d, w, s));
  {  // dPQ = dZq + C dZv
    TcEpilogue e; e.D = dPQ; e.ldd = d; e.d_batch_stride = (int64_t)T * d;
    e.aux = dZq; e.aux_ld = d; e.aux_batch_stride = (int64_t)T * d; e.aux_nbatch = 3 * B; e.aux_mode = TC_AUX_ADD;
    HCA_TRY(launch_gemm_tc(view(cp.C, 0, T, 3 * B, false), view(pZv, 0, N, 3 * B, true), TCP, T, d, N, e, 1, s, 3 * B));
  }
  for (int l = 0; l < 3; ++l) {  // dPV = sum_l dZv_l + C_l^T dZq_l
    TcEpilogue e; e.D = dPV; e.ldd = d; e.d_batch_stride = (int64_t)N * d; e.accumulate = (l > 0);
    e.aux = dZv + l * BN * d; e.aux_ld = d; e.aux_batch_stride = (int64_t)N * d; e.aux_nbatch = B; e.aux_mode = TC_AUX_ADD;
    HCA_TRY(launch_gemm_tc(view(cp.C, l * BT, T, B, true), view(pZq, l * BT, T, B, true), TCP, N, d, T, e, 1, s, B));
  }
  {  // dS = (PQ dZv^T + dZq PV^T) * (1 - C^2): two operand pairs chained along K in one accumulator
    TcEpilogue e; e.D = dS; e.ldd = N; e.d_batch_stride = (int64_t)T * N;
    e.aux = C; e.aux_ld = N; e.aux_batch_stride = (int64_t)T * N; e.aux_nbatch = 3 * B; e.aux_mode = TC_AUX_MUL_1MX2;
    const TcOperand a2 = view(pZq, 0, T, 3 * B, false), b2 = view(cp.PV, 0, N, B, false);
    HCA_TRY(launch_gemm_tc(view(cp.PQ, 0, T, 3 * B, false), view(pZv, 0, N, 3 * B, false), TCP, T, N, d, e, 1, s, 3 * B, &a2, &b2, d));
  }
  HCA_TRY(make_planes(pS, dS, N, 3 * BT, N, w, s));
  {  // dQ = dS V + aq x gq
    TcEpilogue e; e.D = dQ; e.ldd = d; e.d_batch_stride = (int64_t)T * d;
    e.rowv = aq; e.rowv_batch_stride = T; e.r1col = gqhat; e.r1col_batch_stride = d;
    HCA_TRY(launch_gemm_tc(view(pS, 0, T, 3 * B, false), view(cp.V, 0, N, B, true), TCP, T, d, N, e, 1, s, 3 * B));
  }
  HCA_TRY(make_planes(pPQg, dPQ, d, 3 * BT, d, w, s));
  {  // dQ += dPQ Wq
    TcEpilogue e; e.D = dQ; e.ldd = d; e.accumulate = 1;
    HCA_TRY(launch_gemm_tc(view(pPQg, 0, (int)(3 * BT), 1, false), view(cp.Wq, 0, d, 1, true), TCP, (int)(3 * BT), d, d, e, 1, s));
  }
  {  // dWq = dPQ^T Q over batch, time and the three levels at once (K = 3*B*T, split-K)
    const int sk = tc_splitk(d, d, (int)(3 * BT));
    if (sk > 1) HCA_TRY(zero_async(dWq, (size_t)d * d * 4, s));
    TcEpilogue e; e.D = dWq; e.ldd = d;
    HCA_TRY(launch_gemm_tc(view(pPQg, 0, (int)(3 * BT), 1, true), view(cp.Q, 0, (int)(3 * BT), 1, true), TCP, d, d, (int)(3 * BT), e, sk, s));
  }
  HCA_TRY(zero_async(dbq, (size_t)d * 4, s));
  HCA_TRY(launch_colsum(dPQ, d, 3 * BT, d, dbq, s));
  HCA_TRY(make_planes(pPVg, dPV, d, BN, d, w, s));
  {  // dWv = dPV^T V   (K = B*N, split-K)
    const int sk = tc_splitk(d, d, (int)BN);
    if (sk > 1) HCA_TRY(zero_async(dWv, (size_t)d * d * 4, s));
    TcEpilogue e; e.D = dWv; e.ldd = d;
    HCA_TRY(launch_gemm_tc(view(pPVg, 0, (int)BN, 1, true), view(cp.V, 0, (int)BN, 1, true), TCP, d, d, (int)BN, e, sk, s));
  }
  HCA_TRY(zero_async(dbv, (size_t)d * 4, s));
  HCA_TRY(launch_colsum(dPV, d, BN, d, dbv, s));
  if (dV) {  // only when the image features require grad (--vgg_train true)
    {
      TcEpilogue e; e.D = dV; e.ldd = d;
      HCA_TRY(launch_gemm_tc(view(pPVg, 0, (int)BN, 1, false), view(cp.Wv, 0, d, 1, true), TCP, (int)BN, d, d, e, 1, s));
    }
    for (int l = 0; l < 3; ++l) {  // dV += dS_l^T Q_l + av_l x gv_l
      TcEpilogue e; e.D = dV; e.ldd = d; e.d_batch_stride = (int64_t)N * d; e.accumulate = 1;
      e.rowv = av + l * BN; e.rowv_batch_stride = N; e.r1col = gvhat + (int64_t)l * B * d; e.r1col_batch_stride = d;
      HCA_TRY(launch_gemm_tc(view(pS, l * BT, T, B, true), view(cp.Q, l * BT, T, B, true), TCP, N, d, T, e, 1, s, B));
    }
  }
  return 0;
}

bool v_is_dense(int64_t sb, int64_t sn, int64_t sd, int N, int d) { return sd == 1 && sn == d && sb == (int64_t)N * d; }

}  // namespace
}  // namespace hca

extern "C" size_t hca_coattn_workspace(int B, int N, int T, int d, int need_dv) {
  using hca::align_up;
  (void)need_dv;
  const size_t b = (size_t)B;
  size_t s = 0;
  s += align_up(b * N * d * 4);                                // dense copy of V
  s += align_up(3 * b * N * 4) + align_up(3 * b * T * 4);      // sv/dsv, sq/dsq
  s += align_up(3 * b * N * d * 4);                            // dZv
  s += 2 * align_up(3 * b * T * d * 4);                        // dZq, dPQ
  s += align_up(b * N * d * 4);                                // dPV
  s += align_up(3 * b * T * N * 4);                            // dS
  size_t sc = hca::dense_scratch_bytes((int)(b * N), d, d);                    // PV, dV
  sc = std::max(sc, hca::dense_scratch_bytes(d, d, (int)(b * N)));            // dWv (K = B*N)
  sc = std::max(sc, hca::dense_scratch_bytes((int)(3 * b * T), d, d));        // dQ += dPQ Wq
  sc = std::max(sc, hca::dense_scratch_bytes(d, d, (int)(b * T)));            // dWq
  s += sc;
  // tensor-core path: bf16x2 planes of V, Q, C, PV, PQ, dZv, dZq, dS, dPQ, dPV and the two weights
  const size_t pl = 2 * 2 * (6 * b * N * d + 12 * b * T * d + 6 * b * T * (N + 8) + 2 * (size_t)d * d) + 64 * 256;
  return s + pl + 1024;
}

extern "C" int hca_coattn_fwd(const float* V, int64_t v_sb, int64_t v_sn, int64_t v_sd, const float* q0, const float* q1,
                              const float* q2, const float* Wv, const float* bv, const float* Wq, const float* bq,
                              const float* wv, const float* cv, const float* wq, const float* cq, float* vhat, float* qhat,
                              float* PV, float* PQ, float* C, float* av, float* aq, int B, int N, int T, int d, void* ws,
                              size_t ws_bytes, void* stream) {
  using namespace hca;
  cudaStream_t s = (cudaStream_t)stream;
  HCA_CHECK_ARG(V && q0 && q1 && q2 && Wv && bv && Wq && bq && wv && cv && wq && cq, "coattn_fwd: null input");
  HCA_CHECK_ARG(vhat && qhat && PV && PQ && C && av && aq, "coattn_fwd: null output");
  HCA_CHECK_ARG(B > 0 && N > 0 && T > 0 && d > 0 && d % 4 == 0, "coattn_fwd: bad sizes B=%d N=%d T=%d d=%d", B, N, T, d);
  HCA_CHECK_ARG(3 * B <= 65535, "coattn_fwd: batch too large for one call (B=%d)", B);
  Workspace w(ws, ws_bytes);
  const float* Vd = V;
  if (!v_is_dense(v_sb, v_sn, v_sd, N, d)) {
    float* vc = w.take<float>((size_t)B * N * d);
    if (!vc) return set_err(HCA_ERR_WORKSPACE, "coattn_fwd: workspace too small");
    gather_strided_kernel<<<ew_grid((int64_t)B * N * d), 256, 0, s>>>(V, v_sb, v_sn, v_sd, vc, B, N, d);
    HCA_LAUNCHED();
    Vd = vc;
  }
  float* sv = w.take<float>((size_t)3 * B * N);
  float* sq = w.take<float>((size_t)3 * B * T);
  if (!sv || !sq) return set_err(HCA_ERR_WORKSPACE, "coattn_fwd: workspace too small");
  const float* q[3] = {q0, q1, q2};
  const int64_t BT = (int64_t)B * T, BN = (int64_t)B * N;
  if (tc_path_ok(N, d)) {
    HCA_TRY(zero_async(sv, (size_t)3 * B * N * 4, s));
    HCA_TRY(zero_async(sq, (size_t)3 * B * T * 4, s));
    HCA_TRY(coattn_fwd_tc(Vd, q, Wv, bv, Wq, bq, wv, wq, PV, PQ, C, sv, sq, B, N, T, d, w, s));
    const size_t smem_tc = (64 + (size_t)max(N, T)) * sizeof(float);
    attn_finish_kernel<<<3 * B, 256, smem_tc, s>>>(sv, sq, cv, cq, Vd, q0, q1, q2, av, aq, vhat, qhat, B, N, T, d);
    HCA_LAUNCHED();
    return 0;
  }

  // projections
  {
    DenseEpi e; e.bias = bv;
    HCA_TRY(dense_nt(Vd, d, Wv, d, PV, d, (int)BN, d, d, e, w, s));
  }
  for (int l = 0; l < 3; ++l) {
    DenseEpi e; e.bias = bq;
    HCA_TRY(dense_nt(q[l], d, Wq, d, PQ + l * BT * d, d, (int)BT, d, d, e, w, s));
  }
  // affinity C_l = tanh(Q_l V^T), batched over samples
  for (int l = 0; l < 3; ++l) {
    GemmParams g;
    g.A = {q[l], (int64_t)T * d, d, 1, 0};
    g.B = {Vd, (int64_t)N * d, d, 1, 0};
    g.M = T; g.N = N; g.K = d; g.batch = B;
    g.D = C + l * BT * N; g.d_sb = (int64_t)T * N; g.d_sm = N; g.d_sn = 1;
    g.act_tanh = 1;
    HCA_TRY(launch_gemm_ffma(g, false, s));
  }
  HCA_TRY(zero_async(sv, (size_t)3 * B * N * 4, s));
  HCA_TRY(zero_async(sq, (size_t)3 * B * T * 4, s));
  {  // sq[z][t] = sum_j tanh(PQ + C PV)[t][j] * wq[j]
    GemmParams g;
    g.A = {C, (int64_t)T * N, N, 1, 0};
    g.B = {PV, (int64_t)N * d, 1, d, B};
    g.M = T; g.N = d; g.K = N; g.batch = 3 * B;
    g.add = {PQ, (int64_t)T * d, d, 1, 0};
    g.act_tanh = 1;
    g.epi = EPI_ROWDOT; g.colv = wq; g.red_row = sq; g.red_row_sb = T;
    HCA_TRY(launch_gemm_ffma(g, false, s));
  }
  {  // sv[z][n] = sum_j tanh(PV + C^T PQ)[n][j] * wv[j]
    GemmParams g;
    g.A = {C, (int64_t)T * N, 1, N, 0};
    g.B = {PQ, (int64_t)T * d, 1, d, 0};
    g.M = N; g.N = d; g.K = T; g.batch = 3 * B;
    g.add = {PV, (int64_t)N * d, d, 1, B};
    g.act_tanh = 1;
    g.epi = EPI_ROWDOT; g.colv = wv; g.red_row = sv; g.red_row_sb = N;
    HCA_TRY(launch_gemm_ffma(g, false, s));
  }
  const size_t smem = (64 + (size_t)max(N, T)) * sizeof(float);
  attn_finish_kernel<<<3 * B, 256, smem, s>>>(sv, sq, cv, cq, Vd, q0, q1, q2, av, aq, vhat, qhat, B, N, T, d);
  HCA_LAUNCHED();
  return 0;
}

extern "C" int hca_coattn_bwd(const float* V, int64_t v_sb, int64_t v_sn, int64_t v_sd, const float* q0, const float* q1,
                              const float* q2, const float* Wv, const float* Wq, const float* wv, const float* wq,
                              const float* PV, const float* PQ, const float* C, const float* av, const float* aq,
                              const float* gvhat, const float* gqhat, float* dV, float* dQ, float* dWv, float* dbv, float* dWq,
                              float* dbq, float* dwv, float* dcv, float* dwq, float* dcq, int B, int N, int T, int d, void* ws,
                              size_t ws_bytes, void* stream) {
  using namespace hca;
  cudaStream_t s = (cudaStream_t)stream;
  HCA_CHECK_ARG(V && q0 && q1 && q2 && Wv && Wq && wv && wq && PV && PQ && C && av && aq && gvhat && gqhat, "coattn_bwd: null input");
  HCA_CHECK_ARG(dQ && dWv && dbv && dWq && dbq && dwv && dcv && dwq && dcq, "coattn_bwd: null output");
  HCA_CHECK_ARG(B > 0 && N > 0 && T > 0 && d > 0 && d % 4 == 0, "coattn_bwd: bad sizes");
  HCA_CHECK_ARG(3 * B <= 65535, "coattn_bwd: batch too large for one call (B=%d)", B);
  Workspace w(ws, ws_bytes);
  const float* Vd = V;
  if (!v_is_dense(v_sb, v_sn, v_sd, N, d)) {
    float* vc = w.take<float>((size_t)B * N * d);
    if (!vc) return set_err(HCA_ERR_WORKSPACE, "coattn_bwd: workspace too small");
    gather_strided_kernel<<<ew_grid((int64_t)B * N * d), 256, 0, s>>>(V, v_sb, v_sn, v_sd, vc, B, N, d);
    HCA_LAUNCHED();
    Vd = vc;
  }
  const int64_t BT = (int64_t)B * T, BN = (int64_t)B * N;
  float* dsv = w.take<float>((size_t)3 * BN);
  float* dsq = w.take<float>((size_t)3 * BT);
  float* dZv = w.take<float>((size_t)3 * BN * d);
  float* dZq = w.take<float>((size_t)3 * BT * d);
  float* dPQ = w.take<float>((size_t)3 * BT * d);
  float* dPV = w.take<float>((size_t)BN * d);
  float* dS = w.take<float>((size_t)3 * BT * N);
  if (!dS) return set_err(HCA_ERR_WORKSPACE, "coattn_bwd: workspace too small (%zu bytes)", ws_bytes);
  const float* q[3] = {q0, q1, q2};

  HCA_TRY(zero_async(dcv, 4, s));
  HCA_TRY(zero_async(dcq, 4, s));
  HCA_TRY(zero_async(dwv, (size_t)d * 4, s));
  HCA_TRY(zero_async(dwq, (size_t)d * 4, s));
  const size_t smem = (64 + (size_t)max(N, T)) * sizeof(float);
  attn_bwd_prep_kernel<<<3 * B, 256, smem, s>>>(av, aq, Vd, q0, q1, q2, gvhat, gqhat, dsv, dsq, dcv, dcq, B, N, T, d);
  HCA_LAUNCHED();
  if (tc_path_ok(N, d))
    return coattn_bwd_tc(Vd, q, Wv, Wq, wv, wq, PV, PQ, C, av, aq, gvhat, gqhat, dsv, dsq, dZv, dZq, dPQ, dPV, dS, dV, dQ, dWv, dbv,
                         dWq, dbq, dwv, dwq, B, N, T, d, w, s);
  {  // dZv = (dsv x wv) * (1 - Hv^2), Hv recomputed; dwv += Hv^T dsv
    GemmParams g;
    g.A = {C, (int64_t)T * N, 1, N, 0};
    g.B = {PQ, (int64_t)T * d, 1, d, 0};
    g.M = N; g.N = d; g.K = T; g.batch = 3 * B;
    g.add = {PV, (int64_t)N * d, d, 1, B};
    g.act_tanh = 1;
    g.epi = EPI_DZ; g.rowv = dsv; g.rowv_sb = N; g.colv = wv; g.red_col = dwv;
    g.D = dZv; g.d_sb = (int64_t)N * d; g.d_sm = d; g.d_sn = 1;
    HCA_TRY(launch_gemm_ffma(g, false, s));
  }
  {  // dZq likewise
    GemmParams g;
    g.A = {C, (int64_t)T * N, N, 1, 0};
    g.B = {PV, (int64_t)N * d, 1, d, B};
    g.M = T; g.N = d; g.K = N; g.batch = 3 * B;
    g.add = {PQ, (int64_t)T * d, d, 1, 0};
    g.act_tanh = 1;
    g.epi = EPI_DZ; g.rowv = dsq; g.rowv_sb = T; g.colv = wq; g.red_col = dwq;
    g.D = dZq; g.d_sb = (int64_t)T * d; g.d_sm = d; g.d_sn = 1;
    HCA_TRY(launch_gemm_ffma(g, false, s));
  }
  {  // dPQ = dZq + C dZv
    GemmParams g;
    g.A = {C, (int64_t)T * N, N, 1, 0};
    g.B = {dZv, (int64_t)N * d, 1, d, 0};
    g.M = T; g.N = d; g.K = N; g.batch = 3 * B;
    g.add = {dZq, (int64_t)T * d, d, 1, 0};
    g.D = dPQ; g.d_sb = (int64_t)T * d; g.d_sm = d; g.d_sn = 1;
    HCA_TRY(launch_gemm_ffma(g, false, s));
  }
  for (int l = 0; l < 3; ++l) {  // dPV = sum_l dZv_l + C_l^T dZq_l
    GemmParams g;
    g.A = {C + l * BT * N, (int64_t)T * N, 1, N, 0};
    g.B = {dZq + l * BT * d, (int64_t)T * d, 1, d, 0};
    g.M = N; g.N = d; g.K = T; g.batch = B;
    g.add = {dZv + l * BN * d, (int64_t)N * d, d, 1, 0};
    g.D = dPV; g.d_sb = (int64_t)N * d; g.d_sm = d; g.d_sn = 1;
    g.accumulate = (l > 0);
    HCA_TRY(launch_gemm_ffma(g, false, s));
  }
  {  // dS = (PQ dZv^T + dZq PV^T) * (1 - C^2)
    GemmParams g;
    g.A = {PQ, (int64_t)T * d, d, 1, 0};
    g.B = {dZv, (int64_t)N * d, d, 1, 0};
    g.A2 = {dZq, (int64_t)T * d, d, 1, 0};
    g.B2 = {PV, (int64_t)N * d, d, 1, B};
    g.M = T; g.N = N; g.K = d; g.K2 = d; g.batch = 3 * B;
    g.mulx = {C, (int64_t)T * N, N, 1, 0};
    g.D = dS; g.d_sb = (int64_t)T * N; g.d_sm = N; g.d_sn = 1;
    HCA_TRY(launch_gemm_ffma(g, false, s));
  }
  {  // dQ = dS V + aq x gq
    GemmParams g;
    g.A = {dS, (int64_t)T * N, N, 1, 0};
    g.B = {Vd, (int64_t)N * d, 1, d, B};
    g.M = T; g.N = d; g.K = N; g.batch = 3 * B;
    g.r1_row = aq; g.r1r_sb = T; g.r1_col = gqhat; g.r1c_sb = d;
    g.D = dQ; g.d_sb = (int64_t)T * d; g.d_sm = d; g.d_sn = 1;
    HCA_TRY(launch_gemm_ffma(g, false, s));
  }
  {  // dQ += dPQ Wq
    DenseEpi e; e.accumulate = 1;
    HCA_TRY(dense_nn(dPQ, d, Wq, d, dQ, d, (int)(3 * BT), d, d, e, w, s));
  }
  // weight gradients (summed over batch and levels)
  for (int l = 0; l < 3; ++l)
    HCA_TRY(dense_tn(dPQ + l * BT * d, d, q[l], d, dWq, d, d, d, (int)BT, l == 0, w, s));
  HCA_TRY(zero_async(dbq, (size_t)d * 4, s));
  HCA_TRY(launch_colsum(dPQ, d, 3 * BT, d, dbq, s));
  HCA_TRY(dense_tn(dPV, d, Vd, d, dWv, d, d, d, (int)BN, true, w, s));
  HCA_TRY(zero_async(dbv, (size_t)d * 4, s));
  HCA_TRY(launch_colsum(dPV, d, BN, d, dbv, s));
  if (dV) {  // only when the image features require grad (--vgg_train true)
    DenseEpi e;
    HCA_TRY(dense_nn(dPV, d, Wv, d, dV, d, (int)BN, d, d, e, w, s));
    for (int l = 0; l < 3; ++l) {  // dV += dS_l^T Q_l + av_l x gv_l
      GemmParams g;
      g.A = {dS + l * BT * N, (int64_t)T * N, 1, N, 0};
      g.B = {q[l], (int64_t)T * d, 1, d, 0};
      g.M = N; g.N = d; g.K = T; g.batch = B;
      g.r1_row = av + l * BN; g.r1r_sb = N; g.r1_col = gvhat + (int64_t)l * B * d; g.r1c_sb = d;
      g.D = dV; g.d_sb = (int64_t)N * d; g.d_sm = d; g.d_sn = 1;
      g.accumulate = 1;
      HCA_TRY(launch_gemm_ffma(g, false, s));
    }
  }
  return 0;
}
