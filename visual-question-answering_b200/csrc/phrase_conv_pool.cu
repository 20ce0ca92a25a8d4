// PhraseConvPool forward / backward (replaces reference model.py:304-334).
//
// The three Conv1d's (k = 1, 2, 3 with ConstantPad1d (0,0) / (1,0) / (1,1)) are dense contractions over a
// row-shifted view of the word embeddings:
//     Acat[r=(b,t)] = [ x[b,t-1] | x[b,t] | x[b,t+1] ]            (zeros outside [0,T))
//     uni = Acat[:, E:2E] . Wr1^T      bi = Acat[:, 0:2E] . Wr2^T      tri = Acat[:, 0:3E] . Wr3^T
// with Wr_k[o][j*E + c] = W_k[o][c][j] (tap-major repack of the [C_out, C_in, k] conv weight).
// Bias + tanh are fused into the GEMM epilogue; the pool kernel then takes the max over CONSECUTIVE
// channel triples of [uni|bi|tri] (the reshape at model.py:329), records the uint8 argmax (first index
// on ties, like MaxPool2d) and zeroes rows t >= len (model.py:287-292).
//
// The pre-activations must be fp32-grade: an argmax flip re-routes a gradient element,
// so the forward products run on the exact-fp32 GEMM path.  The backward products (dgrad, wgrad) only
// need the 1e-3 budget and may use the tensor-core path.
#include <algorithm>
#include "common.cuh"
#include "gemm_ffma.cuh"
#include "dense.cuh"
#include "gemm_tc.cuh"
#include "util_kernels.cuh"

namespace hca {
namespace {

// Acat[r][j*E + c] = x[b][t+j-1][c]   (float4 granularity)
__global__ void __launch_bounds__(256) im2col3_kernel(const float4* __restrict__ x, float4* __restrict__ acat, int B, int T, int E4) {
  const int64_t total = (int64_t)B * T * 3 * E4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % E4);
    const int j = (int)((i / E4) % 3);
    const int64_t r = i / (3 * E4);
    const int t = (int)(r % T) + j - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t >= 0 && t < T) v = __ldg(x + (r + j - 1) * E4 + c);
    acat[i] = v;
  }
}

// dx[b][t] = dA[r][E:2E] + dA[r+1][0:E] (t+1<T) + dA[r-1][2E:3E] (t>0)
__global__ void __launch_bounds__(256) col2im3_kernel(const float4* __restrict__ dA, float4* __restrict__ dx, int B, int T, int E4) {
  const int64_t total = (int64_t)B * T * E4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % E4);
    const int64_t r = i / E4;
    const int t = (int)(r % T);
    float4 v = dA[(r * 3 + 1) * E4 + c];
    if (t + 1 < T) {
      const float4 u = dA[((r + 1) * 3 + 0) * E4 + c];
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    if (t > 0) {
      const float4 u = dA[((r - 1) * 3 + 2) * E4 + c];
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    dx[i] = v;
  }
}

// wr[o][j*E + c] = w[o][c][j]   (to_conv == false)      or      w[o][c][j] = wr[o][j*E + c]  (to_conv == true)
__global__ void __launch_bounds__(256) repack_conv_w_kernel(const float* __restrict__ src, float* __restrict__ dst, int E, int k, bool to_conv) {
  const int64_t total = (int64_t)E * E * k;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // i indexes the tap-major layout [o][j][c]
    const int c = (int)(i % E);
    const int j = (int)((i / E) % k);
    const int64_t o = i / ((int64_t)E * k);
    const int64_t conv_idx = (o * E + c) * k + j;
    if (to_conv) dst[conv_idx] = src[i];
    else dst[i] = src[conv_idx];
  }
}

// Max-pool argmax must match the reference bit for bit, but the tensor-core conv (bf16x3 operand split, fp32 TMEM
// accumulation) carries ~3e-6 of error.  So the pool kernel records every element whose top-2 gap is below TIE_TOL
// (two orders of magnitude above that error) and fixup_ties_kernel recomputes just those elements in exact fp32.
constexpr float TIE_TOL = 2e-4f;

// cat [R, 3E] (post tanh) -> out [R, E], idx [R, E]; rows t >= len zeroed.  tie_list/tie_count may be null.
__global__ void __launch_bounds__(256) pool3_fwd_kernel(const float* __restrict__ cat, const int64_t* __restrict__ lens,
                                                        float* __restrict__ out, uint8_t* __restrict__ idx, int B, int T, int E,
                                                        int* __restrict__ tie_list, int* __restrict__ tie_count, int tie_cap) {
  const int64_t total = (int64_t)B * T * E;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / E;
    const int e = (int)(i - r * E);
    const int b = (int)(r / T), t = (int)(r % T);
    float best = 0.f;
    int bi = 0;
    if (!lens || t < lens[b]) {
      const float* p = cat + r * 3 * (int64_t)E + 3 * e;
      const float v0 = p[0], v1 = p[1], v2 = p[2];
      best = v0;
      // MaxPool2d semantics: a later element wins only if strictly greater, or is NaN
      if (v1 > best || v1 != v1) { best = v1; bi = 1; }
      if (v2 > best || v2 != v2) { best = v2; bi = 2; }
      if (tie_list) {
        const float lo = fminf(fminf(v0, v1), v2);
        const float mid = v0 + v1 + v2 - best - lo;            // middle value (approximate is fine: only a trigger)
        if (best - mid < TIE_TOL) {
          const int slot = atomicAdd(tie_count, 1);
          if (slot < tie_cap) tie_list[slot] = (int)i;
        }
      }
    }
    out[i] = best;
    idx[i] = (uint8_t)bi;
  }
}

// one warp per listed element: the three pre-activations of its channel triple in plain fp32, then tanh and the max again
__global__ void __launch_bounds__(256) fixup_ties_kernel(const int* __restrict__ tie_list, const int* __restrict__ tie_count, int tie_cap,
                                                         const float* __restrict__ acat, const float* __restrict__ w1,
                                                         const float* __restrict__ wr2, const float* __restrict__ wr3,
                                                         const float* __restrict__ b1, const float* __restrict__ b2,
                                                         const float* __restrict__ b3, float* __restrict__ out,
                                                         uint8_t* __restrict__ idx, int E) {
  const int n = min(*tie_count, tie_cap);
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n; w += warps) {
    const int i = tie_list[w];
    const int64_t r = i / E;
    const int e = i - (int)r * E;
    float v[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int c = 3 * e + j;                    // channel of the concatenated [uni|bi|tri] axis
      const int k = c / E + 1, o = c % E;         // which conv, which output channel
      const float* wrow = (k == 1 ? w1 : (k == 2 ? wr2 : wr3)) + (int64_t)o * k * E;
      const float* arow = acat + r * 3 * (int64_t)E + (k == 1 ? E : 0);
      float acc = 0.f;
      for (int kk = lane; kk < k * E; kk += 32) acc = fmaf(arow[kk], wrow[kk], acc);
      acc = warp_sum(acc);
      v[j] = tanhf(acc + (k == 1 ? b1 : (k == 2 ? b2 : b3))[o]);
    }
    float best = v[0];
    int bi = 0;
    if (v[1] > best || v[1] != v[1]) { best = v[1]; bi = 1; }
    if (v[2] > best || v[2] != v[2]) { best = v[2]; bi = 2; }
    if (lane == 0) {
      out[i] = best;
      idx[i] = (uint8_t)bi;
    }
  }
}

// dcat[r][3e+j] = (j == idx) ? dout * (1 - out^2) : 0 ; masked rows all zero
__global__ void __launch_bounds__(256) pool3_bwd_kernel(const float* __restrict__ out, const uint8_t* __restrict__ idx,
                                                        const float* __restrict__ dout, const int64_t* __restrict__ lens,
                                                        float* __restrict__ dcat, int B, int T, int E) {
  const int64_t total = (int64_t)B * T * E;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / E;
    const int e = (int)(i - r * E);
    const int b = (int)(r / T), t = (int)(r % T);
    float g = 0.f;
    int j = 0;
    if (!lens || t < lens[b]) {
      const float o = out[i];
      g = dout[i] * (1.f - o * o);
      j = idx[i];
    }
    float* p = dcat + r * 3 * (int64_t)E + 3 * e;
    p[0] = j == 0 ? g : 0.f;
    p[1] = j == 1 ? g : 0.f;
    p[2] = j == 2 ? g : 0.f;
  }
}


struct ConvWs {
  Workspace w;
  float *acat, *cat, *dA, *wr2, *wr3, *dwr2, *dwr3;
  int *tie_count, *tie_list;
  int tie_cap;
  bool ok;
  ConvWs(void* p, size_t bytes) : w(p, bytes) {}
};
ConvWs carve(void* ws, size_t bytes, int B, int T, int E) {
  ConvWs c(ws, bytes);
  Workspace& w = c.w;
  const size_t R = (size_t)B * T;
  c.acat = w.take<float>(R * 3 * E);
  c.cat = w.take<float>(R * 3 * E);     // fwd: tanh(conv) ; bwd: dcat
  c.dA = w.take<float>(R * 3 * E);
  c.wr2 = w.take<float>((size_t)E * 2 * E);
  c.wr3 = w.take<float>((size_t)E * 3 * E);
  c.dwr2 = w.take<float>((size_t)E * 2 * E);
  c.dwr3 = w.take<float>((size_t)E * 3 * E);
  c.tie_cap = (int)(R * E / 8 + 1024);
  c.tie_count = w.take<int>(64);
  c.tie_list = w.take<int>((size_t)c.tie_cap);
  c.ok = c.tie_list != nullptr;
  return c;
}

}  // namespace
}  // namespace hca

extern "C" size_t hca_phrase_conv_pool_workspace(int B, int T, int E) {
  using hca::align_up;
  const size_t R = (size_t)B * T;
  return 3 * align_up(R * 3 * E * 4) + 2 * (align_up((size_t)E * 2 * E * 4) + align_up((size_t)E * 3 * E * 4)) + 1024 +
         align_up((R * E / 8 + 1024) * 4) +                                             // near-tie list
         3 * 2 * (align_up(R * 3 * E) + 3 * align_up((size_t)E * 3 * E)) +               // bf16x3 planes of Acat and the weights

         std::max(hca::dense_scratch_bytes(E, 3 * E, (int)R), hca::dense_scratch_bytes((int)R, 3 * E, E));
}

extern "C" int hca_phrase_conv_pool_fwd(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                                        const float* w3, const float* b3, const int64_t* lens, float* out, uint8_t* idx,
                                        int B, int T, int E, void* ws, size_t ws_bytes, void* stream) {
  using namespace hca;
  cudaStream_t s = (cudaStream_t)stream;
  HCA_CHECK_ARG(x && w1 && b1 && w2 && b2 && w3 && b3 && out && idx, "phrase_conv_pool_fwd: null pointer");
  HCA_CHECK_ARG(B > 0 && T > 0 && E > 0 && E % 4 == 0, "phrase_conv_pool_fwd: bad sizes B=%d T=%d E=%d (E %% 4 == 0 required)", B, T, E);
  ConvWs c = carve(ws, ws_bytes, B, T, E);
  if (!c.ok) return set_err(HCA_ERR_WORKSPACE, "phrase_conv_pool_fwd: workspace too small (%zu bytes)", ws_bytes);
  const int R = B * T;
  im2col3_kernel<<<ew_grid((int64_t)R * 3 * E / 4), 256, 0, s>>>((const float4*)x, (float4*)c.acat, B, T, E / 4);
  HCA_LAUNCHED();
  repack_conv_w_kernel<<<ew_grid((int64_t)E * E * 2), 256, 0, s>>>(w2, c.wr2, E, 2, false);
  HCA_LAUNCHED();
  repack_conv_w_kernel<<<ew_grid((int64_t)E * E * 3), 256, 0, s>>>(w3, c.wr3, E, 3, false);
  HCA_LAUNCHED();
  const float* wr[3] = {w1, c.wr2, c.wr3};
  const float* bs[3] = {b1, b2, b3};
  const bool tc = use_tc() && tc_available() && (E % 8 == 0);      // TMA needs 16-byte aligned plane windows
  if (tc) {
    // tensor cores, bf16x3 operand split (6 MMAs per product, fp32-grade): Acat is split once and the three convs read
    // column windows of its planes; bias + tanh fused in the epilogue; near-ties are repaired exactly below
    const int P = 3;
    const int64_t lda = 3 * (int64_t)E, a_stride = (int64_t)R * lda;
    __nv_bfloat16* ap = c.w.take<__nv_bfloat16>((size_t)P * a_stride);
    if (!ap) return set_err(HCA_ERR_WORKSPACE, "phrase_conv_pool_fwd: workspace too small for operand planes");
    HCA_TRY(launch_split_planes(c.acat, lda, R, 3 * E, ap, lda, a_stride, P, s));
    for (int k = 1; k <= 3; ++k) {
      const int64_t ldw = (int64_t)k * E, w_stride = (int64_t)E * ldw;
      __nv_bfloat16* wp = c.w.take<__nv_bfloat16>((size_t)P * w_stride);
      if (!wp) return set_err(HCA_ERR_WORKSPACE, "phrase_conv_pool_fwd: workspace too small for weight planes");
      HCA_TRY(launch_split_planes(wr[k - 1], ldw, E, k * E, wp, ldw, w_stride, P, s));
      TcOperand A, Bw;
      A.planes = ap + (k == 1 ? E : 0); A.ld = lda; A.plane_stride = a_stride; A.rows = R; A.cols = k * E;
      Bw.planes = wp; Bw.ld = ldw; Bw.plane_stride = w_stride; Bw.rows = E; Bw.cols = k * E;
      TcEpilogue ep;
      ep.D = c.cat + (k - 1) * E; ep.ldd = lda; ep.bias = bs[k - 1]; ep.act_tanh = 1;
      HCA_TRY(launch_gemm_tc(A, Bw, P, R, E, k * E, ep, 1, s));
    }
    HCA_TRY(zero_async(c.tie_count, sizeof(int), s));
    pool3_fwd_kernel<<<ew_grid((int64_t)R * E), 256, 0, s>>>(c.cat, lens, out, idx, B, T, E, c.tie_list, c.tie_count, c.tie_cap);
    HCA_LAUNCHED();
    fixup_ties_kernel<<<148 * 2, 256, 0, s>>>(c.tie_list, c.tie_count, c.tie_cap, c.acat, w1, c.wr2, c.wr3, b1, b2, b3, out, idx, E);
    HCA_LAUNCHED();
    return 0;
  }
  // exact-fp32 CUDA-core path: three GEMMs, bias + tanh fused, writing the column blocks of cat [R, 3E]
  for (int k = 1; k <= 3; ++k) {
    GemmParams g;
    const int a_off = (k == 1) ? E : 0;
    g.A = {c.acat + a_off, 0, 3 * (int64_t)E, 1, 0};
    g.B = {wr[k - 1], 0, (int64_t)k * E, 1, 0};
    g.M = R; g.N = E; g.K = k * E;
    g.D = c.cat + (k - 1) * E; g.d_sm = 3 * (int64_t)E; g.d_sn = 1;
    g.bias = bs[k - 1];
    g.act_tanh = 1;
    HCA_TRY(launch_gemm_ffma(g, true, s));
  }
  pool3_fwd_kernel<<<ew_grid((int64_t)R * E), 256, 0, s>>>(c.cat, lens, out, idx, B, T, E, nullptr, nullptr, 0);
  HCA_LAUNCHED();
  return 0;
}

extern "C" int hca_phrase_conv_pool_bwd(const float* x, const float* w1, const float* w2, const float* w3, const float* out,
                                        const uint8_t* idx, const float* dout, const int64_t* lens, float* dx, float* dw1,
                                        float* db1, float* dw2, float* db2, float* dw3, float* db3, int B, int T, int E,
                                        void* ws, size_t ws_bytes, void* stream) {
  using namespace hca;
  cudaStream_t s = (cudaStream_t)stream;
  HCA_CHECK_ARG(x && w1 && w2 && w3 && out && idx && dout && dw1 && db1 && dw2 && db2 && dw3 && db3, "phrase_conv_pool_bwd: null pointer");
  HCA_CHECK_ARG(B > 0 && T > 0 && E > 0 && E % 4 == 0, "phrase_conv_pool_bwd: bad sizes");
  ConvWs c = carve(ws, ws_bytes, B, T, E);
  if (!c.ok) return set_err(HCA_ERR_WORKSPACE, "phrase_conv_pool_bwd: workspace too small (%zu bytes)", ws_bytes);
  const int R = B * T;
  float* dcat = c.cat;
  im2col3_kernel<<<ew_grid((int64_t)R * 3 * E / 4), 256, 0, s>>>((const float4*)x, (float4*)c.acat, B, T, E / 4);
  HCA_LAUNCHED();
  pool3_bwd_kernel<<<ew_grid((int64_t)R * E), 256, 0, s>>>(out, idx, dout, lens, dcat, B, T, E);
  HCA_LAUNCHED();
  // bias gradients: column sums of the three blocks of dcat
  float* dbs[3] = {db1, db2, db3};
  for (int k = 0; k < 3; ++k) {
    HCA_TRY(zero_async(dbs[k], (size_t)E * 4, s));
    HCA_TRY(launch_colsum(dcat + k * E, 3 * (int64_t)E, R, E, dbs[k], s));
  }
  // weight gradients in tap-major layout: dWr_k[o][kk] = sum_r dcat[r][(k-1)E + o] * Acat[r][a_off + kk]
  float* dwr[3] = {dw1, c.dwr2, c.dwr3};
  for (int k = 1; k <= 3; ++k) {
    const int a_off = (k == 1) ? E : 0;
    HCA_TRY(dense_tn(dcat + (k - 1) * E, 3 * (int64_t)E, c.acat + a_off, 3 * (int64_t)E, dwr[k - 1], (int64_t)k * E,
                     /*M=*/E, /*N=*/k * E, /*K=*/R, /*zero_first=*/true, c.w, s));
  }
  repack_conv_w_kernel<<<ew_grid((int64_t)E * E * 2), 256, 0, s>>>(c.dwr2, dw2, E, 2, true);
  HCA_LAUNCHED();
  repack_conv_w_kernel<<<ew_grid((int64_t)E * E * 3), 256, 0, s>>>(c.dwr3, dw3, E, 3, true);
  HCA_LAUNCHED();
  if (dx) {
    // dA[r][kk] = sum_o dcat[r][blk + o] * Wr_k[o][kk], accumulated over the three convs
    repack_conv_w_kernel<<<ew_grid((int64_t)E * E * 2), 256, 0, s>>>(w2, c.wr2, E, 2, false);
    HCA_LAUNCHED();
    repack_conv_w_kernel<<<ew_grid((int64_t)E * E * 3), 256, 0, s>>>(w3, c.wr3, E, 3, false);
    HCA_LAUNCHED();
    const float* wr[3] = {w1, c.wr2, c.wr3};
    for (int k = 3; k >= 1; --k) {   // tri first (covers all 3E columns, plain store), then bi, uni accumulate
      DenseEpi e;
      e.accumulate = (k != 3);
      HCA_TRY(dense_nn(dcat + (k - 1) * E, 3 * (int64_t)E, wr[k - 1], (int64_t)k * E, c.dA + ((k == 1) ? E : 0), 3 * (int64_t)E,
                       /*M=*/R, /*N=*/k * E, /*K=*/E, e, c.w, s));
    }
    col2im3_kernel<<<ew_grid((int64_t)R * E / 4), 256, 0, s>>>((const float4*)c.dA, (float4*)dx, B, T, E / 4);
    HCA_LAUNCHED();
  }
  return 0;
}
