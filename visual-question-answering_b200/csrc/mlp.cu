// MLPClassifier forward / backward (replaces reference model.py:414-434).
//
//   h_w = tanh(W_w (q_w + v_w))            h_p = tanh(W_p [q_p + v_p | h_w])
//   h_s = tanh(W_s [q_s + v_s | h_p])      logits = W_h h_s
// The concatenations are never materialised separately: h_w and h_p are written by the GEMM epilogues
// straight into the right halves of the next layer's input rows (xp, xs), which are also the tensors
// saved for backward.
#include <algorithm>
#include <initializer_list>
#include "common.cuh"
#include "dense.cuh"
#include "util_kernels.cuh"

namespace hca {
namespace {

// xw = q0+v0 ; xp[:, :d] = q1+v1 ; xs[:, :d] = q2+v2      (vhat, qhat are [3,B,d])
__global__ void __launch_bounds__(256) mlp_inputs_kernel(const float4* __restrict__ vhat, const float4* __restrict__ qhat,
                                                         float4* __restrict__ xw, float4* __restrict__ xp, float4* __restrict__ xs,
                                                         int B, int d4) {
  const int64_t per = (int64_t)B * d4, total = 3 * per;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int l = (int)(i / per);
    const int64_t r = (i % per) / d4;
    const int c = (int)(i % d4);
    const float4 a = vhat[i], b = qhat[i];
    const float4 v = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    if (l == 0) xw[r * d4 + c] = v;
    else if (l == 1) xp[r * 2 * d4 + c] = v;
    else xs[r * 2 * d4 + c] = v;
  }
}


}  // namespace
}  // namespace hca

extern "C" size_t hca_mlp_workspace(int B, int d, int mlp, int K) {
  using hca::align_up;
  size_t s = align_up((size_t)B * mlp * 4) + 2 * align_up((size_t)B * d * 4) + 1024;
  size_t sc = 0;
  for (size_t v : {hca::dense_scratch_bytes(B, K, mlp), hca::dense_scratch_bytes(K, mlp, B), hca::dense_scratch_bytes(B, mlp, K),
                   hca::dense_scratch_bytes(mlp, 2 * d, B), hca::dense_scratch_bytes(B, mlp, 2 * d), hca::dense_scratch_bytes(B, 2 * d, mlp),
                   hca::dense_scratch_bytes(d, 2 * d, B), hca::dense_scratch_bytes(B, 2 * d, d)})
    sc = std::max(sc, v);
  s += sc;
  return s;
}

extern "C" int hca_mlp_fwd(const float* vhat, const float* qhat, const float* Ww, const float* bw, const float* Wp,
                           const float* bp, const float* Ws, const float* bs, const float* Wh, const float* bh, float* logits,
                           float* xw, float* xp, float* xs, float* hs, int B, int d, int mlp, int K, void* ws, size_t ws_bytes,
                           void* stream) {
  using namespace hca;
  cudaStream_t s = (cudaStream_t)stream;
  HCA_CHECK_ARG(vhat && qhat && Ww && bw && Wp && bp && Ws && bs && Wh && bh && logits && xw && xp && xs && hs, "mlp_fwd: null pointer");
  HCA_CHECK_ARG(B > 0 && d > 0 && d % 4 == 0 && mlp > 0 && K > 0, "mlp_fwd: bad sizes B=%d d=%d mlp=%d K=%d", B, d, mlp, K);
  Workspace w(ws, ws_bytes);
  mlp_inputs_kernel<<<ew_grid(3LL * B * d / 4), 256, 0, s>>>((const float4*)vhat, (const float4*)qhat, (float4*)xw, (float4*)xp,
                                                           (float4*)xs, B, d / 4);
  HCA_LAUNCHED();
  DenseEpi e;
  e.act_tanh = 1;
  e.bias = bw;
  HCA_TRY(dense_nt(xw, d, Ww, d, xp + d, 2 * d, B, d, d, e, w, s));            // h_w -> xp[:, d:]
  e.bias = bp;
  HCA_TRY(dense_nt(xp, 2 * d, Wp, 2 * d, xs + d, 2 * d, B, d, 2 * d, e, w, s));  // h_p -> xs[:, d:]
  e.bias = bs;
  HCA_TRY(dense_nt(xs, 2 * d, Ws, 2 * d, hs, mlp, B, mlp, 2 * d, e, w, s));      // h_s
  e.act_tanh = 0;
  e.bias = bh;
  HCA_TRY(dense_nt(hs, mlp, Wh, mlp, logits, K, B, K, mlp, e, w, s));
  return 0;
}

extern "C" int hca_mlp_bwd(const float* dlogits, const float* Ww, const float* Wp, const float* Ws, const float* Wh,
                           const float* xw, const float* xp, const float* xs, const float* hs, float* g, float* dWw, float* dbw,
                           float* dWp, float* dbp, float* dWs, float* dbs, float* dWh, float* dbh, int B, int d, int mlp, int K,
                           void* ws, size_t ws_bytes, void* stream) {
  using namespace hca;
  cudaStream_t s = (cudaStream_t)stream;
  HCA_CHECK_ARG(dlogits && Ww && Wp && Ws && Wh && xw && xp && xs && hs, "mlp_bwd: null input");
  HCA_CHECK_ARG(g && dWw && dbw && dWp && dbp && dWs && dbs && dWh && dbh, "mlp_bwd: null output");
  HCA_CHECK_ARG(B > 0 && d > 0 && d % 4 == 0 && mlp > 0 && K > 0, "mlp_bwd: bad sizes");
  Workspace w(ws, ws_bytes);
  float* dzs = w.take<float>((size_t)B * mlp);
  float* dzp = w.take<float>((size_t)B * d);
  float* dzw = w.take<float>((size_t)B * d);
  if (!dzw) return set_err(HCA_ERR_WORKSPACE, "mlp_bwd: workspace too small");
  float* g_w = g;
  float* g_p = g + (size_t)B * d;
  float* g_s = g + (size_t)2 * B * d;

  // W_h
  HCA_TRY(dense_tn(dlogits, K, hs, mlp, dWh, mlp, K, mlp, B, true, w, s));
  HCA_TRY(zero_async(dbh, (size_t)K * 4, s));
  HCA_TRY(launch_colsum(dlogits, K, B, K, dbh, s));
  {  // dzs = (dlogits W_h) * (1 - h_s^2)
    DenseEpi e; e.mulx = hs; e.mulx_ld = mlp;
    HCA_TRY(dense_nn(dlogits, K, Wh, mlp, dzs, mlp, B, mlp, K, e, w, s));
  }
  // W_s
  HCA_TRY(dense_tn(dzs, mlp, xs, 2 * d, dWs, 2 * d, mlp, 2 * d, B, true, w, s));
  HCA_TRY(zero_async(dbs, (size_t)mlp * 4, s));
  HCA_TRY(launch_colsum(dzs, mlp, B, mlp, dbs, s));
  {  // dxs = dzs W_s : left half -> g_s, right half * (1 - h_p^2) -> dzp
    DenseEpi e;
    HCA_TRY(dense_nn(dzs, mlp, Ws, 2 * d, g_s, d, B, d, mlp, e, w, s));
    e.mulx = xs + d; e.mulx_ld = 2 * d;
    HCA_TRY(dense_nn(dzs, mlp, Ws + d, 2 * d, dzp, d, B, d, mlp, e, w, s));
  }
  // W_p
  HCA_TRY(dense_tn(dzp, d, xp, 2 * d, dWp, 2 * d, d, 2 * d, B, true, w, s));
  HCA_TRY(zero_async(dbp, (size_t)d * 4, s));
  HCA_TRY(launch_colsum(dzp, d, B, d, dbp, s));
  {
    DenseEpi e;
    HCA_TRY(dense_nn(dzp, d, Wp, 2 * d, g_p, d, B, d, d, e, w, s));
    e.mulx = xp + d; e.mulx_ld = 2 * d;
    HCA_TRY(dense_nn(dzp, d, Wp + d, 2 * d, dzw, d, B, d, d, e, w, s));
  }
  // W_w
  HCA_TRY(dense_tn(dzw, d, xw, d, dWw, d, d, d, B, true, w, s));
  HCA_TRY(zero_async(dbw, (size_t)d * 4, s));
  HCA_TRY(launch_colsum(dzw, d, B, d, dbw, s));
  {
    DenseEpi e;
    HCA_TRY(dense_nn(dzw, d, Ww, d, g_w, d, B, d, d, e, w, s));
  }
  return 0;
}
