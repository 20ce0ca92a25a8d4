// MLPClassifier forward / backward (replaces reference model.py:414-434).
//
//   h_w = tanh(W_w (q_w + v_w))            h_p = tanh(W_p [q_p + v_p | h_w])
//   h_s = tanh(W_s [q_s + v_s | h_p])      logits = W_h h_s
//
// Batch rows are few (B = 160) and the weights are the large operands, so every product puts the WEIGHT on the 128-row
// M side of the tensor core and the batch on 32-wide N tiles (gemm_tc "transposed" tiles: out/128 x B/32 CTAs instead of
// B/128 x out/128, 2.5x the CTAs streaming the weights), and the epilogue writes the result back as [batch][feature]:
//   * activations never exist as fp32: each epilogue (bias + tanh) emits the bf16 hi/lo planes the next layer reads, and
//     the concatenations are column windows of those plane buffers (h_w lands in xp[:, d:], h_p in xs[:, d:]);
//   * the planes of x_w, x_p, x_s, h_s and of the four weights are the tensors saved for backward (one opaque buffer), so
//     backward converts nothing but dlogits;
//   * backward: dW = dZ^T X as MN-major x MN-major products straight from the saved planes; dX = dZ W with the tanh
//     derivative (1 - h^2, from the saved planes), the plane conversion and the bias gradient (row sums) in the epilogue.
#include <algorithm>
#include "common.cuh"
#include "gemm_tc.cuh"
#include "util_kernels.cuh"

namespace hca {
namespace {

inline int64_t r8(int64_t x) { return (x + 7) / 8 * 8; }

struct Pl {                      // bf16 hi/lo planes [2][rows][ld]
  __nv_bfloat16* p = nullptr;
  int64_t ld = 0, ps = 0;
};

struct Saved {
  Pl Ww, Wp, Ws, Wh, xw, xp, xs, hs;
};

size_t pl_bytes(int64_t rows, int64_t ld) { return align_up((size_t)(2 * rows * ld * 2)); }

size_t saved_bytes(int B, int d, int m, int K) {
  const int64_t m8 = r8(m);
  return pl_bytes(d, d) + pl_bytes(d, 2 * d) + pl_bytes(m, 2 * d) + pl_bytes(K, m8) + pl_bytes(B, d) + 2 * pl_bytes(B, 2 * d) +
         pl_bytes(B, m8) + 256;
}

bool carve(Saved& s, void* buf, size_t bytes, int B, int d, int m, int K) {
  if (!buf || (reinterpret_cast<uintptr_t>(buf) & 255) || bytes < saved_bytes(B, d, m, K)) return false;
  char* c = (char*)buf;
  auto take = [&](Pl& x, int64_t rows, int64_t ld) {
    x.p = (__nv_bfloat16*)c; x.ld = ld; x.ps = rows * ld;
    c += pl_bytes(rows, ld);
  };
  const int64_t m8 = r8(m);
  take(s.Ww, d, d); take(s.Wp, d, 2 * d); take(s.Ws, m, 2 * d); take(s.Wh, K, m8);
  take(s.xw, B, d); take(s.xp, B, 2 * d); take(s.xs, B, 2 * d); take(s.hs, B, m8);
  return true;
}

// operand view of a plane buffer: a [rows, cols] window starting at column col0
TcOperand opnd(const Pl& x, int64_t col0, int rows, int cols, bool mn_major) {
  TcOperand o;
  o.planes = x.p + col0; o.ld = x.ld; o.plane_stride = x.ps; o.rows = rows; o.cols = cols; o.mn_major = mn_major;
  return o;
}
TcPlanes outp(const Pl& x, int64_t col0) {
  TcPlanes t;
  t.p = x.p + col0; t.ld = x.ld; t.plane_stride = x.ps;
  return t;
}

__device__ __forceinline__ void split4(const float (&x)[4], uint2& hi, uint2& lo) {
  __nv_bfloat16 h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h[j] = __float2bfloat16_rn(x[j]);
    l[j] = __float2bfloat16_rn(x[j] - __bfloat162float(h[j]));
  }
  hi = *reinterpret_cast<const uint2*>(h);
  lo = *reinterpret_cast<const uint2*>(l);
}

// planes of  xw = q0+v0 ;  xp[:, :d] = q1+v1 ;  xs[:, :d] = q2+v2      (vhat, qhat are [3,B,d]; d % 4 == 0)
__global__ void __launch_bounds__(256) mlp_inputs_planes_kernel(const float4* __restrict__ vhat, const float4* __restrict__ qhat,
                                                                __nv_bfloat16* __restrict__ xw, int64_t xw_ld, int64_t xw_ps,
                                                                __nv_bfloat16* __restrict__ xp, int64_t xp_ld, int64_t xp_ps,
                                                                __nv_bfloat16* __restrict__ xs, int64_t xs_ld, int64_t xs_ps, int B, int d4) {
  pdl_enter();
  const int64_t per = (int64_t)B * d4, total = 3 * per;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int l = (int)(i / per);
    const int64_t r = (i % per) / d4;
    const int c = (int)(i % d4) * 4;
    const float4 a = vhat[i], b = qhat[i];
    const float x[4] = {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w};
    uint2 hi, lo;
    split4(x, hi, lo);
    __nv_bfloat16* dst = l == 0 ? xw + r * xw_ld + c : (l == 1 ? xp + r * xp_ld + c : xs + r * xs_ld + c);
    const int64_t ps = l == 0 ? xw_ps : (l == 1 ? xp_ps : xs_ps);
    *reinterpret_cast<uint2*>(dst) = hi;
    *reinterpret_cast<uint2*>(dst + ps) = lo;
  }
}

// planes of X [rows, cols] (leading dim ld) + its column sums: block = 32 columns x 8 row lanes over all rows
__global__ void __launch_bounds__(256) split_colsum_kernel(const float* __restrict__ X, int64_t ld, int rows, int cols,
                                                           __nv_bfloat16* __restrict__ planes, int64_t ldp, int64_t ps,
                                                           float* __restrict__ colsum, const ZeroJobs zero) {
  pdl_enter();
  zero_jobs_device(zero);          // (the bias gradients the data-gradient products accumulate into: one launch less)
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float s = 0.f;
  if (c < cols) {
    for (int r = ry; r < rows; r += 8) {
      const float v = X[(int64_t)r * ld + c];
      s += v;
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      planes[(int64_t)r * ldp + c] = h;
      planes[ps + (int64_t)r * ldp + c] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
  red[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && c < cols) {
#pragma unroll
    for (int i = 1; i < 8; ++i) s += red[i][cx];
    colsum[c] = s;
  }
}


// out[b][j] (planes at `dst` and / or fp32 `D`) = act( sum_k X[b][k] W[j][k] + bias[j] )       W K-major [out, in]
int fwd_layer(const Pl& W, int out, int in, const Pl& X, int B, const float* bias, int act_tanh, const Pl* dst, int64_t dst_col0,
              float* D, int64_t ldd, cudaStream_t s) {
  TcEpilogue e;
  e.transposed = 1;
  e.bias = bias;
  e.act_tanh = act_tanh;
  if (dst) e.P = outp(*dst, dst_col0);
  e.D = D; e.ldd = ldd;
  return launch_gemm_tc(opnd(W, 0, out, in, false), opnd(X, 0, B, in, false), 2, out, B, in, e, 1, s);
}

// dX[b][j] = sum_k dZ[b][k] W[k][w_col0 + j]  (j < cols), optionally * (1 - H[b][h_col0 + j]^2); planes -> dst, fp32 -> D,
// row sums (over b) -> red (pre-zeroed)
int bwd_data(const Pl& W, int w_rows, int64_t w_col0, int cols, const Pl& dZ, int B, const Pl* H, int64_t h_col0, const Pl* dst,
             float* D, float* red, cudaStream_t s) {
  TcEpilogue e;
  e.transposed = 1;
  if (H) {
    e.aux_mode = TC_AUX_MUL_1MX2;
    e.auxp = outp(*H, h_col0);
  }
  if (dst) e.P = outp(*dst, 0);
  e.D = D; e.ldd = cols;
  e.red_row = red;
  return launch_gemm_tc(opnd(W, w_col0, w_rows, cols, true), opnd(dZ, 0, B, w_rows, false), 2, cols, B, w_rows, e, 1, s);
}

// dW[j][i] = sum_b dZ[b][j] X[b][i]
int bwd_weight(const Pl& dZ, int out, const Pl& X, int in, int B, float* dW, cudaStream_t s) {
  TcEpilogue e;
  e.D = dW; e.ldd = in;
  return launch_gemm_tc(opnd(dZ, 0, B, out, true), opnd(X, 0, B, in, true), 2, out, in, B, e, 1, s);
}

}  // namespace
}  // namespace hca

extern "C" size_t hca_mlp_saved_bytes(int B, int d, int mlp, int K) { return hca::saved_bytes(B, d, mlp, K); }

extern "C" size_t hca_mlp_workspace(int B, int d, int mlp, int K) {
  using namespace hca;
  return pl_bytes(B, r8(K)) + pl_bytes(B, r8(mlp)) + 2 * pl_bytes(B, d) + 1024;
}

extern "C" int hca_mlp_fwd(const float* vhat, const float* qhat, const float* Ww, const float* bw, const float* Wp,
                           const float* bp, const float* Ws, const float* bs, const float* Wh, const float* bh, float* logits,
                           void* saved, size_t saved_sz, int B, int d, int mlp, int K, void* ws, size_t ws_bytes, void* stream) {
  using namespace hca;
  (void)ws; (void)ws_bytes;
  cudaStream_t s = (cudaStream_t)stream;
  HCA_CHECK_ARG(vhat && qhat && Ww && bw && Wp && bp && Ws && bs && Wh && bh && logits && saved, "mlp_fwd: null pointer");
  HCA_CHECK_ARG(B > 0 && d > 0 && d % 8 == 0 && mlp > 0 && K > 0, "mlp_fwd: bad sizes B=%d d=%d mlp=%d K=%d (d %% 8 == 0 required)", B, d, mlp, K);
  HCA_CHECK_ARG(tc_available(), "mlp_fwd: cuTensorMapEncodeTiled is not available from the driver");
  Saved sv;
  HCA_CHECK_ARG(carve(sv, saved, saved_sz, B, d, mlp, K), "mlp_fwd: `saved` must be 256-byte aligned and hca_mlp_saved_bytes large");
  HCA_LAUNCH_K((mlp_inputs_planes_kernel), ew_grid(3LL * B * d / 4), 256, 0, s, (const float4*)vhat, (const float4*)qhat, sv.xw.p, sv.xw.ld, sv.xw.ps,
                                                                  sv.xp.p, sv.xp.ld, sv.xp.ps, sv.xs.p, sv.xs.ld, sv.xs.ps, B, d / 4);
  HCA_LAUNCHED();
  {  // the four weight matrices -> operand planes, one launch
    SplitBatch sb(s);
    HCA_TRY(sb.add(Ww, d, d, d, sv.Ww.p, sv.Ww.ld, sv.Ww.ps));
    HCA_TRY(sb.add(Wp, 2 * d, d, 2 * d, sv.Wp.p, sv.Wp.ld, sv.Wp.ps));
    HCA_TRY(sb.add(Ws, 2 * d, mlp, 2 * d, sv.Ws.p, sv.Ws.ld, sv.Ws.ps));
    HCA_TRY(sb.add(Wh, mlp, K, mlp, sv.Wh.p, sv.Wh.ld, sv.Wh.ps));
    HCA_TRY(sb.flush());
  }
  HCA_TRY(fwd_layer(sv.Ww, d, d, sv.xw, B, bw, 1, &sv.xp, d, nullptr, 0, s));            // h_w -> xp[:, d:]
  HCA_TRY(fwd_layer(sv.Wp, d, 2 * d, sv.xp, B, bp, 1, &sv.xs, d, nullptr, 0, s));        // h_p -> xs[:, d:]
  HCA_TRY(fwd_layer(sv.Ws, mlp, 2 * d, sv.xs, B, bs, 1, &sv.hs, 0, nullptr, 0, s));      // h_s
  HCA_TRY(fwd_layer(sv.Wh, K, mlp, sv.hs, B, bh, 0, nullptr, 0, logits, K, s));
  return 0;
}

extern "C" int hca_mlp_bwd(const float* dlogits, const void* saved, size_t saved_sz, float* g, float* dWw, float* dbw,
                           float* dWp, float* dbp, float* dWs, float* dbs, float* dWh, float* dbh, int B, int d, int mlp, int K,
                           void* ws, size_t ws_bytes, void* stream) {
  using namespace hca;
  cudaStream_t s = (cudaStream_t)stream;
  HCA_CHECK_ARG(dlogits && saved, "mlp_bwd: null input");
  HCA_CHECK_ARG(g && dWw && dbw && dWp && dbp && dWs && dbs && dWh && dbh, "mlp_bwd: null output");
  HCA_CHECK_ARG(B > 0 && d > 0 && d % 8 == 0 && mlp > 0 && K > 0, "mlp_bwd: bad sizes");
  HCA_CHECK_ARG(tc_available(), "mlp_bwd: cuTensorMapEncodeTiled is not available from the driver");
  Saved sv;
  HCA_CHECK_ARG(carve(sv, const_cast<void*>(saved), saved_sz, B, d, mlp, K), "mlp_bwd: bad `saved` buffer");
  Workspace w(ws, ws_bytes);
  const int64_t K8 = r8(K), m8 = r8(mlp);
  Pl dl, dzs, dzp, dzw;
  dl.p = w.take<__nv_bfloat16>((size_t)2 * B * K8); dl.ld = K8; dl.ps = (int64_t)B * K8;
  dzs.p = w.take<__nv_bfloat16>((size_t)2 * B * m8); dzs.ld = m8; dzs.ps = (int64_t)B * m8;
  dzp.p = w.take<__nv_bfloat16>((size_t)2 * B * d); dzp.ld = d; dzp.ps = (int64_t)B * d;
  dzw.p = w.take<__nv_bfloat16>((size_t)2 * B * d); dzw.ld = d; dzw.ps = (int64_t)B * d;
  if (!dzw.p) return set_err(HCA_ERR_WORKSPACE, "mlp_bwd: workspace too small");
  float* g_w = g;
  float* g_p = g + (size_t)B * d;
  float* g_s = g + (size_t)2 * B * d;

  // The four weight-gradient products feed nothing later in this call: they run on the helper stream, next to the chain of
  // data-gradient products (each 20-64 CTAs: together they still do not fill the 148 SMs).
  SideStream side(s);
  cudaStream_t sw = side.stream();
  // W_h: planes of dlogits + db_h in one pass; dW_h; dz_s = (dlogits W_h) * (1 - h_s^2) with db_s
  {
    ZeroBatch zb(s);
    HCA_TRY(zb.add(dbs, (size_t)mlp * 4)); HCA_TRY(zb.add(dbp, (size_t)d * 4)); HCA_TRY(zb.add(dbw, (size_t)d * 4));
    HCA_LAUNCH_K((split_colsum_kernel), (K + 31) / 32, 256, 0, s, dlogits, K, B, K, dl.p, dl.ld, dl.ps, dbh, zb.take());
    HCA_LAUNCHED();
  }
  HCA_TRY(side.fork());
  HCA_TRY(bwd_weight(dl, K, sv.hs, mlp, B, dWh, sw));
  HCA_TRY(bwd_data(sv.Wh, K, 0, mlp, dl, B, &sv.hs, 0, &dzs, nullptr, dbs, s));
  // W_s: dx_s = dz_s W_s : left half -> g_s, right half * (1 - h_p^2) -> dz_p
  HCA_TRY(side.fork());
  HCA_TRY(bwd_weight(dzs, mlp, sv.xs, 2 * d, B, dWs, sw));
  HCA_TRY(bwd_data(sv.Ws, mlp, d, d, dzs, B, &sv.xs, d, &dzp, nullptr, dbp, s));
  // W_p
  HCA_TRY(side.fork());
  HCA_TRY(bwd_weight(dzp, d, sv.xp, 2 * d, B, dWp, sw));
  HCA_TRY(bwd_data(sv.Ws, mlp, 0, d, dzs, B, nullptr, 0, nullptr, g_s, nullptr, sw));    // off the critical path too
  HCA_TRY(bwd_data(sv.Wp, d, d, d, dzp, B, &sv.xp, d, &dzw, nullptr, dbw, s));
  // W_w
  HCA_TRY(side.fork());
  HCA_TRY(bwd_weight(dzw, d, sv.xw, d, B, dWw, sw));
  HCA_TRY(bwd_data(sv.Wp, d, 0, d, dzp, B, nullptr, 0, nullptr, g_p, nullptr, sw));
  HCA_TRY(bwd_data(sv.Ww, d, 0, d, dzw, B, nullptr, 0, nullptr, g_w, nullptr, s));
  HCA_TRY(side.join());
  return 0;
}
