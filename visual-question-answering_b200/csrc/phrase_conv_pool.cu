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
  pdl_enter();
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
  pdl_enter();
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
  pdl_enter();
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

// Max-pool argmax must match the reference bit for bit, but the tensor-core conv (bf16x2 operand split: per-term relative error
// <= ~2^-16.5, i.e. sigma ~1.3e-5 on a pre-activation of 1536 terms, ~1.8e-5 on a gap of two) is not exact.  So the pool kernel
// records every element whose top-2 gap is below TIE_TOL (> 20 sigma of that error) and fixup_ties_kernel recomputes just those
// elements in exact fp32 (value and index): ~0.1 % of the elements.  (The first version ran the conv as a bf16x3 split, 6 MMAs per
// product, to stay two orders of magnitude under the band; with the exact repair in place 3 MMAs and a wider band do the same job.)
constexpr float TIE_TOL = 4e-4f;

// cat [R, 3E] (post tanh) -> out [R, E], idx [R, E]; rows t >= len zeroed.  tie_list/tie_count may be null.
__global__ void __launch_bounds__(256) pool3_fwd_kernel(const float* __restrict__ cat, const int64_t* __restrict__ lens,
                                                        float* __restrict__ out, uint8_t* __restrict__ idx, int B, int T, int E,
                                                        int* __restrict__ tie_list, int* __restrict__ tie_count, int tie_cap) {
  pdl_enter();
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
                                                         const float* __restrict__ x, int T, const float* __restrict__ w1,
                                                         const float* __restrict__ w2, const float* __restrict__ w3,
                                                         const float* __restrict__ b1, const float* __restrict__ b2,
                                                         const float* __restrict__ b3, float* __restrict__ out,
                                                         uint8_t* __restrict__ idx, int E) {
  pdl_enter();
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
      const float* wrow = (k == 1 ? w1 : (k == 2 ? w2 : w3)) + (int64_t)o * E * k;   // conv layout [C_in][k]
      const int t = (int)(r % T);
      float acc = 0.f;
      // taps of conv k cover x[t-1], x[t] (k = 2), x[t-1..t+1] (k = 3), x[t] (k = 1); zeros outside [0, T)
      for (int tap = 0; tap < k; ++tap) {
        const int tt = t + tap - (k == 1 ? 0 : 1);
        if (tt < 0 || tt >= T) continue;
        const float* xrow = x + (r + tt - t) * (int64_t)E;
        for (int cc = lane; cc < E; cc += 32) acc = fmaf(xrow[cc], wrow[(int64_t)cc * k + tap], acc);
      }
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
  pdl_enter();
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


// ---- tensor-core path helpers: operands are produced directly as bf16 planes (no fp32 im2col / repack round trips) ----
// x = x0 + x1 (+ x2), planes of 4 consecutive elements packed as uint2
template <int P>
__device__ __forceinline__ void split4(float (&x)[4], uint2 (&out)[P]) {
#pragma unroll
  for (int pl = 0; pl < P; ++pl) {
    __nv_bfloat16 h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      h[j] = __float2bfloat16_rn(x[j]);
      x[j] -= __bfloat162float(h[j]);
    }
    out[pl] = *reinterpret_cast<const uint2*>(h);
  }
}
// planes [P][R][3E] of Acat[r][j*E + c] = x[b][t+j-1][c]
template <int P>
__global__ void __launch_bounds__(256) im2col3_planes_kernel(const float4* __restrict__ x, __nv_bfloat16* __restrict__ planes, int64_t ps,
                                                             int B, int T, int E4) {
  pdl_enter();
  const int64_t total = (int64_t)B * T * 3 * E4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % E4);
    const int j = (int)((i / E4) % 3);
    const int64_t r = i / (3 * E4);
    const int t = (int)(r % T) + j - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t >= 0 && t < T) v = __ldg(x + (r + j - 1) * E4 + c);
    float xv[4] = {v.x, v.y, v.z, v.w};
    uint2 o[P];
    split4<P>(xv, o);
#pragma unroll
    for (int pl = 0; pl < P; ++pl) *reinterpret_cast<uint2*>(planes + pl * ps + i * 4) = o[pl];
  }
}
// planes [P][E][k*E] of the tap-major weight Wr[o][j*E + c] = w[o][c][j]  (conv layout [C_out][C_in][k])
template <int P>
__global__ void __launch_bounds__(256) conv_w_planes_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ planes, int64_t ps, int E,
                                                            int k) {
  pdl_enter();
  const int64_t total = (int64_t)E * E * k;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % E);
    const int j = (int)((i / E) % k);
    const int64_t o = i / ((int64_t)E * k);
    float x = w[(o * E + c) * k + j];
#pragma unroll
    for (int pl = 0; pl < P; ++pl) {
      const __nv_bfloat16 h = __float2bfloat16_rn(x);
      planes[pl * ps + i] = h;
      x -= __bfloat162float(h);
    }
  }
}
// The three weights in one launch (hi/lo planes), which also clears the near-tie counter of the pool kernel.
__global__ void __launch_bounds__(256) conv_w_planes3_kernel(const float* __restrict__ w1, const float* __restrict__ w2,
                                                             const float* __restrict__ w3, __nv_bfloat16* __restrict__ p1,
                                                             __nv_bfloat16* __restrict__ p2, __nv_bfloat16* __restrict__ p3, int E,
                                                             int* __restrict__ tie_count) {
  pdl_enter();
  if (tie_count && blockIdx.x == 0 && threadIdx.x == 0) *tie_count = 0;
  const int64_t EE = (int64_t)E * E, total = 6 * EE;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
    const int k = g < EE ? 1 : (g < 3 * EE ? 2 : 3);
    const int64_t i = g - (k == 1 ? 0 : (k == 2 ? EE : 3 * EE));
    const float* w = k == 1 ? w1 : (k == 2 ? w2 : w3);
    __nv_bfloat16* planes = k == 1 ? p1 : (k == 2 ? p2 : p3);
    const int64_t ps = EE * k;
    const int c = (int)(i % E);
    const int j = (int)((i / E) % k);
    const int64_t o = i / ((int64_t)E * k);
    const float x = w[(o * E + c) * k + j];
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    planes[i] = h;
    planes[ps + i] = __float2bfloat16_rn(x - __bfloat162float(h));
  }
}
// Pool backward straight into operand planes: dcat[r][3e+j] = (j == idx) ? dout * (1 - out^2) : 0 as bf16 hi/lo planes
// [2][R][3E], plus the three bias gradients (column sums of dcat over r).  Thread <-> one channel triple e, block <-> a slab
// of rows, so the column sums stay in registers until one atomic per column and block.
constexpr int POOL_BWD_ROWS = 32;
__global__ void __launch_bounds__(256) pool3_bwd_planes_kernel(const float* __restrict__ out, const uint8_t* __restrict__ idx,
                                                               const float* __restrict__ dout, const int64_t* __restrict__ lens,
                                                               __nv_bfloat16* __restrict__ planes, int64_t ps, float* __restrict__ db1,
                                                               float* __restrict__ db2, float* __restrict__ db3, int B, int T, int E) {
  pdl_enter();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t R = (int64_t)B * T;
  const int64_t r0 = (int64_t)blockIdx.y * POOL_BWD_ROWS, r1 = min(R, r0 + POOL_BWD_ROWS);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 4                      // four rows' loads in flight per thread (the rows are independent; one at a time is latency-bound)
  for (int64_t r = r0; r < r1; ++r) {
    const int b = (int)(r / T), t = (int)(r - (int64_t)b * T);
    float g = 0.f;
    int j = 0;
    if (!lens || t < lens[b]) {
      const float o = out[r * E + e];
      g = dout[r * E + e] * (1.f - o * o);
      j = idx[r * E + e];
    }
    const __nv_bfloat16 h = __float2bfloat16_rn(g);
    const __nv_bfloat16 l = __float2bfloat16_rn(g - __bfloat162float(h));
    const __nv_bfloat16 z = __float2bfloat16_rn(0.f);
    __nv_bfloat16* ph = planes + r * 3 * (int64_t)E + 3 * e;
    ph[0] = j == 0 ? h : z; ph[1] = j == 1 ? h : z; ph[2] = j == 2 ? h : z;
    ph[ps] = j == 0 ? l : z; ph[ps + 1] = j == 1 ? l : z; ph[ps + 2] = j == 2 ? l : z;
    s0 += j == 0 ? g : 0.f; s1 += j == 1 ? g : 0.f; s2 += j == 2 ? g : 0.f;
  }
  const float sv[3] = {s0, s1, s2};
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int c = 3 * e + j;                    // channel of the concatenated [uni|bi|tri] axis
    float* db = c < E ? db1 : (c < 2 * E ? db2 : db3);
    atomicAdd(db + (c % E), sv[j]);
  }
}
// tap-major fp32 weight gradient -> conv layout: w[o][c][j] = wr[o][j*E + c]
__global__ void __launch_bounds__(256) unpack_conv_w_kernel(const float* __restrict__ wr, float* __restrict__ w, int E, int k) {
  pdl_enter();
  const int64_t total = (int64_t)E * E * k;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % k);
    const int c = (int)((i / k) % E);
    const int64_t o = i / ((int64_t)E * k);
    w[i] = wr[(o * k + j) * E + c];
  }
}
// both in one launch (bigram and trigram weight gradients)
__global__ void __launch_bounds__(256) unpack_conv_w23_kernel(const float* __restrict__ wr2, float* __restrict__ w2,
                                                              const float* __restrict__ wr3, float* __restrict__ w3, int E) {
  pdl_enter();
  const int64_t n2 = (int64_t)E * E * 2, total = n2 + (int64_t)E * E * 3;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
    const int k = g < n2 ? 2 : 3;
    const int64_t i = g < n2 ? g : g - n2;
    const int j = (int)(i % k);
    const int c = (int)((i / k) % E);
    const int64_t o = i / ((int64_t)E * k);
    (k == 2 ? w2 : w3)[i] = (k == 2 ? wr2 : wr3)[(o * k + j) * E + c];
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
         4 * 2 * align_up(R * 3 * E) + 3 * 3 * 2 * align_up((size_t)E * 3 * E) + 8192 +   // bf16 planes: Acat (x3 fwd) / Acat + dcat (x2 bwd), weights

         std::max(hca::dense_scratch_bytes(E, 3 * E, (int)R), hca::dense_scratch_bytes((int)R, 3 * E, E));
}

namespace hca {
namespace {
// operand planes kept from forward to backward: Acat [2][R][3E], W1 [2][E][E], W2 [2][E][2E], W3 [2][E][3E]  (bf16)
struct ConvSaved {
  __nv_bfloat16* ap = nullptr;
  __nv_bfloat16* wp[3] = {nullptr, nullptr, nullptr};
};
size_t conv_saved_bytes(int B, int T, int E) {
  const size_t R = (size_t)B * T;
  size_t n = align_up(2 * R * 3 * E * 2);
  for (int k = 1; k <= 3; ++k) n += align_up((size_t)2 * E * k * E * 2);
  return n + 256;
}
bool carve_saved(ConvSaved& v, void* buf, size_t bytes, int B, int T, int E) {
  if (!buf || (reinterpret_cast<uintptr_t>(buf) & 255) || bytes < conv_saved_bytes(B, T, E)) return false;
  char* p = (char*)buf;
  const size_t R = (size_t)B * T;
  v.ap = (__nv_bfloat16*)p; p += align_up(2 * R * 3 * E * 2);
  for (int k = 1; k <= 3; ++k) { v.wp[k - 1] = (__nv_bfloat16*)p; p += align_up((size_t)2 * E * k * E * 2); }
  return true;
}
}  // namespace
}  // namespace hca

extern "C" size_t hca_phrase_conv_pool_saved_bytes(int B, int T, int E) { return hca::conv_saved_bytes(B, T, E); }

extern "C" int hca_phrase_conv_pool_fwd(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                                        const float* w3, const float* b3, const int64_t* lens, float* out, uint8_t* idx,
                                        void* fsaved, size_t fsaved_bytes, int B, int T, int E, void* ws, size_t ws_bytes,
                                        void* stream) {
  using namespace hca;
  cudaStream_t s = (cudaStream_t)stream;
  HCA_CHECK_ARG(x && w1 && b1 && w2 && b2 && w3 && b3 && out && idx, "phrase_conv_pool_fwd: null pointer");
  HCA_CHECK_ARG(B > 0 && T > 0 && E > 0 && E % 4 == 0, "phrase_conv_pool_fwd: bad sizes B=%d T=%d E=%d (E %% 4 == 0 required)", B, T, E);
  ConvWs c = carve(ws, ws_bytes, B, T, E);
  if (!c.ok) return set_err(HCA_ERR_WORKSPACE, "phrase_conv_pool_fwd: workspace too small (%zu bytes)", ws_bytes);
  const int R = B * T;
  const float* bs[3] = {b1, b2, b3};
  const bool tc = use_tc() && tc_available() && (E % 8 == 0);      // TMA needs 16-byte aligned plane windows
  if (tc) {
    // tensor cores, bf16x2 operand split (3 MMAs per product): the row-shifted operand Acat and the tap-major
    // weights are written directly as bf16 planes (no fp32 im2col / repack round trip); the three convs read column windows
    // of the Acat planes; bias + tanh fused in the epilogue; near-ties are repaired exactly below
    const int P = 2;
    const int64_t lda = 3 * (int64_t)E, a_stride = (int64_t)R * lda;
    ConvSaved sv;
    if (fsaved) HCA_CHECK_ARG(carve_saved(sv, fsaved, fsaved_bytes, B, T, E), "phrase_conv_pool_fwd: `fsaved` must be 256-byte aligned and hca_phrase_conv_pool_saved_bytes large");
    __nv_bfloat16* ap = fsaved ? sv.ap : c.w.take<__nv_bfloat16>((size_t)P * a_stride);
    if (!ap) return set_err(HCA_ERR_WORKSPACE, "phrase_conv_pool_fwd: workspace too small for operand planes");
    HCA_LAUNCH_K((im2col3_planes_kernel<2>), ew_grid((int64_t)R * 3 * E / 4), 256, 0, s, (const float4*)x, ap, a_stride, B, T, E / 4);
    HCA_LAUNCHED();
    __nv_bfloat16* wps[3];
    for (int k = 1; k <= 3; ++k) {
      wps[k - 1] = fsaved ? sv.wp[k - 1] : c.w.take<__nv_bfloat16>((size_t)P * E * k * E);
      if (!wps[k - 1]) return set_err(HCA_ERR_WORKSPACE, "phrase_conv_pool_fwd: workspace too small for weight planes");
    }
    HCA_LAUNCH_K((conv_w_planes3_kernel), ew_grid((int64_t)6 * E * E), 256, 0, s, w1, w2, w3, wps[0], wps[1], wps[2], E, c.tie_count);
    HCA_LAUNCHED();
    for (int k = 1; k <= 3; ++k) {
      const int64_t ldw = (int64_t)k * E, w_stride = (int64_t)E * ldw;
      __nv_bfloat16* wp = wps[k - 1];
      TcOperand A, Bw;
      A.planes = ap + (k == 1 ? E : 0); A.ld = lda; A.plane_stride = a_stride; A.rows = R; A.cols = k * E;
      Bw.planes = wp; Bw.ld = ldw; Bw.plane_stride = w_stride; Bw.rows = E; Bw.cols = k * E;
      TcEpilogue ep;
      ep.D = c.cat + (k - 1) * E; ep.ldd = lda; ep.bias = bs[k - 1]; ep.act_tanh = 1;
      HCA_TRY(launch_gemm_tc(A, Bw, P, R, E, k * E, ep, 1, s));
    }
    HCA_LAUNCH_K((pool3_fwd_kernel), ew_grid((int64_t)R * E), 256, 0, s, c.cat, lens, out, idx, B, T, E, c.tie_list, c.tie_count, c.tie_cap);
    HCA_LAUNCHED();
    HCA_LAUNCH_K((fixup_ties_kernel), 148 * 2, 256, 0, s, c.tie_list, c.tie_count, c.tie_cap, x, T, w1, w2, w3, b1, b2, b3, out, idx, E);
    HCA_LAUNCHED();
    return 0;
  }
  HCA_LAUNCH_K((im2col3_kernel), ew_grid((int64_t)R * 3 * E / 4), 256, 0, s, (const float4*)x, (float4*)c.acat, B, T, E / 4);
  HCA_LAUNCHED();
  HCA_LAUNCH_K((repack_conv_w_kernel), ew_grid((int64_t)E * E * 2), 256, 0, s, w2, c.wr2, E, 2, false);
  HCA_LAUNCHED();
  HCA_LAUNCH_K((repack_conv_w_kernel), ew_grid((int64_t)E * E * 3), 256, 0, s, w3, c.wr3, E, 3, false);
  HCA_LAUNCHED();
  const float* wr[3] = {w1, c.wr2, c.wr3};
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
  HCA_LAUNCH_K((pool3_fwd_kernel), ew_grid((int64_t)R * E), 256, 0, s, c.cat, lens, out, idx, B, T, E, nullptr, nullptr, 0);
  HCA_LAUNCHED();
  return 0;
}

extern "C" int hca_phrase_conv_pool_bwd(const float* x, const float* w1, const float* w2, const float* w3, const float* out,
                                        const uint8_t* idx, const float* dout, const int64_t* lens, const void* fsaved,
                                        size_t fsaved_bytes, float* dx, float* dw1, float* db1, float* dw2, float* db2, float* dw3,
                                        float* db3, int B, int T, int E, void* ws, size_t ws_bytes, void* stream) {
  using namespace hca;
  cudaStream_t s = (cudaStream_t)stream;
  HCA_CHECK_ARG(x && w1 && w2 && w3 && out && idx && dout && dw1 && db1 && dw2 && db2 && dw3 && db3, "phrase_conv_pool_bwd: null pointer");
  HCA_CHECK_ARG(B > 0 && T > 0 && E > 0 && E % 4 == 0, "phrase_conv_pool_bwd: bad sizes");
  ConvWs c = carve(ws, ws_bytes, B, T, E);
  if (!c.ok) return set_err(HCA_ERR_WORKSPACE, "phrase_conv_pool_bwd: workspace too small (%zu bytes)", ws_bytes);
  const int R = B * T;
  if (use_tc() && tc_available() && (E % 8 == 0)) {
    // tensor-core path: every operand is produced once, directly as bf16 hi/lo planes, and the six products read windows of them
    const int64_t ld3 = 3 * (int64_t)E, pstride = (int64_t)R * ld3;
    // (the planes of Acat and of the weights are the forward's when it left them in `fsaved`: same conversion, same inputs)
    ConvSaved sv;
    if (fsaved) HCA_CHECK_ARG(carve_saved(sv, const_cast<void*>(fsaved), fsaved_bytes, B, T, E), "phrase_conv_pool_bwd: bad `fsaved` buffer");
    __nv_bfloat16* ap = fsaved ? sv.ap : c.w.take<__nv_bfloat16>((size_t)2 * pstride);       // Acat planes
    __nv_bfloat16* dp = c.w.take<__nv_bfloat16>((size_t)2 * pstride);       // dcat planes
    if (!dp || !ap) return set_err(HCA_ERR_WORKSPACE, "phrase_conv_pool_bwd: workspace too small for operand planes");
    if (!fsaved) {
      HCA_LAUNCH_K((im2col3_planes_kernel<2>), ew_grid((int64_t)R * 3 * E / 4), 256, 0, s, (const float4*)x, ap, pstride, B, T, E / 4);
      HCA_LAUNCHED();
    }
    float* dbs[3] = {db1, db2, db3};
    float* dwr[3] = {dw1, c.dwr2, c.dwr3};
    int sks[3];
    {  // bias gradients and the split-K weight-gradient accumulators cleared by one launch
      ZeroBatch zb(s);
      for (int k = 1; k <= 3; ++k) {
        HCA_TRY(zb.add(dbs[k - 1], (size_t)E * 4));
        const int tiles = ((E + 127) / 128) * ((k * E + 127) / 128);
        sks[k - 1] = tiles >= 96 ? 1 : std::max(1, std::min((148 + tiles - 1) / tiles, (R + 255) / 256));
        if (sks[k - 1] > 1) HCA_TRY(zb.add(dwr[k - 1], (size_t)E * k * E * 4));
      }
      HCA_TRY(zb.flush());
    }
    HCA_LAUNCH_K((pool3_bwd_planes_kernel), dim3((E + 255) / 256, (R + POOL_BWD_ROWS - 1) / POOL_BWD_ROWS), 256, 0, s, out, idx, dout, lens, dp, pstride, db1,
                                                                                                       db2, db3, B, T, E);
    HCA_LAUNCHED();
    // weight gradients, tap-major: dWr_k[o][kk] = sum_r dcat[r][(k-1)E + o] * Acat[r][a_off + kk]   (K = R, split-K)
    for (int k = 1; k <= 3; ++k) {
      TcOperand A, Bm;
      A.planes = dp + (k - 1) * E; A.ld = ld3; A.plane_stride = pstride; A.rows = R; A.cols = E; A.mn_major = true;
      Bm.planes = ap + (k == 1 ? E : 0); Bm.ld = ld3; Bm.plane_stride = pstride; Bm.rows = R; Bm.cols = k * E; Bm.mn_major = true;
      const int sk = sks[k - 1];
      TcEpilogue ep;
      ep.D = dwr[k - 1]; ep.ldd = (int64_t)k * E;
      HCA_TRY(launch_gemm_tc(A, Bm, 2, E, k * E, R, ep, sk, s));
    }
    HCA_LAUNCH_K((unpack_conv_w23_kernel), ew_grid((int64_t)E * E * 5), 256, 0, s, c.dwr2, dw2, c.dwr3, dw3, E);
    HCA_LAUNCHED();
    if (dx) {
      // dA[r][a_off + kk] (+)= sum_o dcat[r][(k-1)E + o] * Wr_k[o][kk]: tri first (covers all 3E columns), bi and uni accumulate
      const float* ws_[3] = {w1, w2, w3};
      for (int k = 3; k >= 1; --k) {
        const int64_t ldw = (int64_t)k * E, w_stride = (int64_t)E * ldw;
        __nv_bfloat16* wp = fsaved ? sv.wp[k - 1] : c.w.take<__nv_bfloat16>((size_t)2 * w_stride);
        if (!wp) return set_err(HCA_ERR_WORKSPACE, "phrase_conv_pool_bwd: workspace too small for weight planes");
        if (!fsaved) {
          HCA_LAUNCH_K((conv_w_planes_kernel<2>), ew_grid((int64_t)E * E * k), 256, 0, s, ws_[k - 1], wp, w_stride, E, k);
          HCA_LAUNCHED();
        }
        TcOperand A, Bm;
        A.planes = dp + (k - 1) * E; A.ld = ld3; A.plane_stride = pstride; A.rows = R; A.cols = E;
        Bm.planes = wp; Bm.ld = ldw; Bm.plane_stride = w_stride; Bm.rows = E; Bm.cols = k * E; Bm.mn_major = true;
        TcEpilogue ep;
        ep.D = c.dA + (k == 1 ? E : 0); ep.ldd = ld3; ep.accumulate = (k != 3);
        HCA_TRY(launch_gemm_tc(A, Bm, 2, R, k * E, E, ep, 1, s));
      }
      HCA_LAUNCH_K((col2im3_kernel), ew_grid((int64_t)R * E / 4), 256, 0, s, (const float4*)c.dA, (float4*)dx, B, T, E / 4);
      HCA_LAUNCHED();
    }
    return 0;
  }
  float* dcat = c.cat;
  HCA_LAUNCH_K((im2col3_kernel), ew_grid((int64_t)R * 3 * E / 4), 256, 0, s, (const float4*)x, (float4*)c.acat, B, T, E / 4);
  HCA_LAUNCHED();
  HCA_LAUNCH_K((pool3_bwd_kernel), ew_grid((int64_t)R * E), 256, 0, s, out, idx, dout, lens, dcat, B, T, E);
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
  HCA_LAUNCH_K((repack_conv_w_kernel), ew_grid((int64_t)E * E * 2), 256, 0, s, c.dwr2, dw2, E, 2, true);
  HCA_LAUNCHED();
  HCA_LAUNCH_K((repack_conv_w_kernel), ew_grid((int64_t)E * E * 3), 256, 0, s, c.dwr3, dw3, E, 3, true);
  HCA_LAUNCHED();
  if (dx) {
    // dA[r][kk] = sum_o dcat[r][blk + o] * Wr_k[o][kk], accumulated over the three convs
    HCA_LAUNCH_K((repack_conv_w_kernel), ew_grid((int64_t)E * E * 2), 256, 0, s, w2, c.wr2, E, 2, false);
    HCA_LAUNCHED();
    HCA_LAUNCH_K((repack_conv_w_kernel), ew_grid((int64_t)E * E * 3), 256, 0, s, w3, c.wr3, E, 3, false);
    HCA_LAUNCHED();
    const float* wr[3] = {w1, c.wr2, c.wr3};
    for (int k = 3; k >= 1; --k) {   // tri first (covers all 3E columns, plain store), then bi, uni accumulate
      DenseEpi e;
      e.accumulate = (k != 3);
      HCA_TRY(dense_nn(dcat + (k - 1) * E, 3 * (int64_t)E, wr[k - 1], (int64_t)k * E, c.dA + ((k == 1) ? E : 0), 3 * (int64_t)E,
                       /*M=*/R, /*N=*/k * E, /*K=*/E, e, c.w, s));
    }
    HCA_LAUNCH_K((col2im3_kernel), ew_grid((int64_t)R * E / 4), 256, 0, s, (const float4*)c.dA, (float4*)dx, B, T, E / 4);
    HCA_LAUNCHED();
  }
  return 0;
}
