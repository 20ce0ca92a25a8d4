// ParallelCoAttention, three levels with shared weights (replaces reference model.py:356-397).
//
// Per sample b and level l (V = image features [N,d], Q_l = question level [T,d]):
//   C_l  = tanh(Q_l V^T)               PV = V Wv^T + bv  (level independent: once per step)
//   Hv_l = tanh(PV + C_l^T PQ_l)       PQ_l = Q_l Wq^T + bq
//   Hq_l = tanh(PQ_l + C_l PV)
//   av_l = softmax_N(Hv_l wv + cv)     aq_l = softmax_T(Hq_l wq + cq)   (all T positions, pads included)
//   vhat_l = av_l^T V                  qhat_l = aq_l^T Q_l
// W_b of the reference is dead code (model.py:347 vs :377) and does not appear.
//
// Every contraction runs on tcgen05 (gemm_tc.cuh, bf16x2 operand planes).  The three levels are STACKED along the
// row axis of every question-side matrix -- Q_all[b] = [Q_0; Q_1; Q_2] is [3T, d] -- so that the products whose
// image-side operand is level independent are one M = 3T (or K = 3T) product per sample instead of three M = T ones,
// and the products that contract over T per level put the long image axis (N or d) on the 128-row M side of the
// tensor core and T on a 32-wide N tile (transposed epilogue).  Intermediates are never written as fp32: each GEMM
// epilogue emits the bf16 hi/lo planes its consumers take as operands.  Hv (three levels x [N, d] per sample) is recomputed in
// backward by hv_fused.cu; Hq ([3T, d] per sample) is saved: recomputing it is a per-sample product with 78 rows that is bound by
// the L2 -> SM operand stream, not by its arithmetic.
//
// Saved for backward (one opaque buffer, hca_coattn_saved_bytes): planes of V, Q_all, PV, PQ_all, C_all; av, aq; planes of Hq.
#include <algorithm>
#include "common.cuh"
#include "gemm_tc.cuh"
#include "hv_fused.cuh"
#include "util_kernels.cuh"

namespace hca {
namespace {

inline int64_t round8(int64_t x) { return (x + 7) / 8 * 8; }

// dense copy of a strided [B,N,d] tensor
__global__ void __launch_bounds__(256) gather_strided_kernel(const float* __restrict__ V, int64_t sb, int64_t sn, int64_t sd,
                                                             float* __restrict__ out, int B, int N, int d) {
  pdl_enter();
  const int64_t total = (int64_t)B * N * d;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % d);
    const int n = (int)((i / d) % N);
    const int64_t b = i / ((int64_t)d * N);
    out[i] = V[b * sb + n * sn + c * sd];
  }
}

// The layout the reference's image encoder hands over (model.py:217: a [B, d, N] feature map viewed as [B, N, d], i.e. region stride 1
// and channel stride N) goes STRAIGHT to the operand planes: a 64-channel x 32-region tile is read coalesced along the regions,
// transposed in shared memory and written as bf16 hi / lo rows of the K-major [B*N, d] plane matrices (4-byte stores, 128 bytes per
// warp and plane).  No dense fp32 copy of the features is made on this path (it used to cost a 64 MB write and a 64 MB read per step).
__global__ void __launch_bounds__(256) split_planes_channel_major_kernel(const float* __restrict__ V, int64_t sb, int64_t sd,
                                                                         __nv_bfloat16* __restrict__ planes, int64_t ld, int64_t ps, int N, int d) {
  pdl_enter();
  __shared__ float tile[64][33];
  const int b = blockIdx.z, n0 = blockIdx.x * 32, c0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* src = V + (int64_t)b * sb;
  for (int i = ty; i < 64; i += 8) {
    const int c = c0 + i, n = n0 + tx;
    tile[i][tx] = (c < d && n < N) ? src[(int64_t)c * sd + n] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {                 // one region row per warp and pass: lane = channel pair
    const int n = n0 + i, c = c0 + 2 * tx;
    if (n < N && c < d) {                            // (d % 8 == 0: a channel pair is inside or outside together)
      const float x0 = tile[2 * tx][i], x1 = tile[2 * tx + 1][i];
      const __nv_bfloat162 hh = __floats2bfloat162_rn(x0, x1);
      const __nv_bfloat162 ll = __floats2bfloat162_rn(x0 - __low2float(hh), x1 - __high2float(hh));
      __nv_bfloat16* dst = planes + ((int64_t)b * N + n) * ld + c;
      *reinterpret_cast<__nv_bfloat162*>(dst) = hh;
      *reinterpret_cast<__nv_bfloat162*>(dst + ps) = ll;
    }
  }
}

// one warp: a = softmax(s + c) over L entries; a -> smem and global
__device__ __forceinline__ void softmax_warp(const float* __restrict__ s, float cbias, int L, float* a_sm, float* __restrict__ a_out) {
  const int lane = threadIdx.x & 31;
  float m = -INFINITY;
  for (int i = lane; i < L; i += 32) {
    const float v = s[i] + cbias;
    a_sm[i] = v;
    m = fmaxf(m, v);
  }
  m = warp_max(m);
  float sum = 0.f;
  for (int i = lane; i < L; i += 32) {
    const float e = expf(a_sm[i] - m);
    a_sm[i] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  for (int i = lane; i < L; i += 32) {
    const float a = a_sm[i] * inv;
    a_sm[i] = a;
    a_out[i] = a;
  }
}

// ATTN_SPLIT blocks per sample: the six softmaxes (3 levels x {regions, tokens}; recomputed by every block of the sample,
// they are tiny) and the attention-weighted sums.  Block (b, s) sums the regions n = s (mod ATTN_SPLIT) of V[b] for all three
// levels and, for s < 3, the tokens of question level s; partial sums are added atomically into the zeroed outputs.  Thread
// c owns a float2 column pair so no cross-thread reduction is needed.
//   sv [B][3][N], sq [B][3][T] scores ; av, aq same layout ; vhat, qhat [3][B][d]
constexpr int ATTN_SPLIT = 4;
__global__ void __launch_bounds__(256) attn_finish_kernel(const float* __restrict__ sv, const float* __restrict__ sq,
                                                          const float* __restrict__ cv, const float* __restrict__ cq,
                                                          const __nv_bfloat16* __restrict__ Vp, int64_t v_ld, int64_t v_ps,
                                                          const float* __restrict__ q0,
                                                          const float* __restrict__ q1, const float* __restrict__ q2,
                                                          float* __restrict__ av, float* __restrict__ aq,
                                                          float* __restrict__ vhat, float* __restrict__ qhat,
                                                          int B, int N, int T, int d) {
  pdl_enter();
  extern __shared__ float sm[];
  float* a_sm = sm;              // [3][N]
  float* q_sm = sm + 3 * N;      // [3][T]
  const int b = blockIdx.x, sp = blockIdx.y, w = threadIdx.x >> 5;
  // every block needs the weights; only split 0 publishes them (the other blocks write to a scratch row of shared memory)
  float* dump = sm + 3 * (N + T);
  if (w < 3) softmax_warp(sv + ((int64_t)b * 3 + w) * N, cv[0], N, a_sm + w * N, sp == 0 ? av + ((int64_t)b * 3 + w) * N : dump + w * N);
  else if (w < 6) softmax_warp(sq + ((int64_t)b * 3 + (w - 3)) * T, cq[0], T, q_sm + (w - 3) * T,
                               sp == 0 ? aq + ((int64_t)b * 3 + (w - 3)) * T : dump + 3 * N + (w - 3) * T);
  __syncthreads();
  const int d2 = d >> 1;
  // V as its operand planes (hi + lo = the fp32 value to 2^-17): the same bytes as the fp32 rows, and the only copy of the image
  // features that exists in the layout this loop wants when the caller handed over a channel-major view
  const int64_t vl2 = v_ld >> 1;
  const uint32_t* Vh = reinterpret_cast<const uint32_t*>(Vp + (int64_t)b * N * v_ld);
  const uint32_t* Vl = reinterpret_cast<const uint32_t*>(Vp + v_ps + (int64_t)b * N * v_ld);
  auto ldv = [&](int n, int c) {
    const uint32_t h = __ldg(Vh + (int64_t)n * vl2 + c), l = __ldg(Vl + (int64_t)n * vl2 + c);
    return make_float2(__uint_as_float(h << 16) + __uint_as_float(l << 16), __uint_as_float(h & 0xffff0000u) + __uint_as_float(l & 0xffff0000u));
  };
  for (int c = threadIdx.x; c < d2; c += blockDim.x) {
    float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0;
    int n = sp;
    constexpr int UNR = 8;                       // rows in flight per thread (8-byte loads: the loop is latency-bound with fewer)
    for (; n + (UNR - 1) * ATTN_SPLIT < N; n += UNR * ATTN_SPLIT) {
      float2 v[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) v[u] = ldv(n + u * ATTN_SPLIT, c);
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int nn = n + u * ATTN_SPLIT;
        const float w0 = a_sm[nn], w1 = a_sm[N + nn], w2 = a_sm[2 * N + nn];
        a0.x = fmaf(w0, v[u].x, a0.x); a0.y = fmaf(w0, v[u].y, a0.y);
        a1.x = fmaf(w1, v[u].x, a1.x); a1.y = fmaf(w1, v[u].y, a1.y);
        a2.x = fmaf(w2, v[u].x, a2.x); a2.y = fmaf(w2, v[u].y, a2.y);
      }
    }
    for (; n < N; n += ATTN_SPLIT) {
      const float2 v = ldv(n, c);
      const float w0 = a_sm[n], w1 = a_sm[N + n], w2 = a_sm[2 * N + n];
      a0.x = fmaf(w0, v.x, a0.x); a0.y = fmaf(w0, v.y, a0.y);
      a1.x = fmaf(w1, v.x, a1.x); a1.y = fmaf(w1, v.y, a1.y);
      a2.x = fmaf(w2, v.x, a2.x); a2.y = fmaf(w2, v.y, a2.y);
    }
    float* o = vhat + (int64_t)b * d + 2 * c;
    atomicAdd(o, a0.x); atomicAdd(o + 1, a0.y);
    atomicAdd(o + (int64_t)B * d, a1.x); atomicAdd(o + (int64_t)B * d + 1, a1.y);
    atomicAdd(o + (int64_t)2 * B * d, a2.x); atomicAdd(o + (int64_t)2 * B * d + 1, a2.y);
    if (sp < 3) {
      const int l = sp;
      const float2* Q2 = reinterpret_cast<const float2*>((l == 0 ? q0 : (l == 1 ? q1 : q2)) + (int64_t)b * T * d);
      float2 acc = make_float2(0.f, 0.f);
      for (int t = 0; t < T; ++t) {
        const float2 v = __ldg(Q2 + (int64_t)t * d2 + c);
        const float wq = q_sm[l * T + t];
        acc.x = fmaf(wq, v.x, acc.x); acc.y = fmaf(wq, v.y, acc.y);
      }
      reinterpret_cast<float2*>(qhat + ((int64_t)l * B + b) * d)[c] = acc;
    }
  }
}

__device__ __forceinline__ float bf2f_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf2f_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// dots of one row (hi + lo bf16 planes, d columns) with up to 3 vectors held in smem; warp-cooperative
template <int NG>
__device__ __forceinline__ void row_dots(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, int d,
                                         const float* g_sm, int g_stride, float (&out)[NG]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int g = 0; g < NG; ++g) out[g] = 0.f;
  for (int c = lane * 8; c < d; c += 256) {
    const uint4 h = __ldg(reinterpret_cast<const uint4*>(hi + c));
    const uint4 l = __ldg(reinterpret_cast<const uint4*>(lo + c));
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
    float x[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      x[2 * k] = bf2f_lo(hw[k]) + bf2f_lo(lw[k]);
      x[2 * k + 1] = bf2f_hi(hw[k]) + bf2f_hi(lw[k]);
    }
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      // two 16-byte shared loads per vector (lane stride 32 B: conflict-free per quarter warp) instead of eight scalar loads
      // whose lane stride of 8 words put the whole warp on four banks
      const float4 g0 = *reinterpret_cast<const float4*>(g_sm + g * g_stride + c);
      const float4 g1 = *reinterpret_cast<const float4*>(g_sm + g * g_stride + c + 4);
      out[g] = fmaf(x[0], g0.x, out[g]); out[g] = fmaf(x[1], g0.y, out[g]); out[g] = fmaf(x[2], g0.z, out[g]); out[g] = fmaf(x[3], g0.w, out[g]);
      out[g] = fmaf(x[4], g1.x, out[g]); out[g] = fmaf(x[5], g1.y, out[g]); out[g] = fmaf(x[6], g1.z, out[g]); out[g] = fmaf(x[7], g1.w, out[g]);
    }
  }
#pragma unroll
  for (int g = 0; g < NG; ++g) out[g] = warp_sum(out[g]);
}

// Same dots with the vectors held in REGISTERS: lane l always multiplies the columns 8 l + 256 j, so for d <= 512 the 2 x 8 values of
// each vector it needs are loaded once per block and reused for every row (no shared-memory traffic in the row loop at all).
template <int NG>
struct RowDotRegs {
  float g[NG][2][8];
  __device__ __forceinline__ void load(const float* g_sm, int g_stride, int d) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int v = 0; v < NG; ++v)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int c = lane * 8 + 256 * j + k;
          g[v][j][k] = c < d ? g_sm[v * g_stride + c] : 0.f;
        }
  }
  // all four 16-byte loads of a row (hi / lo x two column groups) are issued before the first multiply
  __device__ __forceinline__ void load_row(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, int d, uint4 (&h)[2],
                                           uint4 (&l)[2]) const {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = lane * 8 + 256 * j;
      h[j] = make_uint4(0u, 0u, 0u, 0u);
      l[j] = make_uint4(0u, 0u, 0u, 0u);
      if (c < d) {
        h[j] = __ldg(reinterpret_cast<const uint4*>(hi + c));
        l[j] = __ldg(reinterpret_cast<const uint4*>(lo + c));
      }
    }
  }
  __device__ __forceinline__ void fma_row(const uint4 (&h)[2], const uint4 (&l)[2], float (&out)[NG]) const {
#pragma unroll
    for (int v = 0; v < NG; ++v) out[v] = 0.f;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint32_t hw[4] = {h[j].x, h[j].y, h[j].z, h[j].w}, lw[4] = {l[j].x, l[j].y, l[j].z, l[j].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float x0 = bf2f_lo(hw[k]) + bf2f_lo(lw[k]), x1 = bf2f_hi(hw[k]) + bf2f_hi(lw[k]);
#pragma unroll
        for (int v = 0; v < NG; ++v) out[v] = fmaf(x1, g[v][j][2 * k + 1], fmaf(x0, g[v][j][2 * k], out[v]));
      }
    }
  }
  __device__ __forceinline__ void dots(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, int d, float (&out)[NG]) const {
    uint4 h[2], l[2];
    load_row(hi, lo, d, h, l);
    fma_row(h, l, out);
#pragma unroll
    for (int v = 0; v < NG; ++v) out[v] = warp_sum(out[v]);
  }
};

// one warp: ds = a * (da - <a, da>) ; *dc += sum ds
__device__ __forceinline__ void softmax_bwd_warp(const float* __restrict__ a, const float* da_sm, int L, float* __restrict__ ds_out,
                                                 float* __restrict__ dc) {
  const int lane = threadIdx.x & 31;
  float dot = 0.f;
  for (int i = lane; i < L; i += 32) dot = fmaf(a[i], da_sm[i], dot);
  dot = warp_sum(dot);
  float tot = 0.f;
  for (int i = lane; i < L; i += 32) {
    const float v = a[i] * (da_sm[i] - dot);
    ds_out[i] = v;
    tot += v;
  }
  tot = warp_sum(tot);
  if (lane == 0) atomicAdd(dc, tot);
}

// da_v[b][l][n] = V[b][n,:] . g_v[l][b] for the three levels in one pass over V[b] (bf16 hi + lo planes), and
// da_q[b][l][t] = Q_l[b][t,:] . g_q[l][b].  ATTN_SPLIT blocks per sample, each taking every ATTN_SPLIT-th row (one warp per row).
//   Vp planes [2][B*N][d], Qp planes [2][B][3T][d] ; gv, gq [3][B][d] ; dav [B][3][N], daq [B][3][T]
__global__ void __launch_bounds__(256, 3) attn_bwd_dots_kernel(const __nv_bfloat16* __restrict__ Vp, int64_t v_ps,
                                                            const __nv_bfloat16* __restrict__ Qp, int64_t q_ps,
                                                            const float* __restrict__ gv, const float* __restrict__ gq,
                                                            float* __restrict__ dav, float* __restrict__ daq, int B, int N, int T, int d) {
  pdl_enter();
  extern __shared__ __align__(16) float sm16[];
  float* gv_sm = sm16;                // [3][d]
  float* gq_sm = sm16 + 3 * d;        // [3][d]
  const int b = blockIdx.x, sp = blockIdx.y, w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i < 3 * d; i += blockDim.x) {
    const int l = i / d, c = i - l * d;
    gv_sm[i] = gv[((int64_t)l * B + b) * d + c];
    gq_sm[i] = gq[((int64_t)l * B + b) * d + c];
  }
  __syncthreads();
  if (d <= 512) {
    RowDotRegs<3> rd;
    rd.load(gv_sm, d, d);
    // (a two-rows-per-iteration version of this loop measured 8 us SLOWER under ncu: 46.8 vs 38.8 us)
    for (int n = sp * nw + w; n < N; n += nw * ATTN_SPLIT) {
      const __nv_bfloat16* hi = Vp + ((int64_t)b * N + n) * d;
      float o[3];
      rd.dots(hi, hi + v_ps, d, o);
      if (lane == 0) {
        float* dst = dav + (int64_t)b * 3 * N + n;
        dst[0] = o[0]; dst[N] = o[1]; dst[2 * N] = o[2];
      }
    }
  } else {
    for (int n = sp * nw + w; n < N; n += nw * ATTN_SPLIT) {
      const __nv_bfloat16* hi = Vp + ((int64_t)b * N + n) * d;
      float o[3];
      row_dots<3>(hi, hi + v_ps, d, gv_sm, d, o);
      if (lane == 0) {
        float* dst = dav + (int64_t)b * 3 * N + n;
        dst[0] = o[0]; dst[N] = o[1]; dst[2 * N] = o[2];
      }
    }
  }
  for (int r = sp * nw + w; r < 3 * T; r += nw * ATTN_SPLIT) {
    const int l = r / T;
    const __nv_bfloat16* hi = Qp + ((int64_t)b * 3 * T + r) * d;
    float o[1];
    row_dots<1>(hi, hi + q_ps, d, gq_sm + l * d, d, o);
    if (lane == 0) daq[(int64_t)b * 3 * T + r] = o[0];
  }
}
// softmax backward of all six attention vectors of a sample: ds = a * (da - <a, da>) ; dc += sum ds.  One block (6 warps) per sample.
__global__ void __launch_bounds__(192) attn_bwd_softmax_kernel(const float* __restrict__ av, const float* __restrict__ aq,
                                                               const float* __restrict__ dav, const float* __restrict__ daq,
                                                               float* __restrict__ dsv, float* __restrict__ dsq,
                                                               float* __restrict__ dcv, float* __restrict__ dcq, int N, int T) {
  pdl_enter();
  const int b = blockIdx.x, w = threadIdx.x >> 5;
  if (w < 3) {
    const int64_t o = ((int64_t)b * 3 + w) * N;
    softmax_bwd_warp(av + o, dav + o, N, dsv + o, dcv);
  } else {
    const int64_t o = ((int64_t)b * 3 + (w - 3)) * T;
    softmax_bwd_warp(aq + o, daq + o, T, dsq + o, dcq);
  }
}

// dV[b][n][:] += sum_l av[b][l][n] * gv[l][b][:]      (only when the image features need a gradient)
__global__ void __launch_bounds__(256) dv_rank3_kernel(float* __restrict__ dV, const float* __restrict__ av, const float* __restrict__ gv,
                                                       int B, int N, int d) {
  pdl_enter();
  const int d4 = d >> 2;
  const int64_t total = (int64_t)B * N * d4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % d4);
    const int64_t bn = i / d4;
    const int b = (int)(bn / N), n = (int)(bn - (int64_t)b * N);
    float4 acc = reinterpret_cast<float4*>(dV)[i];
#pragma unroll
    for (int l = 0; l < 3; ++l) {
      const float a = av[((int64_t)b * 3 + l) * N + n];
      const float4 g = __ldg(reinterpret_cast<const float4*>(gv + ((int64_t)l * B + b) * d) + c);
      acc.x = fmaf(a, g.x, acc.x); acc.y = fmaf(a, g.y, acc.y); acc.z = fmaf(a, g.z, acc.z); acc.w = fmaf(a, g.w, acc.w);
    }
    reinterpret_cast<float4*>(dV)[i] = acc;
  }
}

// ---------------------------------------------------------------------------------------------------------------
struct Pl {                 // bf16 hi/lo planes of a [rows][cols] matrix, leading dimension ld
  __nv_bfloat16* p = nullptr;
  int64_t ld = 0, ps = 0;   // plane stride (elements)
  int64_t rows = 0;
  int cols = 0;
};
size_t pl_bytes(int64_t rows, int cols) { return align_up((size_t)2 * rows * round8(cols) * 2); }
Pl pl_at(void* base, int64_t rows, int cols) {
  Pl r;
  r.p = (__nv_bfloat16*)base;
  r.ld = round8(cols);
  r.ps = rows * r.ld;
  r.rows = rows;
  r.cols = cols;
  return r;
}
// operand view: batch entry z = rows [row0 + z * rows_per_entry, + rows_used) of the plane matrix
TcOperand opv(const Pl& pl, int64_t row0, int rows_used, int64_t rows_per_entry, int nbatch, bool mn_major, int zdiv = 1) {
  TcOperand o;
  o.planes = pl.p + row0 * pl.ld;
  o.ld = pl.ld;
  o.plane_stride = pl.ps;
  o.batch_stride = rows_per_entry * pl.ld;
  o.nbatch = nbatch;
  o.zdiv = zdiv;
  o.rows = rows_used;
  o.cols = pl.cols;
  o.mn_major = mn_major;
  return o;
}
TcPlanes plv(const Pl& pl, int64_t row0, int64_t rows_per_entry, int nbatch, int zdiv = 1) {
  TcPlanes t;
  t.p = pl.p + row0 * pl.ld;
  t.ld = pl.ld;
  t.plane_stride = pl.ps;
  t.batch_stride = rows_per_entry * pl.ld;
  t.nbatch = nbatch;
  t.zdiv = zdiv;
  return t;
}


// Question-side backward through tanh: dZq[r][j] = dsq[r] * wq[j] * (1 - Hq[r][j]^2) as bf16 hi/lo planes, dwq[j] += sum_r Hq[r][j] dsq[r].
// Hq [rows][d] comes as the planes the forward pass saved.  Block = a slab of HQ_ROWS rows x 128 columns: a lane owns 4 columns, the 8 warps
// take the rows of the slab in turn, and the column sums meet in shared memory -- one atomic per column and block (98 per column at B = 160:
// with one atomic per column and 16-row slab the 400 k atomics onto 16 cache lines WERE the kernel: 41 us for 50 MB).
constexpr int HQ_ROWS = 128, HQ_WARPS = 8;
__global__ void __launch_bounds__(32 * HQ_WARPS) hq_bwd_kernel(const __nv_bfloat16* __restrict__ hq, int64_t ld, int64_t ps,
                                                               const float* __restrict__ dsq, const float* __restrict__ wq,
                                                               __nv_bfloat16* __restrict__ dz, int64_t dz_ld, int64_t dz_ps,
                                                               float* __restrict__ dwq, int64_t rows, int d) {
  pdl_enter();
  __shared__ float red[HQ_WARPS][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.y * 128 + 4 * lane;
  const bool col_ok = c < d;                     // (d % 4 == 0: a lane's 4 columns are inside or outside together)
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (col_ok) {
    const float4 w = *reinterpret_cast<const float4*>(wq + c);
    const float wv[4] = {w.x, w.y, w.z, w.w};
    const int64_t r0 = (int64_t)blockIdx.x * HQ_ROWS, r1 = r0 + HQ_ROWS < rows ? r0 + HQ_ROWS : rows;
#pragma unroll 8
    for (int64_t r = r0 + warp; r < r1; r += HQ_WARPS) {
      const uint2 hh = *reinterpret_cast<const uint2*>(hq + r * ld + c);
      const uint2 hl = *reinterpret_cast<const uint2*>(hq + ps + r * ld + c);
      const float ds = __ldg(dsq + r);
      const float h[4] = {__uint_as_float(hh.x << 16) + __uint_as_float(hl.x << 16), __uint_as_float(hh.x & 0xffff0000u) + __uint_as_float(hl.x & 0xffff0000u),
                          __uint_as_float(hh.y << 16) + __uint_as_float(hl.y << 16), __uint_as_float(hh.y & 0xffff0000u) + __uint_as_float(hl.y & 0xffff0000u)};
      __nv_bfloat16 oh[4], ol[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[j] = fmaf(h[j], ds, acc[j]);
        const float g = ds * wv[j] * (1.f - h[j] * h[j]);
        oh[j] = __float2bfloat16_rn(g);
        ol[j] = __float2bfloat16_rn(g - __bfloat162float(oh[j]));
      }
      *reinterpret_cast<uint2*>(dz + r * dz_ld + c) = *reinterpret_cast<const uint2*>(oh);
      *reinterpret_cast<uint2*>(dz + dz_ps + r * dz_ld + c) = *reinterpret_cast<const uint2*>(ol);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) red[warp][4 * lane + j] = acc[j];
  __syncthreads();
  if (threadIdx.x < 128 && blockIdx.y * 128 + (int)threadIdx.x < d) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < HQ_WARPS; ++w) sum += red[w][threadIdx.x];
    atomicAdd(dwq + blockIdx.y * 128 + threadIdx.x, sum);
  }
}

struct Saved {
  Pl V, Q, PV, PQ, C;
  float *av, *aq;
  Pl Hq;          // question-side hidden state tanh(PQ_all + C_all PV) [B][3T][d]: 25 MB saved against one fabric-bound product recomputed
};
size_t saved_bytes(int B, int N, int T, int d) {
  return 2 * pl_bytes((int64_t)B * N, d) + 3 * pl_bytes((int64_t)B * 3 * T, d) + pl_bytes((int64_t)B * 3 * T, N) +
         align_up((size_t)B * 3 * N * 4) + align_up((size_t)B * 3 * T * 4) + 256;
}
bool carve_saved(Saved& s, void* buf, size_t bytes, int B, int N, int T, int d) {
  if (bytes < saved_bytes(B, N, T, d) || (reinterpret_cast<uintptr_t>(buf) & 255) != 0) return false;
  char* p = (char*)buf;
  const int64_t BN = (int64_t)B * N, BT3 = (int64_t)B * 3 * T;
  s.V = pl_at(p, BN, d); p += pl_bytes(BN, d);
  s.PV = pl_at(p, BN, d); p += pl_bytes(BN, d);
  s.Q = pl_at(p, BT3, d); p += pl_bytes(BT3, d);
  s.PQ = pl_at(p, BT3, d); p += pl_bytes(BT3, d);
  s.C = pl_at(p, BT3, N); p += pl_bytes(BT3, N);
  s.av = (float*)p; p += align_up((size_t)B * 3 * N * 4);
  s.aq = (float*)p; p += align_up((size_t)B * 3 * T * 4);
  s.Hq = pl_at(p, BT3, d);                 // (behind the attention weights: hca_coattn_saved_attention's offsets stay put)
  return true;
}
Pl take_pl(Workspace& w, int64_t rows, int cols) {
  void* p = w.take<char>(pl_bytes(rows, cols));
  return p ? pl_at(p, rows, cols) : Pl();
}
int split_to(const Pl& pl, const float* src, int64_t ld, cudaStream_t s) {
  return launch_split_planes(src, ld, pl.rows, pl.cols, pl.p, pl.ld, pl.ps, 2, s);
}

HvPlanes hvp(const Pl& pl) {
  HvPlanes h;
  h.p = pl.p; h.ld = pl.ld; h.ps = pl.ps;
  return h;
}

bool v_is_dense(int64_t sb, int64_t sn, int64_t sd, int N, int d) { return sd == 1 && sn == d && sb == (int64_t)N * d; }

}  // namespace
}  // namespace hca

extern "C" size_t hca_coattn_saved_bytes(int B, int N, int T, int d) { return hca::saved_bytes(B, N, T, d); }

extern "C" size_t hca_coattn_workspace(int B, int N, int T, int d, int need_dv) {
  using namespace hca;
  (void)need_dv;
  const int64_t BN = (int64_t)B * N, BT3 = (int64_t)B * 3 * T;
  size_t fwd = align_up((size_t)BN * d * 4) + 2 * pl_bytes(d, d) + align_up((size_t)3 * B * (N + T) * 4);
  size_t bwd = 2 * pl_bytes(d, d) + 2 * align_up((size_t)3 * B * (N + T) * 4) + pl_bytes(3 * BN, d) + 2 * pl_bytes(BT3, d) + pl_bytes(BT3, N) +
               align_up((size_t)BN * d * 4) + pl_bytes(BN, d);
  return std::max(fwd, bwd) + 4096;
}

// Byte offsets of the attention weights inside `saved` (fp32, written by hca_coattn_fwd): a_v [B][3][N] (softmax over the N regions,
// per level) and a_q [B][3][T] (softmax over all T token positions, pads included: reference model.py:387-388).  The rest of the
// buffer stays private to the library; this is the hook for attention-map export (README "Inference", left TO-DO by the reference).
extern "C" int hca_coattn_saved_attention(int B, int N, int T, int d, size_t* av_offset, size_t* aq_offset) {
  using namespace hca;
  HCA_CHECK_ARG(B > 0 && N > 0 && T > 0 && d > 0 && av_offset && aq_offset, "coattn_saved_attention: bad arguments");
  const int64_t BN = (int64_t)B * N, BT3 = (int64_t)B * 3 * T;
  const size_t av = 2 * pl_bytes(BN, d) + 2 * pl_bytes(BT3, d) + pl_bytes(BT3, N);
  *av_offset = av;
  *aq_offset = av + align_up((size_t)B * 3 * N * 4);
  return 0;
}

extern "C" int hca_coattn_fwd(const float* V, int64_t v_sb, int64_t v_sn, int64_t v_sd, const float* q0, const float* q1,
                              const float* q2, const float* Wv, const float* bv, const float* Wq, const float* bq,
                              const float* wv, const float* cv, const float* wq, const float* cq, float* vhat, float* qhat,
                              void* saved, size_t saved_sz, int B, int N, int T, int d, void* ws, size_t ws_bytes, void* stream) {
  using namespace hca;
  cudaStream_t s = (cudaStream_t)stream;
  HCA_CHECK_ARG(V && q0 && q1 && q2 && Wv && bv && Wq && bq && wv && cv && wq && cq, "coattn_fwd: null input");
  HCA_CHECK_ARG(vhat && qhat && saved, "coattn_fwd: null output");
  HCA_CHECK_ARG(B > 0 && N > 0 && T > 0 && d > 0 && d % 8 == 0, "coattn_fwd: bad sizes B=%d N=%d T=%d d=%d (d %% 8 == 0 required)", B, N, T, d);
  HCA_CHECK_ARG(tc_available(), "coattn_fwd: cuTensorMapEncodeTiled is not available from the driver");
  Saved sv_;
  HCA_CHECK_ARG(carve_saved(sv_, saved, saved_sz, B, N, T, d), "coattn_fwd: `saved` must be 256-byte aligned and hca_coattn_saved_bytes large");
  Workspace w(ws, ws_bytes);
  const int64_t BN = (int64_t)B * N, BT3 = (int64_t)B * 3 * T;
  const int T3 = 3 * T;
  const Pl& Vp = sv_.V;
  if (v_is_dense(v_sb, v_sn, v_sd, N, d)) {
    HCA_TRY(split_to(Vp, V, d, s));
  } else if (v_sn == 1 && B <= 65535) {       // channel-major feature map (the VGG encoder's [B, d, N] output viewed as [B, N, d])
    HCA_LAUNCH_K((split_planes_channel_major_kernel), dim3((N + 31) / 32, (d + 63) / 64, B), 256, 0, s, V, v_sb, v_sd, Vp.p, Vp.ld, Vp.ps, N, d);
    HCA_LAUNCHED();
  } else {                                    // any other strides: dense copy first
    float* vc = w.take<float>((size_t)BN * d);
    if (!vc) return set_err(HCA_ERR_WORKSPACE, "coattn_fwd: workspace too small");
    HCA_LAUNCH_K((gather_strided_kernel), ew_grid(BN * d), 256, 0, s, V, v_sb, v_sn, v_sd, vc, B, N, d);
    HCA_LAUNCHED();
    HCA_TRY(split_to(Vp, vc, d, s));
  }
  Pl Wvp = take_pl(w, d, d), Wqp = take_pl(w, d, d);
  float* sc = w.take<float>((size_t)3 * B * (N + T));
  if (!Wqp.p || !sc) return set_err(HCA_ERR_WORKSPACE, "coattn_fwd: workspace too small (%zu bytes)", ws_bytes);
  float* svs = sc;                          // [B][3][N]
  float* sqs = sc + (size_t)3 * B * N;      // [B][3][T]
  const Pl &Qp = sv_.Q, &PVp = sv_.PV, &PQp = sv_.PQ, &Cp = sv_.C;
  HCA_TRY(launch_split_planes_stack3(q0, q1, q2, B, T, d, Qp.p, Qp.ld, Qp.ps, s));
  {  // both projection weights -> operand planes, and the two accumulated outputs (scores, attended image features) cleared: one launch
    SplitBatch sb(s);
    HCA_TRY(sb.add(Wv, d, d, d, Wvp.p, Wvp.ld, Wvp.ps));
    HCA_TRY(sb.add(Wq, d, d, d, Wqp.p, Wqp.ld, Wqp.ps));
    ZeroBatch zb(s);
    HCA_TRY(zb.add(sc, (size_t)3 * B * (N + T) * 4));
    HCA_TRY(zb.add(vhat, (size_t)3 * B * d * 4));
    HCA_TRY(sb.flush(zb));
  }
  {  // PV = V Wv^T + bv   (once per step: level independent)
    TcEpilogue e; e.bias = bv; e.P = plv(PVp, 0, BN, 1);
    HCA_TRY(launch_gemm_tc(opv(Vp, 0, (int)BN, BN, 1, false), opv(Wvp, 0, d, d, 1, false), 2, (int)BN, d, d, e, 1, s));
  }
  {  // PQ_all = Q_all Wq^T + bq   (the three levels in one product)
    TcEpilogue e; e.bias = bq; e.P = plv(PQp, 0, BT3, 1);
    HCA_TRY(launch_gemm_tc(opv(Qp, 0, (int)BT3, BT3, 1, false), opv(Wqp, 0, d, d, 1, false), 2, (int)BT3, d, d, e, 1, s));
  }
  {  // C_all[b] = tanh(Q_all[b] V[b]^T)
    TcEpilogue e; e.act_tanh = 1; e.P = plv(Cp, 0, T3, B);
    HCA_TRY(launch_gemm_tc(opv(Qp, 0, T3, T3, B, false), opv(Vp, 0, N, N, B, false), 2, T3, N, d, e, 1, s, B));
  }
  {  // sq[b][(l,t)] = sum_j tanh(PQ_all + C_all PV)[(l,t)][j] wq[j]
    TcEpilogue e; e.mode = TC_EPI_ROWDOT; e.act_tanh = 1; e.colv = wq; e.red_row = sqs; e.red_row_batch_stride = T3;
    e.auxp = plv(PQp, 0, T3, B); e.aux_mode = TC_AUX_ADD;
    e.P = plv(sv_.Hq, 0, T3, B);            // Hq itself, for backward
    HCA_TRY(launch_gemm_tc(opv(Cp, 0, T3, T3, B, false), opv(PVp, 0, N, N, B, true), 2, T3, d, N, e, 1, s, B));
  }
  // sv[b][l][n] = sum_j tanh(PV + C_l^T PQ_l)[n][j] wv[j]: the three levels of a tile in one pass over its PV tile (hv_fused.cu)
  HCA_TRY(launch_hv_scores(hvp(Cp), hvp(PQp), hvp(PVp), wv, svs, B, N, T, d, s));
  const size_t smem = (size_t)6 * (N + T) * sizeof(float);
  HCA_CHECK_ARG(smem <= 48 * 1024, "coattn_fwd: N + T too large for the softmax kernel");
  HCA_LAUNCH_K((attn_finish_kernel), dim3(B, ATTN_SPLIT), 256, smem, s, svs, sqs, cv, cq, Vp.p, Vp.ld, Vp.ps, q0, q1, q2, sv_.av, sv_.aq, vhat, qhat, B, N, T, d);
  HCA_LAUNCHED();
  return 0;
}

extern "C" int hca_coattn_bwd(const float* Wv, const float* Wq, const float* wv, const float* wq, const void* saved, size_t saved_sz,
                              const float* gvhat, const float* gqhat, float* dV, float* dQ, float* dWv, float* dbv, float* dWq,
                              float* dbq, float* dwv, float* dcv, float* dwq, float* dcq, int B, int N, int T, int d, void* ws,
                              size_t ws_bytes, void* stream) {
  using namespace hca;
  cudaStream_t s = (cudaStream_t)stream;
  HCA_CHECK_ARG(Wv && Wq && wv && wq && saved && gvhat && gqhat, "coattn_bwd: null input");
  HCA_CHECK_ARG(dQ && dWv && dbv && dWq && dbq && dwv && dcv && dwq && dcq, "coattn_bwd: null output");
  HCA_CHECK_ARG(B > 0 && N > 0 && T > 0 && d > 0 && d % 8 == 0, "coattn_bwd: bad sizes");
  HCA_CHECK_ARG(tc_available(), "coattn_bwd: cuTensorMapEncodeTiled is not available from the driver");
  Saved sv_;
  HCA_CHECK_ARG(carve_saved(sv_, const_cast<void*>(saved), saved_sz, B, N, T, d), "coattn_bwd: bad `saved` buffer");
  Workspace w(ws, ws_bytes);
  const int64_t BN = (int64_t)B * N, BT3 = (int64_t)B * 3 * T;
  const int T3 = 3 * T;
  const Pl &Vp = sv_.V, &Qp = sv_.Q, &PVp = sv_.PV, &PQp = sv_.PQ, &Cp = sv_.C;
  Pl Wvp = take_pl(w, d, d), Wqp = take_pl(w, d, d);
  float* dsc = w.take<float>((size_t)3 * B * (N + T));
  float* dscr = w.take<float>((size_t)3 * B * (N + T));
  Pl dZv = take_pl(w, 3 * BN, d);          // [B][3][N][d]
  Pl dZq = take_pl(w, BT3, d);             // [B][3T][d]
  Pl dPQ = take_pl(w, BT3, d);
  Pl dS = take_pl(w, BT3, N);              // [B][3T][N]
  Pl dPV = take_pl(w, BN, d);
  if (!dPV.p || !dsc || !dscr) return set_err(HCA_ERR_WORKSPACE, "coattn_bwd: workspace too small (%zu bytes)", ws_bytes);
  float* dsv = dsc;                        // [B][3][N]
  float* dsq = dsc + (size_t)3 * B * N;    // [B][3][T]

  const int sk_wq = tc_splitk(d, d, (int)BT3), sk_wv = tc_splitk(d, d, (int)BN);
  {  // weight planes, and every accumulated output of this call (bias / score-vector gradients, split-K weight gradients) cleared: one launch
    SplitBatch sb(s);
    HCA_TRY(sb.add(Wq, d, d, d, Wqp.p, Wqp.ld, Wqp.ps));
    if (dV) HCA_TRY(sb.add(Wv, d, d, d, Wvp.p, Wvp.ld, Wvp.ps));
    ZeroBatch zb(s);
    HCA_TRY(zb.add(dcv, 4)); HCA_TRY(zb.add(dcq, 4));
    HCA_TRY(zb.add(dwv, (size_t)d * 4)); HCA_TRY(zb.add(dwq, (size_t)d * 4));
    HCA_TRY(zb.add(dbv, (size_t)d * 4)); HCA_TRY(zb.add(dbq, (size_t)d * 4));
    if (sk_wq > 1) HCA_TRY(zb.add(dWq, (size_t)d * d * 4));
    if (sk_wv > 1) HCA_TRY(zb.add(dWv, (size_t)d * d * 4));
    HCA_TRY(sb.flush(zb));
  }
  {
    const size_t smem = (size_t)6 * d * sizeof(float);
    HCA_CHECK_ARG(smem <= 48 * 1024, "coattn_bwd: d too large for the attention-backward kernel");
    float* dav = dscr;                       // [B][3][N]
    float* daq = dscr + (size_t)3 * B * N;   // [B][3][T]
    HCA_LAUNCH_K((attn_bwd_dots_kernel), dim3(B, ATTN_SPLIT), 256, smem, s, Vp.p, Vp.ps, Qp.p, Qp.ps, gvhat, gqhat, dav, daq, B, N, T, d);
    HCA_LAUNCHED();
    HCA_LAUNCH_K((attn_bwd_softmax_kernel), B, 192, 0, s, sv_.av, sv_.aq, dav, daq, dsv, dsq, dcv, dcq, N, T);
    HCA_LAUNCHED();
  }
  {  // dZq_all = (dsq x wq) * (1 - Hq^2) ; dwq += Hq^T dsq      (Hq saved by the forward pass: element-wise, no product)
    const int64_t rows = BT3;
    HCA_CHECK_ARG((rows + HQ_ROWS - 1) / HQ_ROWS <= 0x7fffffff && (d + 127) / 128 <= 65535, "coattn_bwd: too many rows / columns for the Hq backward kernel");
    HCA_LAUNCH_K((hq_bwd_kernel), dim3((unsigned)((rows + HQ_ROWS - 1) / HQ_ROWS), (unsigned)((d + 127) / 128)), 32 * HQ_WARPS, 0, s, sv_.Hq.p, sv_.Hq.ld,
                                                           sv_.Hq.ps, dsq, wq, dZq.p, dZq.ld, dZq.ps, dwq, rows, d);
    HCA_LAUNCHED();
  }
  // dZv_l = (dsv_l x wv) * (1 - Hv_l^2) with Hv_l = tanh(PV + C_l^T PQ_l) recomputed, dwv += Hv_l^T dsv_l, and
  // dPV = sum_l dZv_l + C_all^T dZq_all, dbv += sum_n dPV: one kernel, four TMEM accumulators per tile (hv_fused.cu)
  HCA_TRY(launch_hv_grads(hvp(Cp), hvp(PQp), hvp(PVp), hvp(dZq), wv, dsv, hvp(dZv), hvp(dPV), dwv, dbv, B, N, T, d, s));
  {  // dPQ_l^T [d x T] = dZv_l^T C_l^T + dZq_l^T   (z = 3 b + l; long axis d on M, transposed epilogue) ; dbq += sum_t dPQ
    TcEpilogue e; e.transposed = 1; e.auxp = plv(dZq, 0, T, 3 * B); e.aux_mode = TC_AUX_ADD;
    e.P = plv(dPQ, 0, T, 3 * B); e.red_row = dbq;
    HCA_TRY(launch_gemm_tc(opv(dZv, 0, N, N, 3 * B, true), opv(Cp, 0, T, T, 3 * B, false), 2, d, T, N, e, 1, s, 3 * B));
  }
  {  // dS_l^T [N x T] = (dZv_l PQ_l^T + PV dZq_l^T) * (1 - C_l^T ^2): two operand pairs chained along K, transposed epilogue
    TcEpilogue e; e.transposed = 1; e.auxp = plv(Cp, 0, T, 3 * B); e.aux_mode = TC_AUX_MUL_1MX2;
    e.P = plv(dS, 0, T, 3 * B);
    const TcOperand a2 = opv(PVp, 0, N, N, B, false, 3), b2 = opv(dZq, 0, T, T, 3 * B, false);
    HCA_TRY(launch_gemm_tc(opv(dZv, 0, N, N, 3 * B, false), opv(PQp, 0, T, T, 3 * B, false), 2, N, T, d, e, 1, s, 3 * B, &a2, &b2, d));
  }
  if (T3 <= 128) {
    // dQ_all[b] = dS_all[b] V[b] + dPQ_all[b] Wq + aq x gq   -> written straight into the [3][B][T][d] layout
    TcEpilogue e; e.D = dQ; e.ldd = d; e.d_batch_stride = (int64_t)T * d; e.d_groups = 3; e.d_group_stride = (int64_t)B * T * d;
    e.rowv = sv_.aq; e.rowv_batch_stride = T3; e.r1col = gqhat; e.r1col_batch_stride = d; e.r1_rows_per_group = T;
    e.r1_group_stride = (int64_t)B * d;
    const TcOperand a2 = opv(dPQ, 0, T3, T3, B, false), b2 = opv(Wqp, 0, d, d, 1, true);
    HCA_TRY(launch_gemm_tc(opv(dS, 0, T3, T3, B, false), opv(Vp, 0, N, N, B, true), 2, T3, d, N, e, 1, s, B, &a2, &b2, d));
  } else {
    for (int l = 0; l < 3; ++l) {          // long questions: one product per level
      TcEpilogue e; e.D = dQ + (int64_t)l * B * T * d; e.ldd = d; e.d_batch_stride = (int64_t)T * d;
      e.rowv = sv_.aq + (int64_t)l * T; e.rowv_batch_stride = T3; e.r1col = gqhat + (int64_t)l * B * d; e.r1col_batch_stride = d;
      const TcOperand a2 = opv(dPQ, (int64_t)l * T, T, T3, B, false), b2 = opv(Wqp, 0, d, d, 1, true);
      HCA_TRY(launch_gemm_tc(opv(dS, (int64_t)l * T, T, T3, B, false), opv(Vp, 0, N, N, B, true), 2, T, d, N, e, 1, s, B, &a2, &b2, d));
    }
  }
  {  // dWq = dPQ_all^T Q_all over batch, time and the three levels at once (K = 3*B*T, split-K)
    const int sk = sk_wq;
    TcEpilogue e; e.D = dWq; e.ldd = d;
    HCA_TRY(launch_gemm_tc(opv(dPQ, 0, (int)BT3, BT3, 1, true), opv(Qp, 0, (int)BT3, BT3, 1, true), 2, d, d, (int)BT3, e, sk, s));
  }
  {  // dWv = dPV^T V   (K = B*N, split-K)
    const int sk = sk_wv;
    TcEpilogue e; e.D = dWv; e.ldd = d;
    HCA_TRY(launch_gemm_tc(opv(dPV, 0, (int)BN, BN, 1, true), opv(Vp, 0, (int)BN, BN, 1, true), 2, d, d, (int)BN, e, sk, s));
  }
  if (dV) {  // only when the image features require grad (--vgg_train true)
    {  // dV = dPV Wv
      TcEpilogue e; e.D = dV; e.ldd = d;
      HCA_TRY(launch_gemm_tc(opv(dPV, 0, (int)BN, BN, 1, false), opv(Wvp, 0, d, d, 1, true), 2, (int)BN, d, d, e, 1, s));
    }
    {  // dV[b] += dS_all[b]^T Q_all[b]   (K = 3T stacked)
      TcEpilogue e; e.D = dV; e.ldd = d; e.d_batch_stride = (int64_t)N * d; e.accumulate = 1;
      HCA_TRY(launch_gemm_tc(opv(dS, 0, T3, T3, B, true), opv(Qp, 0, T3, T3, B, true), 2, N, d, T3, e, 1, s, B));
    }
    HCA_LAUNCH_K((dv_rank3_kernel), ew_grid(BN * (d / 4)), 256, 0, s, dV, sv_.av, gvhat, B, N, d);
    HCA_LAUNCHED();
  }
  return 0;
}
