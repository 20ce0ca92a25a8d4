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
// An argmax flip re-routes a gradient element, so the pooled indices must be those of an fp32 evaluation.
// The convs run on tcgen05 as a bf16x2 split (~2^-16 operand precision); every channel triple whose top-2 gap lies inside
// the error band of that split (scaled by the norms of the rows involved) is then recomputed in plain fp32 -- see tie_band().
#include <algorithm>
#include "common.cuh"
#include "gemm_tc.cuh"
#include "util_kernels.cuh"

namespace hca {
namespace {

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

// Max-pool argmax must match the reference bit for bit, but the tensor-core conv (bf16x2 operand split, 3 MMAs) is not exact:
// each operand carries a residual of at most 2^-18 relative and the lo.lo product is dropped, so one term x_i w_i is off by at
// most 3 * 2^-18 |x_i w_i| and a pre-activation by at most 3 * 2^-18 * sum |x_i w_i| <= 1.15e-5 * ||x|| ||w|| (Cauchy-Schwarz; x =
// the taps of the token window, w = the filter row).  tie_band() turns that into a band on the gap of two post-tanh values:
//     EPS = 2^-16 (the bound above with a third to spare for the fp32 accumulation in TMEM)
//     err = EPS * ||x_window|| * ||w_row||                       (largest of the three channels of the triple)
//     band = 2 * err * (slope + err) + 1e-6,  slope = 1 - min(a^2, b^2) >= tanh'(.) at either value
// The band SCALES with the operands (a model whose embeddings or filters have grown 8x gets an 8x..64x wider band), and a saturated
// triple (values within ulps of +-1) has slope ~ 0, so only the 1e-6 floor is left -- enough for the 2e-7 of tanh_fast.  The pool
// kernel lists every triple whose top-2 gap is inside the band; fixup_ties_kernel recomputes the listed triples in plain fp32 from
// the fp32 inputs.  If the list overflows its capacity (R*E/4 entries) nothing is dropped: the fix-up kernel then re-derives the
// condition for EVERY element from `cat` and repairs those (slower, still exact) -- the repair is fail-safe, never silent.
constexpr float TIE_EPS = 1.52587890625e-5f;      // 2^-16

// squared norms the band needs, one warp per row: xn2[r] = ||x[r,:]||^2 for the R token rows, wn[c] = ||W_k[o,:,:]|| for the 3E
// channels c = (k-1) E + o of [uni|bi|tri]
__global__ void __launch_bounds__(256) conv_norms_kernel(const float* __restrict__ x, const float* __restrict__ w1, const float* __restrict__ w2,
                                                         const float* __restrict__ w3, float* __restrict__ xn2, float* __restrict__ wn, int R, int E) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < R + 3 * E; row += warps) {
    const float* src;
    int n;
    if (row < R) {
      src = x + (int64_t)row * E;
      n = E;
    } else {
      const int c = row - R, k = c / E + 1, o = c - (k - 1) * E;
      src = (k == 1 ? w1 : (k == 2 ? w2 : w3)) + (int64_t)o * E * k;
      n = E * k;
    }
    float acc = 0.f;
    for (int i = lane * 4; i < n; i += 128) {             // E % 4 == 0: rows are 16-byte aligned
      const float4 v = __ldg(reinterpret_cast<const float4*>(src + i));
      acc = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, acc))));
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      if (row < R) xn2[row] = acc;
      else wn[row - R] = sqrtf(acc);
    }
  }
}

// band on the top-2 gap of the triple e of row r (see above); a, b = the two largest post-tanh values
__device__ __forceinline__ float tie_band(const float* __restrict__ xn2, const float* __restrict__ wn, int64_t r, int t, int T, int e, int E,
                                          float a, float b) {
  const float x0 = xn2[r], xm = t > 0 ? xn2[r - 1] : 0.f, xp = t + 1 < T ? xn2[r + 1] : 0.f;
  float err = 0.f;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int c = 3 * e + j;
    const int k = c < E ? 1 : (c < 2 * E ? 2 : 3);
    const float win = k == 1 ? x0 : (k == 2 ? x0 + xm : x0 + xm + xp);
    err = fmaxf(err, sqrtf(win) * wn[c]);
  }
  err *= TIE_EPS;
  const float slope = 1.f - fminf(a * a, b * b);
  return 2.f * err * (slope + err) + 1e-6f;
}
// the two largest of three and the first index attaining the maximum (MaxPool2d semantics: a later element wins only if strictly
// greater, or is NaN)
__device__ __forceinline__ void top2_of3(float v0, float v1, float v2, float& best, float& mid, int& bi) {
  best = v0;
  bi = 0;
  if (v1 > best || v1 != v1) { best = v1; bi = 1; }
  if (v2 > best || v2 != v2) { best = v2; bi = 2; }
  const float lo = fminf(fminf(v0, v1), v2);
  mid = v0 + v1 + v2 - best - lo;                          // middle value (approximate is fine: only a trigger)
}

// cat [R, 3E] (post tanh) -> out [R, E], idx [R, E]; rows t >= len zeroed.  Near-ties go to tie_list (count keeps counting past tie_cap).
// A thread takes FOUR consecutive channel triples: three float4 of cat (48 contiguous bytes), one float4 of out, one uchar4 of idx (E % 8 == 0).
__global__ void __launch_bounds__(256) pool3_fwd_kernel(const float* __restrict__ cat, const int64_t* __restrict__ lens,
                                                        float* __restrict__ out, uint8_t* __restrict__ idx, int B, int T, int E,
                                                        const float* __restrict__ xn2, const float* __restrict__ wn,
                                                        int* __restrict__ tie_list, int* __restrict__ tie_count, int tie_cap) {
  pdl_enter();
  const int64_t total4 = (int64_t)B * T * (E / 4);
  const int E4 = E / 4;
  for (int64_t i4 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i4 < total4; i4 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i4 / E4;
    const int e0 = (int)(i4 - r * E4) * 4;
    const int b = (int)(r / T), t = (int)(r % T);
    float best[4] = {0.f, 0.f, 0.f, 0.f};
    int bi[4] = {0, 0, 0, 0};
    if (!lens || t < lens[b]) {
      const float4* p4 = reinterpret_cast<const float4*>(cat + r * 3 * (int64_t)E + 3 * e0);
      const float4 c0 = p4[0], c1 = p4[1], c2 = p4[2];
      const float v[12] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c2.x, c2.y, c2.z, c2.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float mid;
        top2_of3(v[3 * q], v[3 * q + 1], v[3 * q + 2], best[q], mid, bi[q]);
        if (!(best[q] - mid >= tie_band(xn2, wn, r, t, T, e0 + q, E, best[q], mid))) {      // (NaN lands here too)
          const int slot = atomicAdd(tie_count, 1);
          if (slot < tie_cap) tie_list[slot] = (int)(r * E + e0 + q);
        }
      }
    }
    *reinterpret_cast<float4*>(out + r * E + e0) = make_float4(best[0], best[1], best[2], best[3]);
    *reinterpret_cast<uchar4*>(idx + r * E + e0) = make_uchar4((unsigned char)bi[0], (unsigned char)bi[1], (unsigned char)bi[2], (unsigned char)bi[3]);
  }
}

// One pre-activation of the concatenated [uni|bi|tri] axis (channel c of token row r) in plain fp32 from the fp32 inputs, by one warp;
// every lane returns tanh(sum + bias).
__device__ __forceinline__ float exact_channel(int64_t r, int c, const float* __restrict__ x, int T, const float* __restrict__ w1,
                                               const float* __restrict__ w2, const float* __restrict__ w3, const float* __restrict__ b1,
                                               const float* __restrict__ b2, const float* __restrict__ b3, int E) {
  const int lane = threadIdx.x & 31;
  const int t = (int)(r % T);
  const float* x0 = x + r * (int64_t)E;
  const bool has_m = t > 0, has_p = t + 1 < T;
  const int k = c / E + 1, o = c % E;             // which conv, which output channel
  const float* wrow = (k == 1 ? w1 : (k == 2 ? w2 : w3)) + (int64_t)o * E * k;   // conv layout [C_in][k]: lane cc reads its k taps contiguously
  // taps: k = 1: x[t] ; k = 2: x[t-1], x[t] ; k = 3: x[t-1], x[t], x[t+1] ; zeros outside [0, T)
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  if (k == 1) {
#pragma unroll 4
    for (int cc = lane; cc < E; cc += 32) a0 = fmaf(x0[cc], wrow[cc], a0);
  } else if (k == 2) {
#pragma unroll 4
    for (int cc = lane; cc < E; cc += 32) {
      const float2 w = *reinterpret_cast<const float2*>(wrow + 2 * cc);
      if (has_m) a0 = fmaf(x0[cc - E], w.x, a0);
      a1 = fmaf(x0[cc], w.y, a1);
    }
  } else {
#pragma unroll 4
    for (int cc = lane; cc < E; cc += 32) {
      const float wa = wrow[3 * cc], wb = wrow[3 * cc + 1], wc = wrow[3 * cc + 2];
      if (has_m) a0 = fmaf(x0[cc - E], wa, a0);
      a1 = fmaf(x0[cc], wb, a1);
      if (has_p) a2 = fmaf(x0[cc + E], wc, a2);
    }
  }
  const float acc = warp_sum(a0 + a1 + a2);
  return tanhf(acc + (k == 1 ? b1 : (k == 2 ? b2 : b3))[o]);
}

// Repairs the listed near-ties: one block of 3 warps per entry, one warp per channel of the triple (the three dot products of up to 3E
// terms run side by side instead of one after the other: the kernel is a latency chain, ~1000 entries for 148 SMs).  When more
// near-ties were found than the list holds, the list is ignored and every element is re-examined instead: each thread of the first
// warp re-derives the band condition of one element from `cat`, the block then repairs the flagged ones in turn.
__global__ void __launch_bounds__(96) fixup_ties_kernel(const int* __restrict__ tie_list, const int* tie_count, int tie_cap,
                                                        const float* __restrict__ cat, const int64_t* __restrict__ lens,
                                                        const float* __restrict__ xn2, const float* __restrict__ wn,
                                                        const float* __restrict__ x, int B, int T, const float* __restrict__ w1,
                                                        const float* __restrict__ w2, const float* __restrict__ w3,
                                                        const float* __restrict__ b1, const float* __restrict__ b2,
                                                        const float* __restrict__ b3, float* __restrict__ out,
                                                        uint8_t* __restrict__ idx, int E, int* __restrict__ stats) {
  pdl_enter();
  // volatile: a load through a `const __restrict__` pointer is an INVARIANT load to the compiler, which may (and in one build did)
  // hoist it above griddepcontrol.wait -- i.e. read the counter while the pool kernel is still counting (tests/test_modules_cpu.py
  // scans the SASS of every kernel for global loads ahead of the wait)
  const int found = *reinterpret_cast<const volatile int*>(tie_count);
  if (stats && blockIdx.x == 0 && threadIdx.x == 0) { stats[0] = found; stats[1] = tie_cap; }
  __shared__ float s_v[3];
  __shared__ unsigned s_mask;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  auto repair = [&](int64_t i) {                 // all 96 threads; i = element index r * E + e
    const int64_t r = i / E;
    const int e = (int)(i - r * E);
    const float v = exact_channel(r, 3 * e + w, x, T, w1, w2, w3, b1, b2, b3, E);
    if (lane == 0) s_v[w] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float best, mid;
      int bi;
      top2_of3(s_v[0], s_v[1], s_v[2], best, mid, bi);
      out[i] = best;
      idx[i] = (uint8_t)bi;
    }
    __syncthreads();
  };
  if (found <= tie_cap) {
    for (int j = blockIdx.x; j < found; j += gridDim.x) repair(tie_list[j]);
    return;
  }
  const int64_t total = (int64_t)B * T * E;
  for (int64_t base = (int64_t)blockIdx.x * 32; base < total; base += (int64_t)gridDim.x * 32) {
    if (w == 0) {
      const int64_t i = base + lane;
      bool flag = false;
      if (i < total) {
        const int64_t r = i / E;
        const int e = (int)(i - r * E);
        const int b = (int)(r / T), t = (int)(r % T);
        if (!lens || t < lens[b]) {
          const float* p = cat + r * 3 * (int64_t)E + 3 * e;
          float best, mid;
          int bi;
          top2_of3(p[0], p[1], p[2], best, mid, bi);
          flag = !(best - mid >= tie_band(xn2, wn, r, t, T, e, E, best, mid));
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, flag);
      if (lane == 0) s_mask = m;
    }
    __syncthreads();
    unsigned m = s_mask;
    __syncthreads();
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      repair(base + src);
    }
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
// The three weights as ONE block-structured operand, hi/lo planes [2][3E][3E]: row (k-1) E + o holds the taps of output channel o of the k-gram
// filter at the columns of the Acat taps it multiplies -- unigram: [E, 2E) (the token itself), bigram: [0, 2E), trigram: [0, 3E) -- and ZEROS
// elsewhere, so that the three convolutions are one product over K = 3E whose zero blocks the GEMM skips (TcEpilogue::kwin_*), and the input
// gradient is one product with the same matrix as the MN-major operand.  Also: the concatenated bias, and the pool kernel's near-tie counter.
__global__ void __launch_bounds__(256) conv_w_planes3_kernel(const float* __restrict__ w1, const float* __restrict__ w2,
                                                             const float* __restrict__ w3, const float* __restrict__ b1,
                                                             const float* __restrict__ b2, const float* __restrict__ b3,
                                                             __nv_bfloat16* __restrict__ planes, float* __restrict__ bcat, int E,
                                                             int* __restrict__ tie_count, int write_zeros) {
  pdl_enter();
  if (tie_count && blockIdx.x == 0 && threadIdx.x == 0) *tie_count = 0;
  // thread <-> (row of Wcat, input channel c): its k taps w[o][c][0..k) are contiguous in the conv layout (coalesced reads) and go to k tap
  // blocks of the row (each a coalesced stream over c); the row's structural zeros are written by the same thread when a tile may read them
  const int64_t E3 = 3 * (int64_t)E, total = E3 * E, ps = E3 * E3;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(g % E);
    const int row = (int)(g / E);
    const int k = row / E + 1, o = row - (k - 1) * E;          // k-gram filter, its output channel
    const float* w = (k == 1 ? w1 : (k == 2 ? w2 : w3)) + ((int64_t)o * E + c) * k;
    const int first = k == 1 ? 1 : 0;                          // tap block of the filter's tap 0
#pragma unroll
    for (int jb = 0; jb < 3; ++jb) {
      const int j = jb - first;
      const bool has = j >= 0 && j < k;
      if (!has && !write_zeros) continue;      // (E % 128 == 0: no tile of the products straddles a block boundary, the zero blocks are never read)
      const float x = has ? w[j] : 0.f;
      const __nv_bfloat16 h = __float2bfloat16_rn(x);
      const int64_t dst = (int64_t)row * E3 + (int64_t)jb * E + c;
      planes[dst] = h;
      planes[ps + dst] = __float2bfloat16_rn(x - __bfloat162float(h));
    }
    if (bcat && c == 0) bcat[row] = (k == 1 ? b1 : (k == 2 ? b2 : b3))[o];
  }
}
// Pool backward straight into operand planes: dcat[r][3e+j] = (j == idx) ? dout * (1 - out^2) : 0 as bf16 hi/lo planes
// [2][R][3E], plus the three bias gradients (column sums of dcat over r).  Thread <-> FOUR consecutive channel triples (one float4 of out /
// dout, 12 consecutive bf16 = three 8-byte stores per plane: a warp writes 768 contiguous bytes), block = 64 such threads x 4 row lanes over
// a slab of rows; the column sums meet in shared memory: one atomic per column and block.
constexpr int POOL_BWD_ROWS = 32;
__global__ void __launch_bounds__(256) pool3_bwd_planes_kernel(const float* __restrict__ out, const uint8_t* __restrict__ idx,
                                                               const float* __restrict__ dout, const int64_t* __restrict__ lens,
                                                               __nv_bfloat16* __restrict__ planes, int64_t ps, float* __restrict__ db1,
                                                               float* __restrict__ db2, float* __restrict__ db3, int B, int T, int E) {
  pdl_enter();
  __shared__ float red[4][64 * 12];
  const int ct = threadIdx.x & 63, rl = threadIdx.x >> 6;          // column thread, row lane
  const int e0 = (blockIdx.x * 64 + ct) * 4;                        // first of this thread's 4 channel triples
  const int64_t R = (int64_t)B * T;
  const int64_t r0 = (int64_t)blockIdx.y * POOL_BWD_ROWS, r1 = min(R, r0 + POOL_BWD_ROWS);
  float sum[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) sum[j] = 0.f;
  if (e0 < E) {
#pragma unroll 4                      // several rows' loads in flight per thread (the rows are independent)
    for (int64_t r = r0 + rl; r < r1; r += 4) {
      const int b = (int)(r / T), t = (int)(r - (int64_t)b * T);
      float g[4] = {0.f, 0.f, 0.f, 0.f};
      int jj[4] = {0, 0, 0, 0};
      if (!lens || t < lens[b]) {
        const float4 o = *reinterpret_cast<const float4*>(out + r * E + e0);
        const float4 d = *reinterpret_cast<const float4*>(dout + r * E + e0);
        const uchar4 ix = *reinterpret_cast<const uchar4*>(idx + r * E + e0);
        g[0] = d.x * (1.f - o.x * o.x); g[1] = d.y * (1.f - o.y * o.y); g[2] = d.z * (1.f - o.z * o.z); g[3] = d.w * (1.f - o.w * o.w);
        jj[0] = ix.x; jj[1] = ix.y; jj[2] = ix.z; jj[3] = ix.w;
      }
      __nv_bfloat16 hi[12], lo[12];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const __nv_bfloat16 h = __float2bfloat16_rn(g[q]);
        const __nv_bfloat16 l = __float2bfloat16_rn(g[q] - __bfloat162float(h));
        const __nv_bfloat16 z = __float2bfloat16_rn(0.f);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const bool on = jj[q] == j;
          hi[3 * q + j] = on ? h : z;
          lo[3 * q + j] = on ? l : z;
          sum[3 * q + j] += on ? g[q] : 0.f;
        }
      }
      __nv_bfloat16* ph = planes + r * 3 * (int64_t)E + 3 * e0;       // (3 e0 is a multiple of 12: 8-byte aligned)
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        reinterpret_cast<uint2*>(ph)[v] = reinterpret_cast<const uint2*>(hi)[v];
        reinterpret_cast<uint2*>(ph + ps)[v] = reinterpret_cast<const uint2*>(lo)[v];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 12; ++j) red[rl][ct * 12 + j] = sum[j];
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 12; i += blockDim.x) {
    const int c = blockIdx.x * 64 * 12 + i;                         // channel of the concatenated [uni|bi|tri] axis
    if (c < 3 * E) {
      const float v = red[0][i] + red[1][i] + red[2][i] + red[3][i];
      float* db = c < E ? db1 : (c < 2 * E ? db2 : db3);
      atomicAdd(db + (c % E), v);
    }
  }
}
// block-structured fp32 weight gradient dWcat [3E][3E] (row (k-1) E + o, column = Acat tap column) -> the three conv layouts
// w_k[o][c][j] = dWcat[(k-1) E + o][(j + (k == 1)) E + c], one launch; threads walk dWcat's rows (coalesced reads, stride-k writes)
__global__ void __launch_bounds__(256) unpack_conv_w_kernel(const float* __restrict__ dwcat, float* __restrict__ w1, float* __restrict__ w2,
                                                            float* __restrict__ w3, int E) {
  pdl_enter();
  // thread <-> (row of dWcat, input channel c): k coalesced reads (one per tap block), k contiguous writes in the conv layout
  const int64_t E3 = 3 * (int64_t)E, total = E3 * E;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(g % E);
    const int row = (int)(g / E);
    const int k = row / E + 1, o = row - (k - 1) * E;
    float* w = (k == 1 ? w1 : (k == 2 ? w2 : w3)) + ((int64_t)o * E + c) * k;
    const float* src = dwcat + (int64_t)row * E3 + (int64_t)(k == 1 ? 1 : 0) * E + c;
    for (int j = 0; j < k; ++j) w[j] = src[(int64_t)j * E];
  }
}

int g_tie_cap_override = 0;      // test hook (hca_set_option("pool_tie_cap", n)): a tiny list forces the overflow path of the tie repair

struct ConvWs {
  Workspace w;
  float *cat, *dA, *dwcat, *xn2, *wn, *bcat;
  int *tie_count, *tie_list;
  int tie_cap;
  bool ok;
  ConvWs(void* p, size_t bytes) : w(p, bytes) {}
};
size_t tie_cap_default(size_t R, int E) { return R * E / 4 + 1024; }
ConvWs carve(void* ws, size_t bytes, int B, int T, int E) {
  ConvWs c(ws, bytes);
  Workspace& w = c.w;
  const size_t R = (size_t)B * T;
  c.cat = w.take<float>(R * 3 * E);     // fwd: tanh(conv)
  c.dA = w.take<float>(R * 3 * E);
  c.dwcat = w.take<float>((size_t)9 * E * E);
  c.xn2 = w.take<float>(R);
  c.wn = w.take<float>((size_t)3 * E);
  c.bcat = w.take<float>((size_t)3 * E);
  c.tie_cap = (int)tie_cap_default(R, E);
  c.tie_count = w.take<int>(64);
  c.tie_list = w.take<int>((size_t)c.tie_cap);
  c.ok = c.tie_list != nullptr;
  if (g_tie_cap_override > 0 && g_tie_cap_override < c.tie_cap) c.tie_cap = g_tie_cap_override;
  return c;
}

}  // namespace
void set_pool_tie_cap(int n) { g_tie_cap_override = n; }
}  // namespace hca

extern "C" size_t hca_phrase_conv_pool_workspace(int B, int T, int E) {
  using hca::align_up;
  const size_t R = (size_t)B * T;
  return 2 * align_up(R * 3 * E * 4) + align_up((size_t)9 * E * E * 4) + align_up(R * 4) +
         2 * align_up((size_t)3 * E * 4) + 1024 + align_up(hca::tie_cap_default(R, E) * 4) +      // near-tie list
         2 * 2 * align_up(R * 3 * E * 2) + align_up((size_t)2 * 3 * E * 3 * E * 2) + 8192;          // bf16 planes when no `fsaved` is given
}

namespace hca {
namespace {
// operand planes kept from forward to backward: Acat [2][R][3E] and the block-structured weight matrix Wcat [2][3E][3E] (bf16); the last
// 256 bytes hold the tie-repair statistics of the forward call (int32: listed near-ties, list capacity)
struct ConvSaved {
  __nv_bfloat16* ap = nullptr;
  __nv_bfloat16* wcat = nullptr;
  int* stats = nullptr;
};
size_t conv_saved_bytes(int B, int T, int E) {
  const size_t R = (size_t)B * T;
  return align_up(2 * R * 3 * E * 2) + align_up((size_t)2 * 3 * E * 3 * E * 2) + 256;
}
bool carve_saved(ConvSaved& v, void* buf, size_t bytes, int B, int T, int E) {
  if (!buf || (reinterpret_cast<uintptr_t>(buf) & 255) || bytes < conv_saved_bytes(B, T, E)) return false;
  char* p = (char*)buf;
  const size_t R = (size_t)B * T;
  v.ap = (__nv_bfloat16*)p; p += align_up(2 * R * 3 * E * 2);
  v.wcat = (__nv_bfloat16*)p; p += align_up((size_t)2 * 3 * E * 3 * E * 2);
  v.stats = (int*)p;
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
  HCA_CHECK_ARG(B > 0 && T > 0 && E > 0 && E % 8 == 0, "phrase_conv_pool_fwd: bad sizes B=%d T=%d E=%d (E %% 8 == 0 required: 16-byte aligned operand rows)", B, T, E);
  HCA_CHECK_ARG((int64_t)B * T * E < (1LL << 31), "phrase_conv_pool_fwd: B*T*E must stay below 2^31");
  HCA_CHECK_ARG(tc_available(), "phrase_conv_pool_fwd: cuTensorMapEncodeTiled is not available from the driver");
  ConvWs c = carve(ws, ws_bytes, B, T, E);
  if (!c.ok) return set_err(HCA_ERR_WORKSPACE, "phrase_conv_pool_fwd: workspace too small (%zu bytes)", ws_bytes);
  const int R = B * T;
  // tensor cores, bf16x2 operand split (3 MMAs per product): the row-shifted operand Acat and the tap-major weights are written
  // directly as bf16 planes (no fp32 im2col / repack round trip); the three convs read column windows of the Acat planes; bias +
  // tanh fused in the epilogue; near-ties are repaired exactly below
  const int P = 2;
  const int64_t lda = 3 * (int64_t)E, a_stride = (int64_t)R * lda;
  ConvSaved sv;
  if (fsaved) HCA_CHECK_ARG(carve_saved(sv, fsaved, fsaved_bytes, B, T, E), "phrase_conv_pool_fwd: `fsaved` must be 256-byte aligned and hca_phrase_conv_pool_saved_bytes large");
  __nv_bfloat16* ap = fsaved ? sv.ap : c.w.take<__nv_bfloat16>((size_t)P * a_stride);
  if (!ap) return set_err(HCA_ERR_WORKSPACE, "phrase_conv_pool_fwd: workspace too small for operand planes");
  HCA_LAUNCH_K((im2col3_planes_kernel<2>), ew_grid((int64_t)R * 3 * E / 4), 256, 0, s, (const float4*)x, ap, a_stride, B, T, E / 4);
  HCA_LAUNCHED();
  const int64_t E3 = 3 * (int64_t)E;
  __nv_bfloat16* wcat = fsaved ? sv.wcat : c.w.take<__nv_bfloat16>((size_t)P * E3 * E3);
  if (!wcat) return set_err(HCA_ERR_WORKSPACE, "phrase_conv_pool_fwd: workspace too small for weight planes");
  HCA_LAUNCH_K((conv_w_planes3_kernel), ew_grid(E3 * E), 256, 0, s, w1, w2, w3, b1, b2, b3, wcat, c.bcat, E, c.tie_count, (E % 128) != 0 ? 1 : 0);
  HCA_LAUNCHED();
  HCA_LAUNCH_K((conv_norms_kernel), std::min(148 * 4, (R + 3 * E + 7) / 8), 256, 0, s, x, w1, w2, w3, c.xn2, c.wn, R, E);
  HCA_LAUNCHED();
  {  // cat[r][(k-1) E + o] = tanh(conv_k): ONE product over K = 3E, each filter family contracting only the taps it has
    TcOperand A, Bw;
    A.planes = ap; A.ld = lda; A.plane_stride = a_stride; A.rows = R; A.cols = (int)E3;
    Bw.planes = wcat; Bw.ld = E3; Bw.plane_stride = E3 * E3; Bw.rows = (int)E3; Bw.cols = (int)E3;
    TcEpilogue ep;
    ep.D = c.cat; ep.ldd = lda; ep.bias = c.bcat; ep.act_tanh = 1;
    ep.kwin_ncol = E;
    ep.kwin_lo[0] = E; ep.kwin_hi[0] = 2 * E;          // unigram: the token itself
    ep.kwin_lo[1] = 0; ep.kwin_hi[1] = 2 * E;          // bigram: previous token, token
    ep.kwin_lo[2] = 0; ep.kwin_hi[2] = 3 * E;          // trigram
    HCA_TRY(launch_gemm_tc(A, Bw, P, R, (int)E3, (int)E3, ep, 1, s));
  }
  HCA_LAUNCH_K((pool3_fwd_kernel), ew_grid((int64_t)R * E / 4), 256, 0, s, c.cat, lens, out, idx, B, T, E, c.xn2, c.wn, c.tie_list, c.tie_count,
               c.tie_cap);
  HCA_LAUNCHED();
  HCA_LAUNCH_K((fixup_ties_kernel), 148 * 20, 96, 0, s, c.tie_list, c.tie_count, c.tie_cap, c.cat, lens, c.xn2, c.wn, x, B, T, w1, w2, w3, b1,
               b2, b3, out, idx, E, fsaved ? sv.stats : (int*)nullptr);
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
  HCA_CHECK_ARG(B > 0 && T > 0 && E > 0 && E % 8 == 0, "phrase_conv_pool_bwd: bad sizes (E %% 8 == 0 required)");
  HCA_CHECK_ARG(tc_available(), "phrase_conv_pool_bwd: cuTensorMapEncodeTiled is not available from the driver");
  ConvWs c = carve(ws, ws_bytes, B, T, E);
  if (!c.ok) return set_err(HCA_ERR_WORKSPACE, "phrase_conv_pool_bwd: workspace too small (%zu bytes)", ws_bytes);
  const int R = B * T;
  // every operand is produced once, directly as bf16 hi/lo planes, and the six products read windows of them
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
  // weight gradients as ONE product: dWcat[oc][kk] = sum_r dcat[r][oc] * Acat[r][kk] (K = R, split-K), of which only the blocks a filter
  // family has taps in are computed (row group k: the tap columns of its window) -- 6 of the 9 E x E blocks
  const int64_t E3w = 3 * (int64_t)E;
  TcEpilogue epw;
  epw.D = c.dwcat; epw.ldd = E3w;
  epw.nwin_nrow = E;
  epw.nwin_lo[0] = E; epw.nwin_hi[0] = 2 * E;
  epw.nwin_lo[1] = 0; epw.nwin_hi[1] = 2 * E;
  epw.nwin_lo[2] = 0; epw.nwin_hi[2] = 3 * E;
  // (split-K: the pair-mode launcher re-sizes the split for its own tile grid when the shape qualifies -- any value > 1 says "D is cleared")
  int skw = tc_splitk_tiles(tc_count_tiles((int)E3w, (int)E3w, epw), R);
  if (skw < 2 && R >= 2048) skw = 2;
  {  // bias gradients and the split-K weight-gradient accumulator cleared by one launch
    ZeroBatch zb(s);
    HCA_TRY(zb.add(db1, (size_t)E * 4)); HCA_TRY(zb.add(db2, (size_t)E * 4)); HCA_TRY(zb.add(db3, (size_t)E * 4));
    if (skw > 1) HCA_TRY(zb.add(c.dwcat, (size_t)9 * E * E * 4));
    HCA_TRY(zb.flush());
  }
  HCA_LAUNCH_K((pool3_bwd_planes_kernel), dim3((E / 4 + 63) / 64, (R + POOL_BWD_ROWS - 1) / POOL_BWD_ROWS), 256, 0, s, out, idx, dout, lens, dp, pstride, db1,
                                                                                                     db2, db3, B, T, E);
  HCA_LAUNCHED();
  {
    TcOperand A, Bm;
    A.planes = dp; A.ld = ld3; A.plane_stride = pstride; A.rows = R; A.cols = (int)E3w; A.mn_major = true;
    Bm.planes = ap; Bm.ld = ld3; Bm.plane_stride = pstride; Bm.rows = R; Bm.cols = (int)E3w; Bm.mn_major = true;
    HCA_TRY(launch_gemm_tc(A, Bm, 2, (int)E3w, (int)E3w, R, epw, skw, s));
  }
  HCA_LAUNCH_K((unpack_conv_w_kernel), ew_grid((int64_t)E * E * 3), 256, 0, s, c.dwcat, dw1, dw2, dw3, E);
  HCA_LAUNCHED();
  if (dx) {
    // dA[r][kk] = sum_oc dcat[r][oc] * Wcat[oc][kk]: ONE product with the block-structured weight matrix as the MN-major operand; tap
    // block 0 (previous token) only receives from the bigram and trigram filters, block 2 (next token) only from the trigram
    const int64_t E3 = 3 * (int64_t)E;
    __nv_bfloat16* wcat = fsaved ? sv.wcat : c.w.take<__nv_bfloat16>((size_t)2 * E3 * E3);
    if (!wcat) return set_err(HCA_ERR_WORKSPACE, "phrase_conv_pool_bwd: workspace too small for weight planes");
    if (!fsaved) {
      HCA_LAUNCH_K((conv_w_planes3_kernel), ew_grid(E3 * E), 256, 0, s, w1, w2, w3, (const float*)nullptr, (const float*)nullptr,
                                                                   (const float*)nullptr, wcat, (float*)nullptr, E, (int*)nullptr, (E % 128) != 0 ? 1 : 0);
      HCA_LAUNCHED();
    }
    TcOperand A, Bm;
    A.planes = dp; A.ld = ld3; A.plane_stride = pstride; A.rows = R; A.cols = (int)E3;
    Bm.planes = wcat; Bm.ld = E3; Bm.plane_stride = E3 * E3; Bm.rows = (int)E3; Bm.cols = (int)E3; Bm.mn_major = true;
    TcEpilogue ep;
    ep.D = c.dA; ep.ldd = ld3;
    ep.kwin_ncol = E;
    ep.kwin_lo[0] = E; ep.kwin_hi[0] = 3 * E;
    ep.kwin_lo[1] = 0; ep.kwin_hi[1] = 3 * E;
    ep.kwin_lo[2] = 2 * E; ep.kwin_hi[2] = 3 * E;
    HCA_TRY(launch_gemm_tc(A, Bm, 2, R, (int)E3, (int)E3, ep, 1, s));
    HCA_LAUNCH_K((col2im3_kernel), ew_grid((int64_t)R * E / 4), 256, 0, s, (const float4*)c.dA, (float4*)dx, B, T, E / 4);
    HCA_LAUNCHED();
  }
  return 0;
}
