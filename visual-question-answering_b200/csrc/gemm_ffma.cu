// fp32 CUDA-core GEMM with strided operands and fused epilogues (see gemm_ffma.cuh).
#include "gemm_ffma.cuh"

namespace hca {

namespace {

__device__ __forceinline__ const float* op_base(const Operand& o, int z) {
  return o.p + (int64_t)(o.zmod ? z % o.zmod : z) * o.sb;
}
__device__ __forceinline__ float mat_at(const MatRef& r, int z, int m, int n) {
  return r.p[(int64_t)(r.zmod ? z % r.zmod : z) * r.sb + (int64_t)m * r.sm + (int64_t)n * r.sn];
}

template <int TM, int TN>
__global__ void __launch_bounds__(256) gemm_ffma_kernel(const GemmParams p) {
  pdl_enter();
  constexpr int BM = 16 * TM, BN = 16 * TN, BK = 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int z = blockIdx.z / p.splitk, ks = blockIdx.z % p.splitk;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int npass = p.A2.p ? 2 : 1;
  for (int pass = 0; pass < npass; ++pass) {
    const Operand& a = pass ? p.A2 : p.A;
    const Operand& b = pass ? p.B2 : p.B;
    const int K = pass ? p.K2 : p.K;
    const float* Ap = op_base(a, z);
    const float* Bp = op_base(b, z);
    const int kchunk = ((K + BK - 1) / BK + p.splitk - 1) / p.splitk * BK;
    const int kbeg = ks * kchunk;
    const int kend = min(K, kbeg + kchunk);
    const bool a_kfast = (a.sk == 1), b_kfast = (b.sk == 1);
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const int idx = tid + 256 * i;
        const int kk = a_kfast ? (idx % BK) : (idx / BM);
        const int mm = a_kfast ? (idx / BK) : (idx % BM);
        const int gm = m0 + mm, gk = k0 + kk;
        As[kk][mm] = (gm < p.M && gk < kend) ? Ap[(int64_t)gm * a.sr + (int64_t)gk * a.sk] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < TN; ++i) {
        const int idx = tid + 256 * i;
        const int kk = b_kfast ? (idx % BK) : (idx / BN);
        const int nn = b_kfast ? (idx / BK) : (idx % BN);
        const int gn = n0 + nn, gk = k0 + kk;
        Bs[kk][nn] = (gn < p.N && gk < kend) ? Bp[(int64_t)gn * b.sr + (int64_t)gk * b.sk] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float av[TM], bv[TN];
#pragma unroll
        for (int i = 0; i < TM; i += 4) *(float4*)&av[i] = *(const float4*)&As[kk][ty * TM + i];
#pragma unroll
        for (int j = 0; j < TN; j += 4) *(float4*)&bv[j] = *(const float4*)&Bs[kk][tx * TN + j];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // ---------------------------------------------------------------- epilogue
  float rowpart[TM];
  float colpart[TN];
#pragma unroll
  for (int i = 0; i < TM; ++i) rowpart[i] = 0.f;
#pragma unroll
  for (int j = 0; j < TN; ++j) colpart[j] = 0.f;

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= p.M) continue;
    const float rv = p.rowv ? p.rowv[(int64_t)z * p.rowv_sb + m] : 0.f;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[n];
      if (p.add.p) v += mat_at(p.add, z, m, n);
      if (p.act_tanh) v = tanhf(v);
      if (p.epi == EPI_STORE) {
        if (p.r1_row) v += p.r1_row[(int64_t)z * p.r1r_sb + m] * p.r1_col[(int64_t)z * p.r1c_sb + n];
        if (p.mulx.p) {
          const float x = mat_at(p.mulx, z, m, n);
          v *= (1.f - x * x);
        }
        float* dst = p.D + (int64_t)z * p.d_sb + (int64_t)m * p.d_sm + (int64_t)n * p.d_sn;
        if (p.splitk > 1) atomicAdd(dst, v);
        else if (p.accumulate) *dst += v;
        else *dst = v;
      } else if (p.epi == EPI_ROWDOT) {
        rowpart[i] = fmaf(v, p.colv[n], rowpart[i]);
      } else {  // EPI_DZ
        const float dz = rv * p.colv[n] * (1.f - v * v);
        p.D[(int64_t)z * p.d_sb + (int64_t)m * p.d_sm + (int64_t)n * p.d_sn] = dz;
        colpart[j] = fmaf(v, rv, colpart[j]);
      }
    }
  }
  if (p.epi == EPI_ROWDOT) {
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      float s = rowpart[i];
      s += __shfl_xor_sync(0xffffffffu, s, 8);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      const int m = m0 + ty * TM + i;
      if (tx == 0 && m < p.M) atomicAdd(p.red_row + (int64_t)z * p.red_row_sb + m, s);
    }
  } else if (p.epi == EPI_DZ) {
    // reduce the column partials over the 16 row-groups through shared memory, then one atomic per column
    float(*red)[BN + 4] = Bs;  // [16][BN+4], free after the main loop's trailing barrier
#pragma unroll
    for (int j = 0; j < TN; ++j) red[ty][tx * TN + j] = colpart[j];
    __syncthreads();
    for (int c = tid; c < BN; c += 256) {
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < 16; ++r) s += red[r][c];
      if (n0 + c < p.N) atomicAdd(p.red_col + n0 + c, s);
    }
  }
}

}  // namespace

int launch_gemm_ffma(const GemmParams& p, bool big, cudaStream_t stream) {
  HCA_CHECK_ARG(p.M > 0 && p.N > 0 && p.K > 0 && p.batch > 0 && p.splitk > 0, "gemm_ffma: bad sizes M=%d N=%d K=%d batch=%d", p.M, p.N, p.K, p.batch);
  HCA_CHECK_ARG(p.splitk == 1 || (p.epi == EPI_STORE && !p.act_tanh && !p.mulx.p), "gemm_ffma: split-K needs a linear epilogue");
  const int BM = big ? 128 : 64, BN = big ? 128 : 64;
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, p.batch * p.splitk);
  HCA_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "gemm_ffma: grid too large");
  if (big) HCA_LAUNCH_K((gemm_ffma_kernel<8, 8>), grid, 256, 0, stream, p);
  else HCA_LAUNCH_K((gemm_ffma_kernel<4, 4>), grid, 256, 0, stream, p);
  HCA_LAUNCHED();
  return 0;
}

}  // namespace hca
