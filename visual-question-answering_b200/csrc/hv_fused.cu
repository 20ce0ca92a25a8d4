// Region-side hidden state of the co-attention, all three question levels in ONE kernel (reference model.py:380-381,387):
//
//   Hv_l[b] = tanh(PV[b] + C_l[b]^T PQ_l[b])            [N, d]   l = word, phrase, sentence
//
// forward  (hv_scores):  sv[b][l][n] = Hv_l[n,:] . wv                                    (the region attention scores)
// backward (hv_grads) :  dZv_l = (dsv_l (x) wv) * (1 - Hv_l^2)   -> bf16 hi/lo planes (operand of dPQ, dS)
//                        dwv  += Hv_l^T dsv_l
//                        dPV   = sum_l dZv_l + C_all^T dZq_all    -> bf16 hi/lo planes (operand of dWv, dV) ; dbv += sum_n dPV
//
// Each level is a rank-T (T = 26) update of the same PV tile followed by tanh and element-wise work: per-level launches of the
// generic GEMM spend their time in the epilogue and re-read PV, re-stage vectors and round-trip an fp32 accumulator through
// HBM three times.  Here one CTA tile (128 regions x 128 channels of one sample) keeps FOUR fp32 accumulators in TMEM
// (3 x C_l^T PQ_l with K = T, and C_all^T dZq_all with K = 3T -- all 512 columns), loads its PV tile once per 32-column
// chunk and produces every output of the tile from registers.  Same machinery as gemm_tc.cu: TMA-fed tcgen05.mma on bf16x2
// planes issued from one elected thread, two epilogue warp groups alternating 32-column chunks, swizzled staging + TMA stores.
#include <cuda.h>
#include <cuda_bf16.h>
#include <algorithm>
#include <cstring>
#include "common.cuh"
#include "gemm_tc.cuh"
#include "hv_fused.cuh"
#include "tc_ptx.cuh"

namespace hca {
namespace {
using namespace ptx;

constexpr int V_BM = 128, V_BN = 128, V_BK = 32;
constexpr int V_STAGES = 3;
constexpr int V_THREADS = 64 + 256;
constexpr uint32_t V_CHUNK = 64 * V_BK * 2;              // one TMA box of an MN-major operand: [32 k][64 mn] = 4 KB
constexpr uint32_t V_TILE = 2 * V_CHUNK;                 // 128 mn: 8 KB per plane
constexpr uint32_t V_STAGE_BYTES = 4 * V_TILE;           // A hi, A lo, B hi, B lo
constexpr uint32_t V_STG = V_BM * 128;                   // one staging / addend buffer: two [128 rows][32 bf16] plane tiles = 16 KB

struct HvMaps {
  CUtensorMap Cl, PQl;     // per-level operands: (cols, T rows, plane, 3B), box (64, 32, 1, 1), 128-byte swizzle
  CUtensorMap Ca, DZq;     // stacked operands:   (cols, 3T rows, plane, B)
  CUtensorMap PV;          // addend planes       (d, N, plane, B), box (32, 128, 1, 1), 64-byte swizzle
  CUtensorMap DZv, DPV;    // outputs             (d, N, plane, 3B) / (d, N, plane, B), box (32, 128, 1, 1), 64-byte swizzle
};
struct HvParams {
  int B, N, T, d;
  int tiles_m, tiles_n, total_tiles;
  int kbl, kbs;                 // k-blocks per level (ceil(T / 32)) and of the stacked product (ceil(3T / 32)); kbs = 0 forward
  const float* wv;              // [d]
  const float* rowv;            // backward: dsv [B][3][N]
  float* sv;                    // forward: scores [B][3][N], accumulated atomically (pre-zeroed)
  float* dwv;                   // backward: [d], accumulated atomically (pre-zeroed)
  float* dbv;                   // backward: [d], accumulated atomically (pre-zeroed)
};

__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ float col_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < off; ++j) {
      const float send = up ? v[j] : v[j + off];
      const float keep = up ? v[j + off] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

template <bool BWD>
__global__ void __launch_bounds__(V_THREADS, 1) hv_kernel(const __grid_constant__ HvMaps maps, const HvParams p) {
  pdl_enter();
  constexpr int NACC = BWD ? 4 : 3;
  constexpr uint32_t TMEM_COLS = 512;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bars[2 * V_STAGES + 2 + 2];      // full[3], empty[3], tmem_full, tmem_empty, aux[2]
  __shared__ uint32_t tmem_ptr_smem;
  __shared__ __align__(16) float wv_sm[2][V_BN];
  __shared__ float colred_sm[2][2][V_BN];                           // [group][dwv | dbv][column]
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
  auto empty_bar = [&](int s) { return smem_u32(&bars[V_STAGES + s]); };
  const uint32_t tmem_full = smem_u32(&bars[2 * V_STAGES]), tmem_empty = smem_u32(&bars[2 * V_STAGES + 1]);
  auto aux_bar = [&](int g) { return smem_u32(&bars[2 * V_STAGES + 2 + g]); };
  // smem: operand ring, then per group: two staging buffers (backward) and one addend buffer
  const uint32_t ring = smem_base;
  const uint32_t epi_base = smem_base + V_STAGES * V_STAGE_BYTES;
  constexpr uint32_t EPI_PER_GROUP = (BWD ? 2u : 0u) * V_STG + V_STG;
  const int kb_tile = 3 * p.kbl + p.kbs;

  if (threadIdx.x == 0) {
    for (int s = 0; s < V_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 8);
    mbar_init(aux_bar(0), 1);
    mbar_init(aux_bar(1), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_ptr_smem), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_ptr_smem, 0);

  if (warp == 0) {
    // ============================================================ TMA producer (one elected lane)
    if (elect_one_sync()) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int n0 = (t % p.tiles_n) * V_BN, m0 = ((t / p.tiles_n) % p.tiles_m) * V_BM, b = t / (p.tiles_n * p.tiles_m);
        for (int kb = 0; kb < kb_tile; ++kb) {
          mbar_spin(empty_bar(s), ph ^ 1);
          mbar_expect_tx(full_bar(s), V_STAGE_BYTES);
          const bool stacked = kb >= 3 * p.kbl;
          const int l = stacked ? 0 : kb / p.kbl;
          const int k0 = (stacked ? kb - 3 * p.kbl : kb - l * p.kbl) * V_BK;
          const CUtensorMap* ma = stacked ? &maps.Ca : &maps.Cl;
          const CUtensorMap* mb = stacked ? &maps.DZq : &maps.PQl;
          const int z = stacked ? b : 3 * b + l;
          const uint32_t st = ring + (uint32_t)s * V_STAGE_BYTES;
#pragma unroll
          for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              tma_load_4d(st + pl * V_TILE + c * V_CHUNK, ma, full_bar(s), m0 + c * 64, k0, pl, z);
              tma_load_4d(st + 2 * V_TILE + pl * V_TILE + c * V_CHUNK, mb, full_bar(s), n0 + c * 64, k0, pl, z);
            }
          }
          if (++s == V_STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer (one elected lane)
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(V_BN >> 3) << 17) |
                               ((uint32_t)(V_BM >> 4) << 24);
    constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    constexpr uint32_t lbo = (V_CHUNK >> 4) << 16, kstep = 2048u >> 4;
    if (elect_one_sync()) {
      const uint32_t tm = *reinterpret_cast<volatile uint32_t*>(&tmem_ptr_smem);
      const uint32_t a_base = ((ring & 0x3FFFFu) >> 4) | lbo;
      const uint32_t b_base = (((ring + 2 * V_TILE) & 0x3FFFFu) >> 4) | lbo;
      int s = 0;
      uint32_t ph = 0;
      int tile_it = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++tile_it) {
        mbar_spin(tmem_empty, (uint32_t)((tile_it & 1) ^ 1));        // the epilogue has drained the accumulators of the previous tile
        tc_fence_after();
        for (int kb = 0; kb < kb_tile; ++kb) {
          mbar_spin(full_bar(s), ph);
          tc_fence_after();
          const bool stacked = kb >= 3 * p.kbl;
          const int l = stacked ? 3 : kb / p.kbl;
          const int kin = stacked ? kb - 3 * p.kbl : kb - l * p.kbl;          // k-block index inside its product
          const int kleft = (stacked ? 3 * p.T : p.T) - kin * V_BK;
          const int nks = min(V_BK / 16, (kleft + 15) / 16);
          const uint32_t d_tmem = tm + (uint32_t)(l * V_BN);
          const uint32_t au = a_base + (uint32_t)s * (V_STAGE_BYTES >> 4), bu = b_base + (uint32_t)s * (V_STAGE_BYTES >> 4);
#pragma unroll
          for (int ks = 0; ks < V_BK / 16; ++ks) {
            if (ks < nks) {
              const uint32_t acc0 = (kin | ks) != 0 ? 1u : 0u;
              umma_bf16_one<desc_hi, idesc>(d_tmem, au + ks * kstep, bu + ks * kstep, acc0);                               // hi . hi
              umma_bf16_one<desc_hi, idesc>(d_tmem, au + ks * kstep, bu + (V_TILE >> 4) + ks * kstep, 1u);                 // hi . lo
              umma_bf16_one<desc_hi, idesc>(d_tmem, au + (V_TILE >> 4) + ks * kstep, bu + ks * kstep, 1u);                 // lo . hi
            }
          }
          umma_commit(empty_bar(s));
          if (++s == V_STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(tmem_full);
      }
    }
  } else {
    // ============================================================ epilogue: two groups alternate the 32-column chunks
    const int eg = (warp - 2) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int et = threadIdx.x - 64 - eg * 128;
    const bool leader = (et == 0);
    const uint32_t bar_id = 1u + (uint32_t)eg;
    auto epi_barrier = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory"); };
    const uint32_t stg_base = epi_base + (uint32_t)eg * EPI_PER_GROUP;              // backward: 2 staging buffers
    const uint32_t aux_base = stg_base + (BWD ? 2u : 0u) * V_STG;
    float* const wv_s = wv_sm[eg];
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    int staged_n0 = -1;
    uint32_t aux_n = 0, nstore = 0;
    // (at, ac): next (tile, chunk) of this group whose PV tile has not been requested; one load in flight per group
    int at = blockIdx.x, ac = eg;
    auto tile_n0 = [&](int t) { return (t % p.tiles_n) * V_BN; };
    auto nch = [&](int n0) { return min(V_BN / 32, (p.d - n0 + 31) / 32); };
    auto settle = [&]() {
      while (at < p.total_tiles && ac >= nch(tile_n0(at))) { at += gridDim.x; ac = eg; }
    };
    auto issue_aux = [&]() {
      const int n0 = tile_n0(at), m0 = ((at / p.tiles_n) % p.tiles_m) * V_BM, b = at / (p.tiles_n * p.tiles_m);
      mbar_expect_tx(aux_bar(eg), V_STG);
      tma_load_4d(aux_base, &maps.PV, aux_bar(eg), n0 + ac * 32, m0, 0, b);
      tma_load_4d(aux_base + V_STG / 2, &maps.PV, aux_bar(eg), n0 + ac * 32, m0, 1, b);
    };
    auto flush_colred = [&](int n0) {
      if constexpr (BWD) {
        for (int j = et; j < V_BN; j += 128) {
          if (((j >> 5) & 1) == eg && n0 + j < p.d) {
            atomicAdd(p.dwv + n0 + j, colred_sm[eg][0][j]);
            atomicAdd(p.dbv + n0 + j, colred_sm[eg][1][j]);
          }
        }
      }
    };
    settle();
    if (leader && at < p.total_tiles) issue_aux();
    int tile_it = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++tile_it) {
      const int n0 = tile_n0(t), m0 = ((t / p.tiles_n) % p.tiles_m) * V_BM, b = t / (p.tiles_n * p.tiles_m);
      const int row = m0 + r;
      const bool row_ok = row < p.N;
      if (n0 != staged_n0) {                       // per-column vectors: restaged only when the column block changes
        epi_barrier();
        if (staged_n0 >= 0) flush_colred(staged_n0);
        for (int j = et; j < V_BN; j += 128) {
          wv_s[j] = (n0 + j < p.d) ? __ldg(p.wv + n0 + j) : 0.f;
          colred_sm[eg][0][j] = 0.f;
          colred_sm[eg][1][j] = 0.f;
        }
        staged_n0 = n0;
        epi_barrier();
      }
      float rv0 = 0.f, rv1 = 0.f, rv2 = 0.f, rd0 = 0.f, rd1 = 0.f, rd2 = 0.f;     // (scalars: indexed arrays would live in local memory)
      if (BWD && row_ok) {
        rv0 = __ldg(p.rowv + ((int64_t)b * 3 + 0) * p.N + row);
        rv1 = __ldg(p.rowv + ((int64_t)b * 3 + 1) * p.N + row);
        rv2 = __ldg(p.rowv + ((int64_t)b * 3 + 2) * p.N + row);
      }
      mbar_wait(tmem_full, (uint32_t)(tile_it & 1), 21);
      tc_fence_after();
      const int nchunks = nch(n0);
#pragma unroll 1
      for (int c = eg; c < nchunks; c += 2) {
        const int col0 = n0 + c * 32;
        // this thread's row of the PV tile (hi + lo planes, 64-byte swizzle)
        mbar_wait(aux_bar(eg), aux_n & 1u, 22);
        float ax[32];
        {
          const uint32_t src = aux_base + (uint32_t)r * 64u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t h[4], l4[4];
            const uint32_t o = (uint32_t)((j ^ ((r >> 1) & 3)) * 16);
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(h[0]), "=r"(h[1]), "=r"(h[2]), "=r"(h[3]) : "r"(src + o));
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(l4[0]), "=r"(l4[1]), "=r"(l4[2]), "=r"(l4[3]) : "r"(src + V_STG / 2 + o));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              ax[8 * j + 2 * k] = bf_lo(h[k]) + bf_lo(l4[k]);
              ax[8 * j + 2 * k + 1] = bf_hi(h[k]) + bf_hi(l4[k]);
            }
          }
        }
        epi_barrier();                              // the tile is in registers: request the group's next one
        ++aux_n;
        ac += 2;
        settle();
        if (leader && at < p.total_tiles) issue_aux();
        float wvv[32];
        {
          const float4* w4 = reinterpret_cast<const float4*>(wv_s + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 w = w4[j];
            wvv[4 * j] = w.x; wvv[4 * j + 1] = w.y; wvv[4 * j + 2] = w.z; wvv[4 * j + 3] = w.w;
          }
        }
        float sum[32];
        if constexpr (BWD) {
          uint32_t v[32];
          __syncwarp();
          tmem_ld32(lane_addr + (uint32_t)(3 * V_BN + c * 32), v);
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[j] = __uint_as_float(v[j]);
        }
#pragma unroll 1
        for (int l = 0; l < 3; ++l) {
          uint32_t v[32];
          __syncwarp();
          tmem_ld32(lane_addr + (uint32_t)(l * V_BN + c * 32), v);
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = tanh_fast(__uint_as_float(v[j]) + ax[j]);
          if constexpr (!BWD) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) acc = fmaf(f[j], wvv[j], acc);          // wv is 0 beyond d
            if (l == 0) rd0 += acc;
            else if (l == 1) rd1 += acc;
            else rd2 += acc;
          } else {
            const float rvl = l == 0 ? rv0 : (l == 1 ? rv1 : rv2);
            float part[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              part[j] = f[j] * rvl;
              f[j] = rvl * wvv[j] * (1.f - f[j] * f[j]);
              sum[j] += f[j];
            }
            const float cs = col_reduce32(part, lane);
            atomicAdd(&colred_sm[eg][0][c * 32 + lane], cs);
            // dZv_l chunk -> bf16 hi/lo planes -> swizzled staging -> TMA store
            const uint32_t sbuf = stg_base + (nstore & 1u) * V_STG;
            if (leader) tma_store_wait_read<1>();                               // the store that last used this buffer has read it
            epi_barrier();
            {
              const uint32_t sb = sbuf + (uint32_t)r * 64u;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint32_t h[4], lo[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float x0 = f[8 * j + 2 * k], x1 = f[8 * j + 2 * k + 1];
                  const __nv_bfloat162 hh = __floats2bfloat162_rn(x0, x1);
                  const __nv_bfloat162 ll = __floats2bfloat162_rn(x0 - __low2float(hh), x1 - __high2float(hh));
                  h[k] = *reinterpret_cast<const uint32_t*>(&hh);
                  lo[k] = *reinterpret_cast<const uint32_t*>(&ll);
                }
                const uint32_t o = (uint32_t)((j ^ ((r >> 1) & 3)) * 16);
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sb + o), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sb + V_STG / 2 + o), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
              }
            }
            fence_proxy_async_smem();
            epi_barrier();
            if (leader) {
              tma_store_4d(&maps.DZv, sbuf, col0, m0, 0, 3 * b + l);
              tma_store_4d(&maps.DZv, sbuf + V_STG / 2, col0, m0, 1, 3 * b + l);
              tma_store_commit();
            }
            ++nstore;
          }
        }
        if constexpr (BWD) {
          // dPV chunk = sum_l dZv_l + C_all^T dZq_all: planes out, column sums -> dbv
          float part[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) part[j] = row_ok ? sum[j] : 0.f;
          const float cs = col_reduce32(part, lane);
          atomicAdd(&colred_sm[eg][1][c * 32 + lane], cs);
          const uint32_t sbuf = stg_base + (nstore & 1u) * V_STG;
          if (leader) tma_store_wait_read<1>();
          epi_barrier();
          {
            const uint32_t sb = sbuf + (uint32_t)r * 64u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t h[4], lo[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float x0 = sum[8 * j + 2 * k], x1 = sum[8 * j + 2 * k + 1];
                const __nv_bfloat162 hh = __floats2bfloat162_rn(x0, x1);
                const __nv_bfloat162 ll = __floats2bfloat162_rn(x0 - __low2float(hh), x1 - __high2float(hh));
                h[k] = *reinterpret_cast<const uint32_t*>(&hh);
                lo[k] = *reinterpret_cast<const uint32_t*>(&ll);
              }
              const uint32_t o = (uint32_t)((j ^ ((r >> 1) & 3)) * 16);
              asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sb + o), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
              asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sb + V_STG / 2 + o), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
            }
          }
          fence_proxy_async_smem();
          epi_barrier();
          if (leader) {
            tma_store_4d(&maps.DPV, sbuf, col0, m0, 0, b);
            tma_store_4d(&maps.DPV, sbuf + V_STG / 2, col0, m0, 1, b);
            tma_store_commit();
          }
          ++nstore;
        }
      }
      // all of this warp's reads of the accumulators are done: hand them back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);
      if constexpr (!BWD) {
        if (row_ok && eg < nchunks) {
          atomicAdd(p.sv + ((int64_t)b * 3 + 0) * p.N + row, rd0);
          atomicAdd(p.sv + ((int64_t)b * 3 + 1) * p.N + row, rd1);
          atomicAdd(p.sv + ((int64_t)b * 3 + 2) * p.N + row, rd2);
        }
      }
    }
    if (BWD && staged_n0 >= 0) {
      epi_barrier();
      flush_colred(staged_n0);
    }
    if (BWD && leader) tma_store_wait_read<0>();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

int make_map(CUtensorMap* tm, const HvPlanes& pl, int cols, int rows, int64_t batch_stride, int nbatch, int box_cols, int box_rows, int swz) {
  const uint64_t dims[4] = {(uint64_t)cols, (uint64_t)rows, 2, (uint64_t)nbatch};
  const uint64_t str[3] = {(uint64_t)pl.ld * 2, (uint64_t)pl.ps * 2, (uint64_t)batch_stride * 2};
  const uint32_t box[4] = {(uint32_t)box_cols, (uint32_t)box_rows, 1, 1};
  return tc_make_tmap(tm, true, 4, pl.p, dims, str, box, swz);
}

template <bool BWD>
int launch(const HvPlanes& C, const HvPlanes& PQ, const HvPlanes& PV, const HvPlanes* dZq, const HvPlanes* dZv, const HvPlanes* dPV,
           HvParams p, cudaStream_t s) {
  HvMaps maps;
  const int T = p.T, N = p.N, d = p.d, B = p.B;
  HCA_TRY(make_map(&maps.Cl, C, N, T, (int64_t)T * C.ld, 3 * B, 64, V_BK, 3));
  HCA_TRY(make_map(&maps.PQl, PQ, d, T, (int64_t)T * PQ.ld, 3 * B, 64, V_BK, 3));
  HCA_TRY(make_map(&maps.PV, PV, d, N, (int64_t)N * PV.ld, B, 32, V_BM, 2));
  if (BWD) {
    HCA_TRY(make_map(&maps.Ca, C, N, 3 * T, (int64_t)3 * T * C.ld, B, 64, V_BK, 3));
    HCA_TRY(make_map(&maps.DZq, *dZq, d, 3 * T, (int64_t)3 * T * dZq->ld, B, 64, V_BK, 3));
    HCA_TRY(make_map(&maps.DZv, *dZv, d, N, (int64_t)N * dZv->ld, 3 * B, 32, V_BM, 2));
    HCA_TRY(make_map(&maps.DPV, *dPV, d, N, (int64_t)N * dPV->ld, B, 32, V_BM, 2));
  } else {
    maps.Ca = maps.Cl; maps.DZq = maps.PQl; maps.DZv = maps.PV; maps.DPV = maps.PV;
  }
  p.tiles_m = (N + V_BM - 1) / V_BM;
  p.tiles_n = (d + V_BN - 1) / V_BN;
  const int64_t total = (int64_t)B * p.tiles_m * p.tiles_n;
  HCA_CHECK_ARG(total < (1LL << 30), "hv: too many tiles");
  p.total_tiles = (int)total;
  p.kbl = (T + V_BK - 1) / V_BK;
  p.kbs = BWD ? (3 * T + V_BK - 1) / V_BK : 0;
  const size_t smem = (size_t)V_STAGES * V_STAGE_BYTES + 2 * ((BWD ? 2 : 0) * (size_t)V_STG + V_STG) + 1024;
  static bool attr_set[64][2] = {};                  // function attributes are per device
  bool& attr_done = attr_set[current_device()][BWD ? 1 : 0];
  if (!attr_done) {
    HCA_CUDA(cudaFuncSetAttribute(hv_kernel<BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
    cudaGetLastError();
    sms = 148;
  }
  const int ctas = (int)std::min<int64_t>(sms, total);
  HCA_LAUNCH_K((hv_kernel<BWD>), ctas, V_THREADS, smem, s, maps, p);
  HCA_LAUNCHED();
  return 0;
}

bool planes_ok(const HvPlanes& t) {
  return t.p && (t.ld % 8) == 0 && (t.ps % 8) == 0 && ((reinterpret_cast<uintptr_t>(t.p) & 15) == 0);
}

}  // namespace

int launch_hv_scores(const HvPlanes& C, const HvPlanes& PQ, const HvPlanes& PV, const float* wv, float* sv, int B, int N, int T, int d,
                     cudaStream_t s) {
  HCA_CHECK_ARG(planes_ok(C) && planes_ok(PQ) && planes_ok(PV) && wv && sv && B > 0 && N > 0 && T > 0 && d > 0, "hv_scores: bad arguments");
  HvParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.N = N; p.T = T; p.d = d; p.wv = wv; p.sv = sv;
  return launch<false>(C, PQ, PV, nullptr, nullptr, nullptr, p, s);
}

int launch_hv_grads(const HvPlanes& C, const HvPlanes& PQ, const HvPlanes& PV, const HvPlanes& dZq, const float* wv, const float* dsv,
                    const HvPlanes& dZv, const HvPlanes& dPV, float* dwv, float* dbv, int B, int N, int T, int d, cudaStream_t s) {
  HCA_CHECK_ARG(planes_ok(C) && planes_ok(PQ) && planes_ok(PV) && planes_ok(dZq) && planes_ok(dZv) && planes_ok(dPV) && wv && dsv && dwv && dbv,
                "hv_grads: bad arguments");
  HvParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.N = N; p.T = T; p.d = d; p.wv = wv; p.rowv = dsv; p.dwv = dwv; p.dbv = dbv;
  return launch<true>(C, PQ, PV, &dZq, &dZv, &dPV, p, s);
}

}  // namespace hca
