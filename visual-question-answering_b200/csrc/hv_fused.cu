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
// HBM three times.  Here one CTA tile (128 regions x 64 channels of one sample) keeps its accumulators in TMEM -- 3 x C_l^T PQ_l
// (K = T) and, backward, C_all^T dZq_all = sum_l C_l^T dZq_l riding on the same k-blocks -- in TWO sets, and produces every output
// of the tile from registers.  TMA-fed tcgen05.mma on bf16x2 planes issued from one elected thread; 16 epilogue warps, each a
// self-contained worker on 32 rows x 32 columns (own PV slice by TMA, own staging + TMA stores), TMEM loads one step ahead of the
// math, tanh in pairs (3 MUFU per 2 values).  Measured history of the design in DESIGN.md section 4.
#include <cuda.h>
#include <cuda_bf16.h>
#include <algorithm>
#include <cstring>
#include "common.cuh"
#include "gemm_tc.cuh"
#include "hv_fused.cuh"
#include "tc_ptx.cuh"

#ifndef HCA_TC_TIMELINE
#define HCA_TC_TIMELINE 0
#endif

namespace hca {
namespace {
using namespace ptx;

constexpr int V_BM = 128, V_BK = 32;
constexpr uint32_t V_CHUNK = 64 * V_BK * 2;              // one TMA box of an MN-major operand: [32 k][64 mn] = 4 KB
constexpr uint32_t V_TILE_A = 2 * V_CHUNK;               // 128 regions: 8 KB per plane
// Tile shapes (BN channels per tile; NG groups of 4 epilogue warps -- one warp per TMEM lane quadrant; NBUF accumulator sets in TMEM):
//   BN = 64 : a tile's 4 accumulators take 256 TMEM columns, so the CTA keeps TWO sets: the MMAs of tile i + 1 run while the epilogue of
//             tile i is still reading.  Two teams of 2 groups each own the tiles of one parity.  (With one set the 16 epilogue warps sat
//             idle for the 2.5 k-cycle MMA phase of every tile -- per-tile timeline of profiles/timeline_hv.py -- and two CTAs per SM with
//             one set each locked into the same phase: both in their MUFU-bound epilogue, then both waiting.)
//   (BN = 128 with one set of 512 columns and the forward's per-level hand-over measured the same: 52.9 / 111.7 us against 54.4 / 111.4 us.)
template <int BN> struct HvShape;
template <> struct HvShape<64> { static constexpr int NG = 4, NBUF = 2, STAGES_FWD = 6, STAGES_BWD = 3; };
constexpr int V_BN = 64;

struct HvMaps {
  CUtensorMap Cl, PQl;     // per-level operands: (cols, T rows, plane, 3B), box (64, 32, 1, 1), 128-byte swizzle
  CUtensorMap DZq;         // backward: dZq_l, same shape as PQl
  CUtensorMap PV;          // addend planes       (d, N, plane, B), box (32, 32, 1, 1), 64-byte swizzle
  CUtensorMap DZv, DPV;    // outputs             (d, N, plane, 3B) / (d, N, plane, B), box (16, 32, 1, 1), 32-byte swizzle
};
struct HvParams {
  int B, N, T, d;
  int tiles_m, tiles_n, total_tiles;
  int kbl;                      // k-blocks per level (ceil(T / 32))
  const float* wv;              // [d]
  const float* rowv;            // backward: dsv [B][3][N]
  float* sv;                    // forward: scores [B][3][N], accumulated atomically (pre-zeroed)
  float* dwv;                   // backward: [d], accumulated atomically (pre-zeroed)
  float* dbv;                   // backward: [d], accumulated atomically (pre-zeroed)
  long long* timeline;          // debug build (HCA_BUILD_TIMELINE=1): clock64 stamps of CTA 0, [tile][16]
};

__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
// tanh of TWO values with three MUFU operations (the epilogues here are bound by the MUFU pipe: 16 lanes / clk / SM, measured): one reciprocal
// serves both -- 1 / a = b / (a b), 1 / b = a / (a b) with a = 1 + e^2x, b = 1 + e^2y.  The exponents are clamped at 2^60 (tanh is 1 to the last
// bit beyond |x| = 9) so the product stays finite; min.NaN keeps a NaN a NaN.  Absolute error ~5e-7.
__device__ __forceinline__ void tanh_fast2(float x, float y, float& tx, float& ty) {
  float ex, ey, cx, cy, r;
  asm("min.NaN.f32 %0, %1, 0f42700000;" : "=f"(cx) : "f"(x * 2.8853900817779268f));      // 60.0
  asm("min.NaN.f32 %0, %1, 0f42700000;" : "=f"(cy) : "f"(y * 2.8853900817779268f));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(cx));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ey) : "f"(cy));
  const float a = ex + 1.f, b = ey + 1.f;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a * b));
  tx = fmaf(-2.f, r * b, 1.f);
  ty = fmaf(-2.f, r * a, 1.f);
}
// column sums of 16 per-lane values over the 32 lanes of a warp: on return lanes 2 j and 2 j + 1 both hold sum_lanes v[j]
__device__ __forceinline__ float col_reduce16(float (&v)[16], int lane) {
#pragma unroll
  for (int off = 16; off >= 2; off >>= 1) {
    const bool up = (lane & off) != 0;
    const int n = off >> 1;
#pragma unroll
    for (int j = 0; j < n; ++j) {
      const float send = up ? v[j] : v[j + n];
      const float keep = up ? v[j + n] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}
// TMEM load of 16 columns WITHOUT the wait: the registers are written asynchronously until tmem_ld_wait16(v) -- which names them as
// in/out operands, so nothing that reads v can be scheduled ahead of the wait and the registers stay reserved in between.
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                 "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :
               : "memory");
}

template <bool BWD, int BN>
__global__ void __launch_bounds__(64 + 128 * HvShape<BN>::NG, 1) hv_kernel(const __grid_constant__ HvMaps maps, const HvParams p) {
  pdl_enter();
  constexpr int NG = HvShape<BN>::NG, NBUF = HvShape<BN>::NBUF, GPT = NG / NBUF;     // GPT: groups per team
  static_assert(GPT * 32 == BN, "one 32-column chunk per group of a team");
  constexpr int V_STAGES = BWD ? HvShape<BN>::STAGES_BWD : HvShape<BN>::STAGES_FWD;
  constexpr uint32_t TMEM_COLS = 512, ACC_COLS = 4 * BN;           // per accumulator set: 3 (forward) / 4 (backward) x BN columns
  constexpr uint32_t V_TILE_B = (BN / 64) * V_CHUNK;               // B operand: BN channels per plane
  // one k-block: A = C_l^T (hi, lo), B = PQ_l (hi, lo) and, backward, B2 = dZq_l (hi, lo): C_all^T dZq_all = sum_l C_l^T dZq_l reuses the A tiles
  constexpr uint32_t V_STAGE_BYTES = 2 * V_TILE_A + (BWD ? 4 : 2) * V_TILE_B;
  constexpr uint32_t NSTG = BWD ? 1u : 0u;                         // staging half-chunks per warp (a level-half of math separates two stores)
  constexpr uint32_t EPI_PER_WARP = 4096u + NSTG * 2048u;            // PV slice (4 KB) + backward staging (2 KB each)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bars[2 * V_STAGES + 6 * NBUF + 4 * NG];     // full[S], empty[S], acc_full[NBUF][3], acc_empty[NBUF][3], aux[epilogue warp]
  __shared__ uint32_t tmem_ptr_smem;
  __shared__ __align__(16) float wv_sm[NG][BN];
  __shared__ float colred_sm[NG][2][BN];                          // [group][dwv | dbv][column]
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
  auto empty_bar = [&](int s) { return smem_u32(&bars[V_STAGES + s]); };
  // accumulator hand-over between the MMA warp and the epilogue warps.  Forward: one barrier pair per (accumulator set, LEVEL) -- the
  // epilogue walks the levels in order and returns a level's accumulator as soon as it has read it, so the MMAs of the set's next tile
  // start while the remaining levels are still being worked on.  Backward (all levels of a column half are needed together): index 0 only.
  auto acc_full = [&](int bf, int l) { return smem_u32(&bars[2 * V_STAGES + bf * 3 + l]); };
  auto acc_empty = [&](int bf, int l) { return smem_u32(&bars[2 * V_STAGES + 3 * NBUF + bf * 3 + l]); };
  auto aux_bar = [&](int g) { return smem_u32(&bars[2 * V_STAGES + 6 * NBUF + g]); };
  // smem: operand ring, then per group: two staging buffers (backward) and one addend buffer
  const uint32_t ring = smem_base;
  const uint32_t epi_base = smem_base + V_STAGES * V_STAGE_BYTES;
  const int kb_tile = 3 * p.kbl;
  long long* const tl = (HCA_TC_TIMELINE && blockIdx.x == 0) ? p.timeline : nullptr;      // [tile][16], first 12 tiles of CTA 0

  if (threadIdx.x == 0) {
    for (int s = 0; s < V_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int bf = 0; bf < NBUF; ++bf) {
      for (int l = 0; l < 3; ++l) {
        mbar_init(acc_full(bf, l), 1);
        mbar_init(acc_empty(bf, l), 4 * GPT);
      }
    }
    for (int g = 0; g < 4 * NG; ++g) mbar_init(aux_bar(g), 1);
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
      int ptile = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++ptile) {
        const int n0 = (t % p.tiles_n) * BN, m0 = ((t / p.tiles_n) % p.tiles_m) * V_BM, b = t / (p.tiles_n * p.tiles_m);
        for (int kb = 0; kb < kb_tile; ++kb) {
          mbar_spin(empty_bar(s), ph ^ 1);
          if (tl && ptile < 12 && (kb == 0 || kb == kb_tile - 1)) tl[ptile * 16 + (kb == 0 ? 0 : 1)] = clock64();
          mbar_expect_tx(full_bar(s), V_STAGE_BYTES);
          const int l = kb / p.kbl;
          const int k0 = (kb - l * p.kbl) * V_BK;
          const int z = 3 * b + l;
          const uint32_t st = ring + (uint32_t)s * V_STAGE_BYTES;
#pragma unroll
          for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
            for (int c = 0; c < 2; ++c) tma_load_4d(st + pl * V_TILE_A + c * V_CHUNK, &maps.Cl, full_bar(s), m0 + c * 64, k0, pl, z);
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) {
              tma_load_4d(st + 2 * V_TILE_A + pl * V_TILE_B + c * V_CHUNK, &maps.PQl, full_bar(s), n0 + c * 64, k0, pl, z);
              if constexpr (BWD) tma_load_4d(st + 2 * V_TILE_A + (2 + pl) * V_TILE_B + c * V_CHUNK, &maps.DZq, full_bar(s), n0 + c * 64, k0, pl, z);
            }
          }
          if (++s == V_STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer (one elected lane)
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) |
                               ((uint32_t)(V_BM >> 4) << 24);
    constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    constexpr uint32_t lbo = (V_CHUNK >> 4) << 16, kstep = 2048u >> 4;
    if (elect_one_sync()) {
      const uint32_t tm = *reinterpret_cast<volatile uint32_t*>(&tmem_ptr_smem);
      const uint32_t a_base = ((ring & 0x3FFFFu) >> 4) | lbo;
      const uint32_t b_base = (((ring + 2 * V_TILE_A) & 0x3FFFFu) >> 4) | lbo;
      int s = 0;
      uint32_t ph = 0;
      int tile_it = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++tile_it) {
        const int bf = tile_it % NBUF;                                // accumulator set of this tile
        const uint32_t free_parity = (uint32_t)(((tile_it / NBUF) & 1) ^ 1);
        if constexpr (BWD) {
          mbar_spin(acc_empty(bf, 0), free_parity);                    // its team has drained the set's previous tile
          tc_fence_after();
        }
        if (tl && tile_it < 12) tl[tile_it * 16 + 2] = clock64();
        for (int kb = 0; kb < kb_tile; ++kb) {
          mbar_spin(full_bar(s), ph);
          tc_fence_after();
          if (tl && tile_it < 12 && (kb == 0 || kb == kb_tile - 1)) tl[tile_it * 16 + (kb == 0 ? 3 : 4)] = clock64();
          const int l = kb / p.kbl;
          const int kin = kb - l * p.kbl;                                     // k-block index inside its product
          if constexpr (!BWD) {
            if (kin == 0) {
              mbar_spin(acc_empty(bf, l), free_parity);                  // the team has read this level's accumulator of the previous tile
              tc_fence_after();
            }
          }
          const int kleft = p.T - kin * V_BK;
          const int nks = min(V_BK / 16, (kleft + 15) / 16);
          const uint32_t d_tmem = tm + (uint32_t)(bf * (int)ACC_COLS + l * BN);
          const uint32_t au = a_base + (uint32_t)s * (V_STAGE_BYTES >> 4), bu = b_base + (uint32_t)s * (V_STAGE_BYTES >> 4);
#pragma unroll
          for (int ks = 0; ks < V_BK / 16; ++ks) {
            if (ks < nks) {
              const uint32_t acc0 = (kin | ks) != 0 ? 1u : 0u;
              umma_bf16_one<desc_hi, idesc>(d_tmem, au + ks * kstep, bu + ks * kstep, acc0);                               // hi . hi
              umma_bf16_one<desc_hi, idesc>(d_tmem, au + ks * kstep, bu + (V_TILE_B >> 4) + ks * kstep, 1u);               // hi . lo
              umma_bf16_one<desc_hi, idesc>(d_tmem, au + (V_TILE_A >> 4) + ks * kstep, bu + ks * kstep, 1u);               // lo . hi
              if constexpr (BWD) {      // accumulator 3 += C_l^T dZq_l
                const uint32_t d3 = tm + (uint32_t)(bf * (int)ACC_COLS + 3 * BN), b2 = bu + (2 * V_TILE_B >> 4);
                umma_bf16_one<desc_hi, idesc>(d3, au + ks * kstep, b2 + ks * kstep, (l | kin | ks) != 0 ? 1u : 0u);
                umma_bf16_one<desc_hi, idesc>(d3, au + ks * kstep, b2 + (V_TILE_B >> 4) + ks * kstep, 1u);
                umma_bf16_one<desc_hi, idesc>(d3, au + (V_TILE_A >> 4) + ks * kstep, b2 + ks * kstep, 1u);
              }
            }
          }
          umma_commit(empty_bar(s));
          if constexpr (!BWD) {
            if (kin == p.kbl - 1) umma_commit(acc_full(bf, l));         // this level's accumulator is complete
          }
          if (++s == V_STAGES) { s = 0; ph ^= 1; }
        }
        if constexpr (BWD) umma_commit(acc_full(bf, 0));
        if (tl && tile_it < 12) tl[tile_it * 16 + 5] = clock64();
      }
    }
  } else {
    // ============================================================ epilogue: 4 NG warps, every one a self-contained worker
    // Warp w reads TMEM lane quadrant q = w % 4 (rows 32 q .. 32 q + 31 of the tile) and is one of NG warps on that quadrant; the 32-column
    // chunks of the CTA's tiles go round-robin to the NG "slots".  A warp fetches its own 32-row slice of the PV tile (own mbarrier), and
    // stores its own rows of the outputs from a private staging buffer (own TMA bulk groups): no barrier between warps in steady state.
    const int wi = warp - 2;
    const int group = wi >> 2;                                       // 4 warps, one per lane quadrant
    const int team = group / GPT, slot = group % GPT;                // the team owns the tiles of one parity (one accumulator set)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t bar_id = 1u + (uint32_t)group;                    // the 4 warps of a group meet only when the column block changes
    auto slot_barrier = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory"); };
    const int st = (int)threadIdx.x - 64 - group * 128;              // thread index inside the group
    const int tile_first = blockIdx.x + team * gridDim.x, tile_step = NBUF * gridDim.x;
    const uint32_t aux_w = epi_base + (uint32_t)wi * EPI_PER_WARP;   // PV slice: [32 rows][64 B] hi, then lo (64-byte swizzle)
    const uint32_t stg_w = aux_w + 4096u;                            // backward: 2 x ([32 rows][32 B] hi + lo), 32-byte swizzle
    const uint32_t auxb = aux_bar(wi);
    float* const wv_s = wv_sm[group];
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)team * ACC_COLS;
    int staged_n0 = -1;
    uint32_t aux_n = 0, nstore = 0;
    auto tile_n0 = [&](int t) { return (t % p.tiles_n) * BN; };
    auto tile_m0 = [&](int t) { return ((t / p.tiles_n) % p.tiles_m) * V_BM; };
    auto nch = [&](int n0) { return min(BN / 32, (p.d - n0 + 31) / 32); };
    auto live = [&](int t) { return tile_m0(t) + q * 32 < p.N; };     // a quadrant entirely beyond the last region has no work
    // The warp's chunk of a tile is chunk `slot` (GPT = BN / 32 groups per team: one chunk each).  `at`: next tile of the team in which
    // this warp has work and whose PV slice has not been requested yet; one load in flight per warp.
    int at = tile_first - tile_step;
    auto next_assigned = [&]() {
      at += tile_step;
      while (at < p.total_tiles && !(live(at) && slot < nch(tile_n0(at)))) at += tile_step;
    };
    auto issue_aux = [&]() {
      const int n0 = tile_n0(at), m0 = tile_m0(at), b = at / (p.tiles_n * p.tiles_m);
      mbar_expect_tx(auxb, 4096u);
      tma_load_4d(aux_w, &maps.PV, auxb, n0 + slot * 32, m0 + q * 32, 0, b);
      tma_load_4d(aux_w + 2048u, &maps.PV, auxb, n0 + slot * 32, m0 + q * 32, 1, b);
    };
    auto flush_colred = [&](int n0) {
      if constexpr (BWD) {
        for (int j = st; j < BN; j += 128) {
          if (n0 + j < p.d) {
            const float a = colred_sm[group][0][j], c2 = colred_sm[group][1][j];
            if (a != 0.f) atomicAdd(p.dwv + n0 + j, a);
            if (c2 != 0.f) atomicAdd(p.dbv + n0 + j, c2);
          }
        }
      }
    };
    // 16 fp32 of this thread's row -> bf16 hi / lo halves in the warp's staging buffer -> one TMA store per plane (rows beyond N clipped)
    auto store_half = [&](const float (&x)[16], const CUtensorMap* map, int col, int row0, int z) {
      const uint32_t sbuf = stg_w + (NSTG > 1 ? (nstore & 1u) : 0u) * 2048u;
      if (lane == 0) tma_store_wait_read<(BWD ? (int)NSTG - 1 : 0)>();       // the store that last used this buffer has been read
      __syncwarp();
      const uint32_t sb = sbuf + (uint32_t)lane * 32u;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint32_t h[4], lo[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float x0 = x[8 * j + 2 * k], x1 = x[8 * j + 2 * k + 1];
          const __nv_bfloat162 hh = __floats2bfloat162_rn(x0, x1);
          const __nv_bfloat162 ll = __floats2bfloat162_rn(x0 - __low2float(hh), x1 - __high2float(hh));
          h[k] = *reinterpret_cast<const uint32_t*>(&hh);
          lo[k] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        const uint32_t o = (uint32_t)((j ^ ((lane >> 2) & 1)) * 16);
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sb + o), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sb + 1024u + o), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_4d(map, sbuf, col, row0, 0, z);
        tma_store_4d(map, sbuf + 1024u, col, row0, 1, z);
        tma_store_commit();
      }
      ++nstore;
    };
    next_assigned();
    if (lane == 0 && at < p.total_tiles) issue_aux();
    // this thread's 16 PV values of column half hf (hi + lo planes, 64-byte swizzle)
    auto load_ax = [&](int hf, float (&ax)[16]) {
      const uint32_t src = aux_w + (uint32_t)lane * 64u;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint32_t h[4], l4[4];
        const uint32_t o = (uint32_t)(((2 * hf + j) ^ ((lane >> 1) & 3)) * 16);
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(h[0]), "=r"(h[1]), "=r"(h[2]), "=r"(h[3]) : "r"(src + o));
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(l4[0]), "=r"(l4[1]), "=r"(l4[2]), "=r"(l4[3]) : "r"(src + 2048u + o));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          ax[8 * j + 2 * k] = bf_lo(h[k]) + bf_lo(l4[k]);
          ax[8 * j + 2 * k + 1] = bf_hi(h[k]) + bf_hi(l4[k]);
        }
      }
    };
    auto request_next_slice = [&]() {           // the slice has been read for the last time: request the warp's next one
      __syncwarp();
      ++aux_n;
      next_assigned();
      if (lane == 0 && at < p.total_tiles) issue_aux();
    };
    auto release_acc = [&](int l) {             // this warp's reads of accumulator l (backward: of the whole set) are complete
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(team, l));
    };
    int tile_it = 0;
    for (int t = tile_first; t < p.total_tiles; t += tile_step, ++tile_it) {
      const int n0 = tile_n0(t), m0 = tile_m0(t), b = t / (p.tiles_n * p.tiles_m);
      const int row = m0 + r;
      const bool row_ok = row < p.N;
      const int c = slot;                                            // the warp's chunk of the tile
      const bool has_work = live(t) && c < nch(n0);                  // (a quadrant beyond the last region / a chunk beyond d: only the hand-over)
      const uint32_t par = (uint32_t)(tile_it & 1);
      const bool stamp = tl && slot == 0 && q == 0 && lane == 0 && tile_it * NBUF + team < 12;
      long long* const tls = tl + (tile_it * NBUF + team) * 16;
      if (n0 != staged_n0) {                       // per-column vectors: restaged only when the column block changes
        slot_barrier();
        if (staged_n0 >= 0) flush_colred(staged_n0);
        for (int j = st; j < BN; j += 128) {
          wv_s[j] = (n0 + j < p.d) ? __ldg(p.wv + n0 + j) : 0.f;
          colred_sm[group][0][j] = 0.f;
          colred_sm[group][1][j] = 0.f;
        }
        staged_n0 = n0;
        slot_barrier();
      }
      const int col0 = n0 + c * 32;
      const float4* const w4 = reinterpret_cast<const float4*>(wv_s + c * 32);     // (read at the point of use: fewer live registers)
      if constexpr (!BWD) {
        if (!has_work) {
          // (warps without work still take part in the hand-over, level by level: an arrival must never run ahead of the other warps')
          for (int l = 0; l < 3; ++l) {
            mbar_wait(acc_full(team, l), par, 21);
            release_acc(l);
          }
          continue;
        }
        float rd0 = 0.f, rd1 = 0.f, rd2 = 0.f;
        // six (level, half) steps of 16 columns, levels outermost, TMEM loads one step ahead of the math (back to back in every warp,
        // the warps of a tile moved in lockstep between the TMEM read port and the MUFU pipe, and the two added up)
        uint32_t vv[2][16];
        mbar_wait(acc_full(team, 0), par, 21);
        tc_fence_after();
        if (stamp) tls[6] = clock64();
        tmem_ld16_nowait(lane_addr + (uint32_t)(c * 32), vv[0]);
        mbar_wait(auxb, aux_n & 1u, 22);             // the warp's slice of the PV tile (hi + lo planes) has landed
        if (stamp) tls[7] = clock64();
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const int l = i >> 1, hf = i & 1;
          float ax[16];
          load_ax(hf, ax);                             // (re-read per step: 4 LDS.128 against 16 registers held across the levels)
          if (i == 5) request_next_slice();
          tmem_ld_wait16(vv[i & 1]);
          if (hf == 1) release_acc(l);                 // level l has been read: its accumulator may take the set's next tile
          if (i + 1 < 6) {
            if (hf == 1) {
              mbar_wait(acc_full(team, l + 1), par, 21);
              tc_fence_after();
            }
            tmem_ld16_nowait(lane_addr + (uint32_t)(((i + 1) >> 1) * BN + c * 32 + ((i + 1) & 1) * 16), vv[(i + 1) & 1]);
          }
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) {          // wv is 0 beyond d
            const float4 w = w4[4 * hf + j];
            float t0, t1, t2, t3;
            tanh_fast2(__uint_as_float(vv[i & 1][4 * j]) + ax[4 * j], __uint_as_float(vv[i & 1][4 * j + 1]) + ax[4 * j + 1], t0, t1);
            tanh_fast2(__uint_as_float(vv[i & 1][4 * j + 2]) + ax[4 * j + 2], __uint_as_float(vv[i & 1][4 * j + 3]) + ax[4 * j + 3], t2, t3);
            acc = fmaf(t0, w.x, acc);
            acc = fmaf(t1, w.y, acc);
            acc = fmaf(t2, w.z, acc);
            acc = fmaf(t3, w.w, acc);
          }
          if (l == 0) rd0 += acc;
          else if (l == 1) rd1 += acc;
          else rd2 += acc;
        }
        if (stamp) tls[8] = clock64();
        if (row_ok) {
          atomicAdd(p.sv + ((int64_t)b * 3 + 0) * p.N + row, rd0);
          atomicAdd(p.sv + ((int64_t)b * 3 + 1) * p.N + row, rd1);
          atomicAdd(p.sv + ((int64_t)b * 3 + 2) * p.N + row, rd2);
        }
      } else {
        float rv0 = 0.f, rv1 = 0.f, rv2 = 0.f;       // (scalars: indexed arrays would live in local memory)
        if (row_ok) {
          rv0 = __ldg(p.rowv + ((int64_t)b * 3 + 0) * p.N + row);
          rv1 = __ldg(p.rowv + ((int64_t)b * 3 + 1) * p.N + row);
          rv2 = __ldg(p.rowv + ((int64_t)b * 3 + 2) * p.N + row);
        }
        // (warps without work wait too: an arrival must never run a tile ahead of the other warps' arrivals)
        mbar_wait(acc_full(team, 0), par, 21);
        tc_fence_after();
        if (stamp) tls[6] = clock64();
        if (has_work) {
          mbar_wait(auxb, aux_n & 1u, 22);           // the warp's slice of the PV tile (hi + lo planes) has landed
          if (stamp) tls[7] = clock64();
          // 16 columns at a time, the three levels inside: running dPV sum (16) + one level's values (16 + 16) + the next level's
          // accumulator values in flight (16) stay in registers; the PV values are re-read from shared memory per level
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t v[16];
            float sum[16];
            tmem_ld16_nowait(lane_addr + (uint32_t)(3 * BN + c * 32 + hf * 16), v);
            tmem_ld_wait16(v);
#pragma unroll
            for (int j = 0; j < 16; ++j) sum[j] = __uint_as_float(v[j]);
            tmem_ld16_nowait(lane_addr + (uint32_t)(c * 32 + hf * 16), v);
            float pacc[16];                              // sum_l Hv_l * dsv_l of this row: reduced over the rows once per half
#pragma unroll
            for (int j = 0; j < 16; ++j) pacc[j] = 0.f;
#pragma unroll 1
            for (int l = 0; l < 3; ++l) {
              float f[16];
              {
                float ax[16];
                load_ax(hf, ax);
                if (hf == 1 && l == 2) request_next_slice();
                tmem_ld_wait16(v);
#pragma unroll
                for (int j = 0; j < 16; j += 2) tanh_fast2(__uint_as_float(v[j]) + ax[j], __uint_as_float(v[j + 1]) + ax[j + 1], f[j], f[j + 1]);
              }
              if (l < 2) tmem_ld16_nowait(lane_addr + (uint32_t)((l + 1) * BN + c * 32 + hf * 16), v);     // lands during the rest of this level
              const float rvl = l == 0 ? rv0 : (l == 1 ? rv1 : rv2);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 w = w4[4 * hf + j];
                const float wj[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float h = f[4 * j + k];
                  pacc[4 * j + k] = fmaf(h, rvl, pacc[4 * j + k]);
                  const float g = rvl * wj[k] * (1.f - h * h);
                  f[4 * j + k] = g;
                  sum[4 * j + k] += g;
                }
              }
              store_half(f, &maps.DZv, col0 + hf * 16, m0 + q * 32, 3 * b + l);       // dZv_l: operand planes of dPQ and dS
            }
            {
              const float cs = col_reduce16(pacc, lane);          // lanes 2 j, 2 j + 1 hold column j
              if ((lane & 1) == 0) atomicAdd(&colred_sm[group][0][c * 32 + hf * 16 + (lane >> 1)], cs);
            }
            if (hf == 1) release_acc(0);               // every accumulator of the set has been read: the MMAs of its next tile may start
            // dPV = sum_l dZv_l + C_all^T dZq_all: planes out, column sums -> dbv
            float part[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) part[j] = row_ok ? sum[j] : 0.f;
            const float cs = col_reduce16(part, lane);
            if ((lane & 1) == 0) atomicAdd(&colred_sm[group][1][c * 32 + hf * 16 + (lane >> 1)], cs);
            store_half(sum, &maps.DPV, col0 + hf * 16, m0 + q * 32, b);
          }
          if (stamp) tls[8] = clock64();
        } else {
          release_acc(0);
        }
      }
    }
    if (BWD && staged_n0 >= 0) {
      slot_barrier();
      flush_colred(staged_n0);
    }
    if (BWD && lane == 0) tma_store_wait_read<0>();
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

template <bool BWD, int BN>
int launch(const HvPlanes& C, const HvPlanes& PQ, const HvPlanes& PV, const HvPlanes* dZq, const HvPlanes* dZv, const HvPlanes* dPV,
           HvParams p, cudaStream_t s) {
  constexpr int NG = HvShape<BN>::NG, STAGES = BWD ? HvShape<BN>::STAGES_BWD : HvShape<BN>::STAGES_FWD, CTAS = 1;
  HvMaps maps;
  const int T = p.T, N = p.N, d = p.d, B = p.B;
  HCA_TRY(make_map(&maps.Cl, C, N, T, (int64_t)T * C.ld, 3 * B, 64, V_BK, 3));
  HCA_TRY(make_map(&maps.PQl, PQ, d, T, (int64_t)T * PQ.ld, 3 * B, 64, V_BK, 3));
  HCA_TRY(make_map(&maps.PV, PV, d, N, (int64_t)N * PV.ld, B, 32, 32, 2));          // one epilogue warp's slice: 32 rows x 32 columns
  if (BWD) {
    HCA_TRY(make_map(&maps.DZq, *dZq, d, T, (int64_t)T * dZq->ld, 3 * B, 64, V_BK, 3));
    HCA_TRY(make_map(&maps.DZv, *dZv, d, N, (int64_t)N * dZv->ld, 3 * B, 16, 32, 1));    // staging half-chunk: 32 rows x 16 columns
    HCA_TRY(make_map(&maps.DPV, *dPV, d, N, (int64_t)N * dPV->ld, B, 16, 32, 1));
  } else {
    maps.DZq = maps.PQl; maps.DZv = maps.PV; maps.DPV = maps.PV;
  }
  p.tiles_m = (N + V_BM - 1) / V_BM;
  p.tiles_n = (d + BN - 1) / BN;
  const int64_t total = (int64_t)B * p.tiles_m * p.tiles_n;
  HCA_CHECK_ARG(total < (1LL << 30), "hv: too many tiles");
  p.total_tiles = (int)total;
  p.kbl = (T + V_BK - 1) / V_BK;
  p.timeline = HCA_TC_TIMELINE ? tc_timeline_buffer() : nullptr;
  // (mirrors the kernel: operand ring + per epilogue warp a PV slice and, backward, one staging half-chunk)
  const size_t stage_bytes = 2 * (size_t)V_TILE_A + (BWD ? 4 : 2) * (size_t)(BN / 64) * V_CHUNK;
  const size_t smem = (size_t)STAGES * stage_bytes + (size_t)4 * NG * (4096 + (BWD ? 2048 : 0)) + 1024;
  static bool attr_set[64] = {};                     // function attributes are per device (and per instantiation: this static is)
  bool& attr_done = attr_set[current_device()];
  if (!attr_done) {
    HCA_CUDA(cudaFuncSetAttribute(hv_kernel<BWD, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
    cudaGetLastError();
    sms = 148;
  }
  const int ctas = (int)std::min<int64_t>((int64_t)sms * CTAS, total);
  HCA_LAUNCH_K((hv_kernel<BWD, BN>), ctas, 64 + 128 * NG, smem, s, maps, p);
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
  return launch<false, V_BN>(C, PQ, PV, nullptr, nullptr, nullptr, p, s);
}

int launch_hv_grads(const HvPlanes& C, const HvPlanes& PQ, const HvPlanes& PV, const HvPlanes& dZq, const float* wv, const float* dsv,
                    const HvPlanes& dZv, const HvPlanes& dPV, float* dwv, float* dbv, int B, int N, int T, int d, cudaStream_t s) {
  HCA_CHECK_ARG(planes_ok(C) && planes_ok(PQ) && planes_ok(PV) && planes_ok(dZq) && planes_ok(dZv) && planes_ok(dPV) && wv && dsv && dwv && dbv,
                "hv_grads: bad arguments");
  HvParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.N = N; p.T = T; p.d = d; p.wv = wv; p.rowv = dsv; p.dwv = dwv; p.dbv = dbv;
  return launch<true, V_BN>(C, PQ, PV, &dZq, &dZv, &dPV, p, s);
}

}  // namespace hca
