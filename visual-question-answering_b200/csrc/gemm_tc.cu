// tcgen05 / TMEM / TMA split-bf16 GEMM, persistent with a double-buffered TMEM accumulator (see gemm_tc.cuh).  sm_100a only.
#include <cuda.h>
#include <cuda_bf16.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include "gemm_tc.cuh"
#include "tc_ptx.cuh"

namespace hca {
namespace {
using namespace ptx;

#ifndef HCA_TC_TIMELINE
#define HCA_TC_TIMELINE 0
#endif
constexpr int BM = 128;               // UMMA M (cta_group::1): TMEM lane i <-> output row i
constexpr int UMMA_K = 16;            // bf16
constexpr int MAX_STAGES = 6;
constexpr int NUM_EG = 2;             // epilogue warp groups (4 warps each: one per TMEM lane quadrant)
constexpr int NUM_THREADS = 64 + 128 * NUM_EG;   // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2..5 / 6..9 epilogue groups 0 / 1
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr uint32_t CHUNK_BYTES = BM * 128;    // one [128 rows][32 fp32] staging tile = two [128 rows][32 bf16] plane tiles
constexpr int NWIN_TM_MAX = 32;               // M tiles of a product with N windows (block-structured outputs)

struct TcParams {
  int M, N, K, K2;
  int stages, kb1, kb_total, kb_per_split, splitk;
  int tiles_m, tiles_n, total_tiles;
  int a_nb, a_zd, b_nb, b_zd, a2_nb, a2_zd, b2_nb, b2_zd;   // operand batch entry = (z / zd) % nb
  uint32_t off_store, off_pstore, off_aux;                   // byte offsets of the epilogue staging areas from the smem base
  int store_nbuf;                                            // staging buffers per output kind and epilogue group (2 = double-buffered)
  int n_eg;                                                  // active epilogue groups: 2 for epilogue-bound products (alternate 32-column chunks)
  int kwin_ncol, kwin_lo[4], kwin_hi[4];                     // K windows per group of output columns, in k-blocks (kwin_ncol = 0: off)
  int nwin_on, tiles_mn;                                     // N windows per group of output rows: only tiles_mn (M tile, N tile) pairs exist;
  short nw_first[NWIN_TM_MAX], nw_pre[NWIN_TM_MAX + 1];      // M tile mt owns the N tiles nw_first[mt] + [0, nw_pre[mt + 1] - nw_pre[mt])
  // fp32 output
  float* D;
  int64_t ldd, d_sb;
  int d_zd, d_rpg;          // rows per group (= M when there is one group)
  int64_t d_gs;
  int accumulate, atomic;
  int tma_store;            // fp32 output goes through smem + TMA tile store / reduce-add
  int pl_direct;            // bf16 planes are stored straight from registers (64 B per row and plane), not staged for a TMA store
  // bf16 hi/lo planes output
  __nv_bfloat16* P;
  int64_t p_ld, p_ps, p_sb;
  const float* bias; int64_t bias_sb;
  int act_tanh;
  const float* mulx; int64_t mulx_ld;
  int mode, aux_kind /* 0 none, 1 fp32 tile, 2 bf16 hi/lo planes */, aux_mode, aux_nb, aux_zd;
  const __nv_bfloat16* auxp; int64_t auxp_ld, auxp_ps, auxp_sb;   // transposed epilogue: read directly
  const float* rowv; int64_t rowv_sb;
  const float* colv;
  const float* r1col; int64_t r1col_sb; int r1_rpg; int64_t r1_gs;
  float* red_row; int64_t red_row_sb;
  float* red_col;
  int dbg;                  // debug bit flags (HCA_TC_DBG): 1 = no TMA (MMA on garbage)
  long long* timeline;      // optional [ncta][64] clock64 stamps (debug / profiling), nullptr normally
  int timeline_ctas;
};

long long* g_timeline = nullptr;
int g_timeline_ctas = 0;
int g_timeline_target = -1;      // record only the launch with this index (counted from the last tc_set_timeline); -1 = every launch
int g_timeline_seen = 0;

struct TcMaps {            // all TMA descriptors of one launch
  CUtensorMap A, B, A2, B2;   // operands, 4-D bf16: (cols, rows, plane, batch), 128-byte swizzle
  CUtensorMap D;              // fp32 output, 4-D: (N, rows in group, groups, batch), box 32 columns, 128-byte swizzle
  CUtensorMap DP;             // bf16 planes output, 4-D: (N, M, plane, batch), box 32 columns, 64-byte swizzle
  CUtensorMap AUX;            // addend tile: fp32 3-D (N, M, batch) 128-byte swizzle, or bf16 planes 4-D like DP
};

struct TileCoord {
  int z, m0, n0, kb_begin, num_kb;
};

__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// ------------------------------------------------------------------------------------------------------ kernel
// Column sums over the 32 lanes of a warp of 32 per-lane values: on return lane j holds sum_lanes v[j].  Butterfly
// "transpose-reduce": 31 shuffles instead of the 160 of 32 independent warp reductions.
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

// CG = 2: CTA-pair mode (cluster of 2, tcgen05 cta_group::2).  The pair computes a 256 x BN tile: each CTA loads its own 128 rows
// of A and HALF of the B tile, the leader issues M = 256 MMAs that read both halves, and each CTA drains its own 128 accumulator
// lanes.  Per output element only half the operand bytes cross the L2 -> SM fabric -- the measured limit of the big K-major
// products (profiles/r1c_pv_gemm_ncu_full.md: 83 B/clk/SM of operand fill against a ~43 B/clk/SM chip-wide L2 cap).
// EPI = 0: "lean" epilogue (bias, fp32 / planes output, split-K reduce-add; EPI = 2 adds tanh): the feature blocks of the full epilogue (addend tiles,
// row dots, dZ, rank-1 terms, column sums) are compiled out.  The profile of the full variant on a plain product showed the chunk loop
// executing ~340 of ~3200 instructions, with 22 % of its stall samples waiting for instruction fetch and 7 % resolving branches.
template <int BN, int P, bool A_MN, bool B_MN, int BK, int CG = 1, int EPI = 1>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ TcMaps maps, const TcParams p) {
  pdl_trigger();      // (the wait follows the set-up below: barrier init, descriptor prefetch and TMEM allocation touch no dependent data)
  static_assert(BK == 64 || (BK == 32 && A_MN && B_MN), "BK = 32 k-blocks are for MN-major operand pairs (short contractions)");
  static_assert(CG == 1 || (CG == 2 && BN == 256 && A_MN == B_MN && BK == 64), "CTA-pair mode: (K,K) or (MN,MN) operands, 256-wide tiles");
  constexpr int A_TILE_BYTES = BM * BK * 2;
  constexpr int B_ROWS = BN / CG;                         // rows of the B tile this CTA loads
  constexpr int B_TILE_BYTES = B_ROWS * BK * 2;
  constexpr uint32_t MN_CHUNK_BYTES = 64 * BK * 2;       // one TMA box of an MN-major tile: [BK k][64 mn]
  constexpr uint32_t stage_bytes = P * (A_TILE_BYTES + B_TILE_BYTES);
  constexpr int CHUNKS = BN / 32;
  constexpr uint32_t TMEM_COLS = 2 * BN;                 // two accumulator buffers (64 or 256 columns: powers of two >= 32)
  constexpr int BNV = BN < 128 ? 128 : BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B tiles need 1024-byte alignment
  __shared__ __align__(8) uint64_t bars[2 * MAX_STAGES + 4 + NUM_EG];
  __shared__ uint32_t tmem_ptr_smem;
  // per epilogue group: this tile's per-column vectors
  __shared__ __align__(16) float bias_sm[NUM_EG][BNV];      // bias slice (per batch entry)
  __shared__ __align__(16) float colv_sm[NUM_EG][BNV];      // colv (ROWDOT)
  __shared__ __align__(16) float r1_sm[NUM_EG][CG == 2 ? 1 : 4][BNV];   // rank-1 column vectors, one per row group (not in pair mode)
  __shared__ float colred_sm[NUM_EG][BNV];                  // column partial sums of the group's four warps

  // warp index through a shuffle: provably warp-uniform for ptxas, so the role branches below are uniform control flow and
  // the MMA descriptors live in uniform registers (otherwise every tcgen05.mma pays an ELECT + VOTEU + 4x R2UR.BROADCAST
  // sequence, ~250 cycles per instruction, measured)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  // pair mode: rank inside the pair, and the persistent schedule walks pair tiles with the pair index / pair count
  const int rank = CG == 2 ? (int)cluster_ctarank() : 0;
  const int sched0 = CG == 2 ? (int)cluster_id_x() : (int)blockIdx.x;
  const int sched_step = CG == 2 ? (int)cluster_nclusters_x() : (int)gridDim.x;

  auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
  auto empty_bar = [&](int s) { return smem_u32(&bars[MAX_STAGES + s]); };
  auto tmem_full_bar = [&](int a) { return smem_u32(&bars[2 * MAX_STAGES + a]); };
  auto tmem_empty_bar = [&](int a) { return smem_u32(&bars[2 * MAX_STAGES + 2 + a]); };
  auto aux_bar = [&](int g) { return smem_u32(&bars[2 * MAX_STAGES + 4 + g]); };
  auto a_tile = [&](int s, int pl) { return smem_base + s * stage_bytes + pl * A_TILE_BYTES; };
  auto b_tile = [&](int s, int pl) { return smem_base + s * stage_bytes + P * A_TILE_BYTES + pl * B_TILE_BYTES; };
  auto decode = [&](int t) {
    TileCoord c;
    int nt, mt;
    if (p.nwin_on) {                  // block-structured output: walk the (M tile, N tile) pairs that exist
      const int u = t % p.tiles_mn;
      t /= p.tiles_mn;
      mt = 0;
      while (mt + 1 < p.tiles_m && p.nw_pre[mt + 1] <= u) ++mt;
      nt = p.nw_first[mt] + (u - p.nw_pre[mt]);
    } else {
      nt = t % p.tiles_n;
      t /= p.tiles_n;
      mt = t % p.tiles_m;
      t /= p.tiles_m;
    }
    const int ks = t % p.splitk;
    c.z = t / p.splitk;
    c.m0 = mt * (BM * CG) + rank * BM;
    c.n0 = nt * BN;
    c.kb_begin = ks * p.kb_per_split;
    c.num_kb = min(p.kb_total, c.kb_begin + p.kb_per_split) - c.kb_begin;
    if (p.kwin_ncol > 0) {            // block-structured contraction: the union of the K windows of the column groups this tile touches
      const int g0 = c.n0 / p.kwin_ncol, g1 = (min(p.N, c.n0 + BN) - 1) / p.kwin_ncol;
      int lo = p.kwin_lo[g0], hi = p.kwin_hi[g0];
      for (int g = g0 + 1; g <= g1; ++g) {
        lo = min(lo, p.kwin_lo[g]);
        hi = max(hi, p.kwin_hi[g]);
      }
      c.kb_begin = lo;
      c.num_kb = hi - lo;
    }
    return c;
  };

  // (clock64 stamps for profiles/timeline_*.py: compiled in only with -DHCA_TC_TIMELINE=1, see build.py)
  long long* tl = (HCA_TC_TIMELINE && p.timeline && (int)blockIdx.x < p.timeline_ctas) ? p.timeline + (size_t)blockIdx.x * 64 : nullptr;
  if (tl && threadIdx.x == 0) {
    tl[0] = clock64();
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    tl[7] = smid;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      // one arrival per epilogue warp that drains the buffer (every group takes part in every tile)
      // (pair mode: the leader's barrier collects the epilogue warps of both CTAs)
      mbar_init(tmem_empty_bar(a), 4 * p.n_eg * CG);
    }
    for (int g = 0; g < NUM_EG; ++g) mbar_init(aux_bar(g), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.A) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.B) : "memory");
    if (p.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.D) : "memory");
    if (p.P && BN != 32) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.DP) : "memory");
    if (p.aux_kind && BN != 32) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.AUX) : "memory");
  }
  if (warp == 1) {
    if constexpr (CG == 2) tmem_alloc_2sm(smem_u32(&tmem_ptr_smem), TMEM_COLS);
    else tmem_alloc(smem_u32(&tmem_ptr_smem), TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();     // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  pdl_wait();                                    // every producer kernel has completed: operands, addends and outputs may be touched
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_ptr_smem, 0);
  if (tl && threadIdx.x == 0) tl[1] = clock64();

  if (warp == 0) {
    // ===================================================================== TMA producer (one elected lane, uniform datapath)
    if (elect_one_sync() && !(p.dbg & 1)) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = sched0; t < p.total_tiles; t += sched_step) {
        const TileCoord tc = decode(t);
        for (int it = 0; it < tc.num_kb; ++it) {
          mbar_spin(empty_bar(s), ph ^ 1);
          if (tl && (p.dbg & 4) && t == sched0 && it < 16) tl[8 + it] = clock64();        // (HCA_TC_DBG=4: per-k-block stamps of the first tile)
          if constexpr (CG == 2) {
            // both CTAs' loads report to the leader's barrier (the MMA issuer waits there): it expects the bytes of both
            if (rank == 0) mbar_expect_tx(full_bar(s), 2 * stage_bytes);
            const int k0 = (tc.kb_begin + it) * BK;
#pragma unroll
            for (int pl = 0; pl < P; ++pl) {
              if constexpr (!A_MN) {
                tma_load_4d_2sm(a_tile(s, pl), &maps.A, full_bar(s), k0, tc.m0, pl, 0);
                tma_load_4d_2sm(b_tile(s, pl), &maps.B, full_bar(s), k0, tc.n0 + rank * B_ROWS, pl, 0);
              } else {                                        // MN-major pair (weight gradients): [BK k][64 mn] boxes
#pragma unroll
                for (int c = 0; c < BM / 64; ++c)
                  tma_load_4d_2sm(a_tile(s, pl) + c * MN_CHUNK_BYTES, &maps.A, full_bar(s), tc.m0 + c * 64, k0, pl, 0);
#pragma unroll
                for (int c = 0; c < B_ROWS / 64; ++c)
                  tma_load_4d_2sm(b_tile(s, pl) + c * MN_CHUNK_BYTES, &maps.B, full_bar(s), tc.n0 + rank * B_ROWS + c * 64, k0, pl, 0);
              }
            }
            if (++s == p.stages) { s = 0; ph ^= 1; }
            continue;
          }
          mbar_expect_tx(full_bar(s), stage_bytes);
          const int kb = tc.kb_begin + it;
          const bool second = kb >= p.kb1;                    // chained second operand pair
          const CUtensorMap* ma = second ? &maps.A2 : &maps.A;
          const CUtensorMap* mb = second ? &maps.B2 : &maps.B;
          const int za = second ? (tc.z / p.a2_zd) % p.a2_nb : (tc.z / p.a_zd) % p.a_nb;
          const int zb = second ? (tc.z / p.b2_zd) % p.b2_nb : (tc.z / p.b_zd) % p.b_nb;
          const int k0 = (second ? kb - p.kb1 : kb) * BK;
#pragma unroll
          for (int pl = 0; pl < P; ++pl) {
            if (!A_MN) {
              tma_load_4d(a_tile(s, pl), ma, full_bar(s), k0, tc.m0, pl, za);
            } else {
#pragma unroll
              for (int c = 0; c < BM / 64; ++c) tma_load_4d(a_tile(s, pl) + c * MN_CHUNK_BYTES, ma, full_bar(s), tc.m0 + c * 64, k0, pl, za);
            }
            if (!B_MN) {
              tma_load_4d(b_tile(s, pl), mb, full_bar(s), k0, tc.n0, pl, zb);
            } else if (BN >= 64) {
#pragma unroll
              for (int c = 0; c < BN / 64; ++c) tma_load_4d(b_tile(s, pl) + c * MN_CHUNK_BYTES, mb, full_bar(s), tc.n0 + c * 64, k0, pl, zb);
            }
          }
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
      if constexpr (CG == 2) {
        // drain: every MMA-completion signal multicast to this CTA's barriers has landed before the CTA may exit
        for (int i = 0; i < p.stages; ++i) {
          mbar_spin(empty_bar(s), ph ^ 1);
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (one elected lane, uniform datapath)
    // instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 [4,6), a=bf16 [7,10), b=bf16 [10,13),
    // a_major bit 15, b_major bit 16, N>>3 [17,23), M>>4 [24,29)
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                               ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * CG) >> 4) << 24);
    // smem descriptor halves (cute::UMMA::SmemDescriptor): hi = SBO(1024 B) | version 1 | SWIZZLE_128B, constant;
    // lo = (addr >> 4) | (LBO >> 4) << 16.  K-major: LBO unused (16 B), K slice = +32 B.  MN-major: LBO = one [BK k][64 mn]
    // box between 64-wide MN chunks, K slice of 16 rows = +2048 B.
    constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    constexpr uint32_t a_lbo = (A_MN ? (MN_CHUNK_BYTES >> 4) : 1u) << 16, b_lbo = (B_MN ? (MN_CHUNK_BYTES >> 4) : 1u) << 16;
    constexpr uint32_t a_kstep = A_MN ? (2048u >> 4) : (32u >> 4), b_kstep = B_MN ? (2048u >> 4) : (32u >> 4);
    // The whole loop runs in ONE elected thread with inline waits: inside such a region every value is trivially warp-uniform,
    // so ptxas builds the descriptors with uniform-datapath adds and issues the UTCHMMAs back to back.  (Per-instruction
    // elect / predicate forms cost ELECT + VOTEU + 4-5 R2UR per MMA, ~75 cycles each: that, not L2 bandwidth, bound the mainloop.)
    if (rank == 0 && elect_one_sync()) {
      const uint32_t tm = *reinterpret_cast<volatile uint32_t*>(&tmem_ptr_smem);
      const uint32_t a_base = ((smem_base & 0x3FFFFu) >> 4) | a_lbo;
      const uint32_t b_base = (((smem_base + P * A_TILE_BYTES) & 0x3FFFFu) >> 4) | b_lbo;
      int s = 0;
      uint32_t ph = 0;
      int tile_it = 0;
      for (int t = sched0; t < p.total_tiles; t += sched_step, ++tile_it) {
        const TileCoord tc = decode(t);
        const int acc = tile_it & 1;
        mbar_spin(tmem_empty_bar(acc), ((tile_it >> 1) & 1) ^ 1);      // the epilogue has drained this accumulator buffer
        tc_fence_after();
        const uint32_t d_tmem = tm + (uint32_t)(acc * BN);
        for (int it = 0; it < tc.num_kb; ++it) {
          if (!(p.dbg & 1)) mbar_spin(full_bar(s), ph);
          tc_fence_after();
          if (tl && tile_it == 0 && it == 0) tl[2] = clock64();
          if (tl && (p.dbg & 4) && tile_it == 0 && it < 16) tl[24 + it] = clock64();
          const uint32_t au = a_base + (uint32_t)s * (stage_bytes >> 4), bu = b_base + (uint32_t)s * (stage_bytes >> 4);
          // K slices of this block that hold data (the tail of K is zero-filled by TMA: skip those MMAs)
          const int kb = tc.kb_begin + it;
          const int kleft = kb >= p.kb1 ? p.K2 - (kb - p.kb1) * BK : p.K - kb * BK;
          const int nks = min(BK / UMMA_K, (kleft + UMMA_K - 1) / UMMA_K);
#pragma unroll
          for (int ks = 0; ks < BK / UMMA_K; ++ks) {
            if (ks < nks) {
#pragma unroll
              for (int i = 0; i < P; ++i) {
#pragma unroll
                for (int j = 0; j < P - i; ++j) {
                  if constexpr (CG == 2)
                    umma_bf16_one_2sm<desc_hi, idesc>(d_tmem, au + i * (A_TILE_BYTES >> 4) + ks * a_kstep,
                                                      bu + j * (B_TILE_BYTES >> 4) + ks * b_kstep, (ks | i | j) != 0 ? 1u : (it == 0 ? 0u : 1u));
                  else
                  umma_bf16_one<desc_hi, idesc>(d_tmem, au + i * (A_TILE_BYTES >> 4) + ks * a_kstep, bu + j * (B_TILE_BYTES >> 4) + ks * b_kstep,
                                                (ks | i | j) != 0 ? 1u : (it == 0 ? 0u : 1u));
                }
              }
            }
          }
          if (tl && (p.dbg & 4) && tile_it == 0 && it < 16) tl[40 + it] = clock64();
          if constexpr (CG == 2) umma_commit_2sm(empty_bar(s));   // frees the stage in both CTAs
          else umma_commit(empty_bar(s));       // frees the smem stage once the MMAs above have read it
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        if constexpr (CG == 2) umma_commit_2sm(tmem_full_bar(acc));   // accumulator halves complete -> both epilogues
        else umma_commit(tmem_full_bar(acc));   // accumulator complete -> epilogue
        if (tl && tile_it == 0) tl[3] = clock64();
      }
    }
  } else if (((warp - 2) >> 2) < p.n_eg) {
    // ===================================================================== epilogue: TMEM -> registers -> (smem -> TMA) global
    // Two groups of four warps.  Wide tiles: group g takes the 32-column chunks c = g, g + n_eg, ... of EVERY tile (its own
    // staging / addend buffers, named barrier and TMA-store queue), so the latency of one group's TMEM loads, MUFU chains,
    // barriers and stores is covered by the other's arithmetic.  BN = 32 tiles: the groups alternate tiles.
    const int eg = (warp - 2) >> 2;
    const int neg = p.n_eg;
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int r = q * 32 + lane;            // row inside the tile
    const int et = threadIdx.x - 64 - eg * 128;   // 0..127 inside the group
    const bool leader = (et == 0);
    const uint32_t bar_id = 1u + (uint32_t)eg;
    auto epi_barrier = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory"); };
    int tile_it = 0;
    if constexpr (BN == 32) {
      // ------------------------------------------------------------------ transposed epilogue (direct, coalesced global I/O)
      // The groups split the COLUMNS of every tile (CW = 32 / groups each).  A tile's epilogue is one warp per scheduler walking its
      // columns with nothing to hide a latency behind: per-launch timeline of the classifier products (profiles/timeline_mlp.py):
      // 7.6 k cycles for the epilogue of ONE 128 x 32 tile with a single group, against 6.5 k for the whole K = 512 mainloop.  (Splitting
      // that mainloop over a cluster with a DSMEM reduction was built and measured too: no change of the step, removed.)
      auto run = [&](auto cw_tag) {
        constexpr int CW = decltype(cw_tag)::value;
        const int c0 = CW == 32 ? 0 : eg * CW;
        for (int t = sched0; t < p.total_tiles; t += sched_step, ++tile_it) {
          const TileCoord tc = decode(t);
          const int acc = tile_it & 1;
          const int row = tc.m0 + r;
          const bool row_ok = row < p.M;
          mbar_wait(tmem_full_bar(acc), (tile_it >> 1) & 1, 3);
          tc_fence_after();
          if (tl && et == 0 && eg == 0 && tile_it == 0) tl[4] = clock64();
          uint32_t v[CW];
          __syncwarp();
          tmem_ld_cols<CW>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), v);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
          if (tc.num_kb == 0) {
#pragma unroll
            for (int j = 0; j < CW; ++j) v[j] = 0u;
          }
          const int ncols = min(CW, p.N - tc.n0 - c0);
          if (!row_ok || ncols <= 0) continue;
          const int64_t col_first = tc.n0 + c0;
          // addend / factor tile (bf16 planes, addressed [n][m]): every load of the thread's columns in flight at once
          float ax[CW];
          if (p.auxp) {
            const int za = (tc.z / p.aux_zd) % p.aux_nb;
            const __nv_bfloat16* a = p.auxp + (int64_t)za * p.auxp_sb + col_first * p.auxp_ld + row;
            __nv_bfloat16 ah[CW], al[CW];
#pragma unroll
            for (int j = 0; j < CW; ++j) {
              const bool in = j < ncols;
              ah[j] = in ? a[0] : __float2bfloat16_rn(0.f);
              al[j] = in ? a[p.auxp_ps] : __float2bfloat16_rn(0.f);
              a += p.auxp_ld;
            }
#pragma unroll
            for (int j = 0; j < CW; ++j) ax[j] = __bfloat162float(ah[j]) + __bfloat162float(al[j]);
          } else {
#pragma unroll
            for (int j = 0; j < CW; ++j) ax[j] = 0.f;
          }
          // transposed tiles: the bias is indexed by the tile ROW m (the output feature: these products put the weight on M)
          const float bm = p.bias ? __ldg(p.bias + (int64_t)tc.z * p.bias_sb + row) : 0.f;
          const bool aux_add = p.aux_mode == TC_AUX_ADD, aux_mul = p.aux_mode == TC_AUX_MUL_1MX2, act = p.act_tanh != 0;
          float f[CW];
          float rsum = 0.f;
#pragma unroll
          for (int j = 0; j < CW; ++j) {
            float x = __uint_as_float(v[j]) + bm;
            if (aux_add) x += ax[j];
            if (act) x = tanh_fast(x);
            if (aux_mul) x *= (1.f - ax[j] * ax[j]);
            f[j] = x;
            rsum += j < ncols ? x : 0.f;
          }
          if (p.D) {
            float* d = p.D + (int64_t)(tc.z / p.d_zd) * p.d_sb + col_first * p.ldd + row;
            if (p.atomic) {
#pragma unroll
              for (int j = 0; j < CW; ++j) { if (j < ncols) atomicAdd(d, f[j]); d += p.ldd; }
            } else if (p.accumulate) {
#pragma unroll
              for (int j = 0; j < CW; ++j) { if (j < ncols) *d += f[j]; d += p.ldd; }
            } else {
#pragma unroll
              for (int j = 0; j < CW; ++j) { if (j < ncols) *d = f[j]; d += p.ldd; }
            }
          }
          if (p.P) {
            __nv_bfloat16* d = p.P + (int64_t)tc.z * p.p_sb + col_first * p.p_ld + row;
#pragma unroll
            for (int j = 0; j < CW; ++j) {
              if (j < ncols) {
                const __nv_bfloat16 h = __float2bfloat16_rn(f[j]);
                d[0] = h;
                d[p.p_ps] = __float2bfloat16_rn(f[j] - __bfloat162float(h));
              }
              d += p.p_ld;
            }
          }
          if (p.red_row) atomicAdd(p.red_row + row, rsum);
          if (tl && et == 0 && eg == 0 && tile_it == 0) tl[5] = clock64();
        }
      };
      if (neg == 2) run(std::integral_constant<int, 16>{});
      else run(std::integral_constant<int, 32>{});
    } else {
      // ------------------------------------------------------------------ standard epilogue
      constexpr bool FULL = (EPI == 1);
      const int aux_kind = FULL ? p.aux_kind : 0;
      const int aux_mode = FULL ? p.aux_mode : (int)TC_AUX_NONE;
      const int mode = FULL ? p.mode : (int)TC_EPI_STORE;
      const bool act_tanh = (EPI != 0) && p.act_tanh != 0;       // EPI = 2: the lean epilogue + tanh
      const bool has_r1 = FULL && CG == 1 && p.r1col != nullptr;
      const float* const mulx = FULL ? p.mulx : nullptr;
      const float* const rowv = FULL ? p.rowv : nullptr;
      const float* const colv = FULL ? p.colv : nullptr;
      const bool do_f32 = (p.D != nullptr) && mode != TC_EPI_ROWDOT;
      const bool do_pl = (p.P != nullptr);          // (ROWDOT may keep f as planes too: the question-side hidden state saved for backward)
      const bool pl_direct = p.pl_direct != 0;
      const bool stage_tma = (do_f32 && p.tma_store) || (do_pl && !pl_direct);
      const bool want_colred = FULL && (p.red_col != nullptr);
      float* const bias_s = bias_sm[eg];
      float* const colv_s = colv_sm[eg];
      float* const colred_s = colred_sm[eg];
      const uint32_t store_base = smem_base + p.off_store + (uint32_t)(eg * p.store_nbuf) * CHUNK_BYTES;
      const uint32_t pstore_base = smem_base + p.off_pstore + (uint32_t)(eg * p.store_nbuf) * CHUNK_BYTES;
      const uint32_t aux_base = smem_base + p.off_aux + (uint32_t)eg * CHUNK_BYTES;
      int tli = 8;                            // debug timeline: stamps of this thread's first chunks (tl[8..63])
      const bool tlt = tl && eg == 0 && et == 32 && !(p.dbg & 4);
#define HCA_TL_STAMP() do { if (tlt && tli < 64) tl[tli++] = clock64(); } while (0)
      int staged_n0 = -1;                     // n0 of the per-column vectors currently in shared memory
      // this group's column sums -> global, one atomic per column (only the columns of its own chunks), then cleared
      auto flush_colred = [&](int n0) {
        for (int j = et; j < BN; j += 128) {
          if (((j >> 5) % neg) == eg && n0 + j < p.N) atomicAdd(p.red_col + n0 + j, colred_s[j]);
        }
      };
      uint32_t gc = 0;                        // chunks this group has processed (staging double-buffer index)
      uint32_t aux_n = 0;                     // addend tiles this group has consumed (mbarrier phase)
      // (at, ac): the next (tile, chunk) whose addend tile has not been requested yet; one load in flight per group.
      // atc caches the decoded coordinates of tile `at` (the decode costs several integer divisions)
      int at = sched0, ac = eg;
      TileCoord atc = decode(at < p.total_tiles ? at : 0);
      auto settle = [&]() {
        while (at < p.total_tiles && ac >= min(CHUNKS, (p.N - atc.n0 + 31) / 32)) {
          at += sched_step;
          ac = eg;
          if (at < p.total_tiles) atc = decode(at);
        }
      };
      auto issue_aux = [&]() {
        const TileCoord& c = atc;
        const int za = (c.z / p.aux_zd) % p.aux_nb;
        mbar_expect_tx(aux_bar(eg), CHUNK_BYTES);
        if (aux_kind == 1) {
          tma_load_3d(aux_base, &maps.AUX, aux_bar(eg), c.n0 + ac * 32, c.m0, za);
        } else {
          tma_load_4d(aux_base, &maps.AUX, aux_bar(eg), c.n0 + ac * 32, c.m0, 0, za);
          tma_load_4d(aux_base + CHUNK_BYTES / 2, &maps.AUX, aux_bar(eg), c.n0 + ac * 32, c.m0, 1, za);
        }
      };
      if (aux_kind) {
        settle();
        if (leader && at < p.total_tiles) issue_aux();
      }
      for (int t = sched0; t < p.total_tiles; t += sched_step, ++tile_it) {
        const TileCoord tc = decode(t);
        const int acc = tile_it & 1;
        const int row = tc.m0 + r;
        const bool row_ok = row < p.M;
        // per-column vectors of this tile: restaged only when they change (a persistent CTA usually keeps its n0: the round-robin
        // stride is a multiple of tiles_n), which also lets the column sums accumulate in shared memory across its tiles
        const bool restage = tc.n0 != staged_n0 || (p.bias && p.bias_sb != 0) || has_r1;
        if (restage) {
          epi_barrier();                      // the previous tile's readers of the per-column vectors are done
          if (want_colred && staged_n0 >= 0 && tc.n0 != staged_n0) flush_colred(staged_n0);
          const float* bias = p.bias ? p.bias + (int64_t)tc.z * p.bias_sb : nullptr;
          for (int j = et; j < BN; j += 128) {
            const bool ok = tc.n0 + j < p.N;
            bias_s[j] = (bias && ok) ? __ldg(bias + tc.n0 + j) : 0.f;
            colv_s[j] = (colv && ok) ? __ldg(colv + tc.n0 + j) : 0.f;
            if (tc.n0 != staged_n0) colred_s[j] = 0.f;
            if constexpr (CG == 1 && FULL) {
              if (p.r1col) {
                const int ng = p.r1_rpg > 0 ? min(4, (p.M + p.r1_rpg - 1) / p.r1_rpg) : 1;
                for (int g = 0; g < ng; ++g)
                  r1_sm[eg][g][j] = ok ? __ldg(p.r1col + (int64_t)tc.z * p.r1col_sb + (int64_t)g * p.r1_gs + tc.n0 + j) : 0.f;
              }
            }
          }
          staged_n0 = tc.n0;
          epi_barrier();
        }
        const float rv = (rowv && row_ok) ? __ldg(rowv + (int64_t)tc.z * p.rowv_sb + row) : 0.f;
        const int rgroup = (has_r1 && p.r1_rpg > 0) ? min(3, row / p.r1_rpg) : 0;
        mbar_wait(tmem_full_bar(acc), (tile_it >> 1) & 1, 3);
        tc_fence_after();
        if (tl && et == 0 && eg == 0 && tile_it == 0) tl[4] = clock64();
        const int nchunks = min(CHUNKS, (p.N - tc.n0 + 31) / 32);           // uniform across the CTA
        if (eg >= nchunks) {                                                // no chunk of this tile for this group
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CG == 2) mbar_arrive_leader(tmem_empty_bar(acc));
            else mbar_arrive(tmem_empty_bar(acc));
          }
        }
        float* drow = nullptr;
        if (p.D && !p.tma_store) {
          const int g = row / p.d_rpg;
          drow = p.D + (int64_t)(tc.z / p.d_zd) * p.d_sb + (int64_t)g * p.d_gs + (int64_t)(row - g * p.d_rpg) * p.ldd;
        }
        const float* xrow = mulx ? mulx + (int64_t)row * p.mulx_ld : nullptr;
        float rowdot = 0.f;
#pragma unroll 1
        for (int c = eg; c < nchunks; c += neg, ++gc) {
          const int col0 = tc.n0 + c * 32;
          const uint32_t sbuf = p.store_nbuf == 2 ? (gc & 1u) : 0u;
          if (stage_tma) {
            if (leader) {                                             // the TMA store that last used this staging buffer has read it
              if (p.store_nbuf == 2) tma_store_wait_read<1>();
              else tma_store_wait_read<0>();
            }
            epi_barrier();
          }
          HCA_TL_STAMP();                                             // 0: staging buffer free
          if (aux_kind) mbar_wait(aux_bar(eg), aux_n & 1u, 4);
          HCA_TL_STAMP();                                             // 1: addend tile landed
          uint32_t v[32];
          __syncwarp();                                               // tcgen05.ld is warp-collective
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), v);
          if (c + neg >= nchunks) {                                   // this group's last read of the accumulator: hand it back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (CG == 2) mbar_arrive_leader(tmem_empty_bar(acc));
              else mbar_arrive(tmem_empty_bar(acc));
            }
          }
          float ax[32];
          if (aux_kind == 1) {                                      // this thread's row of the fp32 addend tile (128-byte swizzle)
            const uint32_t src = aux_base + (uint32_t)r * 128u;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(ax[4 * j]), "=f"(ax[4 * j + 1]), "=f"(ax[4 * j + 2]), "=f"(ax[4 * j + 3])
                           : "r"(src + (uint32_t)((j ^ (r & 7)) * 16)));
            }
          } else if (aux_kind == 2) {                               // hi + lo rows of the bf16 plane tiles (64-byte swizzle)
            const uint32_t src = aux_base + (uint32_t)r * 64u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t h[4], l[4];
              const uint32_t o = (uint32_t)((j ^ ((r >> 1) & 3)) * 16);
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(h[0]), "=r"(h[1]), "=r"(h[2]), "=r"(h[3]) : "r"(src + o));
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(l[0]), "=r"(l[1]), "=r"(l[2]), "=r"(l[3])
                           : "r"(src + CHUNK_BYTES / 2 + o));
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                ax[8 * j + 2 * k] = bf_lo(h[k]) + bf_lo(l[k]);
                ax[8 * j + 2 * k + 1] = bf_hi(h[k]) + bf_hi(l[k]);
              }
            }
          }
          if (aux_kind) {
            // the addend tile is in registers: request the group's next one now, so that its latency hides behind the
            // arithmetic and the stores of this chunk (and behind the other group's work)
            epi_barrier();
            ++aux_n;
            ac += neg;
            settle();
            if (leader && at < p.total_tiles) issue_aux();
          }
          HCA_TL_STAMP();                                             // 2: accumulator + addend in registers, next addend requested
          float f[32];
          {                                                           // bias slice of this chunk: 8 x LDS.128
            const float4* b4 = reinterpret_cast<const float4*>(bias_s + c * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = b4[j];
              f[4 * j] = b.x; f[4 * j + 1] = b.y; f[4 * j + 2] = b.z; f[4 * j + 3] = b.w;
            }
          }
          if (tc.num_kb != 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += __uint_as_float(v[j]);
          }
          if (aux_kind && aux_mode == TC_AUX_ADD) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += ax[j];
          }
          if (act_tanh) {                                           // uniform branches: no predicated-off code on the common path
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = tanh_fast(f[j]);
          }
          if constexpr (CG == 1 && FULL) {
            if (p.r1col) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = fmaf(rv, r1_sm[eg][rgroup][c * 32 + j], f[j]);
            }
          }
          if (aux_kind && aux_mode == TC_AUX_MUL_1MX2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] *= (1.f - ax[j] * ax[j]);
          }
          if (xrow) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (row_ok && col0 + j < p.N) {
                const float h = xrow[col0 + j];
                f[j] *= (1.f - h * h);
              }
            }
          }
          if (mode == TC_EPI_ROWDOT) {
#pragma unroll
            for (int j = 0; j < 32; ++j) rowdot = fmaf(f[j], colv_s[c * 32 + j], rowdot);   // colv_s is 0 beyond N
            if (!do_pl) continue;
          }
          if (want_colred) {
            float part[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) part[j] = row_ok ? f[j] : 0.f;
            const float cs = col_reduce32(part, lane);
            atomicAdd(&colred_s[c * 32 + lane], cs);
          }
          if (stage_tma) {
            if (do_f32 && p.tma_store) {
              // Each thread owns one output row; the 32-column chunk is staged as a [128 rows][128 B] tile in the TMA 128-byte
              // swizzle (16-byte chunk j of row r at chunk j ^ (r & 7): conflict-free float4 stores) and handed to the TMA engine.
              const uint32_t sb = store_base + sbuf * CHUNK_BYTES + (uint32_t)r * 128u;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sb + (uint32_t)((j ^ (r & 7)) * 16)), "f"(f[4 * j]),
                             "f"(f[4 * j + 1]), "f"(f[4 * j + 2]), "f"(f[4 * j + 3])
                             : "memory");
              }
            }
            if (do_pl && !pl_direct) {
              // hi / lo bf16 planes of the same chunk: two [128 rows][64 B] tiles in the 64-byte swizzle
              const uint32_t sb = pstore_base + sbuf * CHUNK_BYTES + (uint32_t)r * 64u;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint32_t h[4], l[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float x0 = f[8 * j + 2 * k], x1 = f[8 * j + 2 * k + 1];
                  const __nv_bfloat162 hh = __floats2bfloat162_rn(x0, x1);
                  const __nv_bfloat162 ll = __floats2bfloat162_rn(x0 - __low2float(hh), x1 - __high2float(hh));
                  h[k] = *reinterpret_cast<const uint32_t*>(&hh);
                  l[k] = *reinterpret_cast<const uint32_t*>(&ll);
                }
                const uint32_t o = (uint32_t)((j ^ ((r >> 1) & 3)) * 16);
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sb + o), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sb + CHUNK_BYTES / 2 + o), "r"(l[0]), "r"(l[1]), "r"(l[2]),
                             "r"(l[3])
                             : "memory");
              }
            }
            HCA_TL_STAMP();                                           // 3: math done, tile staged
            fence_proxy_async_smem();                                 // generic-proxy smem writes -> visible to the TMA engine
            epi_barrier();
            HCA_TL_STAMP();                                           // 4: group barrier passed
            if (leader) {
              if (do_f32 && p.tma_store) {
                const uint32_t sb = store_base + sbuf * CHUNK_BYTES;
                if (p.atomic || p.accumulate) tma_reduce_add_4d(&maps.D, sb, col0, tc.m0, 0, tc.z / p.d_zd);
                else tma_store_4d(&maps.D, sb, col0, tc.m0, 0, tc.z / p.d_zd);
              }
              if (do_pl && !pl_direct) {
                const uint32_t sb = pstore_base + sbuf * CHUNK_BYTES;
                tma_store_4d(&maps.DP, sb, col0, tc.m0, 0, tc.z);
                tma_store_4d(&maps.DP, sb + CHUNK_BYTES / 2, col0, tc.m0, 1, tc.z);
              }
              tma_store_commit();
            }
          }
          if (do_pl && pl_direct) {
            // hi / lo planes through a WARP-PRIVATE staging tile (32 rows x 64 B, 64-byte swizzle) and coalesced 16-byte global
            // stores: lane l writes segment l & 3 of row 8 i + (l >> 2), so one instruction covers 8 rows x 64 contiguous bytes (whole
            // sectors).  Only __syncwarp orders it: no group barrier, no proxy fence, and no wait for a TMA store that sits in the
            // engine's queue behind the mainloop's operand loads (the timeline showed 0.5-1.8k cycles per chunk there).
            const uint32_t wst = smem_base + p.off_pstore + (uint32_t)((eg * 4 + q) * 2048);
            uint32_t hl[2][16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const float x0 = f[2 * k], x1 = f[2 * k + 1];
              const __nv_bfloat162 hh = __floats2bfloat162_rn(x0, x1);
              const __nv_bfloat162 ll = __floats2bfloat162_rn(x0 - __low2float(hh), x1 - __high2float(hh));
              hl[0][k] = *reinterpret_cast<const uint32_t*>(&hh);
              hl[1][k] = *reinterpret_cast<const uint32_t*>(&ll);
            }
            const int rr = lane >> 2, cc = lane & 3;
            __nv_bfloat16* pbase = p.P + (int64_t)tc.z * p.p_sb + (int64_t)(tc.m0 + q * 32) * p.p_ld + col0 + cc * 8;
#pragma unroll
            for (int pl = 0; pl < 2; ++pl) {
              __syncwarp();                                           // the previous readers of the staging tile are done
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(wst + (uint32_t)lane * 64u + (uint32_t)((j ^ ((lane >> 1) & 3)) * 16)),
                             "r"(hl[pl][4 * j]), "r"(hl[pl][4 * j + 1]), "r"(hl[pl][4 * j + 2]), "r"(hl[pl][4 * j + 3])
                             : "memory");
              }
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int R = 8 * i + rr;
                uint4 v;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                             : "r"(wst + (uint32_t)R * 64u + (uint32_t)((cc ^ ((R >> 1) & 3)) * 16)));
                if (tc.m0 + q * 32 + R < p.M) {
                  __nv_bfloat16* dst = pbase + (int64_t)pl * p.p_ps + (int64_t)R * p.p_ld;
                  const int colg = col0 + cc * 8;
                  if (colg + 8 <= p.N) {
                    *reinterpret_cast<uint4*>(dst) = v;
                  } else {
                    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int e2 = 0; e2 < 8; ++e2)
                      if (colg + e2 < p.N) reinterpret_cast<unsigned short*>(dst)[e2] = (unsigned short)(w[e2 >> 1] >> ((e2 & 1) * 16));
                  }
                }
              }
            }
          }
          if (!stage_tma) {
            HCA_TL_STAMP();                                           // 3 / 4: math done, planes stored (no TMA staging on this path)
            HCA_TL_STAMP();
          }
          if (drow && row_ok) {                                       // unaligned fp32 output (e.g. K = 1001 logits): direct stores
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = col0 + j;
              if (col < p.N) {
                if (p.atomic) atomicAdd(drow + col, f[j]);
                else if (p.accumulate) drow[col] += f[j];
                else drow[col] = f[j];
              }
            }
          }
        }
        if (mode == TC_EPI_ROWDOT && row_ok && eg < nchunks) atomicAdd(p.red_row + (int64_t)tc.z * p.red_row_sb + row, rowdot);
        if (tl && et == 0 && eg == 0 && tile_it == 0) tl[5] = clock64();
      }
      if (want_colred && staged_n0 >= 0) {
        epi_barrier();
        flush_colred(staged_n0);
      }
      if (stage_tma && leader) tma_store_wait_read<0>();   // smem must stay valid until the engine has read it
#undef HCA_TL_STAMP
    }
    tc_fence_before();
  }
  __syncthreads();
  if (tl && threadIdx.x == 0) tl[6] = clock64();
  if constexpr (CG == 2) cluster_sync_all();   // remote arrivals and the leader's reads of the peer's shared memory are over
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// fp32 -> bf16 planes
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ src, int64_t ld, int64_t rows, int cols,
                                                           __nv_bfloat16* __restrict__ planes, int64_t ldp, int64_t plane_stride,
                                                           int P) {
  pdl_enter();
  const int c4n = (cols + 3) / 4;
  const int64_t total = rows * c4n;
  const bool aligned = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c4n;
    const int c = (int)(i - r * c4n) * 4;
    float x[4];
    if (aligned && c + 4 <= cols) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(src + r * ld + c));
      x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) x[j] = (c + j < cols) ? src[r * ld + c + j] : 0.f;
    }
    for (int pl = 0; pl < P; ++pl) {
      __nv_bfloat16 h[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        h[j] = __float2bfloat16_rn(x[j]);
        x[j] -= __bfloat162float(h[j]);
      }
      __nv_bfloat16* dst = planes + pl * plane_stride + r * ldp + c;     // ldp % 8 == 0 and c % 4 == 0: 8-byte aligned
      *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(h);
    }
  }
}

struct SplitJobs {
  SplitBatch::Job job[SplitBatch::MAX];
  long long start[SplitBatch::MAX + 1];      // prefix sums of the jobs' float4 counts
  int count;
};
__global__ void __launch_bounds__(256) split_planes_multi_kernel(const SplitJobs jobs, const ZeroJobs zero) {
  pdl_enter();
  zero_jobs_device(zero);
  const long long total = jobs.start[jobs.count];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int j = 0;
    while (j + 1 < jobs.count && i >= jobs.start[j + 1]) ++j;
    const SplitBatch::Job& b = jobs.job[j];
    const long long k = i - jobs.start[j];
    const int c4n = b.cols >> 2;
    const long long r = k / c4n;
    const int c = (int)(k - r * c4n) * 4;
    const float4 t = __ldg(reinterpret_cast<const float4*>(b.src + r * b.ld + c));
    float x[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int pl = 0; pl < 2; ++pl) {
      __nv_bfloat16 h[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        h[q] = __float2bfloat16_rn(x[q]);
        x[q] -= __bfloat162float(h[q]);
      }
      *reinterpret_cast<uint2*>(b.dst + pl * b.ps + r * b.ldp + c) = *reinterpret_cast<const uint2*>(h);
    }
  }
}

// three [B][T][cols] sources -> stacked planes [2][B][3T][ldp]   (cols % 4 == 0, sources 16-byte aligned)
__global__ void __launch_bounds__(256) split_planes_stack3_kernel(const float4* __restrict__ s0, const float4* __restrict__ s1,
                                                                  const float4* __restrict__ s2, int B, int T, int c4n,
                                                                  __nv_bfloat16* __restrict__ planes, int64_t ldp, int64_t plane_stride) {
  pdl_enter();
  const int64_t per = (int64_t)B * T * c4n, total = 3 * per;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int l = (int)(i / per);
    const int64_t k = i - l * per;
    const int64_t bt = k / c4n;
    const int c = (int)(k - bt * c4n) * 4;
    const int b = (int)(bt / T), t = (int)(bt - (int64_t)b * T);
    const float4 v = __ldg((l == 0 ? s0 : (l == 1 ? s1 : s2)) + k);
    float x[4] = {v.x, v.y, v.z, v.w};
    const int64_t row = ((int64_t)b * 3 + l) * T + t;
#pragma unroll
    for (int pl = 0; pl < 2; ++pl) {
      __nv_bfloat16 h[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        h[j] = __float2bfloat16_rn(x[j]);
        x[j] -= __bfloat162float(h[j]);
      }
      *reinterpret_cast<uint2*>(planes + pl * plane_stride + row * ldp + c) = *reinterpret_cast<const uint2*>(h);
    }
  }
}

// ------------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    else
      cudaGetLastError();
  }
  return fn;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = 148;
    }
  }
  return n;
}

// 4-D map over bf16 planes: dims (cols, rows, P, batch); box (box_cols, box_rows, 1, 1); OOB -> zeros
int make_tmap_bf16(CUtensorMap* tm, const void* base, int64_t ld, int64_t plane_stride, int64_t batch_stride, int nbatch, int rows,
                   int cols, int P, int box_cols, int box_rows, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encoder();
  if (!enc) return set_err(HCA_ERR_CUDA, "gemm_tc: cuTensorMapEncodeTiled is not available from the driver");
  const int nb = nbatch > 0 ? nbatch : 1;
  cuuint64_t gdim[4] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)P, (cuuint64_t)nb};
  cuuint64_t gstr[3] = {(cuuint64_t)ld * 2, (cuuint64_t)plane_stride * 2, (cuuint64_t)(nb > 1 ? batch_stride : plane_stride * P) * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_err(HCA_ERR_CUDA, "gemm_tc: cuTensorMapEncodeTiled (bf16) failed (%d) cols=%d rows=%d ld=%lld ps=%lld bs=%lld P=%d nb=%d", (int)r,
                   cols, rows, (long long)ld, (long long)plane_stride, (long long)batch_stride, P, nb);
  return 0;
}
int make_tmap(CUtensorMap* tm, const TcOperand& o, int P, int box_rows) {
  return make_tmap_bf16(tm, o.planes, o.ld, o.plane_stride, o.batch_stride, o.nbatch, o.rows, o.cols, P, 64, box_rows,
                        CU_TENSOR_MAP_SWIZZLE_128B);
}

// 3-D fp32 map over X [batch][M, N] (leading dim ld): box = 32 columns (128 B) x 128 rows x 1, 128-byte swizzle
int make_tmap_f32(CUtensorMap* tm, const float* X, int64_t ld, int64_t batch_stride, int nbatch, int M, int N) {
  EncodeTiledFn enc = get_encoder();
  if (!enc) return set_err(HCA_ERR_CUDA, "gemm_tc: cuTensorMapEncodeTiled is not available from the driver");
  const int nb = nbatch > 0 ? nbatch : 1;
  cuuint64_t gdim[3] = {(cuuint64_t)N, (cuuint64_t)M, (cuuint64_t)nb};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 4, (cuuint64_t)(nb > 1 ? batch_stride : (int64_t)M * ld) * 4};
  cuuint32_t box[3] = {32, (cuuint32_t)BM, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)X, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_err(HCA_ERR_CUDA, "gemm_tc: cuTensorMapEncodeTiled (fp32 tile) failed (%d) M=%d N=%d ld=%lld nb=%d", (int)r, M, N,
                   (long long)ld, nb);
  return 0;
}

// 4-D fp32 output map: dims (N, rows per group, groups, batch); box (32, min(rows per group, 128), groups (if > 1), 1)
int make_tmap_f32_out(CUtensorMap* tm, const float* X, int64_t ld, int64_t group_stride, int groups, int64_t batch_stride, int nbatch,
                      int rows_per_group, int N) {
  EncodeTiledFn enc = get_encoder();
  if (!enc) return set_err(HCA_ERR_CUDA, "gemm_tc: cuTensorMapEncodeTiled is not available from the driver");
  const int nb = nbatch > 0 ? nbatch : 1;
  const int64_t gs = groups > 1 ? group_stride : (int64_t)rows_per_group * ld;
  const int64_t bs = nb > 1 ? batch_stride : gs * groups;
  cuuint64_t gdim[4] = {(cuuint64_t)N, (cuuint64_t)rows_per_group, (cuuint64_t)groups, (cuuint64_t)nb};
  cuuint64_t gstr[3] = {(cuuint64_t)ld * 4, (cuuint64_t)gs * 4, (cuuint64_t)bs * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)(groups > 1 ? rows_per_group : BM), (cuuint32_t)groups, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)X, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_err(HCA_ERR_CUDA, "gemm_tc: cuTensorMapEncodeTiled (fp32 out) failed (%d) rpg=%d N=%d ld=%lld groups=%d gs=%lld bs=%lld nb=%d",
                   (int)r, rows_per_group, N, (long long)ld, groups, (long long)gs, (long long)bs, nb);
  return 0;
}

bool f32_tma_ok(const float* p, int64_t ld, int64_t s1, int64_t s2) {
  return ((ld & 3) == 0) && ((s1 & 3) == 0) && ((s2 & 3) == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0);
}
bool planes_ok(const TcPlanes& t) {
  return t.p && (t.ld % 8) == 0 && (t.plane_stride % 8) == 0 && (t.batch_stride % 8) == 0 && ((reinterpret_cast<uintptr_t>(t.p) & 15) == 0);
}

}  // namespace

bool tc_available() { return get_encoder() != nullptr; }

// N-tile range [first, first + count) of the M tile covering rows [m0, m0 + bm) under the N windows of `e`
static void nwin_range(const TcEpilogue& e, int M, int N, int m0, int bm, int bn, int& first, int& count) {
  const int tiles_n = (N + bn - 1) / bn;
  if (e.nwin_nrow <= 0) { first = 0; count = tiles_n; return; }
  const int g0 = m0 / e.nwin_nrow, g1 = (std::min(M, m0 + bm) - 1) / e.nwin_nrow;
  int lo = N, hi = 0;
  for (int g = g0; g <= g1 && g < 4; ++g) {
    lo = std::min(lo, std::max(0, e.nwin_lo[g]));
    hi = std::max(hi, std::min(N, e.nwin_hi[g]));
  }
  if (g1 >= 4) { lo = 0; hi = N; }
  if (hi <= lo) { first = 0; count = 0; return; }
  first = lo / bn;
  count = std::min(tiles_n, (hi + bn - 1) / bn) - first;
}
int tc_count_tiles(int M, int N, const TcEpilogue& e, int bm, int bn) {
  int total = 0;
  for (int m0 = 0; m0 < M; m0 += bm) {
    int first, count;
    nwin_range(e, M, N, m0, bm, bn, first, count);
    total += count;
  }
  return total;
}
int tc_splitk(int M, int N, int K) { return tc_splitk_tiles(((M + 127) / 128) * ((N + 127) / 128), K); }
int tc_splitk_tiles(int tiles, int K) {
  const int sms = num_sms();
  if (tiles * 3 >= sms * 2) return 1;             // the tiles alone fill most of a wave
  int sk = sms / tiles;                           // floor: one work item per SM at most
  const int maxk = (K + 255) / 256;               // keep at least 4 k-blocks per slice
  if (sk > maxk) sk = maxk;
  return sk < 1 ? 1 : sk;
}

int tc_make_tmap(void* tm, bool bf16, int rank, const void* base, const uint64_t* dims, const uint64_t* strides_bytes,
                 const uint32_t* box, int swizzle) {
  EncodeTiledFn enc = get_encoder();
  if (!enc) return set_err(HCA_ERR_CUDA, "tc_make_tmap: cuTensorMapEncodeTiled is not available from the driver");
  HCA_CHECK_ARG(rank >= 2 && rank <= 5 && swizzle >= 0 && swizzle <= 3, "tc_make_tmap: bad rank / swizzle");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], estr[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; estr[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  static const CUtensorMapSwizzle swz[4] = {CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_SWIZZLE_64B,
                                            CU_TENSOR_MAP_SWIZZLE_128B};
  CUresult r = enc((CUtensorMap*)tm, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                   const_cast<void*>(base), gdim, gstr, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz[swizzle],
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_err(HCA_ERR_CUDA, "tc_make_tmap: cuTensorMapEncodeTiled failed (%d) rank=%d dims=%llu,%llu box=%u,%u", (int)r, rank,
                   (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
  return 0;
}

void tc_set_timeline(long long* buf, int nctas, int launch_index) {
  g_timeline = buf;
  g_timeline_ctas = nctas;
  g_timeline_target = launch_index;
  g_timeline_seen = 0;
}

long long* tc_timeline_buffer() { return g_timeline; }

int launch_split_planes(const float* src, int64_t ld, int64_t rows, int cols, __nv_bfloat16* planes, int64_t ldp,
                        int64_t plane_stride, int P, cudaStream_t s) {
  HCA_CHECK_ARG(src && planes && rows > 0 && cols > 0 && P >= 1 && P <= 3 && (ldp % 8) == 0 && (plane_stride % 8) == 0,
                "split_planes: bad arguments");
  const int64_t total = rows * ((cols + 3) / 4);
  HCA_LAUNCH_K((split_planes_kernel), ew_grid(total), 256, 0, s, src, ld, rows, cols, planes, ldp, plane_stride, P);
  HCA_LAUNCHED();
  return 0;
}

int SplitBatch::add(const float* src, int64_t ld, int64_t rows, int cols, __nv_bfloat16* planes, int64_t ldp, int64_t plane_stride) {
  HCA_CHECK_ARG(src && planes && rows > 0 && cols > 0, "SplitBatch: bad arguments");
  const bool vec = (cols % 4) == 0 && (ld % 4) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (ldp % 4) == 0 && (plane_stride % 4) == 0 &&
                   (reinterpret_cast<uintptr_t>(planes) & 7) == 0;
  if (!vec) return launch_split_planes(src, ld, rows, cols, planes, ldp, plane_stride, 2, stream);      // odd shapes: the general kernel
  if (count == MAX) HCA_TRY(flush());
  job[count++] = Job{src, planes, (long long)ld, (long long)rows, (long long)ldp, (long long)plane_stride, cols};
  return 0;
}

int SplitBatch::flush() {
  ZeroBatch none(stream);
  return flush(none);
}
int SplitBatch::flush(ZeroBatch& also_clear) {
  if (count == 0) return also_clear.flush();
  SplitJobs jobs;
  memset(&jobs, 0, sizeof(jobs));
  long long total = 0;
  for (int i = 0; i < count; ++i) {
    jobs.job[i] = job[i];
    jobs.start[i] = total;
    total += job[i].rows * (job[i].cols / 4);
  }
  jobs.start[count] = total;
  jobs.count = count;
  count = 0;
  unsigned long long zwords = 0;
  for (int i = 0; i < also_clear.count; ++i) zwords += also_clear.words[i];
  HCA_LAUNCH_K((split_planes_multi_kernel), ew_grid(std::max<long long>(total, (long long)(zwords / 4 + 1))), 256, 0, stream, jobs, also_clear.take());
  HCA_LAUNCHED();
  return 0;
}

int launch_split_planes_stack3(const float* s0, const float* s1, const float* s2, int B, int T, int cols, __nv_bfloat16* planes,
                               int64_t ldp, int64_t plane_stride, cudaStream_t s) {
  HCA_CHECK_ARG(s0 && s1 && s2 && planes && B > 0 && T > 0 && cols > 0 && (cols % 4) == 0 && (ldp % 8) == 0 && (plane_stride % 8) == 0,
                "split_planes_stack3: bad arguments");
  HCA_CHECK_ARG(((reinterpret_cast<uintptr_t>(s0) | reinterpret_cast<uintptr_t>(s1) | reinterpret_cast<uintptr_t>(s2)) & 15) == 0,
                "split_planes_stack3: sources must be 16-byte aligned");
  const int64_t total = 3LL * B * T * (cols / 4);
  HCA_LAUNCH_K((split_planes_stack3_kernel), ew_grid(total), 256, 0, s, (const float4*)s0, (const float4*)s1, (const float4*)s2, B, T, cols / 4, planes,
                                                           ldp, plane_stride);
  HCA_LAUNCHED();
  return 0;
}

int launch_gemm_tc(const TcOperand& A, const TcOperand& B, int P, int M, int N, int K, const TcEpilogue& e, int splitk,
                   cudaStream_t s, int batch, const TcOperand* A2, const TcOperand* B2, int K2) {
  HCA_CHECK_ARG(P >= 2 && P <= 3 && M > 0 && N > 0 && K > 0 && splitk >= 1 && batch >= 1, "gemm_tc: bad sizes");
  HCA_CHECK_ARG(batch == 1 || splitk == 1, "gemm_tc: split-K is for un-batched products");
  HCA_CHECK_ARG((A2 == nullptr) == (B2 == nullptr), "gemm_tc: the chained operand pair needs both A2 and B2");
  HCA_CHECK_ARG(!e.transposed || P == 2, "gemm_tc: the transposed epilogue is instantiated for P = 2");
  // CTA-pair mode (cta_group::2, 256 x 256 tiles per pair): the large K-major products with a plain epilogue, where the L2 -> SM
  // operand fill is the limit.  HCA_TC_PAIR=0 disables it, =1 also takes smaller M (tests).
  const bool pair_epi = !e.transposed && P == 2 && !A2 && batch == 1 && e.mode == TC_EPI_STORE && e.aux_mode == TC_AUX_NONE && !e.r1col &&
                        !e.mulx && e.d_groups <= 1 && A.nbatch <= 1 && B.nbatch <= 1 && num_sms() >= 2;
  // (a) large K-major projections, (b) split-K weight gradients (both operands MN-major, plain fp32 reduce-add epilogue, K long
  // enough to leave every pair several k-blocks): there the mainloop's operand fill is the limit and the pair halves it
  bool pair_wgrad = pair_epi && splitk > 1 /* the caller has cleared D */ && A.mn_major && B.mn_major && (M % 256) == 0 && (N % 256) == 0 && K >= 2048 && !e.bias && !e.act_tanh &&
                    !e.red_col && !e.P.p && e.D != nullptr && f32_tma_ok(e.D, e.ldd, e.d_batch_stride, 0);
  if (HCA_ENV_INT("HCA_TC_PAIR_WGRAD", 1) == 0) pair_wgrad = false;
  bool pair = pair_epi && !A.mn_major && !B.mn_major && splitk == 1 && (N % 256) == 0 && M >= 8192 && K >= 256 && e.kwin_ncol == 0;
  if (pair) {
    // wave quantisation: a pair tile is four single tiles of MMA time on two SMs.  Take the pair schedule only when its last,
    // partly filled wave does not cost more than the fabric traffic it saves (PV, M = 31360: 4 pair waves vs 7 single waves;
    // PQ, M = 12480: 2 pair waves vs 3 single ones -- measured slower as a pair)
    const int64_t t2 = (int64_t)((M + 255) / 256) * (N / 256), t1 = (int64_t)((M + 127) / 128) * ((N + 127) / 128);
    const int64_t nc = num_sms() / 2, w2 = (t2 + nc - 1) / nc, w1 = (t1 + num_sms() - 1) / num_sms();
    if ((double)(2 * w2) > 1.15 * (double)w1) pair = false;
  }
  {
    const int pair_env = HCA_ENV_INT("HCA_TC_PAIR", -1);
    if (pair_env == 0) pair = false;
    if (pair_env == 1 && e.kwin_ncol == 0 && !e.transposed && P == 2 && !A.mn_major && !B.mn_major && !A2 && splitk == 1 && batch == 1 && (N % 256) == 0 &&
        e.mode == TC_EPI_STORE && e.aux_mode == TC_AUX_NONE && !e.r1col && !e.mulx && e.d_groups <= 1 && A.nbatch <= 1 && B.nbatch <= 1)
      pair = true;
  }
  if (HCA_ENV_INT("HCA_TC_PAIR", -1) == 0) pair_wgrad = false;
  if (pair_wgrad) {
    pair = true;
    // the split is chosen for the pair grid: one 256 x 256 tile per pair and wave, at least 4 k-blocks per split
    const int t2 = std::max(1, tc_count_tiles(M, N, e, 256, 256)), nc = num_sms() / 2, kbt = (K + 63) / 64;
    int sk = nc / t2;                    // ONE wave: tiles * slices <= pairs (a ceil here left 2 of 76 items for a second wave: 2x the time)
    if (sk > kbt / 4) sk = kbt / 4;
    splitk = sk < 1 ? 1 : sk;
  }
  const int BN = e.transposed ? 32 : (pair ? 256 : 128);
  const int CGn = pair ? 2 : 1;
  const TcOperand* ops[4] = {&A, &B, A2, B2};
  for (const TcOperand* o : ops) {
    if (!o) continue;
    HCA_CHECK_ARG((o->ld % 8) == 0 && (o->plane_stride % 8) == 0 && (o->batch_stride % 8) == 0,
                  "gemm_tc: plane leading dimensions / strides must be multiples of 8 elements (TMA 16-byte strides)");
    HCA_CHECK_ARG((reinterpret_cast<uintptr_t>(o->planes) & 15) == 0, "gemm_tc: planes must be 16-byte aligned");
    HCA_CHECK_ARG(o->zdiv >= 1, "gemm_tc: zdiv must be >= 1");
  }
  HCA_CHECK_ARG(!(A.mn_major && !B.mn_major) || e.transposed, "gemm_tc: (MN-major A, K-major B) is only instantiated for the transposed epilogue");
  HCA_CHECK_ARG(!(e.transposed && B.mn_major), "gemm_tc: the transposed (BN = 32) kernel takes a K-major B");
  if (A2) HCA_CHECK_ARG(A2->mn_major == A.mn_major && B2->mn_major == B.mn_major && K2 > 0, "gemm_tc: chained pair must share the layouts");
  // short contractions over MN-major pairs (K = T tokens per level) use 32-deep k-blocks: half the smem, TMA and MMA work
  const int BK = (A.mn_major && B.mn_major && P == 2 && BN == 128 && !A2 && K <= 96) ? 32 : 64;
  TcMaps maps;
  HCA_TRY(make_tmap(&maps.A, A, P, A.mn_major ? BK : BM));
  HCA_TRY(make_tmap(&maps.B, B, P, B.mn_major ? BK : BN / CGn));
  if (A2) {
    HCA_TRY(make_tmap(&maps.A2, *A2, P, A2->mn_major ? BK : BM));
    HCA_TRY(make_tmap(&maps.B2, *B2, P, B2->mn_major ? BK : BN));
  } else {
    maps.A2 = maps.A;
    maps.B2 = maps.B;
  }
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K; p.K2 = A2 ? K2 : 0;
  const bool rowdot = e.mode == TC_EPI_ROWDOT;
  const bool want_f32 = e.D != nullptr && !rowdot;
  const bool want_pl = e.P.p != nullptr;
  HCA_CHECK_ARG(rowdot || want_f32 || want_pl || e.red_col || (e.transposed && e.red_row), "gemm_tc: no output requested");
  const int groups = e.d_groups > 1 ? e.d_groups : 1;
  HCA_CHECK_ARG(groups == 1 || (M % groups == 0 && M <= BM && !e.transposed), "gemm_tc: grouped output rows need M <= 128, M %% groups == 0");
  const int rpg = M / groups;
  bool tma_store = false;
  if (want_f32 && !e.transposed) {
    tma_store = f32_tma_ok(e.D, e.ldd, e.d_batch_stride, groups > 1 ? e.d_group_stride : 0);
    const int nb_d = (batch + (e.d_zdiv > 1 ? e.d_zdiv : 1) - 1) / (e.d_zdiv > 1 ? e.d_zdiv : 1);
    if (tma_store) HCA_TRY(make_tmap_f32_out(&maps.D, e.D, e.ldd, e.d_group_stride, groups, e.d_batch_stride, nb_d, rpg, N));
  }
  if (!tma_store) maps.D = maps.A;                                     // unused placeholder
  if (want_pl) {
    HCA_CHECK_ARG(planes_ok(e.P), "gemm_tc: output planes need 16-byte aligned rows / strides");
    if (!e.transposed)
      HCA_TRY(make_tmap_bf16(&maps.DP, e.P.p, e.P.ld, e.P.plane_stride, e.P.batch_stride, batch, M, N, 2, 32, BM, CU_TENSOR_MAP_SWIZZLE_64B));
  }
  if (!want_pl || e.transposed) maps.DP = maps.A;
  int aux_kind = 0;
  if (e.aux_mode != TC_AUX_NONE) {
    if (e.transposed) {
      HCA_CHECK_ARG(e.auxp.p != nullptr, "gemm_tc: the transposed epilogue takes its addend as bf16 planes");
      aux_kind = 2;
      maps.AUX = maps.A;
    } else if (e.auxp.p) {
      HCA_CHECK_ARG(planes_ok(e.auxp), "gemm_tc: aux planes need 16-byte aligned rows / strides");
      aux_kind = 2;
      HCA_TRY(make_tmap_bf16(&maps.AUX, e.auxp.p, e.auxp.ld, e.auxp.plane_stride, e.auxp.batch_stride, e.auxp.nbatch, M, N, 2, 32, BM,
                             CU_TENSOR_MAP_SWIZZLE_64B));
    } else {
      HCA_CHECK_ARG(e.aux && f32_tma_ok(e.aux, e.aux_ld, e.aux_batch_stride, 0), "gemm_tc: aux tile must have 16-byte aligned rows");
      aux_kind = 1;
      HCA_TRY(make_tmap_f32(&maps.AUX, e.aux, e.aux_ld, e.aux_batch_stride, e.aux_nbatch, M, N));
    }
  } else {
    maps.AUX = maps.A;
  }
  // shared memory carve: pipeline stages first, then the epilogue staging areas this launch needs (per epilogue group)
  const uint32_t stage_bytes = (uint32_t)P * (uint32_t)(BM * BK * 2 + (BN / CGn) * BK * 2);
  // (static shared memory + the 1024-byte alignment slack take ~9 KB of the 227 KB; ~10.5 KB with the 256-wide per-column vectors)
  const uint32_t avail = SMEM_LIMIT - (pair ? 11264 : 9216);
  // (only behind a long mainloop: there a TMA store queues behind the operand loads; the single-k-block products keep the
  // double-buffered asynchronous TMA stores their epilogue-bound tiles were tuned with)
  p.kb1 = (K + BK - 1) / BK;
  p.kb_total = p.kb1 + (A2 ? (K2 + BK - 1) / BK : 0);
  bool pl_direct = want_pl && !e.transposed && p.kb_total >= 4;
  if (HCA_ENV_INT("HCA_TC_PLDIRECT", 1) == 0) pl_direct = false;      // (re-measured on the final code of round 2: step 1.549 -> 1.540 ms with it; HCA_TC_PLDIRECT=0 for the A/B)
  const int n_out = (!e.transposed && want_f32 && tma_store ? 1 : 0) + (!e.transposed && want_pl && !pl_direct ? 1 : 0);
  const bool has_aux_buf = !e.transposed && aux_kind;
  auto stages_for = [&](int neg, int nbuf) {
    const uint32_t epi = (uint32_t)neg * ((uint32_t)(n_out * nbuf) + (has_aux_buf ? 1u : 0u)) * CHUNK_BYTES + (pl_direct ? (uint32_t)neg * 8192u : 0u);
    return epi >= avail ? 0 : (int)((avail - epi) / stage_bytes);
  };
  // Two epilogue groups when the epilogue is the long pole (fused math / plane conversion behind a short mainloop) and the
  // operand pipeline still gets the stages it needs; one group (deeper pipeline) for plain mainloop-bound products.
  // (measured on the PV projection, K = 512, planes out: one group + a third operand stage 56.7 us, two groups + two stages 67.2 us --
  // a plane-conversion epilogue hides behind 8 k-blocks of MMAs; only fused math or a short mainloop needs the second group)
  const bool epi_math = e.act_tanh || e.aux_mode != TC_AUX_NONE || e.mode != TC_EPI_STORE || e.transposed || e.r1col;
  const bool epi_heavy = epi_math || (want_pl && p.kb_total < 8);
  int neg = 1, nbuf = 2;
  if (epi_heavy && stages_for(2, 1) >= (p.kb_total == 1 ? 1 : 2)) neg = 2;
  // planes-only output through the warp-private staging costs 8 KB per group: take the second group whenever it is free
  if (neg == 1 && pl_direct && n_out == 0 && std::min(stages_for(2, 1), MAX_STAGES) >= std::min(stages_for(1, 1), MAX_STAGES)) neg = 2;
  { const int ev = HCA_ENV_INT("HCA_TC_EG", 0); if (ev == 1 || ev == 2) neg = ev; }
  // single-k-block tiles (the K = T products) need no operand pipelining beyond the TMEM double buffer: one stage is enough
  // there, which leaves room to double-buffer the epilogue staging -- their long pole
  const int min_stages = p.kb_total == 1 ? 1 : 2;
  if (p.kb_total == 1) {
    if (stages_for(neg, 2) < 1) nbuf = 1;
  } else if (stages_for(neg, 2) < 4 && stages_for(neg, 1) > stages_for(neg, 2)) {
    nbuf = 1;                                     // a deeper operand pipeline beats double-buffered staging
  }
  int stages = stages_for(neg, nbuf);
  if (stages < min_stages && neg == 2) { neg = 1; nbuf = 1; stages = stages_for(1, 1); }
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  { const int ev = HCA_ENV_INT("HCA_TC_STAGES", 0); if (ev >= 1 && ev < stages) stages = ev; }
  HCA_CHECK_ARG(stages >= min_stages, "gemm_tc: tile does not fit its pipeline stages");
  p.stages = stages;
  p.store_nbuf = nbuf;
  p.n_eg = neg;
  uint32_t off = (uint32_t)stages * stage_bytes;
  if (!e.transposed && want_f32 && tma_store) { p.off_store = off; off += (uint32_t)(neg * nbuf) * CHUNK_BYTES; }
  if (!e.transposed && want_pl && !pl_direct) { p.off_pstore = off; off += (uint32_t)(neg * nbuf) * CHUNK_BYTES; }
  if (pl_direct) { p.off_pstore = off; off += (uint32_t)neg * 8192u; }      // 2 KB of warp-private staging per epilogue warp
  if (has_aux_buf) { p.off_aux = off; off += (uint32_t)neg * CHUNK_BYTES; }
  if (splitk > p.kb_total) splitk = p.kb_total;
  p.kb_per_split = (p.kb_total + splitk - 1) / splitk;
  splitk = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;      // no empty split
  p.splitk = splitk;
  if (e.kwin_ncol > 0) {
    HCA_CHECK_ARG(splitk == 1 && batch == 1 && !A2 && !pair && (N + e.kwin_ncol - 1) / e.kwin_ncol <= 4, "gemm_tc: K windows are for un-batched, un-split products with at most 4 column groups");
    p.kwin_ncol = e.kwin_ncol;
    for (int g = 0; g < 4; ++g) {
      const int lo = std::max(0, std::min(e.kwin_lo[g], K)), hi = std::max(lo, std::min(e.kwin_hi[g], K));
      p.kwin_lo[g] = lo / BK;
      p.kwin_hi[g] = (hi + BK - 1) / BK;
    }
  }
  p.tiles_m = (M + BM * CGn - 1) / (BM * CGn);
  p.tiles_n = (N + BN - 1) / BN;
  int64_t tiles_mn = (int64_t)p.tiles_m * p.tiles_n;
  if (e.nwin_nrow > 0 && batch == 1 && p.tiles_m <= NWIN_TM_MAX && tiles_mn < 32768) {      // (otherwise: every tile, which is always correct)
    p.nwin_on = 1;
    p.nw_pre[0] = 0;
    for (int mt = 0; mt < p.tiles_m; ++mt) {
      int first, count;
      nwin_range(e, M, N, mt * BM * CGn, BM * CGn, BN, first, count);
      p.nw_first[mt] = (short)first;
      p.nw_pre[mt + 1] = (short)(p.nw_pre[mt] + count);
    }
    tiles_mn = p.nw_pre[p.tiles_m];
    HCA_CHECK_ARG(tiles_mn > 0, "gemm_tc: the N windows leave no tile");
  }
  p.tiles_mn = (int)tiles_mn;
  const int64_t total = tiles_mn * splitk * batch;
  HCA_CHECK_ARG(total < (1LL << 30), "gemm_tc: too many tiles");
  p.total_tiles = (int)total;
  auto nbf = [](int nb) { return nb > 0 ? nb : 1; };
  p.a_nb = nbf(A.nbatch); p.a_zd = A.zdiv; p.b_nb = nbf(B.nbatch); p.b_zd = B.zdiv;
  p.a2_nb = A2 ? nbf(A2->nbatch) : 1; p.a2_zd = A2 ? A2->zdiv : 1; p.b2_nb = B2 ? nbf(B2->nbatch) : 1; p.b2_zd = B2 ? B2->zdiv : 1;
  p.D = want_f32 ? e.D : nullptr; p.ldd = e.ldd; p.d_sb = e.d_batch_stride; p.d_zd = e.d_zdiv > 1 ? e.d_zdiv : 1;
  p.d_rpg = rpg; p.d_gs = e.d_group_stride;
  p.accumulate = e.accumulate;
  p.atomic = (splitk > 1 || (p.d_zd > 1 && want_f32)) ? 1 : 0;
  p.tma_store = tma_store ? 1 : 0;
  p.pl_direct = pl_direct ? 1 : 0;
  p.P = want_pl ? e.P.p : nullptr; p.p_ld = e.P.ld; p.p_ps = e.P.plane_stride; p.p_sb = e.P.batch_stride;
  p.bias = e.bias; p.bias_sb = e.bias_batch_stride;
  p.act_tanh = e.act_tanh; p.mulx = e.mulx; p.mulx_ld = e.mulx_ld;
  p.mode = e.mode; p.aux_kind = aux_kind; p.aux_mode = e.aux_mode;
  if (aux_kind == 2) {
    p.aux_nb = nbf(e.auxp.nbatch); p.aux_zd = e.auxp.zdiv > 1 ? e.auxp.zdiv : 1;
    p.auxp = e.auxp.p; p.auxp_ld = e.auxp.ld; p.auxp_ps = e.auxp.plane_stride; p.auxp_sb = e.auxp.batch_stride;
  } else {
    p.aux_nb = nbf(e.aux_nbatch); p.aux_zd = e.aux_zdiv > 1 ? e.aux_zdiv : 1;
  }
  p.rowv = e.rowv; p.rowv_sb = e.rowv_batch_stride;
  p.colv = e.colv;
  p.r1col = e.r1col; p.r1col_sb = e.r1col_batch_stride; p.r1_rpg = e.r1_rows_per_group; p.r1_gs = e.r1_group_stride;
  p.red_row = e.red_row; p.red_row_sb = e.red_row_batch_stride;
  p.red_col = e.red_col;
  const bool nonlinear = e.bias || e.act_tanh || e.mulx || e.aux_mode || e.mode != TC_EPI_STORE || e.r1col;
  HCA_CHECK_ARG(!(splitk > 1 && (nonlinear || want_pl || e.red_col)), "gemm_tc: split-K needs a linear fp32 epilogue");
  HCA_CHECK_ARG(!(p.d_zd > 1 && !e.accumulate), "gemm_tc: d_zdiv > 1 means several tiles add into one output: set accumulate");
  HCA_CHECK_ARG(e.mode != TC_EPI_ROWDOT || (e.colv && e.red_row), "gemm_tc: ROWDOT needs colv and red_row");
  HCA_CHECK_ARG(!e.r1col || (e.rowv && e.mode == TC_EPI_STORE), "gemm_tc: the rank-1 term needs rowv and the STORE mode");
  HCA_CHECK_ARG(!e.r1col || e.r1_rows_per_group <= 0 || (M + e.r1_rows_per_group - 1) / e.r1_rows_per_group <= 4,
                "gemm_tc: at most 4 rank-1 row groups");
  if (e.transposed)
    HCA_CHECK_ARG(!e.mulx && e.mode == TC_EPI_STORE && !e.r1col && !e.red_col && groups == 1 && !e.aux,
                  "gemm_tc: the transposed epilogue supports fp32 / planes output, a per-row bias, tanh, auxp and red_row only");
  p.dbg = HCA_ENV_INT("HCA_TC_DBG", 0);
  p.timeline = (g_timeline && (g_timeline_target < 0 || g_timeline_seen == g_timeline_target)) ? g_timeline : nullptr;
  p.timeline_ctas = g_timeline_ctas;
  if (g_timeline) ++g_timeline_seen;
  const size_t smem = (size_t)off + 1024;
  const size_t smem_cap = (size_t)SMEM_LIMIT - (pair ? 10240 : 8192);
  HCA_CHECK_ARG(smem <= smem_cap, "gemm_tc: shared memory carve exceeds the limit");
  int ctas = num_sms();
  if (pair) ctas = 2 * std::min(num_sms() / 2, p.total_tiles);
  else if (p.total_tiles < ctas) ctas = p.total_tiles;
  typedef void (*KernelFn)(const TcMaps, const TcParams);
  KernelFn fn = nullptr;
  const int combo = (A.mn_major ? 2 : 0) + (B.mn_major ? 1 : 0);    // 0 = NT (K,K), 1 = NN (K,MN), 2 = (MN,K), 3 = TN (MN,MN)
  int slot = -1;
  // lean epilogue instantiation (EPI = 0) for products without fused epilogue math; HCA_TC_LEAN=0 forces the full one
  bool lean = !e.transposed && e.aux_mode == TC_AUX_NONE && e.mode == TC_EPI_STORE && !e.r1col && !e.mulx && !e.red_col && !e.rowv && !e.colv;
  if (HCA_ENV_INT("HCA_TC_LEAN", 1) == 0) lean = false;
  const bool lean_tanh = lean && e.act_tanh;
#define HCA_TC_CASE(SLOT, BNN, PP, CC, AMN, BMN, BKK)                                                            \
  if (BN == BNN && P == PP && combo == CC && BK == BKK) {                                                         \
    if (lean_tanh && BNN != 32) { fn = gemm_tc_kernel<BNN, PP, AMN, BMN, BKK, 1, 2>; slot = SLOT + 20; }          \
    else if (lean && BNN != 32) { fn = gemm_tc_kernel<BNN, PP, AMN, BMN, BKK, 1, 0>; slot = SLOT + 10; }          \
    else { fn = gemm_tc_kernel<BNN, PP, AMN, BMN, BKK, 1, 1>; slot = SLOT; }                                      \
  }
  HCA_TC_CASE(0, 128, 2, 0, false, false, 64) HCA_TC_CASE(1, 128, 2, 1, false, true, 64) HCA_TC_CASE(2, 128, 2, 3, true, true, 64)
  HCA_TC_CASE(3, 128, 3, 0, false, false, 64) HCA_TC_CASE(4, 128, 3, 1, false, true, 64) HCA_TC_CASE(5, 128, 3, 3, true, true, 64)
  HCA_TC_CASE(6, 32, 2, 0, false, false, 64) HCA_TC_CASE(7, 32, 2, 2, true, false, 64) HCA_TC_CASE(8, 128, 2, 3, true, true, 32)
#undef HCA_TC_CASE
  if (pair) {
    if (A.mn_major) { fn = gemm_tc_kernel<256, 2, true, true, 64, 2, 0>; slot = 28; }
    else if (lean_tanh) { fn = gemm_tc_kernel<256, 2, false, false, 64, 2, 2>; slot = 29; }
    else if (lean) { fn = gemm_tc_kernel<256, 2, false, false, 64, 2, 0>; slot = 19; }
    else { fn = gemm_tc_kernel<256, 2, false, false, 64, 2, 1>; slot = 9; }
  }
  HCA_CHECK_ARG(fn != nullptr, "gemm_tc: this (BN, P, layout) combination is not instantiated (BN=%d P=%d combo=%d)", BN, P, combo);
  static bool attr_set[64][30] = {};                 // function attributes are per device
  bool& attr_done = attr_set[current_device()][slot];
  if (!attr_done) {
    HCA_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
    attr_done = true;
  }
  if (pair) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    HCA_CUDA(cudaLaunchKernelEx(&cfg, fn, maps, p));
  } else {
    HCA_LAUNCH_K((fn), ctas, NUM_THREADS, smem, s, maps, p);
  }
  HCA_LAUNCHED();
  return 0;
}

}  // namespace hca
