// tcgen05 / TMEM / TMA split-bf16 GEMM (see gemm_tc.cuh for the scheme).  sm_100a only.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdlib>
#include "gemm_tc.cuh"

namespace hca {
namespace {

constexpr int BM = 128;               // UMMA M (cta_group::1): TMEM lane i <-> output row i
constexpr int BK = 64;                // bf16 elements per k-block = 128 bytes = one SWIZZLE_128B span
constexpr int UMMA_K = 16;            // bf16
constexpr int A_TILE_BYTES = BM * BK * 2;
constexpr int MAX_STAGES = 6;
constexpr int NUM_THREADS = 192;      // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2..5 epilogue
constexpr int SMEM_BUDGET = 200 * 1024;

struct TcParams {
  int M, N, K, stages, kb_total, kb_per_split, kb1, splitk;
  // epilogue
  float* D;
  int64_t ldd, d_sb;
  const float* bias; int64_t bias_sb;
  int act_tanh;
  const float* mulx; int64_t mulx_ld;
  int accumulate, atomic;
  int tma_store;            // epilogue goes through smem + TMA tile store (D 16-byte aligned, ldd % 4 == 0)
  int mode, aux_mode, aux_nbatch;
  const float* rowv; int64_t rowv_sb;
  const float* colv;
  const float* r1col; int64_t r1col_sb;
  float* red_row; int64_t red_row_sb;
  float* red_col;
  int a_nb, b_nb, a2_nb, b2_nb;   // batch entries of each operand (entry = z % nb)
  int dbg;                  // debug bit flags (HCA_TC_DBG): 1 = no TMA (MMA on garbage), 2 = sleepy epilogue wait
  long long* timeline;      // optional [ncta][64] clock64 stamps (debug / profiling), nullptr normally
  int timeline_ctas;
};

long long* g_timeline = nullptr;
int g_timeline_ctas = 0;

// ----------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a protocol bug must surface as a trapped kernel (an error the host sees), never as a hang.
__device__ __noinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("hiecoattn gemm_tc: mbarrier wait timed out (tag %d, block %d,%d,%d, thread %d)\n", tag, blockIdx.x, blockIdx.y,
             blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// TMA tile store / reduce-add from shared memory (bulk async group completion)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tm), "r"(src), "r"(c0),
               "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tm), "r"(src),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the 4 epilogue warps only

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// Warp-converged variants: every lane of the MMA warp executes the loop (so ptxas keeps the descriptors in
// uniform registers instead of electing + broadcasting them per instruction) and one elected lane issues.
__device__ __forceinline__ void umma_bf16_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// descriptors passed as 32-bit halves so that ptxas can build them on the uniform datapath
__device__ __forceinline__ void umma_bf16_elect32(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                  uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory matrix descriptor, SWIZZLE_128B (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4 |
//   [46,48) version = 1 | [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Descriptor of the k-th 16-wide K slice of a tile.
//   K-major tile  [rows][64 k] : rows are 128 B apart, 8-row groups 1024 B apart (SBO); a K slice is +32 B inside the row.
//   MN-major tile [64 k][64 mn] per 64-wide MN chunk (one TMA box, 8192 B): k rows 128 B apart, 8-row groups 1024 B apart
//                 (SBO), MN chunks 8192 B apart (LBO); a K slice of 16 rows is +2048 B.
__device__ __forceinline__ uint64_t tile_desc(uint32_t tile_addr, int mn_major, int kslice) {
  return mn_major ? umma_desc(tile_addr + kslice * 2048, 8192, 1024) : umma_desc(tile_addr + kslice * 32, 16, 1024);
}

__device__ __forceinline__ float tanh_acc(float x) { return tanhf(x); }

// ------------------------------------------------------------------------------------------------------ kernel
struct TcMaps {            // all TMA descriptors of one launch (operands 4-D: cols, rows, plane, batch; D / aux 3-D: N, M, batch)
  CUtensorMap A, B, A2, B2, D, AUX;
};

template <int BN, int P, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ TcMaps maps, const TcParams p) {
  constexpr int B_TILE_BYTES = BN * BK * 2;
  constexpr uint32_t stage_bytes = P * (A_TILE_BYTES + B_TILE_BYTES);
  constexpr int CHUNKS = BN / 32;
  constexpr uint32_t CHUNK_BYTES = BM * 128;           // one [128 rows][32 fp32] staging tile
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B tiles need 1024-byte alignment
  __shared__ __align__(8) uint64_t bars[2 * MAX_STAGES + 2];
  __shared__ uint32_t tmem_ptr_smem;
  __shared__ __align__(16) float bias_sm[BN];          // this tile's bias slice (per batch entry)
  __shared__ __align__(16) float colv_sm[BN];          // colv (ROWDOT / DZ) or the rank-1 column vector
  __shared__ float colred_sm[BN];                      // DZ: column partial sums of the four epilogue warps

  // warp index through a shuffle: provably warp-uniform for ptxas, so the role branches below are uniform control flow and
  // the MMA descriptors live in uniform registers (otherwise every tcgen05.mma pays an ELECT + VOTEU + 4x R2UR.BROADCAST
  // sequence, ~250 cycles per instruction, measured)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int z = blockIdx.z / p.splitk, ksplit = blockIdx.z - z * p.splitk;
  const int kb_begin = ksplit * p.kb_per_split;
  const int kb_end = min(p.kb_total, kb_begin + p.kb_per_split);
  const int num_kb = kb_end - kb_begin;

  auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
  auto empty_bar = [&](int s) { return smem_u32(&bars[MAX_STAGES + s]); };
  const uint32_t tmem_full_bar = smem_u32(&bars[2 * MAX_STAGES]);
  const uint32_t aux_bar = smem_u32(&bars[2 * MAX_STAGES + 1]);
  auto a_tile = [&](int s, int pl) { return smem_base + s * stage_bytes + pl * A_TILE_BYTES; };
  auto b_tile = [&](int s, int pl) { return smem_base + s * stage_bytes + P * A_TILE_BYTES + pl * B_TILE_BYTES; };

  const int cta_lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  long long* tl = (p.timeline && cta_lin < p.timeline_ctas) ? p.timeline + (size_t)cta_lin * 64 : nullptr;
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
    mbar_init(tmem_full_bar, 1);
    mbar_init(aux_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.A) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.B) : "memory");
    if (p.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.D) : "memory");
    if (p.aux_mode) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.AUX) : "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_ptr_smem), BN);       // BN fp32 accumulator columns (power of two >= 32)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_ptr_smem, 0);
  if (tl && threadIdx.x == 0) tl[1] = clock64();

  if (warp == 0) {
    // ===================================================================== TMA producer (one lane)
    if (lane == 0 && !(p.dbg & 1)) {
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < num_kb; ++it) {
        mbar_wait(empty_bar(s), ph ^ 1, 1);
        if (tl && it < 12) tl[8 + it] = clock64();          // producer: slot free, issuing TMA for k-block `it`
        mbar_expect_tx(full_bar(s), stage_bytes);
        const int kb = kb_begin + it;
        const bool second = kb >= p.kb1;                    // chained second operand pair
        const CUtensorMap* ma = second ? &maps.A2 : &maps.A;
        const CUtensorMap* mb = second ? &maps.B2 : &maps.B;
        const int za = z % (second ? p.a2_nb : p.a_nb), zb = z % (second ? p.b2_nb : p.b_nb);
        const int k0 = (second ? kb - p.kb1 : kb) * BK;
#pragma unroll
        for (int pl = 0; pl < P; ++pl) {
          if (!A_MN) {
            tma_load_4d(a_tile(s, pl), ma, full_bar(s), k0, m0, pl, za);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_4d(a_tile(s, pl) + c * 8192, ma, full_bar(s), m0 + c * 64, k0, pl, za);
          }
          if (!B_MN) {
            tma_load_4d(b_tile(s, pl), mb, full_bar(s), k0, n0, pl, zb);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_4d(b_tile(s, pl) + c * 8192, mb, full_bar(s), n0 + c * 64, k0, pl, zb);
          }
        }
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (whole warp converged, one lane issues)
    {
      // instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 [4,6), a=bf16 [7,10), b=bf16 [10,13),
      // a_major bit 15, b_major bit 16, N>>3 [17,23), M>>4 [24,29)
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                                 ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      // smem descriptor halves (cute::UMMA::SmemDescriptor): hi = SBO(1024 B) | version 1 | SWIZZLE_128B, constant;
      // lo = (addr >> 4) | (LBO >> 4) << 16.  K-major: LBO unused (16 B), K slice = +32 B.  MN-major: LBO = 8192 B
      // between 64-wide MN chunks, K slice of 16 rows = +2048 B.  Everything below is warp-uniform integer arithmetic,
      // so ptxas keeps it on the uniform datapath (no per-MMA ELECT / R2UR.BROADCAST sequence).
      constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      constexpr uint32_t a_lbo = (A_MN ? (8192u >> 4) : 1u) << 16, b_lbo = (B_MN ? (8192u >> 4) : 1u) << 16;
      constexpr uint32_t a_kstep = A_MN ? (2048u >> 4) : (32u >> 4), b_kstep = B_MN ? (2048u >> 4) : (32u >> 4);
      int s = 0;
      uint32_t ph = 0;
      uint32_t a0 = ((smem_base & 0x3FFFFu) >> 4) | a_lbo;
      uint32_t b0 = (((smem_base + P * A_TILE_BYTES) & 0x3FFFFu) >> 4) | b_lbo;
      for (int it = 0; it < num_kb; ++it) {
        if (!(p.dbg & 1)) mbar_wait(full_bar(s), ph, 2);
        tc_fence_after();
        if (tl && lane == 0 && it == 0) tl[2] = clock64();
        if (tl && lane == 0 && it < 12) tl[24 + it] = clock64();         // MMA warp: k-block `it` landed
        // shuffle-broadcast: tells ptxas the stage base is warp-uniform, so the descriptor variants below are
        // uniform-register adds of compile-time constants
        const uint32_t au = __shfl_sync(0xffffffffu, a0, 0), bu = __shfl_sync(0xffffffffu, b0, 0);
        const uint32_t first = __shfl_sync(0xffffffffu, it == 0 ? 0u : 1u, 0);
#pragma unroll
        for (int ks = 0; ks < BK / UMMA_K; ++ks) {
#pragma unroll
          for (int i = 0; i < P; ++i) {
#pragma unroll
            for (int j = 0; j < P - i; ++j) {
              umma_bf16_elect32(tmem_base, au + i * (A_TILE_BYTES >> 4) + ks * a_kstep, desc_hi,
                                bu + j * (B_TILE_BYTES >> 4) + ks * b_kstep, desc_hi, idesc, (ks | i | j) != 0 ? 1u : first);
            }
          }
        }
        umma_commit_elect(empty_bar(s));      // frees the smem stage once the MMAs above have read it
        if (tl && lane == 0 && it < 12) tl[40 + it] = clock64();         // MMA warp: k-block `it` issued + committed
        a0 += stage_bytes >> 4;
        b0 += stage_bytes >> 4;
        if (++s == p.stages) {
          s = 0;
          ph ^= 1;
          a0 -= p.stages * (stage_bytes >> 4);
          b0 -= p.stages * (stage_bytes >> 4);
        }
      }
      umma_commit_elect(tmem_full_bar);       // accumulator complete -> epilogue
      if (tl && lane == 0) tl[3] = clock64();
    }
  } else {
    // ===================================================================== epilogue: TMEM -> registers -> (smem -> TMA) global
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int r = q * 32 + lane;            // row inside the tile
    const int row = m0 + r;
    const bool row_ok = row < p.M;
    const int et = threadIdx.x - 64;        // 0..127
    {                                       // stage the per-column vectors while the mainloop runs
      const float* bias = p.bias ? p.bias + (int64_t)z * p.bias_sb : nullptr;
      const float* cvec = p.r1col ? p.r1col + (int64_t)z * p.r1col_sb : p.colv;
      for (int j = et; j < BN; j += 128) {
        const bool ok = n0 + j < p.N;
        bias_sm[j] = (bias && ok) ? __ldg(bias + n0 + j) : 0.f;
        colv_sm[j] = (cvec && ok) ? __ldg(cvec + n0 + j) : 0.f;
        colred_sm[j] = 0.f;
      }
      epi_barrier();
    }
    const float rv = (p.rowv && row_ok) ? __ldg(p.rowv + (int64_t)z * p.rowv_sb + row) : 0.f;
    if (p.dbg & 2) {
      while (!mbar_try_wait(tmem_full_bar, 0)) __nanosleep(500);
    } else {
      mbar_wait(tmem_full_bar, 0, 3);
    }
    tc_fence_after();
    if (tl && threadIdx.x == 64) tl[4] = clock64();
    const bool leader = (warp == 2 && lane == 0);
    const int nchunks = min(CHUNKS, (p.N - n0 + 31) / 32);           // uniform across the CTA
    // staging layout inside the (now idle) pipeline stages: [0, 2*CHUNK) store tiles, then one aux tile per chunk
    const uint32_t store_buf0 = smem_base, aux_buf0 = smem_base + 2 * CHUNK_BYTES;
    if (p.aux_mode) {                       // fetch the aux tile (pre-activation addend or (1 - x^2) factor) for all chunks at once
      if (leader) {
        mbar_expect_tx(aux_bar, (uint32_t)nchunks * CHUNK_BYTES);
        for (int c = 0; c < nchunks; ++c) tma_load_3d(aux_buf0 + c * CHUNK_BYTES, &maps.AUX, aux_bar, n0 + c * 32, m0, z % p.aux_nbatch);
      }
      mbar_wait(aux_bar, 0, 4);
    }
    float* drow = p.D ? p.D + (int64_t)z * p.d_sb + (int64_t)row * p.ldd : nullptr;
    const float* xrow = p.mulx ? p.mulx + (int64_t)row * p.mulx_ld : nullptr;
    float rowdot = 0.f;
    const bool do_store = (p.mode != TC_EPI_ROWDOT);
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
      const int col0 = n0 + c * 32;
      const uint32_t stage_buf = store_buf0 + (uint32_t)(c & 1) * CHUNK_BYTES;
      if (do_store && p.tma_store && c >= 2) {                    // staging buffer reuse: its previous TMA store must have read it
        if (leader) tma_store_wait_read<1>();
        epi_barrier();
      }
      uint32_t v[32];
      __syncwarp();                                               // tcgen05.ld is warp-collective
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
      float f[32];
      {                                                           // bias slice of this chunk: 8 x LDS.128
        const float4* b4 = reinterpret_cast<const float4*>(bias_sm + c * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = b4[j];
          f[4 * j] = b.x; f[4 * j + 1] = b.y; f[4 * j + 2] = b.z; f[4 * j + 3] = b.w;
        }
      }
      if (num_kb != 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] += __uint_as_float(v[j]);
      }
      float ax[32];
      if (p.aux_mode) {                                           // this thread's row of the aux tile (same swizzle as the stores)
        const uint32_t src = aux_buf0 + (uint32_t)c * CHUNK_BYTES + (uint32_t)r * 128u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(ax[4 * j]), "=f"(ax[4 * j + 1]), "=f"(ax[4 * j + 2]), "=f"(ax[4 * j + 3])
                       : "r"(src + (uint32_t)((j ^ (r & 7)) * 16)));
        }
        if (p.aux_mode == TC_AUX_ADD) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] += ax[j];
        }
      }
      if (p.act_tanh) {                                           // uniform branches: no predicated-off code on the common path
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = tanh_acc(f[j]);
      }
      if (p.r1col) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = fmaf(rv, colv_sm[c * 32 + j], f[j]);
      }
      if (p.aux_mode == TC_AUX_MUL_1MX2) {
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
      if (p.mode == TC_EPI_ROWDOT) {
#pragma unroll
        for (int j = 0; j < 32; ++j) rowdot = fmaf(f[j], colv_sm[c * 32 + j], rowdot);   // colv_sm is 0 beyond N
        continue;
      }
      if (p.mode == TC_EPI_DZ) {
        // column partials of h * rowv over this warp's 32 rows, then dz = rowv * colv * (1 - h^2)
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float part = warp_sum(f[j] * rv);
          if (lane == j) atomicAdd(&colred_sm[c * 32 + j], part);
          f[j] = rv * colv_sm[c * 32 + j] * (1.f - f[j] * f[j]);
        }
      }
      if (p.tma_store) {
        // Each thread owns one output row; the 32-column chunk is staged as a [128 rows][128 B] tile in the TMA 128-byte
        // swizzle (16-byte chunk j of row r at chunk j ^ (r & 7): conflict-free float4 stores) and handed to the TMA engine.
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t dst = stage_buf + (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) * 16);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(f[4 * j]), "f"(f[4 * j + 1]), "f"(f[4 * j + 2]),
                       "f"(f[4 * j + 3])
                       : "memory");
        }
        fence_proxy_async_smem();                                 // generic-proxy smem writes -> visible to the TMA engine
        epi_barrier();
        if (leader) {
          if (p.atomic || p.accumulate) tma_reduce_add_3d(&maps.D, stage_buf, col0, m0, z);
          else tma_store_3d(&maps.D, stage_buf, col0, m0, z);
          tma_store_commit();
        }
      } else if (row_ok) {                                        // unaligned output (e.g. K = 1001 logits): direct stores
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
    if (do_store && p.tma_store && leader) tma_store_wait_read<0>();   // smem must stay valid until the engine has read it
    if (p.mode == TC_EPI_ROWDOT && row_ok) atomicAdd(p.red_row + (int64_t)z * p.red_row_sb + row, rowdot);
    if (p.mode == TC_EPI_DZ) {
      epi_barrier();
      for (int j = et; j < BN; j += 128)
        if (n0 + j < p.N) atomicAdd(p.red_col + n0 + j, colred_sm[j]);
    }
    tc_fence_before();
    if (tl && threadIdx.x == 64) tl[5] = clock64();
  }
  __syncthreads();
  if (tl && threadIdx.x == 0) tl[6] = clock64();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// fp32 -> bf16 planes
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ src, int64_t ld, int64_t rows, int cols,
                                                           __nv_bfloat16* __restrict__ planes, int64_t ldp, int64_t plane_stride,
                                                           int P) {
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

// 4-D map over bf16 planes: dims (cols, rows, P, batch); box (64, box_rows, 1, 1); 128-byte swizzle; OOB -> zeros
int make_tmap(CUtensorMap* tm, const TcOperand& o, int P, int box_rows) {
  EncodeTiledFn enc = get_encoder();
  if (!enc) return set_err(HCA_ERR_CUDA, "gemm_tc: cuTensorMapEncodeTiled is not available from the driver");
  const int nb = o.nbatch > 0 ? o.nbatch : 1;
  cuuint64_t gdim[4] = {(cuuint64_t)o.cols, (cuuint64_t)o.rows, (cuuint64_t)P, (cuuint64_t)nb};
  cuuint64_t gstr[3] = {(cuuint64_t)o.ld * 2, (cuuint64_t)o.plane_stride * 2,
                        (cuuint64_t)(nb > 1 ? o.batch_stride : o.plane_stride * P) * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)o.planes, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_err(HCA_ERR_CUDA, "gemm_tc: cuTensorMapEncodeTiled failed (%d) cols=%d rows=%d ld=%lld P=%d nb=%d", (int)r, o.cols, o.rows,
                   (long long)o.ld, P, nb);
  return 0;
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

bool f32_tma_ok(const float* p, int64_t ld, int64_t batch_stride) {
  return ((ld & 3) == 0) && ((batch_stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0);
}

}  // namespace

bool tc_available() { return get_encoder() != nullptr; }

void tc_set_timeline(long long* buf, int nctas) {
  g_timeline = buf;
  g_timeline_ctas = nctas;
}

int launch_split_planes(const float* src, int64_t ld, int64_t rows, int cols, __nv_bfloat16* planes, int64_t ldp,
                        int64_t plane_stride, int P, cudaStream_t s) {
  HCA_CHECK_ARG(src && planes && rows > 0 && cols > 0 && P >= 1 && P <= 3 && (ldp % 8) == 0 && (plane_stride % 8) == 0,
                "split_planes: bad arguments");
  const int64_t total = rows * ((cols + 3) / 4);
  split_planes_kernel<<<ew_grid(total), 256, 0, s>>>(src, ld, rows, cols, planes, ldp, plane_stride, P);
  HCA_LAUNCHED();
  return 0;
}

int launch_gemm_tc(const TcOperand& A, const TcOperand& B, int P, int M, int N, int K, const TcEpilogue& e, int splitk,
                   cudaStream_t s, int batch, const TcOperand* A2, const TcOperand* B2, int K2) {
  constexpr int BN = 128;
  HCA_CHECK_ARG(P >= 1 && P <= 3 && M > 0 && N > 0 && K > 0 && splitk >= 1 && batch >= 1, "gemm_tc: bad sizes");
  HCA_CHECK_ARG(batch == 1 || splitk == 1, "gemm_tc: split-K is for un-batched products");
  HCA_CHECK_ARG((A2 == nullptr) == (B2 == nullptr), "gemm_tc: the chained operand pair needs both A2 and B2");
  const TcOperand* ops[4] = {&A, &B, A2, B2};
  for (const TcOperand* o : ops) {
    if (!o) continue;
    HCA_CHECK_ARG((o->ld % 8) == 0 && (o->plane_stride % 8) == 0 && (o->batch_stride % 8) == 0,
                  "gemm_tc: plane leading dimensions / strides must be multiples of 8 elements (TMA 16-byte strides)");
    HCA_CHECK_ARG((reinterpret_cast<uintptr_t>(o->planes) & 15) == 0, "gemm_tc: planes must be 16-byte aligned");
  }
  HCA_CHECK_ARG(!(A.mn_major && !B.mn_major), "gemm_tc: (MN-major A, K-major B) is not instantiated");
  if (A2) HCA_CHECK_ARG(A2->mn_major == A.mn_major && B2->mn_major == B.mn_major && K2 > 0, "gemm_tc: chained pair must share the layouts");
  TcMaps maps;
  HCA_TRY(make_tmap(&maps.A, A, P, A.mn_major ? 64 : BM));
  HCA_TRY(make_tmap(&maps.B, B, P, B.mn_major ? 64 : BN));
  if (A2) {
    HCA_TRY(make_tmap(&maps.A2, *A2, P, A2->mn_major ? 64 : BM));
    HCA_TRY(make_tmap(&maps.B2, *B2, P, B2->mn_major ? 64 : BN));
  } else {
    maps.A2 = maps.A;
    maps.B2 = maps.B;
  }
  const bool need_store = e.mode != TC_EPI_ROWDOT;
  const bool tma_store = need_store && f32_tma_ok(e.D, e.ldd, e.d_batch_stride);
  if (need_store) HCA_CHECK_ARG(e.D != nullptr, "gemm_tc: null output");
  if (tma_store) HCA_TRY(make_tmap_f32(&maps.D, e.D, e.ldd, e.d_batch_stride, batch, M, N));
  else maps.D = maps.A;                                                // unused placeholder
  if (e.aux_mode != TC_AUX_NONE) {
    HCA_CHECK_ARG(e.aux && f32_tma_ok(e.aux, e.aux_ld, e.aux_batch_stride), "gemm_tc: aux tile must have 16-byte aligned rows");
    HCA_TRY(make_tmap_f32(&maps.AUX, e.aux, e.aux_ld, e.aux_batch_stride, e.aux_nbatch, M, N));
  } else {
    maps.AUX = maps.A;
  }
  TcParams p;
  p.M = M; p.N = N; p.K = K;
  const int stage_bytes = P * (A_TILE_BYTES + BN * BK * 2);
  p.stages = SMEM_BUDGET / stage_bytes;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  { const char* ev = getenv("HCA_TC_STAGES"); if (ev && atoi(ev) >= 1 && atoi(ev) < p.stages) p.stages = atoi(ev); }
  HCA_CHECK_ARG(p.stages >= 2, "gemm_tc: tile does not fit two pipeline stages");
  HCA_CHECK_ARG((size_t)p.stages * stage_bytes >= (size_t)(2 + BN / 32) * BM * 128, "gemm_tc: staging tiles do not fit");
  p.kb1 = (K + BK - 1) / BK;
  p.kb_total = p.kb1 + (A2 ? (K2 + BK - 1) / BK : 0);
  if (splitk > p.kb_total) splitk = p.kb_total;
  p.kb_per_split = (p.kb_total + splitk - 1) / splitk;
  splitk = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;      // no empty split
  p.splitk = splitk;
  p.D = e.D; p.ldd = e.ldd; p.d_sb = e.d_batch_stride;
  p.bias = e.bias; p.bias_sb = e.bias_batch_stride;
  p.act_tanh = e.act_tanh; p.mulx = e.mulx; p.mulx_ld = e.mulx_ld;
  p.accumulate = e.accumulate;
  p.atomic = splitk > 1 ? 1 : 0;
  p.tma_store = tma_store ? 1 : 0;
  p.mode = e.mode; p.aux_mode = e.aux_mode; p.aux_nbatch = e.aux_nbatch > 0 ? e.aux_nbatch : 1;
  p.rowv = e.rowv; p.rowv_sb = e.rowv_batch_stride;
  p.colv = e.colv;
  p.r1col = e.r1col; p.r1col_sb = e.r1col_batch_stride;
  p.red_row = e.red_row; p.red_row_sb = e.red_row_batch_stride;
  p.red_col = e.red_col;
  p.a_nb = A.nbatch > 0 ? A.nbatch : 1; p.b_nb = B.nbatch > 0 ? B.nbatch : 1;
  p.a2_nb = A2 ? (A2->nbatch > 0 ? A2->nbatch : 1) : 1; p.b2_nb = B2 ? (B2->nbatch > 0 ? B2->nbatch : 1) : 1;
  HCA_CHECK_ARG(!(p.atomic && (e.bias || e.act_tanh || e.mulx || e.aux_mode || e.mode != TC_EPI_STORE || e.r1col)),
                "gemm_tc: split-K needs a linear epilogue");
  HCA_CHECK_ARG(e.mode != TC_EPI_ROWDOT || (e.colv && e.red_row), "gemm_tc: ROWDOT needs colv and red_row");
  HCA_CHECK_ARG(e.mode != TC_EPI_DZ || (e.colv && e.rowv && e.red_col), "gemm_tc: DZ needs rowv, colv and red_col");
  HCA_CHECK_ARG(!e.r1col || (e.rowv && e.mode == TC_EPI_STORE), "gemm_tc: the rank-1 term needs rowv and the STORE mode");
  { static int dbg = -1; if (dbg < 0) { const char* ev = getenv("HCA_TC_DBG"); dbg = ev ? atoi(ev) : 0; } p.dbg = dbg; }
  p.timeline = g_timeline;
  p.timeline_ctas = g_timeline_ctas;
  const size_t smem = (size_t)p.stages * stage_bytes + 1024;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, batch * splitk);
  HCA_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "gemm_tc: grid too large");
  typedef void (*KernelFn)(const TcMaps, const TcParams);
  KernelFn fn = nullptr;
  const int combo = (A.mn_major ? 2 : (B.mn_major ? 1 : 0));    // 0 = NT, 1 = NN, 2 = TN
#define HCA_TC_CASE(PP, CC, AMN, BMN) if (P == PP && combo == CC) fn = gemm_tc_kernel<BN, PP, AMN, BMN>;
  HCA_TC_CASE(1, 0, false, false) HCA_TC_CASE(1, 1, false, true) HCA_TC_CASE(1, 2, true, true)
  HCA_TC_CASE(2, 0, false, false) HCA_TC_CASE(2, 1, false, true) HCA_TC_CASE(2, 2, true, true)
  HCA_TC_CASE(3, 0, false, false) HCA_TC_CASE(3, 1, false, true) HCA_TC_CASE(3, 2, true, true)
#undef HCA_TC_CASE
  static bool attr_set[3][3] = {};
  if (!attr_set[P - 1][combo]) {
    HCA_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET + 2048));
    attr_set[P - 1][combo] = true;
  }
  fn<<<grid, NUM_THREADS, smem, s>>>(maps, p);
  HCA_LAUNCHED();
  return 0;
}

}  // namespace hca
