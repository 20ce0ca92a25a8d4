// Sentence-level LSTM over variable-length questions (replaces pack_padded_sequence -> nn.LSTM(E, H) -> pad_packed_sequence,
// reference model.py:269,287-296).  Gate order i, f, g, o; h0 = c0 = 0; every sequence stops at its own length; rows
// t >= len of the output are zero.
//
//   z_t = x_t W_ih^T + b_ih + b_hh + h_{t-1} W_hh^T        i,f,o = sigmoid(z), g = tanh(z)
//   c_t = f c_{t-1} + i g                                  h_t = o tanh(c_t)
//
// The input projection of all (b, t) is one tcgen05 GEMM (gemm_tc.cuh).  The recurrence -- T dependent steps of a
// [B, H] x [H, 4H] product -- runs in ONE persistent kernel instead of T library GEMM + cell launches:
//   * CTA (mi, ni) owns 128 batch rows x 16 hidden units.  Its slice of W_hh (the 64 gate rows of its units, as bf16 hi/lo
//     planes, 128 KB at H = 512) is loaded into shared memory ONCE and stays there for all steps; the cell state c of its
//     (row, unit) pairs lives in registers.
//   * per step the CTAs of one row tile exchange h_{t-1} through global memory (bf16 hi/lo planes, written by the cell
//     epilogue, streamed back in by TMA as the A operand); a release/acquire counter per (row tile, step) orders the
//     exchange -- rows of different tiles never wait for each other.
//   * batch on the UMMA M axis, the 4 gates of a unit on 4 adjacent accumulator columns (W rows are permuted to
//     [unit][gate] order), so one epilogue thread holds all four gates of its (row, unit) and no cross-thread exchange
//     is needed; tcgen05.mma kind::f16 on bf16x2 operand planes (3 MMAs per k-step), fp32 accumulation in TMEM.
// Backward runs the same skeleton in reverse time: the streamed operand is the gate gradient of step t+1 (written as bf16
// planes by the cell-backward epilogue -- the very array the weight-gradient GEMMs consume afterwards), the resident operand
// the 16 columns of W_hh belonging to the CTA's units.  dW_ih, dW_hh and dx are three big tcgen05 GEMMs over all (b, t).
#include <cuda.h>
#include <cuda_bf16.h>
#include <algorithm>
#include "common.cuh"
#include "gemm_tc.cuh"
#include "tc_ptx.cuh"
#include "util_kernels.cuh"

#ifndef HCA_TC_TIMELINE
#define HCA_TC_TIMELINE 0
#endif

namespace hca {
namespace {
using namespace ptx;

constexpr int L_BM = 128;                         // batch rows per CTA (UMMA M)
constexpr int L_BK = 64;                          // bf16 per k-block row = 128 B = one SWIZZLE_128B span
constexpr int L_UNITS = 16;                       // hidden units per CTA
constexpr int L_MAX_STAGES = 16;                  // ring slots of the streamed operand (as many as fit: short row tiles -> many small slots)
constexpr uint32_t L_RING_BYTES = 96 * 1024;
constexpr int L_THREADS = 64 + 256;               // warp 0 TMA, warp 1 MMA, warps 2..9 cell epilogue (2 groups x 4 TMEM quadrants)
constexpr int L_MAX_SMEM = 227 * 1024 - 2048;

constexpr int L_CL = 8;                           // cluster size of the multicast variant: 8 CTAs of one row tile share the streamed operand
constexpr int L_TMAX = 128;                       // longest sequence for which the per-step row trimming is tabulated
struct LstmMaps {
  CUtensorMap A[3];   // streamed operand planes, 5-D (64 cols, b, k-block, t, plane), boxes (64, box_rows[i], nkb, 1, 1): full, half, quarter tile
  CUtensorMap W;      // resident operand planes, 4-D (cols, rows, plane, 1), box (64, BN, 1, 1)
  CUtensorMap AS;     // "stacked" view of the streamed operand, dims (64 cols, b, plane, k-block, t), box (64, 64, 2, 1, 1): one k-block
                      // lands as a 128-row tile whose rows 0..63 are the hi plane and 64..127 the lo plane of the same 64 batch rows
  CUtensorMap A1[3];  // the streamed operand again with ONE k-block per box (64, box_rows[i], 1, 1, 1): the pieces of a ring slot that the
                      // CTAs of a cluster load for each other in multicast mode
};

struct LstmParams {
  int B, T, H;
  int tile0;              // first row tile of this launch
  int rpt;                // batch rows per row tile (<= 128, multiple of 8): tile i owns rows [i * rpt, (i + 1) * rpt)
  int box_rows[3];        // rows of the three TMA boxes of the streamed operand
  int stages;             // ring slots in use
  uint32_t plane_bytes;   // bytes of one plane of one ring slot (rpt rows x 128 B); a slot = hi plane + lo plane
  uint32_t lo_off;        // offset of the lo-plane tiles inside a slot (= nkb * plane_bytes)
  uint32_t stage_bytes;   // bytes of one ring slot: nkb k-blocks x (hi tile + lo tile)
  int nkb;                // k-blocks per ring slot: ONE TMA box per plane brings nkb k-blocks (the issue rate of small boxes,
                          // ~450 cycles per slot iteration, bounded the recurrence when every k-block was its own slot)
  int nslots;             // slots per round = ceil(kbn / nkb)
  int tiles_n;            // H / 16
  int kbn;                // k-blocks of the recurrent contraction: fwd ceil(H / 64), bwd ceil(4H / 64)
  int K;                  // contraction length: fwd H, bwd 4H
  const int64_t* lens;
  int* counters;          // [row tiles][T]: CTAs of the row tile that have published step t
  float* act;             // [B][T][4H], gate columns in [unit][gate] order.  fwd: in = x-projection + biases, out = activations
  float* c;               // [B][T][H] cell states
  float* out;             // [B][T][H] (fwd)
  __nv_bfloat16* hp;      // h planes [2][B][T][H]: slot t holds h_{t-1} (slot 0 = 0)
  int64_t hp_ps;
  const float* dout;      // [B][T][H] (bwd)
  __nv_bfloat16* dgp;     // gate-gradient planes [2][B][T][4H] ([unit][gate] column order) (bwd)
  int64_t dgp_ps;
  float* dbias;           // [4H] in [unit][gate] order, accumulated atomically (bwd)
  long long* timeline;    // debug: clock64 stamps of CTA 0 ([round][8]: counter seen, loads issued, first operand landed, MMAs issued,
                          // accumulator seen by the cell threads, cell math done, barrier passed, step published), else nullptr
};
long long* g_lstm_timeline = nullptr;
cudaEvent_t g_rec_events[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};     // [fwd / bwd][start / stop]: see hca_debug_lstm_events

__device__ __forceinline__ float sigmoid_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
  return r;
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// multicast variant: the box lands at the same shared-memory offset of every CTA in `mask`, and each of them gets the bytes
// signalled on ITS barrier at the same offset
__device__ __forceinline__ void tma_load_5d_mc(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3, int c4,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5, %6, %7}], [%2], %8;" ::"r"(dst),
      "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "h"(mask)
      : "memory");
}
// MMA completion -> the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t e;
  asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\tselp.u32 %0, 1, 0, pe;\n\t}" : "=r"(e));
  return e;
}
// tcgen05.mma predicated by an integer flag (elected lane && k-slice holds data), constant descriptor halves as immediates
template <uint32_t DESC_HI, uint32_t IDESC>
__device__ __forceinline__ void umma_bf16_imm_pred(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t accumulate, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "setp.ne.b32 pe, %4, 0;\n\t"
      "setp.ne.b32 p, %3, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %6, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(accumulate), "r"(issue), "n"(DESC_HI), "n"(IDESC)
      : "memory");
}
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void cell_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// 8 fp32 -> 8 bf16 hi + 8 bf16 lo (x = hi + lo to ~2^-17), as two 16-byte vectors
__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(x[2 * k], x[2 * k + 1]);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(x[2 * k] - __low2float(hh), x[2 * k + 1] - __high2float(hh));
    h[k] = *reinterpret_cast<const uint32_t*>(&hh);
    l[k] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// MC = true (launched as clusters of L_CL CTAs = L_CL unit blocks of ONE row tile): the streamed operand of a step is the same for
// all CTAs of a row tile, so every CTA loads 1 / L_CL of each ring slot and multicasts it to the whole cluster.  A slot is free once
// ALL CTAs of the cluster have read it (their MMA completions are multicast to everybody's empty barrier, count L_CL); every CTA still
// arms its own full barrier with the bytes of the whole slot.  Per SM this divides the TMA requests by L_CL and the L2 -> SM traffic of
// the recurrence by the cluster's deduplication (the per-step stream of the [rows x 4H] gate gradients bounded the backward kernel).
//
// STK = true (backward, row tiles of at most 64 rows): the hi and lo planes of the streamed operand are stacked along M -- one 128-row
// tile per k-block (AS map) -- so ONE MMA of width 2 BN per k-step yields x_hi.[W_hi | W_lo] in lanes 0..63 and x_lo.[W_hi | W_lo] in lanes
// 64..127, instead of two MMAs.  The backward recurrence issues 4H / 16 k-steps per step of tiny MMAs (N = 32 / 16) and is bound by their
// issue rate (~47 clock ticks per instruction from the single issuing thread, measured), so halving the instruction count is what counts;
// the x_lo.W_hi block is handed from the warps of lanes 64..127 to the owners of the rows through shared memory.
template <bool BWD, bool MC, bool STK = false>
__global__ void __launch_bounds__(L_THREADS, 1) lstm_rec_kernel(const __grid_constant__ LstmMaps maps, const LstmParams p) {
  static_assert(!STK || (BWD && !MC), "the stacked variant is the backward kernel without multicast");
  pdl_enter();
  constexpr int BN = BWD ? L_UNITS : 4 * L_UNITS;         // accumulator columns: dh of 16 units / 4 gates of 16 units
  constexpr uint32_t W_KB_PLANE = BN * L_BK * 2;          // one plane of one resident k-block
  constexpr uint32_t TMEM_COLS = BWD ? 32 : 128;             // [x.W_hi | x.W_lo] column blocks
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bars[2 * L_MAX_STAGES + 2];
  __shared__ uint32_t tmem_ptr_smem;
  __shared__ int s_maxlen;
  __shared__ int s_nact[L_TMAX];     // rows of this tile still running at step t (1 + the last row with len > t)
  __shared__ __align__(16) float s_xch[STK ? 2 * 64 * 8 : 4];   // stacked variant: x_lo.W_hi blocks, [unit half][row][8 units]
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
  auto empty_bar = [&](int s) { return smem_u32(&bars[L_MAX_STAGES + s]); };
  const uint32_t w_bar = smem_u32(&bars[2 * L_MAX_STAGES]), tmem_full = smem_u32(&bars[2 * L_MAX_STAGES + 1]);
  const uint32_t stage_bytes = p.stage_bytes;

  const int mi = blockIdx.x / p.tiles_n, ni = blockIdx.x - mi * p.tiles_n;
  const int m0 = (p.tile0 + mi) * p.rpt;
  int* const counters = p.counters + (int64_t)(p.tile0 + mi) * p.T;
  // (clock64 stamps for profiles/timeline_lstm.py: compiled in only with -DHCA_TC_TIMELINE=1, see build.py)
  long long* const tl = (HCA_TC_TIMELINE && blockIdx.x == 0) ? p.timeline : nullptr;
  const uint32_t w_base = smem_base;
  const uint32_t ring_base = smem_base + (uint32_t)p.kbn * 2u * W_KB_PLANE;

  if (threadIdx.x == 0) {
    s_maxlen = 0;
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), MC ? L_CL : 1);
    }
    mbar_init(w_bar, 1);
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.A[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.W) : "memory");
  }
  for (int i = threadIdx.x; i < L_TMAX; i += blockDim.x) s_nact[i] = 0;
  if (warp == 1) tmem_alloc(smem_u32(&tmem_ptr_smem), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  if constexpr (MC) cluster_sync_all();      // every CTA's barriers exist before a peer's multicast or commit can reach them
  tc_fence_after();
  // steps this row tile needs: the longest sequence among its rows (rows are independent, so other tiles may run longer)
  if ((int)threadIdx.x < p.rpt) {
    const int b = m0 + (int)threadIdx.x;
    if (b < p.B) {
      const int64_t l64 = p.lens[b];
      const int l = (int)(l64 < 0 ? 0 : (l64 > p.T ? p.T : l64));
      atomicMax(&s_maxlen, l);
      for (int t = 0; t < l && t < L_TMAX; ++t) atomicMax(&s_nact[t], (int)threadIdx.x + 1);
    }
  }
  __syncthreads();
  const int steps = s_maxlen;
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_ptr_smem, 0);
  const int rounds = steps > 0 ? steps - 1 : 0;             // recurrent products: every step but the first one in time order

  if (warp == 0) {
    // ============================================================ TMA producer (one elected lane, inline waits: uniform datapath)
    if (elect_one_sync()) {
      mbar_expect_tx(w_bar, (uint32_t)p.kbn * 2u * W_KB_PLANE);
      for (int kb = 0; kb < p.kbn; ++kb)
        for (int pl = 0; pl < 2; ++pl)
          tma_load_4d(w_base + (uint32_t)(kb * 2 + pl) * W_KB_PLANE, &maps.W, w_bar, kb * L_BK, ni * BN, pl, 0);
      int s = 0;
      uint32_t ph = 0;
      for (int n = 0; n < rounds; ++n) {
        // forward round n computes step t = n + 1 from h_n (slot n + 1, published at step n);
        // backward round n computes step t = steps - 2 - n from the gate gradients of step t + 1 (slot t + 1)
        const int dep = BWD ? steps - 1 - n : n;
        const int slot = BWD ? dep : n + 1;
        const int t_cur = BWD ? steps - 2 - n : n + 1;            // the step this round computes
        // rows of the tile that are still inside their sequence at that step: the others need no operand rows (their
        // accumulator rows are never used), so the smallest of the three boxes that covers the running rows is loaded
        const int nrun = (p.T <= L_TMAX) ? s_nact[t_cur] : p.rpt;
        const int bi = nrun <= p.box_rows[2] ? 2 : (nrun <= p.box_rows[1] ? 1 : 0);
        const uint32_t plane_bytes = (uint32_t)p.box_rows[bi] * L_BK * 2;      // one k-block of one plane in the chosen box
        const int* cnt = counters + dep;
        if (ld_acquire(cnt) < p.tiles_n) {           // bounded spin: a protocol bug must trap, never hang the device
          const long long t0 = clock64();
          while (ld_acquire(cnt) < p.tiles_n) {
            __nanosleep(20);
            if (clock64() - t0 > 4000000000LL) {
              printf("hiecoattn lstm: step counter wait timed out (block %d, round %d)\n", blockIdx.x, n);
              __trap();
            }
          }
        }
        fence_proxy_async_global();                  // peers wrote through the generic proxy; TMA reads through the async proxy
        if (tl && n < 7) tl[n * 8 + 0] = clock64();
        for (int sl = 0; sl < p.nslots; ++sl) {
          mbar_spin(empty_bar(s), ph ^ 1);
          mbar_expect_tx(full_bar(s), STK ? (uint32_t)p.nkb * 16384u : 2u * (uint32_t)p.nkb * plane_bytes);   // (k-blocks beyond K arrive as fill: full box bytes)
          const uint32_t dst = ring_base + (uint32_t)s * stage_bytes;
          if constexpr (STK) {
            // (the expect_tx above counted 2 nkb plane tiles of the chosen row box; the stacked box is nkb tiles of 128 rows)
            tma_load_5d(dst, &maps.AS, full_bar(s), 0, m0, 0, sl * p.nkb, slot);
          } else if constexpr (MC) {
            // this CTA's share of the slot: pieces (plane, k-block) cr, cr + L_CL, ... -- one box each, multicast to the cluster
            const int cr = (int)cluster_ctarank();
            for (int piece = cr; piece < 2 * p.nkb; piece += L_CL) {
              const int pl = piece / p.nkb, j = piece - pl * p.nkb;
              tma_load_5d_mc(dst + (uint32_t)pl * p.lo_off + (uint32_t)j * plane_bytes, &maps.A1[bi], full_bar(s), 0, m0, sl * p.nkb + j, slot, pl,
                             (uint16_t)((1u << L_CL) - 1u));
            }
          } else {
            tma_load_5d(dst, &maps.A[bi], full_bar(s), 0, m0, sl * p.nkb, slot, 0);
            tma_load_5d(dst + p.lo_off, &maps.A[bi], full_bar(s), 0, m0, sl * p.nkb, slot, 1);
          }
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        if (tl && n < 7) tl[n * 8 + 1] = clock64();
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(L_BM >> 4) << 24);
    constexpr uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)(L_BM >> 4) << 24);
    constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    // The whole loop runs in ONE elected thread with inline waits: inside such a region every value is trivially warp-uniform, so
    // ptxas builds the descriptors with uniform-datapath adds and the UTCHMMAs issue back to back (the per-instruction elect /
    // predicate forms cost ~75 cycles per MMA, which bounded the recurrence: its MMAs are small and many).
    if (elect_one_sync()) {
      mbar_spin(w_bar, 0);
      tc_fence_after();
      int s = 0;
      uint32_t ph = 0;
      const uint32_t d_tmem = *reinterpret_cast<volatile uint32_t*>(&tmem_ptr_smem);
      const uint32_t a_ring = ((ring_base & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t w_res = ((w_base & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t a_lo = p.lo_off >> 4;
      for (int n = 0; n < rounds; ++n) {
        // the tiles of a slot are packed with the pitch of the box that was loaded (the producer picks the same box: both read s_nact)
        const int t_cur = BWD ? steps - 2 - n : n + 1;
        const int nrun = (p.T <= L_TMAX) ? s_nact[t_cur] : p.rpt;
        const int bi = nrun <= p.box_rows[2] ? 2 : (nrun <= p.box_rows[1] ? 1 : 0);
        const uint32_t tile16 = STK ? (16384u >> 4) : (((uint32_t)p.box_rows[bi] * L_BK * 2) >> 4);
        for (int sl = 0; sl < p.nslots; ++sl) {
          mbar_spin(full_bar(s), ph);
          tc_fence_after();
          if (tl && n < 7 && sl == 0) tl[n * 8 + 2] = clock64();
          const uint32_t a_slot = a_ring + (uint32_t)s * (stage_bytes >> 4);
          for (int j = 0; j < p.nkb; ++j) {
            const int kb = sl * p.nkb + j;
            const uint32_t au = a_slot + (uint32_t)j * tile16;
            const uint32_t bu = w_res + (uint32_t)kb * (2u * W_KB_PLANE >> 4);
            const int nks = max(0, min(L_BK / 16, (p.K - kb * L_BK + 15) / 16));
#pragma unroll
            for (int ks = 0; ks < L_BK / 16; ++ks) {
              if (ks < nks) {
                // The hi and lo planes of the resident slice lie back to back (BN + BN rows), so ONE MMA of width 2 BN gives
                // x_hi.W_hi (columns 0..BN-1) and x_hi.W_lo (columns BN..2BN-1); a second MMA of width BN adds x_lo.W_hi.
                umma_bf16_one<desc_hi, idesc2>(d_tmem, au + ks * 2, bu + ks * 2, (kb | ks) != 0 ? 1u : 0u);
                if constexpr (!STK) umma_bf16_one<desc_hi, idesc>(d_tmem, au + a_lo + ks * 2, bu + ks * 2, 1u);
              }
            }
          }
          if constexpr (MC) umma_commit_mc(empty_bar(s), (uint16_t)((1u << L_CL) - 1u));   // the slot is free when the whole cluster has read it
          else umma_commit(empty_bar(s));
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        if (tl && n < 7) tl[n * 8 + 3] = clock64();
        umma_commit(tmem_full);
      }
    }
  } else {
    // ============================================================ cell epilogue: thread = (batch row, 8 hidden units)
    const int eg = (warp - 2) >> 2;                 // which half of the CTA's 16 units
    const int q = warp & 3;                         // TMEM lane quadrant
    const int r = q * 32 + lane;
    const int b = m0 + r;
    const bool row_ok = r < p.rpt && b < p.B;
    int len_b = 0;
    if (row_ok) {
      const int64_t l = p.lens[b];
      len_b = (int)(l < 0 ? 0 : (l > p.T ? p.T : l));
    }
    const bool leader = (threadIdx.x == 64);
    const int H = p.H, H4 = 4 * p.H;
    const int u0 = ni * L_UNITS + eg * 8;           // first of this thread's 8 units
    const int g0 = ni * 4 * L_UNITS + eg * 32;      // first of its 32 gate columns ([unit][gate] order)
    const int64_t bt0 = (int64_t)b * p.T;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    if constexpr (!BWD) {
      float cst[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) cst[u] = 0.f;
      for (int t = 0; t < steps; ++t) {
        float z[32];
        if (row_ok) {                                // x-projection + biases of this step (independent of the recurrence)
          const float4* gx = reinterpret_cast<const float4*>(p.act + (bt0 + t) * H4 + g0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 v = __ldg(gx + j);
            z[4 * j] = v.x; z[4 * j + 1] = v.y; z[4 * j + 2] = v.z; z[4 * j + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) z[j] = 0.f;
        }
        if (t > 0) {
          mbar_wait(tmem_full, (uint32_t)((t - 1) & 1), 14);
          tc_fence_after();
          if (tl && leader && t - 1 < 7) tl[(t - 1) * 8 + 4] = clock64();
          uint32_t v[32], v2[32];                   // x.W_hi block and x.W_lo block of this thread's 32 gate columns
          __syncwarp();
          tmem_ld32(lane_addr + (uint32_t)(eg * 32), v);
          tmem_ld32(lane_addr + (uint32_t)(4 * L_UNITS + eg * 32), v2);
          tc_fence_before();
          float acc[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(v[j]) + __uint_as_float(v2[j]);
#pragma unroll
          for (int j = 0; j < 32; ++j) z[j] += acc[j];
        }
        float h[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float ig = sigmoid_fast(z[4 * u]), fg = sigmoid_fast(z[4 * u + 1]), gg = tanh_fast(z[4 * u + 2]),
                      og = sigmoid_fast(z[4 * u + 3]);
          cst[u] = fmaf(fg, cst[u], ig * gg);
          h[u] = og * tanh_fast(cst[u]);
          z[4 * u] = ig; z[4 * u + 1] = fg; z[4 * u + 2] = gg; z[4 * u + 3] = og;
        }
        // publish h_t (slot t + 1) first: it is on the critical path of every CTA of this row tile
        if (t < len_b && t + 1 < p.T) {                 // (rows past their end keep the zero slot: their h is never read)
          uint4 hi, lo;
          split8(h, hi, lo);
          __nv_bfloat16* dst = p.hp + (bt0 + t + 1) * H + u0;
          *reinterpret_cast<uint4*>(dst) = hi;
          *reinterpret_cast<uint4*>(dst + p.hp_ps) = lo;
        }
        if (tl && leader && t >= 1 && t - 1 < 7) tl[(t - 1) * 8 + 5] = clock64();
        fence_proxy_async_global();
        cell_barrier();
        if (tl && leader && t >= 1 && t - 1 < 7) tl[(t - 1) * 8 + 6] = clock64();
        if (leader) {
          red_release_add(counters + t, 1);      // release at gpu scope: covers the CTA's writes ordered before it by the barrier
          if (tl && t >= 1 && t - 1 < 7) tl[(t - 1) * 8 + 7] = clock64();
        }
        if (row_ok) {                                // saved for backward + the module output
          float4* a4 = reinterpret_cast<float4*>(p.act + (bt0 + t) * H4 + g0);
#pragma unroll
          for (int j = 0; j < 8; ++j) a4[j] = make_float4(z[4 * j], z[4 * j + 1], z[4 * j + 2], z[4 * j + 3]);
          float4* c4 = reinterpret_cast<float4*>(p.c + (bt0 + t) * H + u0);
          c4[0] = make_float4(cst[0], cst[1], cst[2], cst[3]);
          c4[1] = make_float4(cst[4], cst[5], cst[6], cst[7]);
          float4* o4 = reinterpret_cast<float4*>(p.out + (bt0 + t) * H + u0);
          const bool valid = t < len_b;
          o4[0] = valid ? make_float4(h[0], h[1], h[2], h[3]) : make_float4(0.f, 0.f, 0.f, 0.f);
          o4[1] = valid ? make_float4(h[4], h[5], h[6], h[7]) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      if (row_ok) {                                  // steps nobody in this row tile reaches: zero output rows
        for (int t = steps; t < p.T; ++t) {
          float4* o4 = reinterpret_cast<float4*>(p.out + (bt0 + t) * H + u0);
          o4[0] = make_float4(0.f, 0.f, 0.f, 0.f);
          o4[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    } else {
      float dcn[8], dbacc[32];
#pragma unroll
      for (int u = 0; u < 8; ++u) dcn[u] = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) dbacc[j] = 0.f;
      for (int t = steps - 1; t >= 0; --t) {
        const bool valid = row_ok && t < len_b;
        float a[32], ct[8], cp[8], dh[8];
        if (valid) {
          const float4* a4 = reinterpret_cast<const float4*>(p.act + (bt0 + t) * H4 + g0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 v = __ldg(a4 + j);
            a[4 * j] = v.x; a[4 * j + 1] = v.y; a[4 * j + 2] = v.z; a[4 * j + 3] = v.w;
          }
          const float4* c4 = reinterpret_cast<const float4*>(p.c + (bt0 + t) * H + u0);
          const float4 c0 = __ldg(c4), c1 = __ldg(c4 + 1);
          ct[0] = c0.x; ct[1] = c0.y; ct[2] = c0.z; ct[3] = c0.w; ct[4] = c1.x; ct[5] = c1.y; ct[6] = c1.z; ct[7] = c1.w;
          if (t > 0) {
            const float4 p0 = __ldg(c4 - H / 4), p1 = __ldg(c4 - H / 4 + 1);
            cp[0] = p0.x; cp[1] = p0.y; cp[2] = p0.z; cp[3] = p0.w; cp[4] = p1.x; cp[5] = p1.y; cp[6] = p1.z; cp[7] = p1.w;
          } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) cp[u] = 0.f;
          }
          const float4* d4 = reinterpret_cast<const float4*>(p.dout + (bt0 + t) * H + u0);
          const float4 d0 = __ldg(d4), d1 = __ldg(d4 + 1);
          dh[0] = d0.x; dh[1] = d0.y; dh[2] = d0.z; dh[3] = d0.w; dh[4] = d1.x; dh[5] = d1.y; dh[6] = d1.z; dh[7] = d1.w;
        }
        if (t < steps - 1) {                         // + W_hh^T dz_{t+1}
          mbar_wait(tmem_full, (uint32_t)((steps - 2 - t) & 1), 15);
          tc_fence_after();
          if (tl && leader && steps - 2 - t < 7) tl[(steps - 2 - t) * 8 + 4] = clock64();
          uint32_t v[8], v2[8];                    // columns 0..15: x.W_hi ; columns 16..31: x.W_lo
          __syncwarp();
          tmem_ld8(lane_addr + (uint32_t)(eg * 8), v);
          float acc[8];
          if constexpr (STK) {
            // lanes 0..63: x_hi.[W_hi | W_lo] of row r ; lanes 64..127: x_lo.[W_hi | W_lo] of row r - 64 (only its W_hi block is used)
            if (q >= 2) {
              float4* x4 = reinterpret_cast<float4*>(s_xch + ((eg * 64 + (r - 64)) * 8));
              x4[0] = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
              x4[1] = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
            } else {
              tmem_ld8(lane_addr + (uint32_t)(L_UNITS + eg * 8), v2);
            }
            tc_fence_before();
            cell_barrier();
            if (q < 2) {
              const float4* x4 = reinterpret_cast<const float4*>(s_xch + ((eg * 64 + r) * 8));
              const float4 a0 = x4[0], a1 = x4[1];
              const float lo[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
              for (int u = 0; u < 8; ++u) acc[u] = __uint_as_float(v[u]) + __uint_as_float(v2[u]) + lo[u];
            } else {
#pragma unroll
              for (int u = 0; u < 8; ++u) acc[u] = 0.f;
            }
          } else {
            tmem_ld8(lane_addr + (uint32_t)(L_UNITS + eg * 8), v2);
            tc_fence_before();
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[u] = __uint_as_float(v[u]) + __uint_as_float(v2[u]);
          }
          if (valid) {
#pragma unroll
            for (int u = 0; u < 8; ++u) dh[u] += acc[u];
          }
        }
        float dz[32];
        if (valid) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float ig = a[4 * u], fg = a[4 * u + 1], gg = a[4 * u + 2], og = a[4 * u + 3];
            const float tc = tanh_fast(ct[u]);
            const float dc = fmaf(dh[u] * og, 1.f - tc * tc, dcn[u]);
            dz[4 * u] = dc * gg * ig * (1.f - ig);
            dz[4 * u + 1] = dc * cp[u] * fg * (1.f - fg);
            dz[4 * u + 2] = dc * ig * (1.f - gg * gg);
            dz[4 * u + 3] = dh[u] * tc * og * (1.f - og);
            dcn[u] = dc * fg;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) dz[j] = 0.f;
#pragma unroll
          for (int u = 0; u < 8; ++u) dcn[u] = 0.f;
        }
        if (row_ok) {                                // gate gradients of step t: operand of step t - 1 and of the weight-gradient GEMMs
          __nv_bfloat16* dst = p.dgp + (bt0 + t) * H4 + g0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float x[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = dz[8 * j + k];
            uint4 hi, lo;
            split8(x, hi, lo);
            *reinterpret_cast<uint4*>(dst + 8 * j) = hi;
            *reinterpret_cast<uint4*>(dst + p.dgp_ps + 8 * j) = lo;
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) dbacc[j] += dz[j];
        const int rn = steps - 2 - t;
        if (tl && leader && rn >= 0 && rn < 7) tl[rn * 8 + 5] = clock64();
        fence_proxy_async_global();
        cell_barrier();
        if (tl && leader && rn >= 0 && rn < 7) tl[rn * 8 + 6] = clock64();
        if (leader) {
          red_release_add(counters + t, 1);      // release at gpu scope: covers the CTA's writes ordered before it by the barrier
          if (tl && rn >= 0 && rn < 7) tl[rn * 8 + 7] = clock64();
        }
      }
      // bias gradient: column sums over this warp's 32 rows, then one atomic per column
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int j = 0; j < off; ++j) {
          const float send = up ? dbacc[j] : dbacc[j + off];
          const float keep = up ? dbacc[j + off] : dbacc[j];
          dbacc[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      if (steps > 0) atomicAdd(p.dbias + g0 + lane, dbacc[0]);
    }
    tc_fence_before();
  }
  __syncthreads();
  if constexpr (MC) cluster_sync_all();      // no peer multicasts into, or signals, a CTA that has left
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- weight layout helpers.  Gate row j' = 64 * ni + 4 * u + g of the permuted matrices is row g * H + 16 * ni + u of the
// PyTorch [4H, .] layout (gates i, f, g, o stacked along rows).
__device__ __forceinline__ int perm_row(int jp, int H) {
  const int ni = jp >> 6, rem = jp & 63;
  return (rem & 3) * H + ni * L_UNITS + (rem >> 2);
}
// planes [2][4H][cols] <- W[perm][cols]
__global__ void __launch_bounds__(256) lstm_split_perm_kernel(const float* __restrict__ W, int H, int cols, __nv_bfloat16* __restrict__ planes,
                                                              int64_t ps) {
  pdl_enter();
  const int64_t total = (int64_t)4 * H * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int jp = (int)(i / cols), c = (int)(i - (int64_t)jp * cols);
    const float x = W[(int64_t)perm_row(jp, H) * cols + c];
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    planes[i] = h;
    planes[ps + i] = __float2bfloat16_rn(x - __bfloat162float(h));
  }
}
// planes [2][H][4H] <- W_hh[perm(j')][k] at [k][j']   (the resident operand of the backward recurrence)
__global__ void __launch_bounds__(256) lstm_split_perm_t_kernel(const float* __restrict__ Whh, int H, __nv_bfloat16* __restrict__ planes,
                                                                int64_t ps) {
  pdl_enter();
  const int64_t total = (int64_t)4 * H * H;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i / (4 * H)), jp = (int)(i - (int64_t)k * 4 * H);
    const float x = Whh[(int64_t)perm_row(jp, H) * H + k];
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    planes[i] = h;
    planes[ps + i] = __float2bfloat16_rn(x - __bfloat162float(h));
  }
}
// forward prologue in one launch: planes of W_ih and W_hh (gate rows permuted) and the permuted bias sum
__global__ void __launch_bounds__(256) lstm_prep_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                                                        const float* __restrict__ b_ih, const float* __restrict__ b_hh, int H, int E,
                                                        __nv_bfloat16* __restrict__ wip, __nv_bfloat16* __restrict__ whp,
                                                        float* __restrict__ biasp) {
  pdl_enter();
  const int64_t n1 = (int64_t)4 * H * E, n2 = (int64_t)4 * H * H, total = n1 + n2 + 4 * H;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
    if (g < n1 + n2) {
      const bool first = g < n1;
      const int64_t i = first ? g : g - n1;
      const int cols = first ? E : H;
      const float* W = first ? w_ih : w_hh;
      __nv_bfloat16* planes = first ? wip : whp;
      const int64_t ps = first ? n1 : n2;
      const int jp = (int)(i / cols), c = (int)(i - (int64_t)jp * cols);
      const float x = W[(int64_t)perm_row(jp, H) * cols + c];
      const __nv_bfloat16 h = __float2bfloat16_rn(x);
      planes[i] = h;
      planes[ps + i] = __float2bfloat16_rn(x - __bfloat162float(h));
    } else {
      const int jp = (int)(g - n1 - n2);
      const int j = perm_row(jp, H);
      biasp[jp] = b_ih[j] + b_hh[j];
    }
  }
}
__global__ void __launch_bounds__(256) lstm_bias_perm_kernel(const float* __restrict__ b_ih, const float* __restrict__ b_hh, int H,
                                                             float* __restrict__ out) {
  pdl_enter();
  const int jp = blockIdx.x * blockDim.x + threadIdx.x;
  if (jp < 4 * H) {
    const int j = perm_row(jp, H);
    out[jp] = b_ih[j] + b_hh[j];
  }
}
// dst[perm(j')][c] = src[j'][c]   (dst2 optional second destination: b_ih and b_hh receive the same gradient)
__global__ void __launch_bounds__(256) lstm_unperm_kernel(const float* __restrict__ src, int H, int cols, float* __restrict__ dst,
                                                          float* __restrict__ dst2) {
  pdl_enter();
  const int64_t total = (int64_t)4 * H * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int jp = (int)(i / cols), c = (int)(i - (int64_t)jp * cols);
    const float v = src[i];
    const int64_t o = (int64_t)perm_row(jp, H) * cols + c;
    dst[o] = v;
    if (dst2) dst2[o] = v;
  }
}

// dW_ih, dW_hh and the bias gradient (both b_ih and b_hh receive it) in one launch
__global__ void __launch_bounds__(256) lstm_unperm3_kernel(const float* __restrict__ dwi, const float* __restrict__ dwh,
                                                           const float* __restrict__ dbp, int H, int E, float* __restrict__ dw_ih,
                                                           float* __restrict__ dw_hh, float* __restrict__ db_ih, float* __restrict__ db_hh) {
  pdl_enter();
  const int64_t n1 = (int64_t)4 * H * E, n2 = (int64_t)4 * H * H, total = n1 + n2 + 4 * H;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
    if (g < n1) {
      const int jp = (int)(g / E), c = (int)(g - (int64_t)jp * E);
      dw_ih[(int64_t)perm_row(jp, H) * E + c] = dwi[g];
    } else if (g < n1 + n2) {
      const int64_t i = g - n1;
      const int jp = (int)(i / H), c = (int)(i - (int64_t)jp * H);
      dw_hh[(int64_t)perm_row(jp, H) * H + c] = dwh[i];
    } else {
      const int jp = (int)(g - n1 - n2);
      const float v = dbp[jp];
      const int o = perm_row(jp, H);
      db_ih[o] = v;
      db_hh[o] = v;
    }
  }
}

struct Saved {
  float* act;              // [B][T][4H]
  float* c;                // [B][T][H]
  __nv_bfloat16* hp;       // [2][B][T][H]
  __nv_bfloat16* xp;       // [2][B][T][E]
  __nv_bfloat16* wip;      // [2][4H][E]: gate-permuted planes of w_ih (the input-gradient product of backward reads them again)
};
size_t saved_bytes(int B, int T, int E, int H) {
  const size_t BT = (size_t)B * T;
  return align_up(BT * 4 * H * 4) + align_up(BT * H * 4) + align_up(2 * BT * H * 2) + align_up(2 * BT * E * 2) +
         align_up((size_t)2 * 4 * H * E * 2) + 256;
}
bool carve_saved(Saved& s, void* buf, size_t bytes, int B, int T, int E, int H) {
  if (bytes < saved_bytes(B, T, E, H) || (reinterpret_cast<uintptr_t>(buf) & 255) != 0) return false;
  const size_t BT = (size_t)B * T;
  char* p = (char*)buf;
  s.act = (float*)p; p += align_up(BT * 4 * H * 4);
  s.c = (float*)p; p += align_up(BT * H * 4);
  s.hp = (__nv_bfloat16*)p; p += align_up(2 * BT * H * 2);
  s.xp = (__nv_bfloat16*)p; p += align_up(2 * BT * E * 2);
  s.wip = (__nv_bfloat16*)p;
  return true;
}

bool shape_ok(int B, int T, int E, int H) { return B > 0 && T > 0 && E > 0 && E % 8 == 0 && H >= 16 && H % 16 == 0 && H <= 512; }

TcOperand operand(const __nv_bfloat16* planes, int64_t ld, int64_t ps, int rows, int cols, bool mn_major) {
  TcOperand o;
  o.planes = planes; o.ld = ld; o.plane_stride = ps; o.rows = rows; o.cols = cols; o.mn_major = mn_major;
  return o;
}
int tc_splitk(int M, int N, int K) {
  const int tiles = ((M + 127) / 128) * ((N + 127) / 128);
  if (tiles >= 96) return 1;
  int sk = (148 + tiles - 1) / tiles;
  const int maxk = (K + 255) / 256;
  if (sk > maxk) sk = maxk;
  return sk < 1 ? 1 : sk;
}

// Row tiling: as many row tiles as fit on the device next to each other (tiles * H/16 CTAs <= SMs), so that every tile is as
// short as possible -- the per-step cost of a CTA is dominated by streaming its tile's rows of the exchanged operand.
struct RowTiling {
  int rpt, tiles, per_launch;
};
RowTiling row_tiling(int B, int H, int max_tiles_per_launch = 1 << 30) {
  int sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
    cudaGetLastError();
    sms = 148;
  }
  RowTiling r;
  r.per_launch = std::max(1, std::min(sms / (H / L_UNITS), max_tiles_per_launch));
  const int launches = ((B + L_BM - 1) / L_BM + r.per_launch - 1) / r.per_launch;
  const int slots = launches * r.per_launch;
  r.rpt = std::min(L_BM, (int)(((B + slots - 1) / slots + 7) / 8 * 8));
  r.tiles = (B + r.rpt - 1) / r.rpt;
  return r;
}
size_t counter_count(int B, int T) { return (size_t)((B + 7) / 8 + 1) * T; }

template <bool BWD>
int launch_rec(const LstmParams& base, const __nv_bfloat16* stream_planes, int64_t stream_ps, int stream_cols,
               const __nv_bfloat16* w_planes, int64_t w_ps, int w_rows, int w_cols, cudaStream_t s) {
  constexpr int BN = BWD ? L_UNITS : 4 * L_UNITS;
  LstmParams p = base;
  p.timeline = g_lstm_timeline;
  p.tiles_n = base.H / L_UNITS;
  p.K = BWD ? 4 * base.H : base.H;
  p.kbn = (p.K + L_BK - 1) / L_BK;
  // Backward, multicast variant (clusters of L_CL unit blocks of one row tile): taken when the unit blocks divide into clusters; the
  // residency requirement (every CTA of a launch resident at once) then applies to whole clusters, so the row tiling is chosen for the
  // number of clusters the device can hold (B200: 15 clusters of 8 at this shared-memory size -> 3 row tiles of 56 rows instead of 4
  // of 40 for H = 512; the MMAs cost the same either way, M is 128 rows per tile).  opt-in: HCA_LSTM_MC=1.
  bool mc = false;
  int mc_tiles = 0;
  if constexpr (BWD) {
    mc = (p.tiles_n % L_CL) == 0;
    // (measured at B = 160, H = 512: 287 us against 263 us without -- the stream is not what bounds the kernel, the MMA issue rate is:
    // the variant stays as an opt-in, HCA_LSTM_MC=1)
    { const char* ev = getenv("HCA_LSTM_MC"); if (!(ev && atoi(ev) == 1)) mc = false; }
    if (mc) {
      static bool mc_attr = false;
      if (!mc_attr) {
        HCA_CUDA(cudaFuncSetAttribute(lstm_rec_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, L_MAX_SMEM));
        mc_attr = true;
      }
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(L_CL * 64));
      cfg.blockDim = dim3(L_THREADS);
      cfg.dynamicSmemBytes = L_MAX_SMEM;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = L_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      int max_clusters = 0;
      if (cudaOccupancyMaxActiveClusters(&max_clusters, lstm_rec_kernel<true, true>, &cfg) != cudaSuccess) { cudaGetLastError(); max_clusters = 0; }
      mc_tiles = max_clusters / (p.tiles_n / L_CL);
      { const char* ev = getenv("HCA_LSTM_DEBUG"); if (ev && atoi(ev) != 0) fprintf(stderr, "lstm bwd: %d clusters of %d can be resident -> %d row tiles per launch\n", max_clusters, L_CL, mc_tiles); }
      if (mc_tiles < 1) mc = false;
    }
  }
  const RowTiling rt = mc ? row_tiling(base.B, base.H, mc_tiles) : row_tiling(base.B, base.H);
  HCA_CHECK_ARG(p.tiles_n <= rt.per_launch * p.tiles_n, "lstm: hidden size %d needs more CTAs per row tile than the device has SMs", base.H);
  p.rpt = rt.rpt;
  p.box_rows[0] = rt.rpt;
  p.box_rows[1] = std::min(rt.rpt, (rt.rpt / 2 + 7) / 8 * 8);
  p.box_rows[2] = std::min(rt.rpt, (rt.rpt / 4 + 7) / 8 * 8);
  // ring slots of the streamed operand: nkb k-blocks x (hi tile + lo tile) of rpt rows x 64 k (whole 8-row swizzle atoms: rpt % 8 == 0);
  // as many k-blocks per slot as still leave two slots (one TMA box per plane and slot: few large boxes, not many small ones)
  p.plane_bytes = (uint32_t)rt.rpt * L_BK * 2;
  // (the MMA always reads 128 rows = 16 KB from a tile base: rows beyond the loaded ones are other tiles' data and only feed
  // unused accumulator rows, but the last tile's read must stay inside the allocation, hence the tail padding)
  const uint32_t tail_pad = 16u * 1024u - p.plane_bytes;
  p.nkb = std::max(1, std::min(p.kbn, (int)((L_RING_BYTES - tail_pad) / (4 * p.plane_bytes))));
  { const char* ev = getenv(BWD ? "HCA_LSTM_NKB_BWD" : "HCA_LSTM_NKB_FWD"); if (ev && atoi(ev) >= 1 && atoi(ev) < p.nkb) p.nkb = atoi(ev); }
  p.nslots = (p.kbn + p.nkb - 1) / p.nkb;
  p.lo_off = (uint32_t)p.nkb * p.plane_bytes;
  p.stage_bytes = 2u * p.lo_off;
  p.stages = std::max(2, std::min(L_MAX_STAGES, (int)((L_RING_BYTES - tail_pad) / p.stage_bytes)));
  // stacked variant (backward, row tiles of <= 64 rows, see the kernel): one 128-row tile (16 KB) per k-block and ring slot
  bool stk = BWD && !mc && rt.rpt <= 64 &&
             (size_t)p.kbn * 2 * BN * L_BK * 2 + 2 * 16384u + 1024 <= (size_t)L_MAX_SMEM - 6144;
  // (measured at B = 160, H = 512: 270 us against 262 us for the two-MMA form, and finer ring slots are slower still (282 / 334 us at 2 / 1
  // k-blocks per slot): the backward step is a latency chain -- producer waits for the slot, TMA lands, issuer waits, MMAs complete, commit
  // -- over a ring that the 128 KB resident W_hh slice leaves only 96 KB for, not an issue-rate or bandwidth limit.  Opt-in: HCA_LSTM_STK=1)
  { const char* ev = getenv("HCA_LSTM_STK"); if (!(ev && atoi(ev) == 1)) stk = false; }
  if (stk) {
    p.nkb = 1;
    p.nslots = p.kbn;
    p.stage_bytes = 16384u;
    p.lo_off = 8192u;
    const size_t room = (size_t)L_MAX_SMEM - 6144 - 1024 - (size_t)p.kbn * 2 * BN * L_BK * 2;
    p.stages = std::max(2, std::min(std::min(L_MAX_STAGES, 6), (int)(room / p.stage_bytes)));
  }
  const size_t smem = (size_t)p.kbn * 2 * BN * L_BK * 2 + (size_t)p.stages * p.stage_bytes + (stk ? 0 : tail_pad) + 1024;
  LstmMaps maps;
  for (int i = 0; i < 3; ++i) {
    // streamed operand [2][B][T][cols] seen as (64 cols of a k-block, b, k-block, t, plane).  Dimension 0 is always a full 64:
    // when cols is not a multiple of 64 the last k-block of a row runs into the following row -- finite data that only feeds
    // k-slices the MMA warp predicates off (K is a multiple of 16); the buffers carry slack behind their last row.
    const uint64_t dims[5] = {(uint64_t)L_BK, (uint64_t)base.B, (uint64_t)((stream_cols + L_BK - 1) / L_BK), (uint64_t)base.T, 2};
    const uint64_t str[4] = {(uint64_t)base.T * stream_cols * 2, (uint64_t)L_BK * 2, (uint64_t)stream_cols * 2, (uint64_t)stream_ps * 2};
    const uint32_t box[5] = {L_BK, (uint32_t)p.box_rows[i], (uint32_t)p.nkb, 1, 1};
    HCA_TRY(tc_make_tmap(&maps.A[i], true, 5, stream_planes, dims, str, box, 3));
    const uint32_t box1[5] = {L_BK, (uint32_t)p.box_rows[i], 1, 1, 1};
    HCA_TRY(tc_make_tmap(&maps.A1[i], true, 5, stream_planes, dims, str, box1, 3));
  }
  if (stk) {
    // (64 cols of a k-block, b, plane, k-block, t): the box (64, 64, 2, 1, 1) lands as [plane][row][64] = hi rows 0..63, lo rows 64..127
    const uint64_t dims[5] = {(uint64_t)L_BK, (uint64_t)base.B, 2, (uint64_t)((stream_cols + L_BK - 1) / L_BK), (uint64_t)base.T};
    const uint64_t str[4] = {(uint64_t)base.T * stream_cols * 2, (uint64_t)stream_ps * 2, (uint64_t)L_BK * 2, (uint64_t)stream_cols * 2};
    const uint32_t box[5] = {L_BK, 64, 2, 1, 1};
    HCA_TRY(tc_make_tmap(&maps.AS, true, 5, stream_planes, dims, str, box, 3));
  } else {
    maps.AS = maps.A[0];
  }
  {  // resident operand [2][rows][cols]: dims (cols, rows, 2, 1)
    const uint64_t dims[4] = {(uint64_t)w_cols, (uint64_t)w_rows, 2, 1};
    const uint64_t str[3] = {(uint64_t)w_cols * 2, (uint64_t)w_ps * 2, (uint64_t)w_ps * 4};
    const uint32_t box[4] = {L_BK, BN, 1, 1};
    HCA_TRY(tc_make_tmap(&maps.W, true, 4, w_planes, dims, str, box, 3));
  }
  HCA_CHECK_ARG(smem <= (size_t)L_MAX_SMEM, "lstm: hidden size %d needs %zu bytes of shared memory", base.H, smem);
  static bool attr_set[2] = {false, false};
  if (!attr_set[BWD ? 1 : 0]) {
    HCA_CUDA(cudaFuncSetAttribute(lstm_rec_kernel<BWD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, L_MAX_SMEM));
    attr_set[BWD ? 1 : 0] = true;
  }
  // every CTA of a launch must be resident at once (the CTAs of a row tile synchronise through global counters); row tiles
  // are independent, so when the batch needs more tiles than fit the launches simply follow each other
  for (int t0 = 0; t0 < rt.tiles; t0 += rt.per_launch) {
    const int nm = std::min(rt.per_launch, rt.tiles - t0);
    p.tile0 = t0;
    if constexpr (BWD) {
      if (mc) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(nm * p.tiles_n));
        cfg.blockDim = dim3(L_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = L_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = pdl_enabled() ? 2 : 1;
        HCA_CUDA(cudaLaunchKernelEx(&cfg, lstm_rec_kernel<true, true>, maps, p));
        HCA_LAUNCHED();
        continue;
      }
    }
    if constexpr (BWD) {
      if (stk) {
        static bool stk_attr = false;
        if (!stk_attr) {
          // (the exchange buffer adds 4 KB of static shared memory: the dynamic limit shrinks by as much)
          HCA_CUDA(cudaFuncSetAttribute(lstm_rec_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, L_MAX_SMEM - 6144));
          stk_attr = true;
        }
        HCA_LAUNCH_K((lstm_rec_kernel<true, false, true>), nm * p.tiles_n, L_THREADS, smem, s, maps, p);
        HCA_LAUNCHED();
        continue;
      }
    }
    if (g_rec_events[BWD ? 1 : 0][0] && t0 == 0) HCA_CUDA(cudaEventRecord(g_rec_events[BWD ? 1 : 0][0], s));
    HCA_LAUNCH_K((lstm_rec_kernel<BWD, false>), nm * p.tiles_n, L_THREADS, smem, s, maps, p);
    HCA_LAUNCHED();
    if (g_rec_events[BWD ? 1 : 0][1] && t0 + rt.per_launch >= rt.tiles) HCA_CUDA(cudaEventRecord(g_rec_events[BWD ? 1 : 0][1], s));
  }
  return 0;
}

}  // namespace
}  // namespace hca

extern "C" int hca_debug_lstm_timeline(void* buf) {
  hca::g_lstm_timeline = (long long*)buf;
  return 0;
}

// bench.py's roofline legs: the recurrence kernel of the next hca_lstm_fwd (which = 0) / hca_lstm_bwd (which = 1) calls is bracketed by
// these two CUDA events on the launching stream (cudaEvent_t handles; nullptr switches it off)
extern "C" int hca_debug_lstm_events(void* ev_start, void* ev_stop, int which) {
  if (which < 0 || which > 1) return hca::set_err(HCA_ERR_ARG, "debug_lstm_events: which = 0 (forward) or 1 (backward)");
  hca::g_rec_events[which][0] = (cudaEvent_t)ev_start;
  hca::g_rec_events[which][1] = (cudaEvent_t)ev_stop;
  return 0;
}

extern "C" int hca_lstm_supported(int B, int T, int E, int H) { return hca::shape_ok(B, T, E, H) && hca::tc_available() ? 1 : 0; }

extern "C" size_t hca_lstm_saved_bytes(int B, int T, int E, int H) { return hca::saved_bytes(B, T, E, H); }

extern "C" size_t hca_lstm_workspace(int B, int T, int E, int H) {
  using hca::align_up;
  const size_t BT = (size_t)B * T, H4 = (size_t)4 * H;
  const size_t fwd = align_up(2 * H4 * E * 2) + align_up(2 * H4 * H * 2) + align_up(H4 * 4) + align_up(hca::counter_count(B, T) * 4);
  const size_t bwd = align_up(2 * BT * H4 * 2) + align_up(2 * H4 * E * 2) + align_up(2 * H4 * H * 2) + align_up(H4 * E * 4) + align_up(H4 * H * 4) +
                     align_up(H4 * 4) + align_up(hca::counter_count(B, T) * 4);
  return std::max(fwd, bwd) + 4096;
}

extern "C" int hca_lstm_fwd(const float* x, const int64_t* lens, const float* w_ih, const float* w_hh, const float* b_ih,
                            const float* b_hh, float* out, void* saved, size_t saved_sz, int B, int T, int E, int H, void* ws,
                            size_t ws_bytes, void* stream) {
  using namespace hca;
  cudaStream_t s = (cudaStream_t)stream;
  HCA_CHECK_ARG(x && lens && w_ih && w_hh && b_ih && b_hh && out && saved, "lstm_fwd: null pointer");
  HCA_CHECK_ARG(shape_ok(B, T, E, H), "lstm_fwd: unsupported sizes B=%d T=%d E=%d H=%d (E %% 8 == 0, H %% 16 == 0, H <= 512)", B, T, E, H);
  HCA_CHECK_ARG(tc_available(), "lstm_fwd: cuTensorMapEncodeTiled is not available from the driver");
  Saved sv;
  HCA_CHECK_ARG(carve_saved(sv, saved, saved_sz, B, T, E, H), "lstm_fwd: `saved` must be 256-byte aligned and hca_lstm_saved_bytes large");
  Workspace w(ws, ws_bytes);
  const int64_t BT = (int64_t)B * T;
  const int H4 = 4 * H;
  __nv_bfloat16* wip = sv.wip;                                   // kept in `saved`: backward's dx product reads them again
  __nv_bfloat16* whp = w.take<__nv_bfloat16>((size_t)2 * H4 * H);
  float* biasp = w.take<float>((size_t)H4);
  int* counters = w.take<int>(counter_count(B, T));
  if (!counters) return set_err(HCA_ERR_WORKSPACE, "lstm_fwd: workspace too small (%zu bytes)", ws_bytes);
  HCA_TRY(launch_split_planes(x, E, BT, E, sv.xp, E, BT * E, 2, s));
  HCA_LAUNCH_K((lstm_prep_kernel), ew_grid((int64_t)H4 * (E + H + 1)), 256, 0, s, w_ih, w_hh, b_ih, b_hh, H, E, wip, whp, biasp);
  HCA_LAUNCHED();
  {
    ZeroBatch zb(s);
    HCA_TRY(zb.add(sv.hp, (size_t)2 * BT * H * 2));            // slot 0 (h_{-1} = 0) and the slots no step reaches
    HCA_TRY(zb.add(counters, counter_count(B, T) * 4));
    HCA_TRY(zb.flush());
  }
  {  // x-projection of every (b, t), gate columns in [unit][gate] order, biases folded in
    TcEpilogue e;
    e.D = sv.act; e.ldd = H4; e.bias = biasp;
    HCA_TRY(launch_gemm_tc(operand(sv.xp, E, BT * E, (int)BT, E, false), operand(wip, E, (int64_t)H4 * E, H4, E, false), 2, (int)BT, H4, E,
                           e, 1, s));
  }
  LstmParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.T = T; p.H = H; p.lens = lens; p.counters = counters;
  p.act = sv.act; p.c = sv.c; p.out = out; p.hp = sv.hp; p.hp_ps = BT * H;
  return launch_rec<false>(p, sv.hp, BT * H, H, whp, (int64_t)H4 * H, H4, H, s);
}

extern "C" int hca_lstm_bwd(const int64_t* lens, const float* w_ih, const float* w_hh, const void* saved, size_t saved_sz,
                            const float* dout, float* dx, float* dw_ih, float* dw_hh, float* db_ih, float* db_hh, int B, int T, int E,
                            int H, void* ws, size_t ws_bytes, void* stream) {
  using namespace hca;
  cudaStream_t s = (cudaStream_t)stream;
  HCA_CHECK_ARG(lens && w_ih && w_hh && saved && dout && dw_ih && dw_hh && db_ih && db_hh, "lstm_bwd: null pointer");
  HCA_CHECK_ARG(shape_ok(B, T, E, H), "lstm_bwd: unsupported sizes B=%d T=%d E=%d H=%d", B, T, E, H);
  HCA_CHECK_ARG(tc_available(), "lstm_bwd: cuTensorMapEncodeTiled is not available from the driver");
  Saved sv;
  HCA_CHECK_ARG(carve_saved(sv, const_cast<void*>(saved), saved_sz, B, T, E, H), "lstm_bwd: bad `saved` buffer");
  Workspace w(ws, ws_bytes);
  const int64_t BT = (int64_t)B * T;
  const int H4 = 4 * H;
  __nv_bfloat16* dgp = w.take<__nv_bfloat16>((size_t)2 * BT * H4);
  const __nv_bfloat16* wip = sv.wip;                             // written by the forward call
  __nv_bfloat16* wtp = w.take<__nv_bfloat16>((size_t)2 * H4 * H);
  float* dwi = w.take<float>((size_t)H4 * E);
  float* dwh = w.take<float>((size_t)H4 * H);
  float* dbp = w.take<float>((size_t)H4);
  int* counters = w.take<int>(counter_count(B, T));
  if (!counters) return set_err(HCA_ERR_WORKSPACE, "lstm_bwd: workspace too small (%zu bytes)", ws_bytes);
  const int sk_wi = tc_splitk(H4, E, (int)BT), sk_wh = tc_splitk(H4, H, (int)BT);
  {
    ZeroBatch zb(s);
    HCA_TRY(zb.add(dgp, (size_t)2 * BT * H4 * 2));             // rows no step writes must read as zero in the GEMMs below
    HCA_TRY(zb.add(dbp, (size_t)H4 * 4));
    HCA_TRY(zb.add(counters, counter_count(B, T) * 4));
    if (sk_wi > 1) HCA_TRY(zb.add(dwi, (size_t)H4 * E * 4));
    if (sk_wh > 1) HCA_TRY(zb.add(dwh, (size_t)H4 * H * 4));
    HCA_TRY(zb.flush());
  }
  HCA_LAUNCH_K((lstm_split_perm_t_kernel), ew_grid((int64_t)H4 * H), 256, 0, s, w_hh, H, wtp, (int64_t)H4 * H);
  HCA_LAUNCHED();
  LstmParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.T = T; p.H = H; p.lens = lens; p.counters = counters;
  p.act = sv.act; p.c = sv.c; p.dout = dout; p.dgp = dgp; p.dgp_ps = BT * H4; p.dbias = dbp;
  HCA_TRY(launch_rec<true>(p, dgp, BT * H4, H4, wtp, (int64_t)H4 * H, H, H4, s));
  const TcOperand dg_mn = operand(dgp, H4, BT * H4, (int)BT, H4, true);
  {  // dW_ih' = dz^T x   (K = B*T, split-K)
    const int sk = sk_wi;
    TcEpilogue e; e.D = dwi; e.ldd = E;
    HCA_TRY(launch_gemm_tc(dg_mn, operand(sv.xp, E, BT * E, (int)BT, E, true), 2, H4, E, (int)BT, e, sk, s));
  }
  {  // dW_hh' = dz^T h_prev
    const int sk = sk_wh;
    TcEpilogue e; e.D = dwh; e.ldd = H;
    HCA_TRY(launch_gemm_tc(dg_mn, operand(sv.hp, H, BT * H, (int)BT, H, true), 2, H4, H, (int)BT, e, sk, s));
  }
  // the three gradients back to PyTorch's gate order, one launch
  HCA_LAUNCH_K((lstm_unperm3_kernel), ew_grid((int64_t)H4 * (E + H + 1)), 256, 0, s, dwi, dwh, dbp, H, E, dw_ih, dw_hh, db_ih, db_hh);
  HCA_LAUNCHED();
  if (dx) {  // dx = dz W_ih
    TcEpilogue e; e.D = dx; e.ldd = E;
    HCA_TRY(launch_gemm_tc(operand(dgp, H4, BT * H4, (int)BT, H4, false), operand(wip, E, (int64_t)H4 * E, H4, E, true), 2, (int)BT, E, H4,
                           e, 1, s));
  }
  return 0;
}
