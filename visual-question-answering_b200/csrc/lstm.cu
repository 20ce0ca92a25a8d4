// Sentence-level LSTM over variable-length questions (replaces pack_padded_sequence -> nn.LSTM(E, H) -> pad_packed_sequence,
// reference model.py:269,287-296).  Gate order i, f, g, o; h0 = c0 = 0; every sequence stops at its own length; rows
// t >= len of the output are zero.
//
//   z_t = x_t W_ih^T + b_ih + b_hh + h_{t-1} W_hh^T        i,f,o = sigmoid(z), g = tanh(z)
//   c_t = f c_{t-1} + i g                                  h_t = o tanh(c_t)
//
// The input projection of all (b, t) is one tcgen05 GEMM (gemm_tc.cuh).  The recurrence -- T dependent steps of a
// [B, H] x [H, 4H] product -- runs in ONE persistent kernel per direction instead of T library GEMM + cell launches.
//
// Orientation: the WEIGHT rows sit on the UMMA M axis and the batch on N.  A CTA owns the cell math of 16 hidden units of one
// row tile (rpt <= 40 batch rows).  Its resident operand is a 64-row slice of the weight matrix as bf16 hi / lo planes stacked
// along M ([W_hi ; W_lo] = 128 rows x K, 128 KB at K = 512, loaded once and kept for all steps); the streamed operand of a step
// is the [rpt x K] activation tile of the row tile, hi and lo planes stacked along N ([x_hi ; x_lo] = 2 rpt rows).  ONE
// tcgen05.mma per 16-wide k-slice (M = 128, N = 2 rpt <= 80, fp32 accumulator in TMEM) therefore yields all four partial
// products W_hi.x_hi, W_hi.x_lo, W_lo.x_hi, W_lo.x_lo; the epilogue adds them.  (The previous orientation put the batch on M:
// 128 MMA rows for 40 valid batch rows, 3.2x the tensor cycles, and a cell epilogue in which 80 of 256 threads had work.)
//   forward : rows = the 64 gate rows of the CTA's 16 units ([unit][gate] order), K = H, stream = h_{t-1}.  The accumulator
//             comes back as D[gate row][batch]; the four TMEM-reader warps add the partial products and lay them out in shared
//             memory as [batch][gate row], where 320 cell threads (batch row x unit pair) pick up the 4 gates of their units.
//   backward: dh = dz_{t+1} . W_hh contracts over 4H gate columns.  With 16 units per CTA every CTA would have to stream all
//             4H columns of the row tile's gate gradients every step (320 KB per CTA and step: that stream bounded the old
//             kernel).  Instead a CLUSTER of 4 CTAs shares 64 units and splits the contraction: CTA j holds W_hh^T[64 units,
//             quarter j of the gate columns] (again 128 KB), streams only its quarter (80 KB, like the forward), and the four
//             partial [64 units x rpt] results are exchanged through DISTRIBUTED SHARED MEMORY (st.shared::cluster + a cluster-
//             scope mbarrier): each CTA receives the three foreign partials of the 16 units whose cell math it owns.
// The CTAs of a row tile exchange h_t (forward) / dz_t (backward) through global bf16 planes -- written by the cell threads,
// ordered by a release/acquire counter per (row tile, step), streamed back in by TMA -- because a row tile spans 32 CTAs, more
// than a cluster holds; rows of different tiles never wait for each other.  The gate gradients are written as the very planes
// the three big weight / input-gradient GEMMs consume afterwards.  All CTAs of a launch must be resident at once (they wait for
// each other): the launcher sizes every launch from the occupancy API and refuses shapes the device cannot hold.
#include <cuda.h>
#include <cuda_bf16.h>
#include <algorithm>
#include <cstring>
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

constexpr int L_BK = 64;                          // bf16 per k-block row = 128 B = one SWIZZLE_128B span
constexpr int L_UNITS = 16;                       // hidden units whose cell math one CTA owns
constexpr int L_WROWS = 64;                       // weight rows per plane in the resident tile (fwd: 4 gates x 16 units; bwd: 64 units)
constexpr int L_RPT_MAX = 40;                     // batch rows per row tile (UMMA N = 2 rpt <= 80; the streamed step operand stays in smem)
constexpr int L_KB_MAX = 8;                       // k-blocks of a CTA's contraction (K <= 512)
constexpr int L_KSPLIT = 4;                       // backward: CTAs per cluster = slices of the 4H contraction
constexpr int L_EPI = 320;                        // cell threads: 8 unit pairs x 40 batch rows
constexpr int L_THREADS = 64 + L_EPI;             // warp 0 TMA, warp 1 MMA, warps 2..11 epilogue (warps 2..5 also read TMEM)
constexpr int L_PSTRIDE = 132;                    // forward: floats per batch row of the [batch][gate row] staging tile
constexpr int L_RSTRIDE = 44;                     // backward: floats per unit row of the staging / receive tiles ([unit][batch])
constexpr uint32_t L_WKB_BYTES = 2u * L_WROWS * L_BK * 2u;                 // one resident k-block: hi tile + lo tile = 16 KB
constexpr uint32_t L_RECV_BYTES = L_KSPLIT * L_UNITS * L_RSTRIDE * 4u;     // backward: partial sums received from the cluster
constexpr int L_MAX_SMEM = 227 * 1024 - 2048;
constexpr uint32_t L_TMEM_COLS = 128;

struct LstmMaps {
  CUtensorMap X;      // streamed operand planes, 4-D (col, b, plane, t), box (64, rpt, 2, 1): lands as [x_hi rows ; x_lo rows] x 64 k
  CUtensorMap W;      // resident operand planes, 4-D (col, row, plane, 1), box (64, 64, 1, 1)
};

struct LstmParams {
  int B, T, H;
  int K;                  // contraction length of one CTA: H (forward: all of h; backward: a quarter of the 4H gate columns)
  int kbn;                // its k-blocks: ceil(K / 64)
  int tile0;              // first row tile of this launch
  int rpt;                // batch rows per row tile (multiple of 8, <= 40): tile i owns rows [i * rpt, (i + 1) * rpt)
  int ctas_per_tile;      // forward H / 16 ; backward 4 * ceil(H / 64)
  uint32_t ring_bytes;    // shared memory of the streamed operand (>= the staging tile that aliases it)
  uint32_t idesc;         // UMMA instruction descriptor (M = 128, N = 2 rpt, bf16 x bf16 -> f32, both operands K-major)
  const int64_t* lens;
  int* counters;          // [row tiles][T]: CTAs of the row tile that have published step t
  float* act;             // [B][T][4H], gate columns in [unit][gate] order.  fwd: in = x-projection + biases, out = activations
  float* c;               // [B][T][H] cell states
  float* out;             // [B][T][H] (fwd)
  __nv_bfloat16* hp;      // h planes [2][B][T][H]: slot t holds h_{t-1} (slot 0 = 0)
  int64_t hp_ps, hp_ld;   // plane stride and row pitch of the h planes (they share rows with the x planes: [x | h] per token)
  const float* dout;      // [B][T][H] (bwd)
  __nv_bfloat16* dgp;     // gate-gradient planes [2][B][T][4H] ([unit][gate] column order) (bwd)
  int64_t dgp_ps;
  float* dbias;           // [4H] in [unit][gate] order, accumulated atomically (bwd)
  long long* timeline;    // debug (HCA_BUILD_TIMELINE=1): clock64 stamps of CTA 0, [round][8]: counter seen, loads issued, first k-block landed,
                          // MMAs issued, accumulator seen by the cell threads, cell math done, barrier passed, step published
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
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// tcgen05.mma from ONE elected thread, instruction descriptor in a register (N depends on the row tiling)
template <uint32_t DESC_HI>
__device__ __forceinline__ void umma_bf16_rt(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "n"(DESC_HI)
      : "memory");
}
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_relaxed_add(int* p, int v) {
  asm volatile("red.relaxed.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Publication of a step: every storing thread makes ITS stores visible at gpu scope (all of them in parallel, ~one L2 round trip in
// total), the epilogue barrier orders those fences before the leader's relaxed increment of the step counter.  (One releasing
// increment by the leader after the barrier had to wait for the stores of all 320 threads on its own: 1.7 k cycles forward, 3.8 k
// backward in the per-step timeline.)
__device__ __forceinline__ void publish_fence() {
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
  asm volatile("fence.proxy.async.global;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_all() { asm volatile("bar.sync 1, 320;" ::: "memory"); }      // the 10 epilogue warps
__device__ __forceinline__ void reader_bar() { asm volatile("bar.sync 2, 128;" ::: "memory"); }       // the 4 TMEM-reader warps
// ---- distributed shared memory (cluster of L_KSPLIT CTAs, backward) ----
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t remote_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
static __device__ __noinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, int tag) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("hiecoattn lstm: cluster barrier wait timed out (tag %d, block %d, thread %d)\n", tag, blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

// 2 fp32 -> bf16 hi pair + bf16 lo pair (x = hi + lo to ~2^-17)
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 hh = __floats2bfloat162_rn(x0, x1);
  const __nv_bfloat162 ll = __floats2bfloat162_rn(x0 - __low2float(hh), x1 - __high2float(hh));
  hi = *reinterpret_cast<const uint32_t*>(&hh);
  lo = *reinterpret_cast<const uint32_t*>(&ll);
}

template <bool BWD>
__global__ void __launch_bounds__(L_THREADS, 1) lstm_rec_kernel(const __grid_constant__ LstmMaps maps, const LstmParams p) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bars[L_KB_MAX + 3];
  __shared__ uint32_t tmem_ptr_smem;
  __shared__ int s_maxlen;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  auto full_bar = [&](int kb) { return smem_u32(&bars[kb]); };
  const uint32_t w_bar = smem_u32(&bars[L_KB_MAX]), tmem_full = smem_u32(&bars[L_KB_MAX + 1]), red_bar = smem_u32(&bars[L_KB_MAX + 2]);

  // which row tile / unit block / contraction slice this CTA is
  int mi, ub, jg = 0, w_row0, w_col0;
  if constexpr (BWD) {
    const int per = p.ctas_per_tile;                      // 4 * unit groups
    mi = blockIdx.x / per;
    const int r = blockIdx.x - mi * per;
    const int ug = r / L_KSPLIT;
    jg = r - ug * L_KSPLIT;                               // == %cluster_ctarank (the cluster is 4 consecutive CTAs)
    ub = ug * L_KSPLIT + jg;                              // unit block whose cell math this CTA owns (may lie beyond H: no cell work then)
    w_row0 = ug * L_WROWS;                                // resident tile: rows = the 64 units of the group, columns = quarter jg of 4H
    w_col0 = jg * p.H;
  } else {
    mi = blockIdx.x / p.ctas_per_tile;
    ub = blockIdx.x - mi * p.ctas_per_tile;
    w_row0 = ub * L_WROWS;                                // resident tile: the 64 gate rows of the unit block, all H columns
    w_col0 = 0;
  }
  const int m0 = (p.tile0 + mi) * p.rpt;
  int* const counters = p.counters + (int64_t)(p.tile0 + mi) * p.T;
  long long* const tl = (HCA_TC_TIMELINE && blockIdx.x == 0) ? p.timeline : nullptr;
  const uint32_t slot_bytes = (uint32_t)p.rpt * 2u * L_BK * 2u;            // one k-block of the streamed operand: [x_hi ; x_lo] rows x 128 B
  const uint32_t w_base = smem_base;
  const uint32_t ring_base = w_base + (uint32_t)p.kbn * L_WKB_BYTES;
  const uint32_t recv_base = ring_base + p.ring_bytes;                     // (backward only)
  // staging tiles alias the ring: they are written after the step's last MMA has completed and read before the step is published,
  // i.e. strictly between two uses of the ring (the next step's loads are issued only after every CTA of the row tile has published)
  float* const stage = reinterpret_cast<float*>(smem_raw + (ring_base - smem_u32(smem_raw)));
  const float* const recv = reinterpret_cast<const float*>(smem_raw + (recv_base - smem_u32(smem_raw)));

  if (threadIdx.x == 0) {
    s_maxlen = 0;
    for (int kb = 0; kb < p.kbn; ++kb) mbar_init(full_bar(kb), 1);
    mbar_init(w_bar, 1);
    mbar_init(tmem_full, 1);
    mbar_init(red_bar, 2 * L_KSPLIT);                     // (backward) one arrival per sending warp (2) of every CTA of the cluster
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.X) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.W) : "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_ptr_smem), L_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  if constexpr (BWD) cluster_sync_all();     // every CTA's barriers exist before a peer can signal them
  tc_fence_after();
  // steps this row tile needs: the longest sequence among its rows (rows are independent, so other tiles may run longer)
  if ((int)threadIdx.x < p.rpt) {
    const int b = m0 + (int)threadIdx.x;
    if (b < p.B) {
      const int64_t l64 = p.lens[b];
      atomicMax(&s_maxlen, (int)(l64 < 0 ? 0 : (l64 > p.T ? p.T : l64)));
    }
  }
  __syncthreads();
  const int steps = s_maxlen;
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_ptr_smem, 0);
  const int rounds = steps > 0 ? steps - 1 : 0;             // recurrent products: every step but the first one in time order

  if (warp == 0) {
    // ============================================================ TMA producer (one elected lane, inline waits: uniform datapath)
    if (elect_one_sync()) {
      mbar_expect_tx(w_bar, (uint32_t)p.kbn * L_WKB_BYTES);
      for (int kb = 0; kb < p.kbn; ++kb)
        for (int pl = 0; pl < 2; ++pl)        // (rows / columns beyond the matrix arrive as zeros: out-of-bounds fill)
          tma_load_4d(w_base + (uint32_t)kb * L_WKB_BYTES + (uint32_t)pl * (L_WKB_BYTES / 2), &maps.W, w_bar, w_col0 + kb * L_BK, w_row0, pl, 0);
      for (int n = 0; n < rounds; ++n) {
        // forward round n computes step t = n + 1 from h_n (slot n + 1, published at step n);
        // backward round n computes step t = steps - 2 - n from the gate gradients of step t + 1 (slot t + 1)
        const int dep = BWD ? steps - 1 - n : n;
        const int slot = BWD ? dep : n + 1;
        const int* cnt = counters + dep;
        if (ld_acquire(cnt) < p.ctas_per_tile) {     // bounded spin: a protocol bug must trap, never hang the device
          const long long t0 = clock64();
          while (ld_acquire(cnt) < p.ctas_per_tile) {
            if (clock64() - t0 > 4000000000LL) {
              printf("hiecoattn lstm: step counter wait timed out (block %d, round %d)\n", blockIdx.x, n);
              __trap();
            }
          }
        }
        fence_proxy_async_all();                     // peers wrote through the generic proxy; TMA reads (and overwrites the staging tiles) through the async proxy
        if (tl && n < 7) tl[n * 8 + 0] = clock64();
        for (int kb = 0; kb < p.kbn; ++kb) {
          mbar_expect_tx(full_bar(kb), slot_bytes);
          tma_load_4d(ring_base + (uint32_t)kb * slot_bytes, &maps.X, full_bar(kb), w_col0 + kb * L_BK, m0, 0, slot);
        }
        if (tl && n < 7) tl[n * 8 + 1] = clock64();
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer (one elected thread: descriptors stay in uniform registers)
    constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    if (elect_one_sync()) {
      mbar_spin(w_bar, 0);
      tc_fence_after();
      const uint32_t d_tmem = *reinterpret_cast<volatile uint32_t*>(&tmem_ptr_smem);
      const uint32_t a_res = ((w_base & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t b_ring = ((ring_base & 0x3FFFFu) >> 4) | (1u << 16);
      for (int n = 0; n < rounds; ++n) {
        for (int kb = 0; kb < p.kbn; ++kb) {
          mbar_spin(full_bar(kb), (uint32_t)(n & 1));
          tc_fence_after();
          if (tl && n < 7 && kb == 0) tl[n * 8 + 2] = clock64();
          const uint32_t au = a_res + (uint32_t)kb * (L_WKB_BYTES >> 4);
          const uint32_t bu = b_ring + (uint32_t)kb * (slot_bytes >> 4);
          const int nks = max(0, min(L_BK / 16, (p.K - kb * L_BK + 15) / 16));
#pragma unroll
          for (int ks = 0; ks < L_BK / 16; ++ks)
            if (ks < nks) umma_bf16_rt<desc_hi>(d_tmem, au + ks * 2, bu + ks * 2, p.idesc, (kb | ks) != 0 ? 1u : 0u);
        }
        if (tl && n < 7) tl[n * 8 + 3] = clock64();
        umma_commit(tmem_full);
      }
    }
  } else {
    // ============================================================ epilogue: 10 warps; thread e = (batch row b, unit pair up)
    const int e = (int)threadIdx.x - 64;
    const int b_loc = e >> 3, up = e & 7;
    const int b = m0 + b_loc;
    const bool row_ok = b_loc < p.rpt && b < p.B;
    const bool unit_ok = ub * L_UNITS < p.H;              // (H % 16 == 0: a unit block is entirely inside or outside)
    const bool cell_ok = row_ok && unit_ok;
    int len_b = 0;
    if (row_ok) {
      const int64_t l = p.lens[b];
      len_b = (int)(l < 0 ? 0 : (l > p.T ? p.T : l));
    }
    const bool leader = (e == 0);
    const bool reader = warp < 6;                         // warps 2..5 own the TMEM lane quadrants 2, 3, 0, 1
    const int lrow = (warp & 3) * 32 + lane;              // accumulator lane (= weight row of the stacked [hi ; lo] tile) of a reader thread
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int H = p.H, H4 = 4 * p.H;
    const int u0 = ub * L_UNITS + 2 * up;                 // first of this thread's 2 units
    const int g0 = 4 * u0;                                // first of its 8 gate columns ([unit][gate] order)
    const int64_t bt0 = (int64_t)b * p.T;
    const int nchunk = p.rpt >> 3;
    if constexpr (!BWD) {
      float cst[2] = {0.f, 0.f};
      for (int t = 0; t < steps; ++t) {
        float z[8];
        if (row_ok) {                                // x-projection + biases of this step (independent of the recurrence)
          const float4* gx = reinterpret_cast<const float4*>(p.act + (bt0 + t) * H4 + g0);
          const float4 v0 = __ldg(gx), v1 = __ldg(gx + 1);
          z[0] = v0.x; z[1] = v0.y; z[2] = v0.z; z[3] = v0.w; z[4] = v1.x; z[5] = v1.y; z[6] = v1.z; z[7] = v1.w;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) z[j] = 0.f;
        }
        if (t > 0) {
          mbar_wait(tmem_full, (uint32_t)((t - 1) & 1), 14);
          tc_fence_after();
          if (tl && leader && t - 1 < 7) tl[(t - 1) * 8 + 4] = clock64();
          if (reader) {
            // D[weight row][2 rpt columns]: columns [0, rpt) = . x_hi, [rpt, 2 rpt) = . x_lo.  stage[b][row] = their sum.
            for (int c = 0; c < nchunk; ++c) {
              uint32_t v[8], w2[8];
              tmem_ld8(lane_addr + (uint32_t)(8 * c), v);
              tmem_ld8(lane_addr + (uint32_t)(p.rpt + 8 * c), w2);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 8; ++j) stage[(8 * c + j) * L_PSTRIDE + lrow] = __uint_as_float(v[j]) + __uint_as_float(w2[j]);
            }
            tc_fence_before();
          }
          epi_bar_all();
          if (row_ok) {                              // rows 0..63 = W_hi . x, rows 64..127 = W_lo . x, [unit][gate] order: 8 gates of the unit pair
            const float4* sp = reinterpret_cast<const float4*>(stage + b_loc * L_PSTRIDE + 8 * up);
            const float4 h0 = sp[0], h1 = sp[1], l0 = sp[L_WROWS / 4], l1 = sp[L_WROWS / 4 + 1];
            z[0] += h0.x + l0.x; z[1] += h0.y + l0.y; z[2] += h0.z + l0.z; z[3] += h0.w + l0.w;
            z[4] += h1.x + l1.x; z[5] += h1.y + l1.y; z[6] += h1.z + l1.z; z[7] += h1.w + l1.w;
          }
        }
        float h[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float ig = sigmoid_fast(z[4 * u]), fg = sigmoid_fast(z[4 * u + 1]), gg = tanh_fast(z[4 * u + 2]),
                      og = sigmoid_fast(z[4 * u + 3]);
          cst[u] = fmaf(fg, cst[u], ig * gg);
          h[u] = og * tanh_fast(cst[u]);
          z[4 * u] = ig; z[4 * u + 1] = fg; z[4 * u + 2] = gg; z[4 * u + 3] = og;
        }
        // publish h_t (slot t + 1) first: it is on the critical path of every CTA of this row tile
        if (row_ok && t < len_b && t + 1 < p.T) {         // (rows past their end keep the zero slot: their h is never used)
          uint32_t hi, lo;
          split2(h[0], h[1], hi, lo);
          __nv_bfloat16* dst = p.hp + (bt0 + t + 1) * p.hp_ld + u0;
          *reinterpret_cast<uint32_t*>(dst) = hi;
          *reinterpret_cast<uint32_t*>(dst + p.hp_ps) = lo;
        }
        if (tl && leader && t >= 1 && t - 1 < 7) tl[(t - 1) * 8 + 5] = clock64();
        publish_fence();
        epi_bar_all();
        if (tl && leader && t >= 1 && t - 1 < 7) tl[(t - 1) * 8 + 6] = clock64();
        if (leader) {
          red_relaxed_add(counters + t, 1);
          if (tl && t >= 1 && t - 1 < 7) tl[(t - 1) * 8 + 7] = clock64();
        }
        if (row_ok) {                                // saved for backward + the module output (off the critical path)
          float4* a4 = reinterpret_cast<float4*>(p.act + (bt0 + t) * H4 + g0);
          a4[0] = make_float4(z[0], z[1], z[2], z[3]);
          a4[1] = make_float4(z[4], z[5], z[6], z[7]);
          *reinterpret_cast<float2*>(p.c + (bt0 + t) * H + u0) = make_float2(cst[0], cst[1]);
          const bool valid = t < len_b;
          *reinterpret_cast<float2*>(p.out + (bt0 + t) * H + u0) = valid ? make_float2(h[0], h[1]) : make_float2(0.f, 0.f);
        }
      }
      if (row_ok) {                                  // steps nobody in this row tile reaches: zero output rows
        for (int t = steps; t < p.T; ++t) *reinterpret_cast<float2*>(p.out + (bt0 + t) * H + u0) = make_float2(0.f, 0.f);
      }
    } else {
      float dcn[2] = {0.f, 0.f}, dbacc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) dbacc[j] = 0.f;
      for (int t = steps - 1; t >= 0; --t) {
        const bool valid = cell_ok && t < len_b;
        float a[8], ct[2] = {0.f, 0.f}, cp[2] = {0.f, 0.f}, dh[2] = {0.f, 0.f};
        if (valid) {
          const float4* a4 = reinterpret_cast<const float4*>(p.act + (bt0 + t) * H4 + g0);
          const float4 v0 = __ldg(a4), v1 = __ldg(a4 + 1);
          a[0] = v0.x; a[1] = v0.y; a[2] = v0.z; a[3] = v0.w; a[4] = v1.x; a[5] = v1.y; a[6] = v1.z; a[7] = v1.w;
          const float2* c2 = reinterpret_cast<const float2*>(p.c + (bt0 + t) * H + u0);
          const float2 c0 = __ldg(c2);
          ct[0] = c0.x; ct[1] = c0.y;
          if (t > 0) {
            const float2 p0 = __ldg(c2 - H / 2);
            cp[0] = p0.x; cp[1] = p0.y;
          }
          const float2 d0 = __ldg(reinterpret_cast<const float2*>(p.dout + (bt0 + t) * H + u0));
          dh[0] = d0.x; dh[1] = d0.y;
        }
        if (t < steps - 1) {                         // + W_hh^T dz_{t+1}: this CTA's quarter of the contraction, then the cluster's other three
          const int n = steps - 2 - t;
          mbar_wait(tmem_full, (uint32_t)(n & 1), 15);
          tc_fence_after();
          if (tl && leader && n < 7) tl[n * 8 + 4] = clock64();
          if (reader) {
            // D[unit row of the stacked [hi ; lo] tile][2 rpt columns].  The lo half (lanes 64..127) goes through the local staging tile to the
            // thread that holds the same unit's hi half, which adds the four partial products and sends the unit's [rpt] partial sums to the CTA
            // that owns the unit's cell math (units 16 q .. 16 q + 15 of the group -> CTA q of the cluster), into ITS receive tile.
            float acc[L_RPT_MAX];                      // this thread's accumulator row, hi + lo column halves added
#pragma unroll
            for (int c = 0; c < L_RPT_MAX / 8; ++c) {
              if (c < nchunk) {
                uint32_t v[8], w2[8];
                tmem_ld8(lane_addr + (uint32_t)(8 * c), v);
                tmem_ld8(lane_addr + (uint32_t)(p.rpt + 8 * c), w2);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[8 * c + j] = __uint_as_float(v[j]) + __uint_as_float(w2[j]);
              }
            }
            tc_fence_before();
            if (lrow >= L_WROWS) {
#pragma unroll
              for (int c = 0; c < L_RPT_MAX / 8; ++c) {
                if (c < nchunk) {
                  float4* sp = reinterpret_cast<float4*>(stage + (lrow - L_WROWS) * L_RSTRIDE + 8 * c);
                  sp[0] = make_float4(acc[8 * c], acc[8 * c + 1], acc[8 * c + 2], acc[8 * c + 3]);
                  sp[1] = make_float4(acc[8 * c + 4], acc[8 * c + 5], acc[8 * c + 6], acc[8 * c + 7]);
                }
              }
            }
            reader_bar();
            if (lrow < L_WROWS) {
              const uint32_t dst_cta = (uint32_t)(lrow / L_UNITS);
              const uint32_t dst = mapa_u32(recv_base + (uint32_t)(((jg * L_UNITS + (lrow % L_UNITS)) * L_RSTRIDE) * 4), dst_cta);
#pragma unroll
              for (int c = 0; c < L_RPT_MAX / 8; ++c) {
                if (c < nchunk) {
                  const float4* sp = reinterpret_cast<const float4*>(stage + lrow * L_RSTRIDE + 8 * c);
                  const float4 s0 = sp[0], s1 = sp[1];
                  st_cluster_v4(dst + (uint32_t)(32 * c), acc[8 * c] + s0.x, acc[8 * c + 1] + s0.y, acc[8 * c + 2] + s0.z, acc[8 * c + 3] + s0.w);
                  st_cluster_v4(dst + (uint32_t)(32 * c + 16), acc[8 * c + 4] + s1.x, acc[8 * c + 5] + s1.y, acc[8 * c + 6] + s1.z, acc[8 * c + 7] + s1.w);
                }
              }
              __syncwarp();
              if (lane == 0) {                       // the warp's stores are ordered before its arrivals (release at cluster scope)
#pragma unroll
                for (uint32_t q = 0; q < (uint32_t)L_KSPLIT; ++q) mbar_arrive_cluster(mapa_u32(red_bar, q));
              }
            }
          }
          mbar_wait_cluster(red_bar, (uint32_t)(n & 1), 16);         // all four partials of this CTA's 16 units have landed
          if (valid) {
#pragma unroll
            for (int q = 0; q < L_KSPLIT; ++q) {
              dh[0] += recv[(q * L_UNITS + 2 * up) * L_RSTRIDE + b_loc];
              dh[1] += recv[(q * L_UNITS + 2 * up + 1) * L_RSTRIDE + b_loc];
            }
          }
        }
        float dz[8];
        if (valid) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
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
          for (int j = 0; j < 8; ++j) dz[j] = 0.f;
          dcn[0] = dcn[1] = 0.f;
        }
        if (cell_ok) {                               // gate gradients of step t: operand of step t - 1 and of the weight-gradient GEMMs
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) split2(dz[2 * k], dz[2 * k + 1], hi[k], lo[k]);
          __nv_bfloat16* dst = p.dgp + (bt0 + t) * H4 + g0;
          *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(dst + p.dgp_ps) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) dbacc[j] += dz[j];
        const int rn = steps - 2 - t;
        if (tl && leader && rn >= 0 && rn < 7) tl[rn * 8 + 5] = clock64();
        publish_fence();
        epi_bar_all();
        if (tl && leader && rn >= 0 && rn < 7) tl[rn * 8 + 6] = clock64();
        if (leader) {
          red_relaxed_add(counters + t, 1);
          if (tl && rn >= 0 && rn < 7) tl[rn * 8 + 7] = clock64();
        }
      }
      // bias gradient: the 8 threads of a batch row hold different unit pairs; sum over the rows of this warp (4 rows), then one atomic
      // per column and warp
      if (unit_ok && steps > 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v = dbacc[j];
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (lane < 8) atomicAdd(p.dbias + g0 + j, v);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if constexpr (BWD) cluster_sync_all();     // no peer writes into, or signals, a CTA that has left
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, L_TMEM_COLS);
  }
}

// ---- weight layout helpers.  Gate row j' = 64 * ni + 4 * u + g of the permuted matrices is row g * H + 16 * ni + u of the
// PyTorch [4H, .] layout (gates i, f, g, o stacked along rows).
__device__ __forceinline__ int perm_row(int jp, int H) {
  const int ni = jp >> 6, rem = jp & 63;
  return (rem & 3) * H + ni * L_UNITS + (rem >> 2);
}
// planes [2][H][4H] <- W_hh[perm(j')][k] at [k][j']   (the resident operand of the backward recurrence): 32 x 32 tiles through shared memory,
// rows of W_hh read along k, plane rows written along j' (both coalesced)
__global__ void __launch_bounds__(256) lstm_split_perm_t_kernel(const float* __restrict__ Whh, int H, __nv_bfloat16* __restrict__ planes,
                                                                int64_t ps, const ZeroJobs zero) {
  pdl_enter();
  zero_jobs_device(zero);          // (the backward call's accumulators and counters: one launch less)
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
  const int tiles_j = 4 * H / 32, tiles_k = (H + 31) / 32;
  for (int t = blockIdx.x; t < tiles_j * tiles_k; t += gridDim.x) {
    const int j0 = (t % tiles_j) * 32, k0 = (t / tiles_j) * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int jp = j0 + ty + 8 * i, k = k0 + tx;
      tile[ty + 8 * i][tx] = k < H ? Whh[(int64_t)perm_row(jp, H) * H + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + ty + 8 * i, jp = j0 + tx;
      if (k < H) {
        const float x = tile[tx][ty + 8 * i];
        const __nv_bfloat16 h = __float2bfloat16_rn(x);
        const int64_t o = (int64_t)k * 4 * H + jp;
        planes[o] = h;
        planes[ps + o] = __float2bfloat16_rn(x - __bfloat162float(h));
      }
    }
    __syncthreads();
  }
}
// forward prologue in one launch: planes of W_ih and W_hh (gate rows permuted) and the permuted bias sum
__global__ void __launch_bounds__(256) lstm_prep_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                                                        const float* __restrict__ b_ih, const float* __restrict__ b_hh, int H, int E,
                                                        __nv_bfloat16* __restrict__ wip, __nv_bfloat16* __restrict__ whp,
                                                        float* __restrict__ biasp, const ZeroJobs zero) {
  pdl_enter();
  zero_jobs_device(zero);          // (the h planes and the step counters of the forward call: one launch less)
  // (four consecutive columns per thread: E % 8 == 0 and H % 16 == 0, so every row is a whole number of float4)
  const int64_t n1 = (int64_t)4 * H * E, n2 = (int64_t)4 * H * H, v1 = n1 / 4, v2 = n2 / 4, total = v1 + v2 + 4 * H;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
    if (g < v1 + v2) {
      const bool first = g < v1;
      const int64_t i = 4 * (first ? g : g - v1);
      const int cols = first ? E : H;
      const float* W = first ? w_ih : w_hh;
      __nv_bfloat16* planes = first ? wip : whp;
      const int64_t ps = first ? n1 : n2;
      const int jp = (int)(i / cols), c = (int)(i - (int64_t)jp * cols);
      const float4 x4 = __ldg(reinterpret_cast<const float4*>(W + (int64_t)perm_row(jp, H) * cols + c));
      const float x[4] = {x4.x, x4.y, x4.z, x4.w};
      __nv_bfloat16 hi[4], lo[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        hi[q] = __float2bfloat16_rn(x[q]);
        lo[q] = __float2bfloat16_rn(x[q] - __bfloat162float(hi[q]));
      }
      *reinterpret_cast<uint2*>(planes + i) = *reinterpret_cast<const uint2*>(hi);
      *reinterpret_cast<uint2*>(planes + ps + i) = *reinterpret_cast<const uint2*>(lo);
    } else {
      const int jp = (int)(g - v1 - v2);
      const int j = perm_row(jp, H);
      biasp[jp] = b_ih[j] + b_hh[j];
    }
  }
}
// dW_ih, dW_hh and the bias gradient (both b_ih and b_hh receive it) in one launch
__global__ void __launch_bounds__(256) lstm_unperm3_kernel(const float* __restrict__ dwc /* [4H][E + H]: dW_ih' | dW_hh' */,
                                                           const float* __restrict__ dbp, int H, int E, float* __restrict__ dw_ih,
                                                           float* __restrict__ dw_hh, float* __restrict__ db_ih, float* __restrict__ db_hh) {
  pdl_enter();
  const int64_t n1 = (int64_t)4 * H * E, n2 = (int64_t)4 * H * H, v1 = n1 / 4, v2 = n2 / 4, total = v1 + v2 + 4 * H;      // (float4 per thread)
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
    if (g < v1) {
      const int64_t i = 4 * g;
      const int jp = (int)(i / E), c = (int)(i - (int64_t)jp * E);
      *reinterpret_cast<float4*>(dw_ih + (int64_t)perm_row(jp, H) * E + c) = *reinterpret_cast<const float4*>(dwc + (int64_t)jp * (E + H) + c);
    } else if (g < v1 + v2) {
      const int64_t i = 4 * (g - v1);
      const int jp = (int)(i / H), c = (int)(i - (int64_t)jp * H);
      *reinterpret_cast<float4*>(dw_hh + (int64_t)perm_row(jp, H) * H + c) = *reinterpret_cast<const float4*>(dwc + (int64_t)jp * (E + H) + E + c);
    } else {
      const int jp = (int)(g - v1 - v2);
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
  __nv_bfloat16* xh;       // [2][B][T][E + H]: x planes in columns [0, E), h planes (slot t = h_{t-1}) in [E, E + H): ONE operand for both
                           // weight gradients of backward (dW_ih' | dW_hh' = dz^T [x | h])
  __nv_bfloat16* wip;      // [2][4H][E]: gate-permuted planes of w_ih (the input-gradient product of backward reads them again)
};
size_t saved_bytes(int B, int T, int E, int H) {
  const size_t BT = (size_t)B * T;
  return align_up(BT * 4 * H * 4) + align_up(BT * H * 4) + align_up(2 * BT * (E + H) * 2) + align_up((size_t)2 * 4 * H * E * 2) + 256;
}
bool carve_saved(Saved& s, void* buf, size_t bytes, int B, int T, int E, int H) {
  if (bytes < saved_bytes(B, T, E, H) || (reinterpret_cast<uintptr_t>(buf) & 255) != 0) return false;
  const size_t BT = (size_t)B * T;
  char* p = (char*)buf;
  s.act = (float*)p; p += align_up(BT * 4 * H * 4);
  s.c = (float*)p; p += align_up(BT * H * 4);
  s.xh = (__nv_bfloat16*)p; p += align_up(2 * BT * (E + H) * 2);
  s.wip = (__nv_bfloat16*)p;
  return true;
}

bool shape_ok(int B, int T, int E, int H) { return B > 0 && T > 0 && E > 0 && E % 8 == 0 && H >= 16 && H % 16 == 0 && H <= 512; }

TcOperand operand(const __nv_bfloat16* planes, int64_t ld, int64_t ps, int rows, int cols, bool mn_major) {
  TcOperand o;
  o.planes = planes; o.ld = ld; o.plane_stride = ps; o.rows = rows; o.cols = cols; o.mn_major = mn_major;
  return o;
}

// ---- launch geometry -------------------------------------------------------------------------------------------------
// Every CTA of a launch waits for the other CTAs of its row tile (and, backward, of its cluster), so a launch may only hold as many
// row tiles as the device can keep RESIDENT at once: the occupancy API says how many CTAs (forward) / clusters of 4 (backward) that
// is for this kernel's shared-memory footprint; a shape that does not even fit one row tile is refused (hca_lstm_supported = 0,
// the module then falls back to cuDNN) instead of being launched into a spin.  Row tiles are independent: when the batch needs
// more tiles than fit, the launches simply follow each other.  Within that limit the tiles are made as short as the CTA count
// allows (the per-step cost of a CTA grows with the rows it streams).
struct Geometry {
  int ctas_per_tile = 0, per_launch = 0, rpt = 0, tiles = 0, kbn = 0;
  size_t smem = 0;
};
// the streamed operand of one step (kbn k-blocks of [x_hi ; x_lo] rows), or the staging tile that aliases it if that is larger
size_t ring_bytes(bool bwd, int kbn, int rpt) {
  const size_t ring = (size_t)kbn * rpt * 2 * L_BK * 2;
  const size_t stage = bwd ? (size_t)L_WROWS * L_RSTRIDE * 4 : (size_t)rpt * L_PSTRIDE * 4;
  return (std::max(ring, stage) + 1023) / 1024 * 1024;
}
size_t rec_smem_bytes(bool bwd, int kbn, int rpt) {
  return (size_t)kbn * L_WKB_BYTES + ring_bytes(bwd, kbn, rpt) + (bwd ? L_RECV_BYTES : 0) + 1024;
}
template <bool BWD>
int resident_ctas(size_t smem) {
  // (cached per device and footprint: the answer only depends on those)
  static int cache[64][2] = {};
  static size_t cache_smem[64][2] = {};
  const int dev = current_device();
  if (cache[dev][BWD ? 1 : 0] > 0 && cache_smem[dev][BWD ? 1 : 0] == smem) return cache[dev][BWD ? 1 : 0];
  if (cudaFuncSetAttribute(lstm_rec_kernel<BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, L_MAX_SMEM) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int n = 0;
  if constexpr (BWD) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(L_KSPLIT * 64);
    cfg.blockDim = dim3(L_THREADS);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = L_KSPLIT; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&clusters, lstm_rec_kernel<true>, &cfg) != cudaSuccess) { cudaGetLastError(); clusters = 0; }
    n = clusters * L_KSPLIT;
  } else {
    int per_sm = 0, sms = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lstm_rec_kernel<false>, L_THREADS, smem) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    n = per_sm * sms;
  }
  { const int cap = HCA_ENV_INT("HCA_LSTM_MAX_CTAS", 0); if (cap > 0 && cap < n) n = cap; }      // tests: pretend the device is smaller
  cache[dev][BWD ? 1 : 0] = n;
  cache_smem[dev][BWD ? 1 : 0] = smem;
  return n;
}
template <bool BWD>
bool geometry(Geometry& g, int B, int H) {
  g.kbn = (H + L_BK - 1) / L_BK;
  g.ctas_per_tile = BWD ? L_KSPLIT * ((H + L_KSPLIT * L_UNITS - 1) / (L_KSPLIT * L_UNITS)) : H / L_UNITS;
  // footprint at the longest row tile decides residency (a shorter tile only makes it smaller)
  const int resident = resident_ctas<BWD>(rec_smem_bytes(BWD, g.kbn, L_RPT_MAX));
  g.per_launch = resident / g.ctas_per_tile;
  if (g.per_launch < 1) return false;
  const int min_tiles = (B + L_RPT_MAX - 1) / L_RPT_MAX;
  const int launches = (min_tiles + g.per_launch - 1) / g.per_launch;
  const int slots = launches * g.per_launch;
  g.rpt = std::min(L_RPT_MAX, (int)(((B + slots - 1) / slots + 7) / 8 * 8));
  g.tiles = (B + g.rpt - 1) / g.rpt;
  g.smem = rec_smem_bytes(BWD, g.kbn, g.rpt);
  return g.smem <= (size_t)L_MAX_SMEM;
}
size_t counter_count(int B, int T) { return (size_t)((B + 7) / 8 + 1) * T; }

template <bool BWD>
int launch_rec(const LstmParams& base, const __nv_bfloat16* stream_planes, int64_t stream_ps, int stream_cols, int64_t stream_ld,
               const __nv_bfloat16* w_planes, int64_t w_ps, int w_rows, int w_cols, cudaStream_t s) {
  Geometry g;
  if (!geometry<BWD>(g, base.B, base.H))
    return set_err(HCA_ERR_ARG, "lstm: hidden size %d needs %d co-resident CTAs per row tile, more than this device can hold", base.H, g.ctas_per_tile);
  LstmParams p = base;
  p.timeline = g_lstm_timeline;
  p.K = base.H;
  p.kbn = g.kbn;
  p.rpt = g.rpt;
  p.ctas_per_tile = g.ctas_per_tile;
  p.ring_bytes = (uint32_t)ring_bytes(BWD, g.kbn, g.rpt);
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * g.rpt) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  LstmMaps maps;
  {  // streamed operand [2][B][T][cols] as (col, b, plane, t); a box brings one k-block of both planes for the rows of a tile.  Rows
     // beyond B and columns beyond `cols` arrive as zeros (out-of-bounds fill).
    const uint64_t dims[4] = {(uint64_t)stream_cols, (uint64_t)base.B, 2, (uint64_t)base.T};
    const uint64_t str[3] = {(uint64_t)base.T * stream_ld * 2, (uint64_t)stream_ps * 2, (uint64_t)stream_ld * 2};
    const uint32_t box[4] = {L_BK, (uint32_t)g.rpt, 2, 1};
    HCA_TRY(tc_make_tmap(&maps.X, true, 4, stream_planes, dims, str, box, 3));
  }
  {  // resident operand [2][rows][cols]: dims (cols, rows, 2, 1)
    const uint64_t dims[4] = {(uint64_t)w_cols, (uint64_t)w_rows, 2, 1};
    const uint64_t str[3] = {(uint64_t)w_cols * 2, (uint64_t)w_ps * 2, (uint64_t)w_ps * 4};
    const uint32_t box[4] = {L_BK, L_WROWS, 1, 1};
    HCA_TRY(tc_make_tmap(&maps.W, true, 4, w_planes, dims, str, box, 3));
  }
  static bool attr_set[64][2] = {};
  bool& attr_done = attr_set[current_device()][BWD ? 1 : 0];
  if (!attr_done) {
    HCA_CUDA(cudaFuncSetAttribute(lstm_rec_kernel<BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, L_MAX_SMEM));
    attr_done = true;
  }
  for (int t0 = 0; t0 < g.tiles; t0 += g.per_launch) {
    const int nm = std::min(g.per_launch, g.tiles - t0);
    p.tile0 = t0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(nm * g.ctas_per_tile));
    cfg.blockDim = dim3(L_THREADS);
    cfg.dynamicSmemBytes = g.smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[3];
    int na = 0;
    if (BWD) {
      attr[na].id = cudaLaunchAttributeClusterDimension;
      attr[na].val.clusterDim.x = L_KSPLIT; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
      ++na;
    }
    if (HCA_ENV_INT("HCA_LSTM_COOP", 0) != 0) {       // experiment: let the driver gang-schedule the grid (cooperative launch)
      attr[na].id = cudaLaunchAttributeCooperative;
      attr[na].val.cooperative = 1;
      ++na;
    } else if (pdl_enabled()) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    if (g_rec_events[BWD ? 1 : 0][0] && t0 == 0) HCA_CUDA(cudaEventRecord(g_rec_events[BWD ? 1 : 0][0], s));
    HCA_CUDA(cudaLaunchKernelEx(&cfg, lstm_rec_kernel<BWD>, maps, p));
    HCA_LAUNCHED();
    if (g_rec_events[BWD ? 1 : 0][1] && t0 + g.per_launch >= g.tiles) HCA_CUDA(cudaEventRecord(g_rec_events[BWD ? 1 : 0][1], s));
  }
  return 0;
}

bool device_holds(int B, int H) {
  Geometry g;
  return geometry<false>(g, B, H) && geometry<true>(g, B, H);
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

extern "C" int hca_lstm_supported(int B, int T, int E, int H) {
  return hca::shape_ok(B, T, E, H) && hca::tc_available() && hca::device_holds(B, H) ? 1 : 0;
}

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
  const int64_t xh_ld = E + H, xh_ps = BT * xh_ld;
  __nv_bfloat16* const hp = sv.xh + E;
  {
    ZeroBatch zb(s);
    HCA_TRY(zb.add(sv.xh, (size_t)2 * xh_ps * 2));             // the h columns: slot 0 (h_{-1} = 0) and the slots no step reaches (the x columns follow)
    HCA_TRY(zb.add(counters, counter_count(B, T) * 4));
    HCA_LAUNCH_K((lstm_prep_kernel), ew_grid((int64_t)H4 * ((E + H) / 4 + 1)), 256, 0, s, w_ih, w_hh, b_ih, b_hh, H, E, wip, whp, biasp, zb.take());
    HCA_LAUNCHED();
  }
  HCA_TRY(launch_split_planes(x, E, BT, E, sv.xh, xh_ld, xh_ps, 2, s));
  {  // x-projection of every (b, t), gate columns in [unit][gate] order, biases folded in
    TcEpilogue e;
    e.D = sv.act; e.ldd = H4; e.bias = biasp;
    HCA_TRY(launch_gemm_tc(operand(sv.xh, xh_ld, xh_ps, (int)BT, E, false), operand(wip, E, (int64_t)H4 * E, H4, E, false), 2, (int)BT, H4, E,
                           e, 1, s));
  }
  LstmParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.T = T; p.H = H; p.lens = lens; p.counters = counters;
  p.act = sv.act; p.c = sv.c; p.out = out; p.hp = hp; p.hp_ps = xh_ps; p.hp_ld = xh_ld;
  return launch_rec<false>(p, hp, xh_ps, H, xh_ld, whp, (int64_t)H4 * H, H4, H, s);
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
  float* dwc = w.take<float>((size_t)H4 * (E + H));           // [4H][E + H]: dW_ih' | dW_hh' (gate rows still permuted)
  float* dbp = w.take<float>((size_t)H4);
  int* counters = w.take<int>(counter_count(B, T));
  if (!counters) return set_err(HCA_ERR_WORKSPACE, "lstm_bwd: workspace too small (%zu bytes)", ws_bytes);
  const int sk_w = tc_splitk(H4, E + H, (int)BT);
  {
    ZeroBatch zb(s);
    HCA_TRY(zb.add(dgp, (size_t)2 * BT * H4 * 2));             // rows no step writes must read as zero in the GEMMs below
    HCA_TRY(zb.add(dbp, (size_t)H4 * 4));
    HCA_TRY(zb.add(counters, counter_count(B, T) * 4));
    if (sk_w > 1) HCA_TRY(zb.add(dwc, (size_t)H4 * (E + H) * 4));
    HCA_LAUNCH_K((lstm_split_perm_t_kernel), std::min(148 * 8, (H4 / 32) * ((H + 31) / 32)), 256, 0, s, w_hh, H, wtp, (int64_t)H4 * H, zb.take());
    HCA_LAUNCHED();
  }
  LstmParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.T = T; p.H = H; p.lens = lens; p.counters = counters;
  p.act = sv.act; p.c = sv.c; p.dout = dout; p.dgp = dgp; p.dgp_ps = BT * H4; p.dbias = dbp;
  HCA_TRY(launch_rec<true>(p, dgp, BT * H4, H4, H4, wtp, (int64_t)H4 * H, H, H4, s));
  const TcOperand dg_mn = operand(dgp, H4, BT * H4, (int)BT, H4, true);
  {  // [dW_ih' | dW_hh'] = dz^T [x | h_prev]: ONE product (K = B*T, split-K) over the planes the forward pass laid out side by side
    TcEpilogue e; e.D = dwc; e.ldd = E + H;
    HCA_TRY(launch_gemm_tc(dg_mn, operand(sv.xh, E + H, BT * (E + H), (int)BT, E + H, true), 2, H4, E + H, (int)BT, e, sk_w, s));
  }
  // the three gradients back to PyTorch's gate order, one launch
  HCA_LAUNCH_K((lstm_unperm3_kernel), ew_grid((int64_t)H4 * ((E + H) / 4 + 1)), 256, 0, s, dwc, dbp, H, E, dw_ih, dw_hh, db_ih, db_hh);
  HCA_LAUNCHED();
  if (dx) {  // dx = dz W_ih
    TcEpilogue e; e.D = dx; e.ldd = E;
    HCA_TRY(launch_gemm_tc(operand(dgp, H4, BT * H4, (int)BT, H4, false), operand(wip, E, (int64_t)H4 * E, H4, E, true), 2, (int)BT, E, H4,
                           e, 1, s));
  }
  return 0;
}
