// Data-parallel tail of the training step as ONE kernel over NVLink 5 / NVSwitch: gradient all-reduce + Adam + parameter
// broadcast (the reference has no multi-GPU path at all, main.py:102-106 is a TODO; the single-GPU update it implies is
// torch.optim.Adam, main.py:180,222).
//
// Every rank keeps its flat gradient buffer g and its flat parameter buffer p in SYMMETRIC memory (same size and layout on
// every GPU, mapped into every peer's address space, optionally also bound to an NVSwitch multicast object).  Rank r owns
// the r-th slice of the element range.  For its slice the kernel
//     1. waits until every rank's gradients are complete                       (flag barrier over peer memory)
//     2. g_sum = sum over ranks of g     -- multimem.ld_reduce.add.v4.f32: the switch adds the eight copies in flight (NVLS);
//                                           without a multicast mapping: eight 16-byte peer loads, added in rank order
//     3. Adam on (p, g_sum, m, v) of the slice; m / v exist only for the owner's slice (sharded optimizer state)
//     4. p_new -> every rank             -- multimem.st.v4.f32 (one store, the switch replicates); else eight peer stores
//     5. waits until every rank's stores have landed                            (second flag barrier)
// so the gradients cross the fabric once (reduce-scatter half), the parameters once (all-gather half), nothing is staged in
// HBM in between and the optimizer costs no pass of its own: 2 x 48.7 MB over the links instead of NCCL's all-reduce of
// 48.7 MB followed by a 350 MB Adam pass per GPU.  Replicas stay bit-identical: every element is updated by exactly one rank.
//
// Flags: one 32-bit word per (channel, phase, CTA, peer) in the symmetric block, driven by compare-and-swap (0 -> 1 by the
// signalling rank with release semantics, 1 -> 0 by the waiting rank with acquire semantics), so they re-arm themselves and the
// kernel can be replayed from a CUDA graph.  CTA b of a rank pairs only with CTA b of the other ranks; the grid never exceeds
// the SM count, so all CTAs are resident and no CTA waits for one that cannot be scheduled.  Waits are bounded: a rank that
// never arrives turns into a trapped kernel (a CUDA error on the host), not a hang.
#include <algorithm>
#include <cstring>
#include "common.cuh"

namespace hca {
namespace {

constexpr int DP_MAX_WORLD = 8;
constexpr int DP_MAX_GRID = 160;                  // CTAs per launch (<= SM count)
constexpr int DP_CHANNELS = 4;                    // independent flag sets: concurrent launches on different streams
constexpr long long DP_TIMEOUT = 6000000000LL;    // ~3 s of SM clocks

struct DpArgs {
  unsigned long long peer[DP_MAX_WORLD];   // base address of every rank's symmetric block, as mapped on THIS device
  unsigned long long mc;                   // multicast address of the block (0: no NVLS, use the peer mappings)
  unsigned long long flags_off, g_off, p_off;   // byte offsets inside the block
  long long begin4, end4;                  // float4 range of this launch (whole ranks' slices are cut from it)
  int rank, world, channel, mode;          // mode bit 0: Adam update + parameter broadcast ; bit 1: write g_sum back to every rank
  float* m;                                // [n] first moment  (only the owner's slice is used)
  float* v;                                // [n] second moment
  const float* coef;                       // [2] step size, 1 / sqrt(bias correction 2)  (adam_prep_kernel)
  float b1, b2, eps;
};

__device__ __forceinline__ unsigned cas_release_sys(unsigned* p, unsigned cmp, unsigned val) {
  unsigned old;
  asm volatile("atom.release.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(p), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ unsigned cas_acquire_sys(unsigned* p, unsigned cmp, unsigned val) {
  unsigned old;
  asm volatile("atom.acquire.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(p), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4* mc) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(mc)
               : "memory");
  return r;
}
__device__ __forceinline__ void multimem_st(float4* mc, const float4& x) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
}
__device__ __forceinline__ float4 ld_peer(const float4* p) {
  float4 r;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_peer(float4* p, const float4& x) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
}

// All ranks' CTA `blockIdx.x` meet here.  Thread q < world signals rank q (sets ITS flag word for (phase, this CTA, my rank)) and then
// consumes the word rank q set for me.  The __syncthreads() before makes the whole CTA's prior work part of what the release publishes.
__device__ __forceinline__ void cta_rank_barrier(const DpArgs& a, int phase) {
  __syncthreads();
  const int q = threadIdx.x;
  if (q < a.world) {
    const size_t word = (((size_t)a.channel * 2 + phase) * DP_MAX_GRID + blockIdx.x) * DP_MAX_WORLD;
    unsigned* theirs = reinterpret_cast<unsigned*>(a.peer[q] + a.flags_off) + word + a.rank;
    unsigned* mine = reinterpret_cast<unsigned*>(a.peer[a.rank] + a.flags_off) + word + q;
    const long long t0 = clock64();
    while (cas_release_sys(theirs, 0u, 1u) != 0u) {         // (still set: the peer has not consumed the previous round yet)
      if (clock64() - t0 > DP_TIMEOUT) { printf("hiecoattn dp: rank %d CTA %d timed out signalling rank %d (phase %d)\n", a.rank, blockIdx.x, q, phase); __trap(); }
    }
    while (cas_acquire_sys(mine, 1u, 0u) != 1u) {
      if (clock64() - t0 > DP_TIMEOUT) { printf("hiecoattn dp: rank %d CTA %d timed out waiting for rank %d (phase %d)\n", a.rank, blockIdx.x, q, phase); __trap(); }
    }
  }
  __syncthreads();
}

template <bool MC>
__global__ void __launch_bounds__(256) dp_reduce_adam_kernel(const DpArgs a) {
  pdl_enter();
  cta_rank_barrier(a, 0);                                  // every rank's backward has written its gradients
  // this rank's slice of [begin4, end4)
  const long long span = a.end4 - a.begin4, per = (span + a.world - 1) / a.world;
  const long long lo = a.begin4 + per * a.rank, hi = min(a.end4, lo + per);
  const bool adam = (a.mode & 1) != 0, wb = (a.mode & 2) != 0;
  float step_size = 0.f, inv_bc2 = 0.f;
  if (adam) { step_size = a.coef[0]; inv_bc2 = a.coef[1]; }
  const float c1 = 1.f - a.b1, c2 = 1.f - a.b2;
  float4* const p_loc = reinterpret_cast<float4*>(a.peer[a.rank] + a.p_off);
  float4* const m4 = reinterpret_cast<float4*>(a.m);
  float4* const v4 = reinterpret_cast<float4*>(a.v);
  constexpr int UNR = 4;                                   // 4 x 16 bytes in flight per thread on the fabric
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += stride * UNR) {
    float4 g[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long i = i0 + u * stride;
      if (i >= hi) break;
      if (MC) {
        g[u] = multimem_ld_reduce_add(reinterpret_cast<const float4*>(a.mc + a.g_off) + i);
      } else {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int q = 0; q < a.world; ++q) {
          const float4 t = ld_peer(reinterpret_cast<const float4*>(a.peer[q] + a.g_off) + i);
          s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
        }
        g[u] = s;
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long i = i0 + u * stride;
      if (i >= hi) break;
      const float4 gg = g[u];
      if (wb) {
        if (MC) multimem_st(reinterpret_cast<float4*>(a.mc + a.g_off) + i, gg);
        else for (int q = 0; q < a.world; ++q) st_peer(reinterpret_cast<float4*>(a.peer[q] + a.g_off) + i, gg);
      }
      if (adam) {
        float4 pp = p_loc[i], mm = m4[i], vv = v4[i];
#define HCA_ADAM1(c)                                         \
  mm.c = fmaf(a.b1, mm.c, c1 * gg.c);                        \
  vv.c = fmaf(a.b2, vv.c, c2 * gg.c * gg.c);                 \
  pp.c -= step_size * (mm.c / fmaf(sqrtf(vv.c), inv_bc2, a.eps));
        HCA_ADAM1(x) HCA_ADAM1(y) HCA_ADAM1(z) HCA_ADAM1(w)
#undef HCA_ADAM1
        m4[i] = mm; v4[i] = vv;
        if (MC) multimem_st(reinterpret_cast<float4*>(a.mc + a.p_off) + i, pp);
        else for (int q = 0; q < a.world; ++q) st_peer(reinterpret_cast<float4*>(a.peer[q] + a.p_off) + i, pp);
      }
    }
  }
  __threadfence_system();                                  // this thread's stores are performed at system scope ...
  cta_rank_barrier(a, 1);                                  // ... before any rank is told that this CTA is done
}

__global__ void adam_prep_only_kernel(long long* __restrict__ step, float* __restrict__ coef, float lr, float b1, float b2) {
  pdl_enter();
  const long long t = step[0] + 1;
  step[0] = t;
  coef[0] = (float)((double)lr / (1.0 - pow((double)b1, (double)t)));
  coef[1] = (float)(1.0 / sqrt(1.0 - pow((double)b2, (double)t)));
}

}  // namespace
}  // namespace hca

extern "C" size_t hca_dp_flags_bytes(void) {
  return (size_t)hca::DP_CHANNELS * 2 * hca::DP_MAX_GRID * hca::DP_MAX_WORLD * sizeof(unsigned);
}

extern "C" int hca_adam_prep(long long* step, float* coef, float lr, float beta1, float beta2, void* stream) {
  using namespace hca;
  HCA_CHECK_ARG(step && coef, "adam_prep: null pointer");
  HCA_LAUNCH_K((adam_prep_only_kernel), 1, 1, 0, (cudaStream_t)stream, step, coef, lr, beta1, beta2);
  HCA_LAUNCHED();
  return 0;
}

extern "C" int hca_dp_reduce_adam(const uint64_t* peer_bases, uint64_t mc_base, size_t flags_off, size_t g_off, size_t p_off,
                                  int64_t begin, int64_t end, int rank, int world, int channel, int max_ctas, float* m, float* v,
                                  const float* coef, float beta1, float beta2, float eps, int mode, void* stream) {
  using namespace hca;
  HCA_CHECK_ARG(peer_bases && world >= 2 && world <= DP_MAX_WORLD && rank >= 0 && rank < world, "dp_reduce_adam: world size 2..%d, rank inside it", DP_MAX_WORLD);
  HCA_CHECK_ARG(channel >= 0 && channel < DP_CHANNELS, "dp_reduce_adam: channel 0..%d", DP_CHANNELS - 1);
  HCA_CHECK_ARG(begin >= 0 && end > begin && begin % 4 == 0 && end % 4 == 0, "dp_reduce_adam: [begin, end) must be a non-empty range of whole float4s");
  HCA_CHECK_ARG((flags_off % 16) == 0 && (g_off % 16) == 0 && (p_off % 16) == 0, "dp_reduce_adam: offsets must be 16-byte aligned");
  HCA_CHECK_ARG((mode & 3) != 0, "dp_reduce_adam: mode must request the Adam update (1), the gradient write-back (2) or both");
  HCA_CHECK_ARG(!(mode & 1) || (m && v && coef), "dp_reduce_adam: the Adam update needs m, v and coef");
  DpArgs a;
  memset(&a, 0, sizeof(a));
  for (int q = 0; q < world; ++q) {
    HCA_CHECK_ARG(peer_bases[q] != 0 && (peer_bases[q] % 16) == 0, "dp_reduce_adam: peer base %d is null or misaligned", q);
    a.peer[q] = peer_bases[q];
  }
  a.mc = mc_base;
  a.flags_off = flags_off; a.g_off = g_off; a.p_off = p_off;
  a.begin4 = begin / 4; a.end4 = end / 4;
  a.rank = rank; a.world = world; a.channel = channel; a.mode = mode;
  a.m = m; a.v = v; a.coef = coef; a.b1 = beta1; a.b2 = beta2; a.eps = eps;
  int sms = 148;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, current_device()) != cudaSuccess || sms <= 0) { cudaGetLastError(); sms = 148; }
  // every rank must launch the SAME grid (CTA b pairs with CTA b): it depends only on the range and on the caller's cap
  const long long per = ((a.end4 - a.begin4) + world - 1) / world;
  long long want = (per + 256 * 4 - 1) / (256 * 4);
  int cap = std::min(std::min(sms, DP_MAX_GRID), max_ctas > 0 ? max_ctas : DP_MAX_GRID);
  const int grid = (int)std::max(1LL, std::min<long long>(want, cap));
  if (mc_base) HCA_LAUNCH_K((dp_reduce_adam_kernel<true>), grid, 256, 0, (cudaStream_t)stream, a);
  else HCA_LAUNCH_K((dp_reduce_adam_kernel<false>), grid, 256, 0, (cudaStream_t)stream, a);
  HCA_LAUNCHED();
  return 0;
}
