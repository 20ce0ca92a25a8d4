// Shared helpers for the hiecoattn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstdlib>
#include <atomic>

#include "../../include/hiecoattn_b200.h"

namespace hca {

// ---- error plumbing (thread-local message, C ABI never throws) -------------------------------------
char* err_buf();
int set_err(int code, const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

#define HCA_CHECK_ARG(cond, ...)                                    \
  do {                                                              \
    if (!(cond)) return ::hca::set_err(HCA_ERR_ARG, __VA_ARGS__);   \
  } while (0)

#define HCA_CUDA(expr)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return ::hca::set_err(HCA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                            __FILE__, __LINE__);                                              \
  } while (0)

// call right after a <<<>>> launch
#define HCA_LAUNCHED()                                                                         \
  do {                                                                                         \
    ::hca::g_launches.fetch_add(1, std::memory_order_relaxed);                                 \
    cudaError_t _e = cudaPeekAtLastError();                                                    \
    if (_e != cudaSuccess) {                                                                   \
      cudaGetLastError();                                                                      \
      return ::hca::set_err(HCA_ERR_CUDA, "kernel launch failed: %s (%s:%d)",                  \
                            cudaGetErrorString(_e), __FILE__, __LINE__);                       \
    }                                                                                          \
  } while (0)

#define HCA_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != 0) return _rc;    \
  } while (0)

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
// grid size for grid-stride elementwise kernels: enough blocks to fill 148 SMs a few times over
static inline int ew_grid(int64_t total) {
  const int64_t g = (total + 255) / 256;
  return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

// bump allocator over the caller's workspace
struct Workspace {
  char* base;
  size_t cap, off;
  Workspace(void* p, size_t bytes) : base((char*)p), cap(bytes), off(0) {}
  template <class T>
  T* take(size_t n) {
    size_t b = align_up(n * sizeof(T));
    if (off + b > cap) return nullptr;
    T* r = (T*)(base + off);
    off += b;
    return r;
  }
};

// ---- device helpers ----------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Tuning knobs for profiling experiments come from the environment and are read ONCE per process (function-local static), never
// on the per-launch path: HCA_ENV_INT("NAME", default).
#define HCA_ENV_INT(name, dflt)                    \
  ([]() -> int {                                   \
    static const int v = []() -> int {             \
      const char* e_ = getenv(name);               \
      return e_ ? atoi(e_) : (dflt);               \
    }();                                           \
    return v;                                      \
  }())
int current_device();      // cudaGetDevice, clamped to [0, 64)

// options
bool pdl_enabled();      // HCA_PDL=0 / hca_set_option("pdl", "0") turns programmatic dependent launch off (plain stream-ordered launches)

// ---- programmatic dependent launch -------------------------------------------------------------------
// The step is a chain of ~90 kernels, many of them a few microseconds long: the gap between a kernel's last CTA retiring and
// the first CTA of its successor running (launch latency, block scheduling, parameter fetch) is a visible share of it.  Every
// kernel of the library is launched with cudaLaunchAttributeProgrammaticStreamSerialization and starts with pdl_enter():
// `launch_dependents` lets the NEXT kernel's CTAs be scheduled as soon as all of this kernel's CTAs have started (they then
// sit in their own pdl_enter), and `wait` blocks until every prerequisite grid has completed and its writes are visible --
// i.e. exactly stream order for all memory effects, minus the launch gap.  Under CUDA-graph capture the edges become
// programmatic edges of the graph.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <class... KArgs, class... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// kernel in parentheses (template commas): HCA_LAUNCH_K((k<1, 2>), grid, block, smem, stream, args...); follow with HCA_LAUNCHED()
#define HCA_LAUNCH_K(kernel, grid, block, smem, stream, ...) \
  (void)::hca::launch_k(kernel, dim3(grid), dim3(block), (size_t)(smem), stream, __VA_ARGS__)

// ---- side stream -------------------------------------------------------------------------------------
// Independent branches of one entry point (e.g. the weight-gradient products of the classifier, which nothing later in the
// call consumes) run on a per-device helper stream: fork() makes the helper wait for the work enqueued on `main` so far,
// join() makes `main` wait for the helper.  Both are event record / wait pairs, so under CUDA-graph capture they become
// parallel branches of the graph.  The helper is created on the first call made OUTSIDE a capture; until then (or with
// HCA_SIDE_STREAM=0) `stream()` is `main` itself and everything stays serial.
struct SideStream {
  explicit SideStream(cudaStream_t main);
  cudaStream_t stream() const { return side_ ? side_ : main_; }
  int fork();                 // helper waits for main (call again after each producer the helper depends on)
  int join();                 // main waits for the helper
 private:
  cudaStream_t main_, side_ = nullptr;
  cudaEvent_t ev_[2] = {nullptr, nullptr};
};

}  // namespace hca
