// Error plumbing, launch counter and runtime options of the C ABI (include/hiecoattn_b200.h).
#include "common.cuh"
#include <cstring>
#include <cstdlib>

namespace hca {

std::atomic<int64_t> g_launches{0};

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

// HCA_PDL=0 turns programmatic dependent launch off (plain stream-ordered launches).  The environment is read once; a profiler leg
// that wants exclusive kernel durations switches at run time through hca_set_option("pdl", "0" / "1").
static std::atomic<int> g_pdl{-1};
bool pdl_enabled() {
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("HCA_PDL");
    v = (e && atoi(e) == 0) ? 0 : 1;
    g_pdl.store(v);
  }
  return v == 1;
}
void set_pool_tie_cap(int n);      // phrase_conv_pool.cu

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  return dev < 0 ? 0 : (dev > 63 ? 63 : dev);
}

namespace {
struct SideRes {
  cudaStream_t s = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  bool ok = false, tried = false;
};
SideRes g_side[64];
}  // namespace

SideStream::SideStream(cudaStream_t main) : main_(main) {
  if (HCA_ENV_INT("HCA_SIDE_STREAM", 1) == 0) return;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return; }
  SideRes& r = g_side[dev];
  if (!r.tried) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(main, &st) != cudaSuccess) { cudaGetLastError(); return; }
    if (st != cudaStreamCaptureStatusNone) return;        // never create resources inside a capture: stay serial this time
    r.tried = true;
    r.ok = cudaStreamCreateWithFlags(&r.s, cudaStreamNonBlocking) == cudaSuccess &&
           cudaEventCreateWithFlags(&r.ev[0], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&r.ev[1], cudaEventDisableTiming) == cudaSuccess;
    if (!r.ok) cudaGetLastError();
  }
  if (r.ok) { side_ = r.s; ev_[0] = r.ev[0]; ev_[1] = r.ev[1]; }
}

int SideStream::fork() {
  if (!side_) return 0;
  HCA_CUDA(cudaEventRecord(ev_[0], main_));
  HCA_CUDA(cudaStreamWaitEvent(side_, ev_[0], 0));
  return 0;
}

int SideStream::join() {
  if (!side_) return 0;
  HCA_CUDA(cudaEventRecord(ev_[1], side_));
  HCA_CUDA(cudaStreamWaitEvent(main_, ev_[1], 0));
  return 0;
}

}  // namespace hca

extern "C" {

int hca_abi_version(void) { return HCA_ABI_VERSION; }
const char* hca_last_error(void) { return hca::err_buf(); }
int64_t hca_launch_count(void) { return hca::g_launches.load(); }

int hca_set_option(const char* name, const char* value) {
  if (!name || !value) return hca::set_err(HCA_ERR_ARG, "hca_set_option: null argument");
  if (strcmp(name, "pdl") == 0) { hca::g_pdl.store(atoi(value) != 0 ? 1 : 0); return 0; }
  if (strcmp(name, "pool_tie_cap") == 0) { hca::set_pool_tie_cap(atoi(value)); return 0; }
  if (strcmp(name, "gemm") == 0 && strcmp(value, "tc") == 0) return 0;      // the only contraction backend there is
  return hca::set_err(HCA_ERR_ARG, "hca_set_option: unknown option %s=%s", name, value);
}

// ---- host staging buffers ---------------------------------------------------------------------------------------
// Page-locked host memory for the loader -> GPU hand-over of a step's inputs (reference main.py:205-208 copies pageable tensors with
// .to(device)).  write_combined != 0 asks for write-combined pages: the CPU only ever WRITES a staging buffer (the loader fills it
// once), and uncached write-combined pages are not snooped on the device's reads, which is what limits eight GPUs pulling 64 MB each
// per step from one host.
int hca_pinned_alloc(size_t bytes, int write_combined, void** out) {
  if (!out || bytes == 0) return hca::set_err(HCA_ERR_ARG, "pinned_alloc: bad arguments");
  void* p = nullptr;
  const unsigned flags = cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0u);
  cudaError_t e = cudaHostAlloc(&p, bytes, flags);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return hca::set_err(HCA_ERR_CUDA, "cudaHostAlloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
  }
  *out = p;
  return 0;
}
int hca_pinned_free(void* p) {
  if (!p) return 0;
  cudaError_t e = cudaFreeHost(p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return hca::set_err(HCA_ERR_CUDA, "cudaFreeHost failed: %s", cudaGetErrorString(e));
  }
  return 0;
}

const char* hca_get_option(const char* name) {
  if (name && strcmp(name, "gemm") == 0) return "tc";
  if (name && strcmp(name, "pdl") == 0) return hca::pdl_enabled() ? "1" : "0";
  return "";
}

}  // extern "C"
