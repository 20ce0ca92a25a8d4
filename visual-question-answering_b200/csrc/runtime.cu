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

static std::atomic<int> g_gemm_mode{-1};  // -1 unset, 0 ffma, 1 tc

bool use_tc() {
  int m = g_gemm_mode.load(std::memory_order_relaxed);
  if (m < 0) {
    const char* e = getenv("HCA_GEMM");
    m = (e && strcmp(e, "ffma") == 0) ? 0 : 1;
    g_gemm_mode.store(m);
  }
  return m == 1;
}

bool pdl_enabled() {
  const char* e = getenv("HCA_PDL");          // read per launch: a profiler leg may switch it off to get exclusive kernel durations
  return !(e && atoi(e) == 0);
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
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("HCA_SIDE_STREAM"); enabled = (e && atoi(e) == 0) ? 0 : 1; }
  if (!enabled) return;
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
  if (strcmp(name, "gemm") == 0) {
    if (strcmp(value, "tc") == 0) { hca::g_gemm_mode.store(1); return 0; }
    if (strcmp(value, "ffma") == 0) { hca::g_gemm_mode.store(0); return 0; }
  }
  return hca::set_err(HCA_ERR_ARG, "hca_set_option: unknown option %s=%s", name, value);
}

const char* hca_get_option(const char* name) {
  if (name && strcmp(name, "gemm") == 0) return hca::use_tc() ? "tc" : "ffma";
  return "";
}

}  // extern "C"
