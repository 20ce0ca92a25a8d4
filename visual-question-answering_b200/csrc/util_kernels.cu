#include "util_kernels.cuh"

namespace hca {
namespace {
// 256 threads = 32 columns x 8 row lanes; one block reduces 32 columns over a 512-row slab.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, int64_t ld, int64_t rows, int cols,
                                                     float* __restrict__ out) {
  pdl_enter();
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const int64_t r0 = (int64_t)blockIdx.y * 512, r1 = min(rows, r0 + 512);
  float s = 0.f;
  if (c < cols)
    for (int64_t r = r0 + ry; r < r1; r += 8) s += X[r * ld + c];
  red[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && c < cols) {
#pragma unroll
    for (int i = 1; i < 8; ++i) s += red[i][cx];
    atomicAdd(out + c, s);
  }
}
}  // namespace

int launch_colsum(const float* X, int64_t ld, int64_t rows, int cols, float* out, cudaStream_t s) {
  dim3 grid((cols + 31) / 32, (unsigned)((rows + 511) / 512));
  HCA_LAUNCH_K((colsum_kernel), grid, 256, 0, s, X, ld, rows, cols, out);
  HCA_LAUNCHED();
  return 0;
}

namespace {
__global__ void __launch_bounds__(256) zero_many_kernel(const ZeroJobs jobs) {
  pdl_enter();
  zero_jobs_device(jobs);
}
}  // namespace

ZeroJobs ZeroBatch::take() {
  ZeroJobs jobs;
  for (int i = 0; i < count; ++i) { jobs.ptr[i] = ptr[i]; jobs.words[i] = words[i]; }
  for (int i = count; i < MAX; ++i) { jobs.ptr[i] = nullptr; jobs.words[i] = 0; }
  jobs.count = count;
  count = 0;
  return jobs;
}

int ZeroBatch::add(void* p, size_t bytes) {
  if (!p || bytes == 0) return 0;
  HCA_CHECK_ARG((reinterpret_cast<uintptr_t>(p) & 3) == 0 && (bytes & 3) == 0, "ZeroBatch: buffers must be 4-byte aligned");
  if (count == MAX) HCA_TRY(flush());
  ptr[count] = p;
  words[count] = bytes >> 2;
  ++count;
  return 0;
}

int ZeroBatch::flush() {
  if (count == 0) return 0;
  unsigned long long total = 0;
  for (int i = 0; i < count; ++i) total += words[i];
  const ZeroJobs jobs = take();
  HCA_LAUNCH_K((zero_many_kernel), ew_grid((int64_t)(total / 4 + 1)), 256, 0, stream, jobs);
  HCA_LAUNCHED();
  return 0;
}

int zero_async(void* p, size_t bytes, cudaStream_t s) {
  HCA_CUDA(cudaMemsetAsync(p, 0, bytes, s));
  return 0;
}
}  // namespace hca
