// Micro-benchmark: MUFU (ex2 / rcp) throughput per SM -- the bound of the tanh-heavy hv_kernel epilogues.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_rate mufu_rate.cu && /tmp/mufu_rate
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, int iters, float seed) {
  float x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = seed + 0.001f * (threadIdx.x + j);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
      if (MODE == 1) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
      if (MODE == 2) {            // the tanh of the epilogues: mul, ex2, add, rcp, fma
        float e, r;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x[j] * 2.8853900817779268f));
        asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
        x[j] = fmaf(-2.f, r, 1.f);
      }
      if (MODE == 3) x[j] = fmaf(x[j], 1.0001f, 0.5f);   // FFMA reference
    }
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, int mufu_per_iter, int threads) {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  float* out; cudaMalloc(&out, sizeof(float) * sms * 1024);
  const int iters = 20000;
  k<MODE><<<sms, threads>>>(out, 100, 0.3f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<sms, threads>>>(out, iters, 0.3f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)threads * 8 * iters * mufu_per_iter;          // per SM
  printf("%-28s threads/SM %4d  %.3f ms  %.2f ops/clk/SM (at %d MHz)\n", name, threads, ms, ops / (ms * 1e-3) / (khz * 1e3), khz / 1000);
  cudaFree(out);
}
int main() {
  for (int threads : {128, 256, 512, 1024}) {
    run<0>("ex2.approx", 1, threads);
    run<1>("rcp.approx", 1, threads);
    run<2>("tanh (ex2+rcp), MUFU ops", 2, threads);
    run<3>("ffma", 1, threads);
  }
  return 0;
}
