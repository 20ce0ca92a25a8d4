// Micro-test: is completion transitive along a chain of programmatically-dependent launches?
// A (long, writes flag at its end) -> B1 .. Bn (each: trigger early, wait, trivial work) -> Z (wait, read flag).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__global__ void A(int* flag, long long spin) {
  pdl_trigger(); pdl_wait();
  if (blockIdx.x == 0 && threadIdx.x == 0) { *flag = 0; }
  long long t0 = clock64();
  while (clock64() - t0 < spin) {}
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) { atomicAdd(flag, 1); }
}
__global__ void Bk(int* scratch, int mode) {
  pdl_trigger();
  if (mode == 1 && blockIdx.x >= 4) return;        // some CTAs leave without waiting
  pdl_wait();
  if (threadIdx.x == 0) scratch[blockIdx.x] = 1;
}
__global__ void Z(const int* __restrict__ flag, int* out, int i) {
  pdl_trigger(); pdl_wait();
  if (blockIdx.x == 0 && threadIdx.x == 0) out[i] = *flag;
}
template <class... KA, class... Args>
void launch(void (*k)(KA...), int grid, int block, cudaStream_t s, bool pdl, Args... a) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, k, static_cast<KA>(a)...);
}
int main() {
  int *flag, *scratch, *out;
  cudaMalloc(&flag, 4); cudaMalloc(&scratch, 4096 * 4); cudaMalloc(&out, 4096 * 4);
  cudaStream_t s; cudaStreamCreate(&s);
  for (int mode = 0; mode < 2; ++mode)
    for (int nb = 1; nb <= 6; nb += 5) {
      const int iters = 400;
      cudaMemset(out, 0xff, 4096 * 4);
      for (int i = 0; i < iters; ++i) {
        launch(A, 148 * 4, 128, s, true, flag, (long long)(2000 + 37 * (i % 50)));
        for (int j = 0; j < nb; ++j) launch(Bk, 16, 64, s, true, scratch, mode);
        launch(Z, 8, 64, s, true, (const int*)flag, out, i);
      }
      cudaStreamSynchronize(s);
      int h[400]; cudaMemcpy(h, out, iters * 4, cudaMemcpyDeviceToHost);
      int bad = 0; for (int i = 0; i < iters; ++i) bad += h[i] != 1;
      printf("mode %d (%s), %d intermediate kernels: %d / %d chains saw a stale flag  (%s)\n", mode, mode ? "some CTAs exit before the wait" : "all CTAs wait", nb, bad, iters, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
