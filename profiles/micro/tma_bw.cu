// Micro-benchmark: per-SM TMA load throughput from L2-resident data (what bounds the split-plane GEMM mainloops).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bw tma_bw.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT;\n\tDONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

constexpr int STAGES = 8;
__global__ void __launch_bounds__(64) bw_kernel(const __grid_constant__ CUtensorMap map, int box_bytes, int boxes_per_stage, int iters, int rows_total, int box_rows, int ncolblk) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bars[2 * STAGES];
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&bars[s]), 1); mbar_init(smem_u32(&bars[STAGES + s]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t stage_bytes = (uint32_t)box_bytes * boxes_per_stage;
  if (threadIdx.x == 0) {
    int s = 0; uint32_t ph = 0;
    int r = (blockIdx.x * 37) % (rows_total / box_rows), c = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(smem_u32(&bars[STAGES + s]), ph ^ 1);
      mbar_expect_tx(smem_u32(&bars[s]), stage_bytes);
      for (int b = 0; b < boxes_per_stage; ++b) {
        tma_load_2d(base + s * stage_bytes + b * box_bytes, &map, smem_u32(&bars[s]), c * 64, r * box_rows);
        if (++c == ncolblk) { c = 0; if (++r == rows_total / box_rows) r = 0; }
      }
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(smem_u32(&bars[s]), ph);
      mbar_arrive(smem_u32(&bars[STAGES + s]));
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fnp, 12000, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fnp;
  const int rows = 8192;
  for (int cols : {64, 512}) {                       // row pitch 128 B (contiguous boxes) vs 1024 B (a [rows][512] bf16 matrix)
    void* d; cudaMalloc(&d, (size_t)rows * cols * 2); cudaMemset(d, 0, (size_t)rows * cols * 2);   // 1 MB / 8 MB: L2 resident
    for (int box_rows : {128, 64, 32}) {
      CUtensorMap map;
      cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows}; cuuint64_t gstr[1] = {(cuuint64_t)cols * 2};
      cuuint32_t box[2] = {64, (cuuint32_t)box_rows}, estr[2] = {1, 1};
      CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
      const int box_bytes = box_rows * 128;
      for (int bps : {1, 2, 4}) {
        if (box_bytes * bps * STAGES > 200 * 1024) continue;
        for (int grid : {1, 16, 64, 148}) {
          const int iters = 4000 / bps;
          cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
          const size_t smem = (size_t)box_bytes * bps * STAGES + 1024;
          cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
          bw_kernel<<<grid, 64, smem>>>(map, box_bytes, bps, 200, rows, box_rows, cols / 64);
          cudaEventRecord(e0);
          bw_kernel<<<grid, 64, smem>>>(map, box_bytes, bps, iters, rows, box_rows, cols / 64);
          cudaEventRecord(e1); cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          cudaError_t err = cudaGetLastError();
          const double bytes = (double)iters * bps * box_bytes;
          printf("pitch %4d B  box %3d rows (%5d B)  boxes/stage %d (in flight %3d KB)  grid %3d : %7.1f GB/s per SM  %7.2f TB/s total  %s\n", cols * 2, box_rows, box_bytes, bps,
                 box_bytes * bps * STAGES / 1024, grid, bytes / (ms * 1e-3) / 1e9, bytes * grid / (ms * 1e-3) / 1e12, err == cudaSuccess ? "" : cudaGetErrorString(err));
        }
      }
    }
    cudaFree(d);
  }
  return 0;
}
