// Micro-benchmark: does the halo kernel's A-operand addressing (start shifted by whole pixels = 128 B, 8-row group stride =
// one patch row = 1280 B) cost tensor-pipe cycles?  kind::f16, M = 128, N = 64 / 128, 36 MMAs per "tile pass" like the kernel.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I doubletake_b200/csrc -o tools/umma_halo_bench.bin tools/umma_halo_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace dtb200::tc;

struct Variant {
  const char* name;
  uint32_t a_shift_bytes;   // added to the A start address per tap index t: shift = a_shift_bytes * f(t)
  uint32_t a_sbo_bytes;     // 8-row group stride of A
  uint32_t swizzle;         // layout type field: 2 = 128B, 4 = 64B, 6 = 32B, 0 = none
  int taps;                 // distinct start addresses cycled through (1 = always the same)
  int n;
};

__global__ void __launch_bounds__(128, 1) bench(long long* out, int iters, Variant v) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (96 << 10) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;   // fp16 1.0
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0) {
    const uint32_t a_u = smem_u32(smem), b_u = a_u + (64 << 10);
    const uint32_t idesc = umma_idesc_f16(128, v.n);
    const uint64_t hi_a = (uint64_t)((v.a_sbo_bytes >> 4) | (1u << 14) | (v.swizzle << 29)) << 32;
    const uint64_t hi_b = (uint64_t)(64u | (1u << 14) | (2u << 29)) << 32;
    long long t0 = clock64();
    if (elect_one()) {
      for (int it = 0; it < iters; ++it) {
        for (int t = 0; t < 9; ++t) {
          const int tt = t % v.taps;
          const uint32_t shift = v.a_shift_bytes * (uint32_t)((tt / 3) * 10 + tt % 3);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t da = hi_a | ((((a_u + shift + ks * 32) & 0x3FFFFu) >> 4) | (1u << 16));
            const uint64_t db = hi_b | ((((b_u + t * 16384 % 32768 + ks * 32) & 0x3FFFFu) >> 4) | (1u << 16));
            umma_f16(tm, da, db, idesc, true);
          }
        }
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  const int iters = 200;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 << 10);
  const Variant vs[] = {
      {"aligned start, SBO 1024 (GEMM layout)      N128", 0, 1024, 2, 1, 128},
      {"aligned start, SBO 1024 (GEMM layout)      N64 ", 0, 1024, 2, 1, 64},
      {"start + 128 B (one pixel), SBO 1024        N128", 128, 1024, 2, 2, 128},
      {"start + 512 B, SBO 1024                    N128", 512, 1024, 2, 2, 128},
      {"aligned start, SBO 1280 (patch rows)       N128", 0, 1280, 2, 1, 128},
      {"aligned start, SBO 2048                    N128", 0, 2048, 2, 1, 128},
      {"aligned start, SBO 3072                    N128", 0, 3072, 2, 1, 128},
      {"halo: 9 shifted starts, SBO 1280           N128", 128, 1280, 2, 9, 128},
      {"halo: 9 shifted starts, SBO 1280           N64 ", 128, 1280, 2, 9, 64},
      {"halo: 9 shifted starts, SBO 2048           N128", 128, 2048, 2, 9, 128},
      {"9 starts shifted by 1024 B, SBO 1024       N128", 1024, 1024, 2, 9, 128},
      {"9 starts shifted by 1024 B, SBO 2048       N128", 1024, 2048, 2, 9, 128},
      {"SWIZZLE_64B  aligned, SBO 512              N128", 0, 512, 4, 1, 128},
      {"SWIZZLE_32B  aligned, SBO 256              N128", 0, 256, 6, 1, 128},
      {"SWIZZLE_32B  shifted 32 B x tap, SBO 320   N128", 32, 320, 6, 9, 128},
      {"SWIZZLE_32B  shifted 32 B x tap, SBO 320   N64 ", 32, 320, 6, 9, 64},
      {"no swizzle   aligned, SBO 128              N128", 0, 128, 0, 1, 128},
  };
  for (const Variant& v : vs) {
    bench<<<148, 128, 100 << 10>>>(d, iters, v);
    cudaDeviceSynchronize();
    bench<<<148, 128, 100 << 10>>>(d, iters, v);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%s : %7.1f clk/MMA (%s)\n", v.name, (double)h / (iters * 36.0), cudaGetErrorString(e));
  }
  return 0;
}
