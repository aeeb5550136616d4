// Micro-benchmark: cycles per tcgen05.mma kind::tf32 / kind::f16 by shape, same vs alternating accumulators.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I doubletake_b200/csrc -o /tmp/umma_bench tools/umma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace dtb200::tc;

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, bool acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
               "l"(a), "l"(b), "r"(idesc), "r"((uint32_t)acc) : "memory");
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// mode 0: tf32, all MMAs into one accumulator; 1: tf32 alternating 2 accumulators; 2: bf16 one accumulator
template <int M, int N, int MODE>
__global__ void __launch_bounds__(128, 1) bench(long long* out, int iters) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (64 << 10) / 4; i += blockDim.x) ((float*)smem)[i] = 0.001f * (i & 63);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0) {
    const uint32_t a_u = smem_u32(smem), b_u = a_u + (16 << 10);
    const uint32_t idesc = MODE == 2 ? idesc_bf16(M, N) : umma_idesc_tf32(M, N);
    long long t0 = clock64();
    if (elect_one()) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint64_t da = umma_desc_k128(a_u + ks * 32), db = umma_desc_k128(b_u + ks * 32);
          uint32_t d = tm + ((MODE == 1 && (ks & 1)) ? 256 : 0);
          if (MODE == 2) umma_f16(d, da, db, idesc, true); else umma_tf32(d, da, db, idesc, true);
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

// pipeline-skeleton costs seen by the issuing warp.  VAR 0: {12 MMA; commit(bar[i&3])} no waits;  1: {12 MMA; commit; wait};
// 2: {commit; wait} (no MMA);  3: {wait-on-completed-barrier} only;  4: {12 MMA; commit(bar[i&3]); wait(bar[(i-3)&3])} (depth 4)
template <int VAR>
__global__ void __launch_bounds__(128, 1) loopbench(long long* out, int iters) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (64 << 10) / 4; i += blockDim.x) ((float*)smem)[i] = 0.001f * (i & 63);
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0) {
    const uint32_t a_u = smem_u32(smem), b_u = a_u + (16 << 10);
    const uint32_t idesc = umma_idesc_tf32(128, 64);
    int ph[4] = {0, 0, 0, 0};
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int s = it & 3;
      if (VAR == 3) { mbar_wait(&bar[0], 1); continue; }
      if (VAR == 5 || VAR == 6) {
        // MMAs first, then the wait, then the commit: does the wait overlap with MMA execution?
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 12; ++ks) {
            uint64_t da = umma_desc_k128(a_u + (ks & 3) * 32), db = umma_desc_k128(b_u + (ks & 3) * 32);
            umma_tf32(tm, da, db, idesc, true);
          }
        }
        __syncwarp();
        if (VAR == 5) mbar_wait(&bar[3], 1);                       // completed barrier (never armed): pure latency
        if (VAR == 6) { volatile uint64_t* vb = &bar[3]; (void)*vb; (void)*vb; }  // two plain shared loads instead
        if (elect_one()) umma_commit(&bar[s % 3]);
        __syncwarp();
        continue;
      }
      if (elect_one()) {
        if (VAR != 2) {
#pragma unroll
          for (int ks = 0; ks < 12; ++ks) {
            uint64_t da = umma_desc_k128(a_u + (ks & 3) * 32), db = umma_desc_k128(b_u + (ks & 3) * 32);
            umma_tf32(tm, da, db, idesc, true);
          }
        }
        umma_commit(&bar[VAR == 0 || VAR == 4 ? s : 0]);
      }
      __syncwarp();
      if (VAR == 1 || VAR == 2) { mbar_wait(&bar[0], ph[0]); ph[0] ^= 1; }
      if (VAR == 4 && it >= 3) { const int w = (it - 3) & 3; mbar_wait(&bar[w], ph[w]); ph[w] ^= 1; }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

template <int VAR>
void runloop(const char* name) {
  long long* d; cudaMalloc(&d, 8);
  int iters = 2000;
  cudaFuncSetAttribute(loopbench<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 << 10);
  loopbench<VAR><<<148, 128, 80 << 10>>>(d, iters);
  cudaDeviceSynchronize();
  loopbench<VAR><<<148, 128, 80 << 10>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-64s %8.1f clk/iter (%s)\n", name, (double)h / iters, cudaGetErrorString(e));
  cudaFree(d);
}

template <int M, int N, int MODE>
void run(const char* name, int grid) {
  long long* d; cudaMalloc(&d, 8);
  int iters = 2000;
  cudaFuncSetAttribute(bench<M, N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 << 10);
  bench<M, N, MODE><<<grid, 128, 80 << 10>>>(d, iters);
  cudaDeviceSynchronize();
  bench<M, N, MODE><<<grid, 128, 80 << 10>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  double per = (double)h / (iters * 4.0);
  double macs = (double)M * N * (MODE == 2 ? 16 : 8);
  printf("%-28s grid %3d: %7.1f clk/MMA  %7.0f MAC/clk/SM  (%s)\n", name, grid, per, macs / per, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  runloop<0>("{12 MMA(N64); commit(bar[i&3])}  no waits");
  runloop<1>("{12 MMA(N64); commit; wait}  serialized");
  runloop<2>("{commit; wait}  no MMA");
  runloop<3>("{wait on already-completed barrier}");
  runloop<4>("{12 MMA; commit(bar[i&3]); wait(bar[(i-3)&3])}  depth 4");
  runloop<5>("{12 MMA; wait(completed); commit}  wait between issue and commit");
  runloop<6>("{12 MMA; 2x LDS; commit}");
  for (int grid : {148}) {
    run<128, 64, 0>("tf32 M128 N64  same acc", grid);
    run<128, 64, 1>("tf32 M128 N64  alt acc", grid);
    run<128, 128, 0>("tf32 M128 N128 same acc", grid);
    run<128, 256, 0>("tf32 M128 N256 same acc", grid);
    run<64, 256, 0>("tf32 M64  N256 same acc", grid);
    run<128, 64, 2>("bf16 M128 N64  same acc", grid);
    run<128, 128, 2>("bf16 M128 N128 same acc", grid);
    run<128, 256, 2>("bf16 M128 N256 same acc", grid);
  }
  return 0;
}
