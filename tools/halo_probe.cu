// Probe for the halo-tile conv design (DESIGN.md §4.3 "next"): can ONE shared-memory patch of pixels (128-byte rows,
// SWIZZLE_128B by absolute address) feed tcgen05.mma for every tap of a 3x3 conv through descriptors that are only SHIFTED
// by whole pixels (start address + shift * 128 B) and whose 8-row core-matrix stride (SBO) is a patch row of 10 pixels
// (1280 B) instead of 1024 B?  Which value must the descriptor's base_offset field (bits 49..51) carry then?
//   A row r (group g = r / 8, i = r % 8) is expected to read pixel  shift + g * (SBO / 128) + i  of the patch.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I doubletake_b200/csrc -o tools/halo_probe.bin tools/halo_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace dtb200::tc;

constexpr int kPatchPixels = 320;  // 40 KB patch
constexpr int kN = 64;

__device__ __forceinline__ uint64_t desc_halo(uint32_t addr, uint32_t sbo_bytes, uint32_t base_offset) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) |
         ((uint64_t)(base_offset & 7) << 49) | (2ull << 61);
}

// mode 0: base_offset = 0; mode 1: base_offset = (start >> 7) & 7
__global__ void probe(const float* __restrict__ a_patch, const float* __restrict__ b_tile, float* __restrict__ out, int shift,
                      int sbo_bytes, int mode) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a_s = smem;                          // patch: pixel p at p * 128, chunk c at (c ^ (p & 7)) * 16
  uint8_t* b_s = smem + kPatchPixels * 128;     // [64][32] standard SWIZZLE_128B K-major tile (1024-aligned)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < kPatchPixels * 32; i += blockDim.x) {
    const int p = i / 32, k = i % 32;
    *(float*)(a_s + sw128_offset(p, k)) = a_patch[i];
  }
  for (int i = tid; i < kN * 32; i += blockDim.x) {
    const int n = i / 32, k = i % 32;
    *(float*)(b_s + sw128_offset(n, k)) = b_tile[i];
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc<64>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_tf32(128, kN);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t a_addr = smem_u32(a_s) + (uint32_t)shift * 128u + ks * 32;
        const uint32_t bo = mode ? ((a_addr >> 7) & 7) : 0;
        umma_tf32(tmem, desc_halo(a_addr, (uint32_t)sbo_bytes, bo), umma_desc_k128(smem_u32(b_s) + ks * 32), idesc, ks != 0);
      }
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  if (warp < 4) {
    const int row = warp * 32 + (tid & 31);
    for (int cc = 0; cc < kN; cc += 32) {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)cc, v);
      for (int j = 0; j < 32; ++j) out[row * kN + cc + j] = v[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<64>(tmem);
  }
}

int main() {
  std::vector<float> a((size_t)kPatchPixels * 32), b((size_t)kN * 32);
  for (int p = 0; p < kPatchPixels; ++p)
    for (int k = 0; k < 32; ++k) a[(size_t)p * 32 + k] = (float)(((p * 7 + k * 3) % 31) - 15);
  for (int n = 0; n < kN; ++n)
    for (int k = 0; k < 32; ++k) b[(size_t)n * 32 + k] = (float)(((n * 5 + k) % 13) - 6);
  float *da, *db, *dout;
  cudaMalloc(&da, a.size() * 4);
  cudaMalloc(&db, b.size() * 4);
  cudaMalloc(&dout, 128 * kN * 4);
  cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
  const int smem = kPatchPixels * 128 + kN * 128 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int shifts[] = {0, 1, 3, 8, 11, 21};
  const int sbos[] = {1024, 1280, 2304};
  std::vector<float> got(128 * kN);
  for (int sbo : sbos)
    for (int shift : shifts)
      for (int mode = 0; mode < 2; ++mode) {
        cudaMemset(dout, 0xff, got.size() * 4);
        probe<<<1, 128, smem>>>(da, db, dout, shift, sbo, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("sbo %4d shift %2d base_offset %s: CUDA error %s\n", sbo, shift, mode ? "auto" : "0   ", cudaGetErrorString(e));
          return 1;
        }
        cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0, first_bad = -1;
        for (int r = 0; r < 128; ++r) {
          const int p = shift + (r / 8) * (sbo / 128) + (r % 8);
          for (int n = 0; n < kN; ++n) {
            float want = 0.f;
            for (int k = 0; k < 32; ++k) want += a[(size_t)p * 32 + k] * b[(size_t)n * 32 + k];
            if (got[r * kN + n] != want) {
              if (first_bad < 0) first_bad = r;
              ++bad;
            }
          }
        }
        printf("sbo %4d shift %2d base_offset %s: %s (%d / %d wrong, first bad row %d)\n", sbo, shift, mode ? "auto" : "0   ",
               bad ? "MISMATCH" : "ok", bad, 128 * kN, first_bad);
      }
  return 0;
}
