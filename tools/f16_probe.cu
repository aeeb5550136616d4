// Probe for the 2-term fp16 split path ("tch"): (1) kind::f16 tcgen05.mma with fp16 SWIZZLE_128B K-major operands, K = 16 per
// instruction (32-byte steps inside the 128-byte row), merged-N issue A_big x [B_big | B_small]^T + A_small x B_big^T with
// the correction products in their own TMEM columns, result = main + corr / 2048 -- compared with a double-precision
// product of the original fp32 operands; (2) the TMA box of a split-plane activation tensor (B, H, W, 2, C) fp16: a 4-D
// map over ONE plane (pixel stride 4C bytes, base + 2C bytes for the small plane), 64-channel box on a 24-channel tensor
// (hardware zero fill of the channel overhang and of the conv halo).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I doubletake_b200/csrc -o tools/f16_probe.bin tools/f16_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace dtb200::tc;

constexpr int kM = 128, kN = 64, kK = 64;  // one K block of the conv kernel: 128 pixels x 64 channels x 64 outputs

__global__ void mma_probe(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                          float* __restrict__ out_main) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a_big = smem;                  // [128][64 fp16] 16 KB
  uint8_t* a_small = a_big + kM * 128;    // 16 KB
  uint8_t* b_big = a_small + kM * 128;    // [64][64 fp16] 8 KB, followed by b_small: 128 contiguous K-major rows
  uint8_t* b_small = b_big + kN * 128;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < kM * kK / 2; i += blockDim.x) {
    const int row = i / (kK / 2), k = (i % (kK / 2)) * 2;
    uint32_t big, small;
    split_half2(a[row * kK + k], a[row * kK + k + 1], big, small);
    *(uint32_t*)(a_big + sw128_offset_h(row, k)) = big;
    *(uint32_t*)(a_small + sw128_offset_h(row, k)) = small;
  }
  for (int i = tid; i < kN * kK / 2; i += blockDim.x) {
    const int row = i / (kK / 2), k = (i % (kK / 2)) * 2;
    uint32_t big, small;
    split_half2(b[row * kK + k], b[row * kK + k + 1], big, small);
    *(uint32_t*)(b_big + sw128_offset_h(row, k)) = big;
    *(uint32_t*)(b_small + sw128_offset_h(row, k)) = small;
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc<128>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(kM, kN), idesc2 = umma_idesc_f16(kM, 2 * kN);
#pragma unroll
      for (int ks = 0; ks < kK / 16; ++ks) {
        const uint32_t ko = ks * 32;  // 16 fp16 = 32 bytes along K inside the swizzled row
        umma_f16(tmem, umma_desc_k128(smem_u32(a_big) + ko), umma_desc_k128(smem_u32(b_big) + ko), idesc2, ks != 0);  // [big x big | big x small]
        umma_f16(tmem + kN, umma_desc_k128(smem_u32(a_small) + ko), umma_desc_k128(smem_u32(b_big) + ko), idesc, true);  // small x big
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
      float v[32], c[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)cc, v);
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(kN + cc), c);
      for (int j = 0; j < 32; ++j) {
        out_main[row * kN + cc + j] = v[j];
        out[row * kN + cc + j] = fmaf(c[j], kHalfSplitInv, v[j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<128>(tmem);
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int kPW = 10, kPH = 18;  // halo patch of the 8 x 16 pixel tile

__global__ void tma_probe(const __grid_constant__ CUtensorMap tm_big, const __grid_constant__ CUtensorMap tm_small,
                          __half* out_big, __half* out_small, int c0, int x0, int y0, int b0) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  constexpr int kPatch = kPW * kPH * 128, kSlot = (kPatch + 1023) / 1024 * 1024;
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, 2 * kPatch);
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                     smem_u32(smem)),
                 "l"(&tm_big), "r"(c0), "r"(x0), "r"(y0), "r"(b0), "r"(smem_u32(&bar))
                 : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                     smem_u32(smem + kSlot)),
                 "l"(&tm_small), "r"(c0), "r"(x0), "r"(y0), "r"(b0), "r"(smem_u32(&bar))
                 : "memory");
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < kPW * kPH * 64; i += blockDim.x) {
    const int p = i / 64, k = i % 64;
    out_big[i] = *(__half*)(smem + sw128_offset_h(p, k));
    out_small[i] = *(__half*)(smem + kSlot + sw128_offset_h(p, k));
  }
}

int main() {
  int failures = 0;
  // ---------------------------------------------------------------- (1) MMA numerics
  {
    std::vector<float> a((size_t)kM * kK), b((size_t)kN * kK);
    srand(7);
    auto rnd = [] { return (float)rand() / RAND_MAX * 2.f - 1.f; };
    for (auto& v : a) v = rnd() * 3.f;
    for (auto& v : b) v = rnd() * 0.2f;
    a[5] = 1e-6f, a[70] = 12345.678f, b[9] = 3e-6f, a[64 * 3 + 1] = -0.f;  // tiny / large operands
    float *da, *db, *dout, *dmain;
    cudaMalloc(&da, a.size() * 4);
    cudaMalloc(&db, b.size() * 4);
    cudaMalloc(&dout, kM * kN * 4);
    cudaMalloc(&dmain, kM * kN * 4);
    cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
    const int smem = 2 * kM * 128 + 2 * kN * 128 + 1024;
    cudaFuncSetAttribute(mma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    mma_probe<<<1, 128, smem>>>(da, db, dout, dmain);
    cudaError_t e = cudaDeviceSynchronize();
    printf("mma probe: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> got(kM * kN), got_main(kM * kN);
    cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(got_main.data(), dmain, got.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0, worst_main = 0;
    for (int r = 0; r < kM; ++r)
      for (int n = 0; n < kN; ++n) {
        double want = 0, mag = 0;
        for (int k = 0; k < kK; ++k) {
          want += (double)a[r * kK + k] * b[n * kK + k];
          mag += fabs((double)a[r * kK + k] * b[n * kK + k]);
        }
        worst = fmax(worst, fabs(got[r * kN + n] - want) / mag);
        worst_main = fmax(worst_main, fabs(got_main[r * kN + n] - want) / mag);
      }
    printf("  split result: max |err| / sum|a b| = %.3e   (plain fp16 operands: %.3e; fp32 eps 6e-8)\n", worst, worst_main);
    failures += !(worst < 1e-6) || !(worst_main < 2e-3) || !(worst_main > 1e-5);
  }
  // ---------------------------------------------------------------- (2) TMA boxes of a split-plane tensor
  {
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres);
    if (!encode) {
      printf("no cuTensorMapEncodeTiled\n");
      return 1;
    }
    for (int variant = 0; variant < 2; ++variant) {
      const int B = 2, H = 20, W = 13, C = variant ? 160 : 24;
      std::vector<__half> h((size_t)B * H * W * 2 * C);
      for (size_t i = 0; i < h.size(); ++i) h[i] = __float2half((float)(i % 2039) + 1.f);
      __half *d, *ob, *os;
      cudaMalloc(&d, h.size() * 2);
      cudaMalloc(&ob, kPW * kPH * 64 * 2);
      cudaMalloc(&os, kPW * kPH * 64 * 2);
      cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
      CUtensorMap tm[2];
      int ok = 1;
      for (int plane = 0; plane < 2; ++plane) {
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
        cuuint32_t box[4] = {64, kPW, kPH, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&tm[plane], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d + (size_t)plane * C, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("tma variant %d plane %d (C=%d): encode -> %d\n", variant, plane, C, (int)r);
        ok &= (r == CUDA_SUCCESS);
      }
      if (!ok) {
        ++failures;
        continue;
      }
      const int c0 = variant ? 128 : 0, x0 = variant ? 4 : -1, y0 = variant ? 5 : -1, b0 = 1;
      const int smem = 2 * 24 * 1024 + 1024;
      cudaFuncSetAttribute(tma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      tma_probe<<<1, 128, smem>>>(tm[0], tm[1], ob, os, c0, x0, y0, b0);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<__half> gb(kPW * kPH * 64), gs(kPW * kPH * 64);
      cudaMemcpy(gb.data(), ob, gb.size() * 2, cudaMemcpyDeviceToHost);
      cudaMemcpy(gs.data(), os, gs.size() * 2, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int p = 0; p < kPW * kPH; ++p)
        for (int k = 0; k < 64; ++k) {
          const int y = y0 + p / kPW, x = x0 + p % kPW, c = c0 + k;
          float wb = 0.f, ws = 0.f;
          if (y >= 0 && y < H && x >= 0 && x < W && c < C) {
            const size_t pix = ((size_t)b0 * H + y) * W + x;
            wb = __half2float(h[pix * 2 * C + c]);
            ws = __half2float(h[pix * 2 * C + C + c]);
          }
          if (__half2float(gb[p * 64 + k]) != wb || __half2float(gs[p * 64 + k]) != ws) {
            if (bad < 4) printf("  mismatch pixel %d k %d: got %g / %g want %g / %g\n", p, k, __half2float(gb[p * 64 + k]),
                                __half2float(gs[p * 64 + k]), wb, ws);
            ++bad;
          }
        }
      printf("  kernel: %s, mismatches %d / %d\n", cudaGetErrorString(e), bad, kPW * kPH * 64);
      failures += (bad != 0) || (e != cudaSuccess);
      cudaFree(d), cudaFree(ob), cudaFree(os);
    }
  }
  printf(failures ? "PROBE FAILED\n" : "PROBE OK\n");
  return failures;
}
