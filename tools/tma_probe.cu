// Probe: cp.async.bulk.tensor.4d (TMA tiled mode) on an NHWC fp32 tensor -> SWIZZLE_128B smem tile; checks the layout the
// conv kernel relies on: row = ty*TW + tx, 16-byte chunk c of row r stored at chunk (c ^ (r & 7)), zero fill outside.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I doubletake_b200/csrc -o tools/tma_probe.bin tools/tma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace dtb200::tc;

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tm, float* out, int c0, int x0, int y0, int b0) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, 128 * 128);
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(smem)), "l"(&tm), "r"(c0), "r"(x0), "r"(y0), "r"(b0), "r"(smem_u32(&bar)) : "memory");
  }
  mbar_wait(&bar, 0);
  // de-swizzle: logical (row, k) lives at row*128 + ((k/4) ^ (row&7))*16 + (k%4)*4
  for (int i = threadIdx.x; i < 128 * 32; i += blockDim.x) {
    int row = i / 32, k = i % 32;
    out[i] = *(float*)(smem + sw128_offset(row, k));
  }
}

int main(int argc, char** argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1;
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres);
  printf("entry point: %s q=%d fn=%p\n", cudaGetErrorString(e), (int)qres, (void*)encode);
  if (!encode) return 1;
  int failures = 0;
  for (int variant = 0; variant < 4; ++variant) {
    if (only >= 0 && variant != only) continue;
    const int B = 2, H = 12, W = 20;
    const int C = (variant == 1) ? 24 : 64;
    const int stride = (variant == 2) ? 2 : 1;
    const int TW = 16, TH = 8;
    std::vector<float> h((size_t)B * H * W * C);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003) * 0.5f + 1.0f;
    float *d, *o;
    cudaMalloc(&d, h.size() * 4);
    cudaMalloc(&o, 128 * 32 * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)TW, (cuuint32_t)TH, 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d (C=%d stride=%d): encode -> %d\n", variant, C, stride, (int)r);
    if (r != CUDA_SUCCESS) { ++failures; continue; }
    const int c0 = (variant == 3) ? 32 : 0, x0 = -1, y0 = (variant == 3) ? 7 : -1, b0 = 1;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 20 << 10);
    probe<<<1, 128, 20 << 10>>>(tm, o, c0, x0, y0, b0);
    e = cudaDeviceSynchronize();
    std::vector<float> got(128 * 32);
    cudaMemcpy(got.data(), o, got.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int row = 0; row < 128; ++row)
      for (int k = 0; k < 32; ++k) {
        int ty = row / TW, tx = row % TW;
        int y = y0 + ty * stride, x = x0 + tx * stride, c = c0 + k;
        float want = 0.f;
        if (y >= 0 && y < H && x >= 0 && x < W && c < C) want = h[(((size_t)b0 * H + y) * W + x) * C + c];
        if (got[row * 32 + k] != want) {
          if (bad < 4) printf("  mismatch row %d k %d: got %g want %g\n", row, k, got[row * 32 + k], want);
          ++bad;
        }
      }
    printf("  kernel: %s, mismatches %d / 4096\n", cudaGetErrorString(e), bad);
    failures += (bad != 0) || (e != cudaSuccess);
    cudaFree(d); cudaFree(o);
  }
  printf(failures ? "PROBE FAILED\n" : "PROBE OK\n");
  return failures;
}
