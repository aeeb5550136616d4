// Shared host/device helpers for the doubletake_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "../../include/doubletake_b200.h"

namespace dtb200 {

extern thread_local char g_error[512];
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const char* fmt, const char* a = "", long long b = 0, long long c = 0) {
  snprintf(g_error, sizeof(g_error), fmt, a, b, c);
  return code;
}

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_error, sizeof(g_error), "%s: %s", what, cudaGetErrorString(e));
    return DTB200_ERR_CUDA;
  }
  return DTB200_OK;
}

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace dtb200

// Non-contracting fp32 primitives: the reference evaluates these as separate rounded torch ops, and nvcc would
// otherwise fuse a*b+c into one FMA.
#define DT_MUL(a, b) __fmul_rn((a), (b))
#define DT_ADD(a, b) __fadd_rn((a), (b))
#define DT_SUB(a, b) __fsub_rn((a), (b))
#define DT_DIV(a, b) __fdiv_rn((a), (b))
#define DT_FMA(a, b, c) __fmaf_rn((a), (b), (c))
