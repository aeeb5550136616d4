// im2col loaders shared by the SIMT and tcgen05 convolution kernels: concat sources, x2 up-sampling on load, zero padding.
#pragma once
#include "common.cuh"

namespace dtb200 {

struct SrcView {
  const float* ptr;
  int c, resample, h, w;  // h,w = stored spatial size of this source
};

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float4 lerp4(float4 a, float4 b, float l0, float l1) {
  float4 r;
  r.x = DT_FMA(l1, b.x, DT_MUL(l0, a.x));
  r.y = DT_FMA(l1, b.y, DT_MUL(l0, a.y));
  r.z = DT_FMA(l1, b.z, DT_MUL(l0, a.z));
  r.w = DT_FMA(l1, b.w, DT_MUL(l0, a.w));
  return r;
}

// ATen area_pixel_compute_source_index(scale=0.5, align_corners=False): src = 0.5*(dst+0.5)-0.5, clamped at 0.
__device__ __forceinline__ void up2_coord(int dst, int size, int& i0, int& i1, float& l0, float& l1) {
  float s = DT_SUB(DT_MUL(0.5f, DT_ADD((float)dst, 0.5f)), 0.5f);
  s = s < 0.f ? 0.f : s;
  i0 = (int)s;
  i1 = i0 + (i0 < size - 1 ? 1 : 0);
  l1 = DT_SUB(s, (float)i0);
  l0 = DT_SUB(1.f, l1);
}

// 4 consecutive channels (c..c+3) of source `sv` at conv-input pixel (iy,ix) of batch b; zero outside the input.
__device__ __forceinline__ float4 load_input4(const SrcView& sv, int b, int iy, int ix, int in_h, int in_w, int c) {
  if (iy < 0 || iy >= in_h || ix < 0 || ix >= in_w) return make_float4(0.f, 0.f, 0.f, 0.f);
  const float* base = sv.ptr + (long long)b * sv.h * sv.w * sv.c + c;
  if (sv.resample == DTB200_RESAMPLE_NONE) return ld4(base + ((long long)iy * sv.w + ix) * sv.c);
  if (sv.resample == DTB200_RESAMPLE_NEAREST_UP2) return ld4(base + ((long long)(iy >> 1) * sv.w + (ix >> 1)) * sv.c);
  int y0, y1, x0, x1;
  float hy0, hy1, wx0, wx1;
  up2_coord(iy, sv.h, y0, y1, hy0, hy1);
  up2_coord(ix, sv.w, x0, x1, wx0, wx1);
  float4 v00 = ld4(base + ((long long)y0 * sv.w + x0) * sv.c), v01 = ld4(base + ((long long)y0 * sv.w + x1) * sv.c);
  float4 v10 = ld4(base + ((long long)y1 * sv.w + x0) * sv.c), v11 = ld4(base + ((long long)y1 * sv.w + x1) * sv.c);
  // ATen upsample_bilinear2d: h0*(w0*v00 + w1*v01) + h1*(w0*v10 + w1*v11)
  return lerp4(lerp4(v00, v01, wx0, wx1), lerp4(v10, v11, wx0, wx1), hy0, hy1);
}

__device__ __forceinline__ float activate(float v, int act, float slope) {
  if (act == DTB200_ACT_LEAKY) return v > 0.f ? v : DT_MUL(v, slope);
  if (act == DTB200_ACT_ELU) return v > 0.f ? v : expm1f(v);
  return v;
}


}  // namespace dtb200
