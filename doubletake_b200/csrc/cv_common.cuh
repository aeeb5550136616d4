// Geometry + sampling helpers shared by the cost-volume kernels (exact fp32 and tcgen05 variants).
// Arithmetic follows SURVEY.md Appendix B; see cost_volume.cu for the reference citations.
#pragma once
#include "common.cuh"

namespace dtb200 {

constexpr int kC = 16;            // matching feature channels (options.py matching_feature_dims)
constexpr int kPixPerWarp = 8;    // 8 pixels x 4 channel-quads
constexpr int kWarps = 8;

struct ViewConst {
  float P[12];  // rows 0..2 of K_src @ src_cam_T_cur_cam   (geometry_utils.py:82-84)
  float t[3];   // cur_cam_T_src_cam[:3,3]                   (geometry_utils.py:178-180)
  float comb, rm, tm;  // pose_distance                      (geometry_utils.py:187-199)
};

// Per-(b,k) constants, computed by the first warps of each block (K <= 16: negligible).
__device__ __forceinline__ void load_view_const(ViewConst& vc, const float* __restrict__ Ks,
                                                const float* __restrict__ ext, const float* __restrict__ pose) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float acc = DT_MUL(Ks[i * 4 + 0], ext[0 * 4 + j]);
      acc = DT_FMA(Ks[i * 4 + 1], ext[1 * 4 + j], acc);
      acc = DT_FMA(Ks[i * 4 + 2], ext[2 * 4 + j], acc);
      acc = DT_FMA(Ks[i * 4 + 3], ext[3 * 4 + j], acc);
      vc.P[i * 4 + j] = acc;
    }
  vc.t[0] = pose[3];
  vc.t[1] = pose[7];
  vc.t[2] = pose[11];
  float trace = DT_ADD(DT_ADD(pose[0], pose[5]), pose[10]);
  float rm = sqrtf(DT_MUL(2.f, DT_SUB(1.f, DT_DIV(fminf(3.f, trace), 3.f))));
  float tt = DT_MUL(vc.t[0], vc.t[0]);
  tt = DT_FMA(vc.t[1], vc.t[1], tt);
  tt = DT_FMA(vc.t[2], vc.t[2], tt);
  float tm = sqrtf(tt);
  vc.rm = rm;
  vc.tm = tm;
  vc.comb = sqrtf(DT_ADD(DT_MUL(tm, tm), DT_MUL(rm, rm)));
}

struct Projected {
  float u, v, zp;
};

// Project3D.forward (geometry_utils.py:77-93) for one point X (already d * ray).
__device__ __forceinline__ Projected project_point(const ViewConst& vc, float X0, float X1, float X2) {
  float q[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float acc = DT_MUL(vc.P[i * 4 + 0], X0);
    acc = DT_FMA(vc.P[i * 4 + 1], X1, acc);
    acc = DT_FMA(vc.P[i * 4 + 2], X2, acc);
    q[i] = DT_ADD(acc, vc.P[i * 4 + 3]);
  }
  Projected p;
  p.zp = DT_ADD(q[2], 1e-8f);
  float s = (fabsf(q[2]) > 1e-8f) ? DT_DIV(1.f, p.zp) : 1.f;
  p.u = DT_MUL(q[0], s);
  p.v = DT_MUL(q[1], s);
  return p;
}

// F.grid_sample(bilinear, zeros, align_corners=False) at pixel coords (u,v), split in two steps so a quad of lanes can
// share the setup: (1) sampling setup = float offset of the nw texel, the four tap weights (nw, ne, sw, se) and a 4-bit
// validity mask (zeros padding: OOB / NaN / huge coordinates fail every range test -- GridSampler.cuh behaviour);
// (2) the gather of this lane's 4 channels.  mesh_hint_volume.py:238-249 + ATen grid_sampler_unnormalize
// ((g+1)*W-1)/2 -- the division by 2 is written as an exact multiplication by 0.5.
struct SampleSetup {
  int off;          // ((y0 * W) + x0) * kC, valid whenever mask != 0
  float w[4];
  int mask;         // bit t set: tap t (nw, ne, sw, se) is inside the image
};

__device__ __forceinline__ SampleSetup sample_setup(float u, float v, int H, int W, float invW, float invH) {
  float gx = DT_SUB(DT_MUL(DT_MUL(2.f, u), invW), 1.f);
  float gy = DT_SUB(DT_MUL(DT_MUL(2.f, v), invH), 1.f);
  float ix = DT_MUL(DT_SUB(DT_MUL(DT_ADD(gx, 1.f), (float)W), 1.f), 0.5f);
  float iy = DT_MUL(DT_SUB(DT_MUL(DT_ADD(gy, 1.f), (float)H), 1.f), 0.5f);
  float x0f = floorf(ix), y0f = floorf(iy);
  float x1f = DT_ADD(x0f, 1.f), y1f = DT_ADD(y0f, 1.f);
  float wx1 = DT_SUB(ix, x0f), wx0 = DT_SUB(x1f, ix);
  float wy1 = DT_SUB(iy, y0f), wy0 = DT_SUB(y1f, iy);
  bool x0ok = (x0f >= 0.f) && (x0f <= (float)(W - 1));
  bool x1ok = (x1f >= 0.f) && (x1f <= (float)(W - 1));
  bool y0ok = (y0f >= 0.f) && (y0f <= (float)(H - 1));
  bool y1ok = (y1f >= 0.f) && (y1f <= (float)(H - 1));
  SampleSetup s;
  s.mask = (x0ok && y0ok ? 1 : 0) | (x1ok && y0ok ? 2 : 0) | (x0ok && y1ok ? 4 : 0) | (x1ok && y1ok ? 8 : 0);
  s.off = s.mask ? ((int)y0f * W + (int)x0f) * kC : 0;
  s.w[0] = DT_MUL(wx0, wy0);
  s.w[1] = DT_MUL(wx1, wy0);
  s.w[2] = DT_MUL(wx0, wy1);
  s.w[3] = DT_MUL(wx1, wy1);
  return s;
}

__device__ __forceinline__ float4 sample_apply(const float* __restrict__ src_view, int q, const SampleSetup& s, int W) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (s.mask == 0) return acc;
  const float* base = src_view + s.off + q * 4;
  auto tap = [&](int bit, int off, float w) {
    if (s.mask & bit) {
      float4 t = __ldg(reinterpret_cast<const float4*>(base + off));
      acc.x = DT_FMA(t.x, w, acc.x);
      acc.y = DT_FMA(t.y, w, acc.y);
      acc.z = DT_FMA(t.z, w, acc.z);
      acc.w = DT_FMA(t.w, w, acc.w);
    }
  };
  tap(1, 0, s.w[0]);
  tap(2, kC, s.w[1]);
  tap(4, W * kC, s.w[2]);
  tap(8, (W + 1) * kC, s.w[3]);
  return acc;
}

// Branch-free form of sample_apply for the tensor-core kernels: the four taps are loaded by predicated instructions (all
// in flight together) and an out-of-image tap contributes fma(0, 0, acc) = acc, i.e. exactly nothing (the weight is zeroed
// too: NaN coordinates fail every range test but leave NaN weights).  Same FMA order as sample_apply.
__device__ __forceinline__ float4 sample_apply_nb(const float* __restrict__ src_view, int q, const SampleSetup& s, int W) {
  const float4* base = reinterpret_cast<const float4*>(src_view + s.off + q * 4);
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool v0 = s.mask & 1, v1 = s.mask & 2, v2 = s.mask & 4, v3 = s.mask & 8;
  const float4 t0 = v0 ? __ldg(base) : z;
  const float4 t1 = v1 ? __ldg(base + kC / 4) : z;
  const float4 t2 = v2 ? __ldg(base + (W * kC) / 4) : z;
  const float4 t3 = v3 ? __ldg(base + ((W + 1) * kC) / 4) : z;
  const float w0 = v0 ? s.w[0] : 0.f, w1 = v1 ? s.w[1] : 0.f, w2 = v2 ? s.w[2] : 0.f, w3 = v3 ? s.w[3] : 0.f;
  float4 acc;
  acc.x = DT_FMA(t3.x, w3, DT_FMA(t2.x, w2, DT_FMA(t1.x, w1, DT_FMA(t0.x, w0, 0.f))));
  acc.y = DT_FMA(t3.y, w3, DT_FMA(t2.y, w2, DT_FMA(t1.y, w1, DT_FMA(t0.y, w0, 0.f))));
  acc.z = DT_FMA(t3.z, w3, DT_FMA(t2.z, w2, DT_FMA(t1.z, w1, DT_FMA(t0.z, w0, 0.f))));
  acc.w = DT_FMA(t3.w, w3, DT_FMA(t2.w, w2, DT_FMA(t1.w, w1, DT_FMA(t0.w, w0, 0.f))));
  return acc;
}

__device__ __forceinline__ float4 sample_quad(const float* __restrict__ src_view, int q, float u, float v, int H, int W,
                                              float invW, float invH) {
  return sample_apply(src_view, q, sample_setup(u, v, H, W, invW, invH), W);
}

// 16-channel dot product of the warped source texel with the current-view feature; result in all 4 quad lanes.
__device__ __forceinline__ float quad_dot(float4 a, float4 c) {
  float p = DT_MUL(a.x, c.x);
  p = DT_FMA(a.y, c.y, p);
  p = DT_FMA(a.z, c.z, p);
  p = DT_FMA(a.w, c.w, p);
  p = DT_ADD(p, __shfl_xor_sync(0xffffffffu, p, 1));
  p = DT_ADD(p, __shfl_xor_sync(0xffffffffu, p, 2));
  return p;
}

__device__ __forceinline__ bool better(float v, int i, float bv, int bi) {
  // torch.argmax semantics: NaN is maximal, first occurrence wins
  bool vn = isnan(v), bn = isnan(bv);
  if (vn != bn) return vn;
  if (!vn && v != bv) return v > bv;
  return i < bi;
}

__device__ __forceinline__ float plane_depth(const dtb200_cost_volume_params& p, int b, int d, int pix) {
  if (p.planes_per_pixel) return p.plane_depths[((long long)b * p.planes + d) * p.height * p.width + pix];
  return p.plane_depths[b * p.planes + d];
}

__device__ __forceinline__ void backproject_ray(const float* __restrict__ invK, int x, int y, float r[3]) {
  // BackprojectDepth (geometry_utils.py:34-39,60): invK[:3,:3] @ (x+0.5, y+0.5, 1)
  float px = (float)x + 0.5f, py = (float)y + 0.5f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float acc = DT_MUL(invK[i * 4 + 0], px);
    acc = DT_FMA(invK[i * 4 + 1], py, acc);
    r[i] = DT_ADD(acc, invK[i * 4 + 2]);
  }
}

__device__ __forceinline__ void write_masks(const dtb200_cost_volume_params& p, int b, int pix, int k, bool depth_ok,
                                            bool bounds_ok, bool& any_d, bool& any_b) {
  any_d |= depth_ok;
  any_b |= bounds_ok;
  if (p.mask_views) p.mask_views[((long long)b * p.views + k) * p.height * p.width + pix] = depth_ok && bounds_ok;
}

__device__ __forceinline__ float leaky01(float x) { return x > 0.f ? x : DT_MUL(x, 0.01f); }
// the same function in two instructions (slope < 1: max(x, 0.01 x) picks x for x > 0 and 0.01 x otherwise; NaN stays NaN)
__device__ __forceinline__ float leaky01_fast(float x) { return fmaxf(x, DT_MUL(x, 0.01f)); }

}  // namespace dtb200
