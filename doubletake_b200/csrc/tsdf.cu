// TSDF fusion of depth maps + sampling of the fused volume, sm_100a.
//
// Reference: tools/tsdf.py -- TSDFFuser.integrate_depth (:414-558), TSDF.sample_tsdf (:277-337),
// TSDF.generate_voxel_coords (:155-166).  The reference keeps everything in fp16 and runs ~40 torch ops per frame over
// the voxels of the frustum's bounding box (gather -> project -> grid_sample -> masks -> scatter).  Every one of those ops
// is elementwise per voxel, evaluated in fp32 and rounded to fp16, so the whole update is ONE pass here:
//   * a thread owns 8 consecutive voxels along Z (one 16-byte vector of values and one of weights; dims are multiples of 8)
//   * voxel coordinates are regenerated in registers (fp16(origin + index * voxel_size), bit-identical to the stored
//     grid) unless the caller supplies its own grid -- 6 of the 10 bytes per voxel never cross HBM
//   * the launch scans only the index box that covers the frames' frustum boxes (the reference's default volume is a
//     20 m cube of 128 M voxels of which a frame touches ~1e5), the exact fp16 box test stays per voxel
//   * up to 8 frames are integrated in order inside the pass (a voxel depends only on its own previous state), and the
//     value / weight vectors are loaded lazily: voxels outside every frame's frustum box cost no memory traffic at all
// HBM-bound by construction: 8 bytes per touched voxel (4 read + 4 written), ~60 flops.
// hr() = "round to fp16": floats below always hold fp16 values exactly where the reference holds an fp16 tensor.
#include <cuda_fp16.h>

#include "common.cuh"

namespace dtb200 {

__device__ __forceinline__ float hr(float x) { return __half2float(__float2half_rn(x)); }
// torch.clamp / np.clip semantics: NaN propagates (fminf / fmaxf would drop it)
__device__ __forceinline__ float clamp_nan(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

struct Half8 {
  uint4 raw;
  __device__ __forceinline__ float get(int j) const {
    const uint32_t w = (&raw.x)[j >> 1];
    return __half2float(__ushort_as_half((unsigned short)((j & 1) ? (w >> 16) : (w & 0xffffu))));
  }
  __device__ __forceinline__ void set(int j, float v) {
    const uint32_t hbits = __half_as_ushort(__float2half_rn(v));
    uint32_t& w = (&raw.x)[j >> 1];
    w = (j & 1) ? ((w & 0x0000ffffu) | (hbits << 16)) : ((w & 0xffff0000u) | hbits);
  }
};

// F.grid_sample(mode="nearest", padding_mode="zeros", align_corners=False) index of one axis from the fp16 normalised
// coordinate g (tools/tsdf.py:476-483); returns false when the tap falls into the zero padding.
__device__ __forceinline__ bool nearest_index(float g, int size, int semantics, int& idx) {
  float i;
  if (semantics == DTB200_TSDF_SEMANTICS_ATEN_CPU) {
    // c10::Half arithmetic: ((g + 1) * size - 1) / 2 with one fp16 rounding per operation
    i = hr(DT_ADD(g, 1.f));
    i = hr(DT_MUL(i, (float)size));
    i = hr(DT_SUB(i, 1.f));
    i = hr(DT_DIV(i, 2.f));
  } else if (semantics == DTB200_TSDF_SEMANTICS_ATEN_CUDA_HALF_INDEX) {
    // GridSampler.cuh with a scalar_t index: `coord + 1.f` promotes to fp32; the result is stored back into an fp16 scalar
    i = hr(DT_DIV(DT_SUB(DT_MUL(DT_ADD(g, 1.f), (float)size), 1.f), 2.f));
  } else {
    // GridSampler.cu with opmath_t (fp32) coordinates: the index never goes back to fp16
    i = DT_DIV(DT_SUB(DT_MUL(DT_ADD(g, 1.f), (float)size), 1.f), 2.f);
  }
  float n = rintf(i);  // nearbyint, ties to even
  if (semantics == DTB200_TSDF_SEMANTICS_ATEN_CPU) {
    if (!isfinite(n)) n = 0.f;  // the x86 build converts NaN / +-inf to integer 0: such voxels read row / column 0
  } else {
    if (isnan(n)) n = 0.f;      // cvt.rzi saturates: +-inf stays out of bounds, NaN becomes 0
  }
  if (!(n >= 0.f && n <= (float)(size - 1))) return false;
  idx = (int)n;
  return true;
}

template <bool kGenCoords>
__global__ void __launch_bounds__(256) tsdf_integrate_kernel(const dtb200_tsdf_integrate_params p) {
  const int Y = p.dims[1], Z = p.dims[2];
  // the launch covers the index box [vox_begin, vox_end) only (a conservative cover of the frames' frustum boxes)
  const int sy = p.vox_end[1] - p.vox_begin[1], sz8 = (p.vox_end[2] - p.vox_begin[2]) >> 3;
  const long long nsub = (long long)(p.vox_end[0] - p.vox_begin[0]) * sy * sz8;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= nsub) return;
  const int izv = (p.vox_begin[2] >> 3) + (int)(t % sz8);
  const long long r = t / sz8;
  const int iy = p.vox_begin[1] + (int)(r % sy), ix = p.vox_begin[0] + (int)(r / sy);
  const long long vec = ((long long)ix * Y + iy) * (Z >> 3) + izv;  // index of this thread's 8-voxel vector in the volume
  const long long N = (long long)p.dims[0] * Y * Z;

  float cx[8], cy[8], cz[8];
  if (kGenCoords) {
    // TSDF.generate_voxel_coords + .half(): fp32 origin + fp32(index) * fp32(voxel_size), rounded once
    const float x = hr(DT_ADD(p.origin[0], DT_MUL((float)ix, p.voxel_size)));
    const float y = hr(DT_ADD(p.origin[1], DT_MUL((float)iy, p.voxel_size)));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      cx[j] = x, cy[j] = y;
      cz[j] = hr(DT_ADD(p.origin[2], DT_MUL((float)(izv * 8 + j), p.voxel_size)));
    }
  } else {
    const uint4* c = reinterpret_cast<const uint4*>(p.voxel_coords);
    Half8 hx, hy, hz;
    hx.raw = __ldg(c + vec), hy.raw = __ldg(c + (N >> 3) + vec), hz.raw = __ldg(c + 2 * (N >> 3) + vec);
#pragma unroll
    for (int j = 0; j < 8; ++j) cx[j] = hx.get(j), cy[j] = hy.get(j), cz[j] = hz.get(j);
  }

  Half8 val, wgt;
  bool loaded = false, dirty = false;
  uint4* vptr = reinterpret_cast<uint4*>(p.values) + vec;
  uint4* wptr = reinterpret_cast<uint4*>(p.weights) + vec;

  for (int b = 0; b < p.num_frames; ++b) {
    const dtb200_tsdf_frame& fr = p.frames[b];
    const unsigned short* depth = reinterpret_cast<const unsigned short*>(fr.depth);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // voxels strictly inside the frustum's bounding box (tools/tsdf.py:451-459)
      const bool in_box = cx[j] > fr.box_min[0] && cx[j] < fr.box_max[0] && cy[j] > fr.box_min[1] && cy[j] < fr.box_max[1] &&
                          cz[j] > fr.box_min[2] && cz[j] < fr.box_max[2];
      if (!in_box) continue;
      // project_to_camera (:398-410): fp16 (3x4) @ (4xN) with fp32 accumulation in k order (products of two fp16 numbers
      // are exact in fp32, so the FMA chain equals the sequential sum), one fp16 rounding
      float cam[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        float acc = DT_MUL(fr.P[i * 4 + 0], cx[j]);
        acc = DT_FMA(fr.P[i * 4 + 1], cy[j], acc);
        acc = DT_FMA(fr.P[i * 4 + 2], cz[j], acc);
        acc = DT_ADD(acc, fr.P[i * 4 + 3]);
        cam[i] = hr(acc);
      }
      const float vz = cam[2];
      const float px = hr(DT_DIV(cam[0], vz)), py = hr(DT_DIV(cam[1], vz));
      // 2 * pix / img_size - 1 (:472-473)
      const float gx = hr(DT_SUB(hr(DT_DIV(hr(DT_MUL(2.f, px)), (float)p.img_w)), 1.f));
      const float gy = hr(DT_SUB(hr(DT_DIV(hr(DT_MUL(2.f, py)), (float)p.img_h)), 1.f));
      int xi = 0, yi = 0;
      float sd = 0.f;
      if (nearest_index(gx, p.img_w, p.semantics, xi) && nearest_index(gy, p.img_h, p.semantics, yi)) {
        const int pix = yi * p.img_w + xi;
        sd = (fr.mask && !fr.mask[pix]) ? -1.f : __half2float(__ushort_as_half(__ldg(depth + pix)));
      }
      // InfiniTAM confidence (:486-493)
      float conf = hr(DT_DIV(hr(DT_SUB(sd, p.min_depth)), p.depth_range));
      conf = hr(DT_SUB(1.f, conf));
      conf = clamp_nan(conf, 0.25f, 1.f);
      conf = hr(DT_MUL(conf, conf));
      const float dist = hr(DT_SUB(sd, vz));
      const float tsdf = clamp_nan(hr(DT_DIV(dist, p.truncation)), -1.f, 1.f);
      const bool valid = vz > 0.f && dist > p.trunc_check_h && sd > 0.f && vz < p.max_depth_h && conf > 0.f;
      if (!valid) continue;
      if (!loaded) {
        val.raw = *vptr, wgt.raw = *wptr;
        loaded = true;
      }
      // weighted running average (:537-558)
      const float old_v = val.get(j), old_w = wgt.get(j);
      const float new_w = hr(DT_DIV(hr(DT_MUL(conf, 2.5f)), 100.f));
      const float total = hr(DT_ADD(old_w, new_w));
      const float num = hr(DT_ADD(hr(DT_MUL(old_v, old_w)), hr(DT_MUL(tsdf, new_w))));
      val.set(j, DT_DIV(num, total));
      wgt.set(j, total > 1.f ? 1.f : total);  // torch.clamp(max=1.0); NaN propagates
      dirty = true;
    }
  }
  if (dirty) {
    *vptr = val.raw;
    *wptr = wgt.raw;
  }
}

// One sample of TSDF.sample_tsdf's arithmetic (tools/tsdf.py:291-337, CPU branch: fp32 coordinates, align_corners=True, zeros
// padding) at world point `pt`: mode 0 = trilinear in ATen's corner order, 1 = nearest.  Shared by the sample kernel and the
// ray caster, so a ray-cast weight is bit-identical to sample_tsdf at the same point.
struct TsdfGrid {
  int X, Y, Z;
  float o[3];
  float voxel_size;
};
__device__ __forceinline__ void tsdf_index(const TsdfGrid& g, const float pt[3], float idx[3]) {
  const int dims[3] = {g.X, g.Y, g.Z};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float v = DT_SUB(pt[a], g.o[a]);
    v = DT_DIV(v, g.voxel_size);
    v = DT_DIV(v, (float)(dims[a] - 1));
    const float gg = DT_SUB(DT_MUL(v, 2.f), 1.f);
    idx[a] = DT_MUL(DT_DIV(DT_ADD(gg, 1.f), 2.f), (float)(dims[a] - 1));  // align_corners=True un-normalisation
  }
}
__device__ __forceinline__ float tsdf_sample_at(const __half* __restrict__ volume, const TsdfGrid& g, const float idx[3], int mode,
                                                float* min_corner = nullptr) {
  const int X = g.X, Y = g.Y, Z = g.Z;
  auto fetch = [&](float a0, float a1, float a2, float& v) -> bool {  // volume axis 0 / 1 / 2 index, zeros padding
    if (!(a0 >= 0.f && a0 <= (float)(X - 1) && a1 >= 0.f && a1 <= (float)(Y - 1) && a2 >= 0.f && a2 <= (float)(Z - 1))) return false;
    v = __half2float(volume[((long long)(int)a0 * Y + (int)a1) * Z + (int)a2]);
    return true;
  };
  if (mode == 1) {
    float v = 0.f;
    return fetch(rintf(idx[0]), rintf(idx[1]), rintf(idx[2]), v) ? v : 0.f;
  }
  // grid_sample's x is the LAST volume axis (the reference swaps the point's components, :309): x = axis 2, z = axis 0
  const float x = idx[2], y = idx[1], z = idx[0];
  const float x0 = floorf(x), y0 = floorf(y), z0 = floorf(z);
  const float x1 = DT_ADD(x0, 1.f), y1 = DT_ADD(y0, 1.f), z1 = DT_ADD(z0, 1.f);
  const float wx0 = DT_SUB(x1, x), wx1 = DT_SUB(x, x0), wy0 = DT_SUB(y1, y), wy1 = DT_SUB(y, y0);
  const float wz0 = DT_SUB(z1, z), wz1 = DT_SUB(z, z0);
  float acc = 0.f, lo = 3e38f;
  auto corner = [&](float cx, float cy, float cz, float wx, float wy, float wz) {
    float v;
    if (fetch(cz, cy, cx, v)) {
      acc = DT_ADD(acc, DT_MUL(v, DT_MUL(DT_MUL(wx, wy), wz)));
      lo = fminf(lo, v);
    } else {
      lo = fminf(lo, 0.f);   // a corner in the zero padding counts as unobserved
    }
  };
  // ATen grid_sampler_3d order: tnw, tne, tsw, tse, bnw, bne, bsw, bse
  corner(x0, y0, z0, wx0, wy0, wz0);
  corner(x1, y0, z0, wx1, wy0, wz0);
  corner(x0, y1, z0, wx0, wy1, wz0);
  corner(x1, y1, z0, wx1, wy1, wz0);
  corner(x0, y0, z1, wx0, wy0, wz1);
  corner(x1, y0, z1, wx1, wy0, wz1);
  corner(x0, y1, z1, wx0, wy1, wz1);
  corner(x1, y1, z1, wx1, wy1, wz1);
  if (min_corner) *min_corner = lo;
  return acc;
}

__global__ void tsdf_sample_kernel(const __half* __restrict__ volume, TsdfGrid g, const float* __restrict__ pts, float* __restrict__ out,
                                   long long n, int mode) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float pt[3] = {pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]};
  float idx[3];
  tsdf_index(g, pt, idx);
  out[i] = tsdf_sample_at(volume, g, idx, mode);
}

// ----------------------------------------------------------------------------------------------------------------------
// Rendered-depth hint by ray casting the TSDF (the mesh-free form): replaces, for the incremental loop
// of reference test_incremental.py:186-252, marching cubes (tools/marching_cubes/marching_cubes.cu) + the PyTorch3D depth
// rasteriser (utils/rendering_utils.py:25-53) + BackprojectDepth + TSDF.sample_tsdf + the threshold / NaN / mask rules
// (:238-252) by ONE kernel: one thread per hint pixel marches its camera ray through the volume, stops at the first
// front-facing zero crossing between two OBSERVED samples (all eight voxels around either sample carry a weight > 0 -- the
// reference meshes only cubes of active voxels, and a trilinear blend of observed and never-observed (-1) voxels would
// fake a crossing at every frustum boundary), interpolates the crossing depth, samples the fused confidence there with sample_tsdf's exact arithmetic,
// and writes depth_hint_b1hw (NaN where there is no surface or the confidence is below the threshold),
// depth_hint_mask_b1hw and sampled_weights_b1hw (0 where invalid).
// Marching rule (restated by oracle/oracle_tsdf.py::raycast_hint): start at z_near; step 2 voxels while |tsdf| >= 0.99 or the
// sample is outside the volume / unobserved, half a voxel otherwise (the truncation band is +-3 voxels wide, so a 2-voxel
// step cannot jump over it); depth is the camera-space z of the crossing, z* = z0 + (z1 - z0) * v0 / (v0 - v1).
// ----------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ray_point(const float* ray, const float* Rt, float z, float pt[3]) {
  const float c0 = DT_MUL(z, ray[0]), c1 = DT_MUL(z, ray[1]), c2 = DT_MUL(z, ray[2]);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float acc = DT_MUL(Rt[i * 4 + 0], c0);
    acc = DT_FMA(Rt[i * 4 + 1], c1, acc);
    acc = DT_FMA(Rt[i * 4 + 2], c2, acc);
    pt[i] = DT_ADD(acc, Rt[i * 4 + 3]);
  }
}

__global__ void __launch_bounds__(128) tsdf_raycast_kernel(const dtb200_tsdf_raycast_params p, TsdfGrid g) {
  const int x = blockIdx.x * 16 + (threadIdx.x & 15), y = blockIdx.y * 8 + (threadIdx.x >> 4);
  const int b = blockIdx.z;
  if (x >= p.width || y >= p.height) return;
  const __half* values = reinterpret_cast<const __half*>(p.values);
  const __half* weights = reinterpret_cast<const __half*>(p.weights);
  const float* invK = p.invK + b * 16;
  const float* Rt = p.world_T_cam + b * 16;
  float ray[3];
  {
    const float px = (float)x + 0.5f, py = (float)y + 0.5f;   // pixel centres, BackprojectDepth (utils/geometry_utils.py:34-39)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float acc = DT_MUL(invK[i * 4 + 0], px);
      acc = DT_FMA(invK[i * 4 + 1], py, acc);
      ray[i] = DT_ADD(acc, invK[i * 4 + 2]);
    }
  }
  const float big_step = DT_MUL(2.f, g.voxel_size), small_step = DT_MUL(0.5f, g.voxel_size);
  float z = p.z_near, z_prev = 0.f, v_prev = -1.f, w_prev = 0.f;
  float hit = -1.f;
  for (int it = 0; it < p.max_steps && z <= p.z_far; ++it) {
    float pt[3], idx[3];
    ray_point(ray, Rt, z, pt);
    tsdf_index(g, pt, idx);
    const bool inside = idx[0] >= 0.f && idx[0] <= (float)(g.X - 1) && idx[1] >= 0.f && idx[1] <= (float)(g.Y - 1) && idx[2] >= 0.f &&
                        idx[2] <= (float)(g.Z - 1);
    float v = -1.f, w = 0.f;
    if (inside) {
      float w_min;
      tsdf_sample_at(weights, g, idx, 0, &w_min);
      if (w_min > 0.f) w = 1.f, v = tsdf_sample_at(values, g, idx, 0);   // w: "observed" flag of this sample
    }
    if (v_prev > 0.f && v <= 0.f && w_prev > 0.f && w > 0.f) {
      hit = DT_ADD(z_prev, DT_DIV(DT_MUL(DT_SUB(z, z_prev), v_prev), DT_SUB(v_prev, v)));
      break;
    }
    z_prev = z, v_prev = v, w_prev = w;
    z = DT_ADD(z, fabsf(v) >= 0.99f ? big_step : small_step);
  }
  const long long o = ((long long)b * p.height + y) * p.width + x;
  float hint = __int_as_float(0x7fc00000), sw = 0.f;
  if (hit > 0.f) {
    float pt[3], idx[3];
    ray_point(ray, Rt, hit, pt);
    tsdf_index(g, pt, idx);
    sw = tsdf_sample_at(weights, g, idx, 0);
    // test_incremental.py:238-252: depth = camera z of the rendered point; hint[weights < threshold] = NaN; mask = ~isnan;
    // weights[~mask] = 0
    if (!(sw < p.weight_threshold)) hint = DT_MUL(hit, ray[2]);
    else sw = 0.f;
  }
  p.depth_hint[o] = hint;
  p.hint_mask[o] = isnan(hint) ? 0.f : 1.f;
  p.sampled_weights[o] = isnan(hint) ? 0.f : sw;
}

}  // namespace dtb200

using namespace dtb200;

extern "C" int dtb200_tsdf_integrate(const dtb200_tsdf_integrate_params* p, dtb200_stream_t stream) {
  if (!p || !p->values || !p->weights) return fail(DTB200_ERR_INVALID, "tsdf_integrate: null volume%s");
  if (p->num_frames < 1 || p->num_frames > DTB200_TSDF_MAX_FRAMES)
    return fail(DTB200_ERR_INVALID, "tsdf_integrate: num_frames must be 1..%s%lld", "", (long long)DTB200_TSDF_MAX_FRAMES);
  for (int a = 0; a < 3; ++a)
    if (p->dims[a] < 8 || (p->dims[a] & 7))
      return fail(DTB200_ERR_INVALID, "tsdf_integrate: volume dims must be positive multiples of 8 (TSDF.VOX_MOD), got %s%lld", "",
                  (long long)p->dims[a]);
  if (p->img_h < 1 || p->img_w < 1) return fail(DTB200_ERR_INVALID, "tsdf_integrate: bad image size%s");
  if (p->semantics != DTB200_TSDF_SEMANTICS_ATEN_CPU && p->semantics != DTB200_TSDF_SEMANTICS_ATEN_CUDA &&
      p->semantics != DTB200_TSDF_SEMANTICS_ATEN_CUDA_HALF_INDEX)
    return fail(DTB200_ERR_INVALID, "tsdf_integrate: unknown semantics%s");
  for (int b = 0; b < p->num_frames; ++b)
    if (!p->frames[b].depth) return fail(DTB200_ERR_INVALID, "tsdf_integrate: frame %s%lld has no depth map", "", (long long)b);
  if ((reinterpret_cast<uintptr_t>(p->values) | reinterpret_cast<uintptr_t>(p->weights) |
       reinterpret_cast<uintptr_t>(p->voxel_coords)) & 15)
    return fail(DTB200_ERR_INVALID, "tsdf_integrate: volume pointers must be 16-byte aligned%s");
  for (int a = 0; a < 3; ++a)
    if (p->vox_begin[a] < 0 || p->vox_end[a] > p->dims[a] || p->vox_begin[a] > p->vox_end[a] ||
        (a == 2 && ((p->vox_begin[a] | p->vox_end[a]) & 7)))
      return fail(DTB200_ERR_INVALID, "tsdf_integrate: vox_begin / vox_end must be a box inside dims with z bounds multiples of 8%s");
  const long long nvec = (long long)(p->vox_end[0] - p->vox_begin[0]) * (p->vox_end[1] - p->vox_begin[1]) *
                         ((p->vox_end[2] - p->vox_begin[2]) / 8);
  if (nvec == 0) return DTB200_OK;  // no voxel can be inside any frame's frustum box
  const long long blocks = (nvec + 255) / 256;
  if (blocks > 0x7fffffffLL) return fail(DTB200_ERR_UNSUPPORTED, "tsdf_integrate: volume too large%s");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (p->voxel_coords)
    tsdf_integrate_kernel<false><<<(unsigned)blocks, 256, 0, s>>>(*p);
  else
    tsdf_integrate_kernel<true><<<(unsigned)blocks, 256, 0, s>>>(*p);
  return check_launch("tsdf_integrate_kernel");
}

extern "C" int dtb200_tsdf_sample(const void* volume, const int32_t* dims, const float* origin_h, float voxel_size,
                                  const float* world_points, float* out, int64_t num_points, int32_t mode,
                                  dtb200_stream_t stream) {
  if (!volume || !dims || !origin_h || !world_points || !out) return fail(DTB200_ERR_INVALID, "tsdf_sample: null argument%s");
  if (mode != 0 && mode != 1) return fail(DTB200_ERR_INVALID, "tsdf_sample: mode must be 0 (trilinear) or 1 (nearest)%s");
  if (dims[0] < 2 || dims[1] < 2 || dims[2] < 2) return fail(DTB200_ERR_INVALID, "tsdf_sample: volume dims must be >= 2%s");
  if (num_points <= 0) return DTB200_OK;
  const long long blocks = (num_points + 255) / 256;
  TsdfGrid g;
  g.X = dims[0], g.Y = dims[1], g.Z = dims[2];
  g.o[0] = origin_h[0], g.o[1] = origin_h[1], g.o[2] = origin_h[2];
  g.voxel_size = voxel_size;
  tsdf_sample_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(volume), g, world_points, out, (long long)num_points, mode);
  return check_launch("tsdf_sample_kernel");
}

extern "C" int dtb200_tsdf_raycast(const dtb200_tsdf_raycast_params* p, dtb200_stream_t stream) {
  if (!p || !p->values || !p->weights || !p->invK || !p->world_T_cam || !p->depth_hint || !p->hint_mask || !p->sampled_weights)
    return fail(DTB200_ERR_INVALID, "tsdf_raycast: null argument%s");
  if (p->batch < 1 || p->height < 1 || p->width < 1) return fail(DTB200_ERR_INVALID, "tsdf_raycast: bad image size%s");
  if (p->dims[0] < 2 || p->dims[1] < 2 || p->dims[2] < 2) return fail(DTB200_ERR_INVALID, "tsdf_raycast: volume dims must be >= 2%s");
  if (!(p->voxel_size > 0.f) || !(p->z_near > 0.f) || !(p->z_far > p->z_near) || p->max_steps < 1)
    return fail(DTB200_ERR_INVALID, "tsdf_raycast: bad voxel size / depth range / step bound%s");
  TsdfGrid g;
  g.X = p->dims[0], g.Y = p->dims[1], g.Z = p->dims[2];
  g.o[0] = p->origin_h[0], g.o[1] = p->origin_h[1], g.o[2] = p->origin_h[2];
  g.voxel_size = p->voxel_size;
  dim3 grid((p->width + 15) / 16, (p->height + 7) / 8, p->batch);
  tsdf_raycast_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*p, g);
  return check_launch("tsdf_raycast_kernel");
}
