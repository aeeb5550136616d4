// Fused plane-sweep cost volumes for sm_100a: warp + match (+ metadata MLP + hint MLP) + arg-max + mask, one launch.
//
// Replaces the per-plane Python loops of the reference
//   modules/cost_volume.py:285-320, modules/feature_volume.py:186-352, modules/mesh_hint_volume.py:209-393
// and the geometry they call (utils/geometry_utils.py:55-93,178-199).  The arithmetic follows SURVEY.md
// Appendix B step by step (pixel centres +0.5, invK*pix then *d, (K@T) then P*X, z+1e-8, 1/(z+eps),
// 2*p*(1/W)-1, ATen's ((g+1)*W-1)/2, floor, (x1-x)(y1-y) weights, zeros padding), with FMA contraction disabled
// on the coordinate path (DT_* macros) so the sampling positions are those of the reference's separate torch ops.
//
// Thread mapping (both kernels): 4 lanes per (pixel, plane) -- lane q owns channels 4q..4q+3, so one bilinear tap is
// ONE 16-byte load per lane and the quad reads a contiguous 64-byte NHWC texel; 8 pixels x 4 lanes = one warp.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "cv_common.cuh"
#include "tc_common.cuh"

namespace dtb200 {

// ============================================================================================================
// DOT: CostVolumeManager (cost_volume.py:219-363).  Block = 8 warps; warp w -> pixel group (w / S), plane subset
// (w % S); S-way plane split keeps enough warps in flight when H*W is small.
// ============================================================================================================
template <int S>
__global__ void __launch_bounds__(kWarps * 32) cv_dot_kernel(const dtb200_cost_volume_params p) {
  constexpr int kGroups = kWarps / S;  // pixel groups per block
  __shared__ ViewConst s_vc[DTB200_MAX_VIEWS];
  __shared__ float s_best[kWarps][kPixPerWarp];
  __shared__ int s_besti[kWarps][kPixPerWarp];

  const int b = blockIdx.y;
  const int HW = p.height * p.width;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < p.views) {
    int k = threadIdx.x;
    load_view_const(s_vc[k], p.src_Ks + ((long long)b * p.views + k) * 16,
                    p.src_extrinsics + ((long long)b * p.views + k) * 16, p.src_poses + ((long long)b * p.views + k) * 16);
  }
  __syncthreads();

  const int group = warp / S, sub = warp % S;
  const int q = lane & 3;
  const int pix = (blockIdx.x * kGroups + group) * kPixPerWarp + (lane >> 2);
  const bool live = pix < HW;
  const int pixc = live ? pix : HW - 1;
  const int y = pixc / p.width, x = pixc - y * p.width;
  const float invW = 1.f / (float)p.width, invH = 1.f / (float)p.height;

  float r[3];
  backproject_ray(p.cur_invK + b * 16, x, y, r);
  float4 cur;
  {
    const float* c = p.cur_feats + ((long long)b * kC + q * 4) * HW + pixc;
    cur = make_float4(c[0], c[HW], c[2 * HW], c[3 * HW]);
  }
  float best = 0.f;
  int besti = 0x7fffffff;
  for (int d = sub; d < p.planes; d += S) {
    float depth = plane_depth(p, b, d, pixc);
    float X0 = DT_MUL(depth, r[0]), X1 = DT_MUL(depth, r[1]), X2 = DT_MUL(depth, r[2]);
    float total = 0.f;
    bool any_d = false, any_b = false;
    const bool last = (d == p.planes - 1);
    // Views in groups of four: lane q of the quad does the projection / sampling setup of view 4g + q ONCE and the quad shares
    // it by shuffles (round 1 recomputed it on all four lanes: the kernel was instruction-issue bound); then every lane
    // gathers its 4 channels of each view with branch-free predicated taps.  Same arithmetic per (pixel, plane, view) and the
    // same summation order over views as before: bit-identical results.
    for (int k0 = 0; k0 < p.views; k0 += 4) {
      SampleSetup mine;
      mine.off = 0, mine.mask = 0, mine.w[0] = mine.w[1] = mine.w[2] = mine.w[3] = 0.f;
      float my_m = 0.f;
      const int my_k = k0 + q;
      if (my_k < p.views) {
        const Projected pr = project_point(s_vc[my_k], X0, X1, X2);
        mine = sample_setup(pr.u, pr.v, p.height, p.width, invW, invH);
        const bool depth_ok = pr.zp > 0.f;
        my_m = depth_ok ? 1.f : 0.f;
        if (last && live) {
          const bool bounds = (pr.u > 2.f) && (pr.u < (float)(p.width - 2)) && (pr.v > 2.f) && (pr.v < (float)(p.height - 2));
          write_masks(p, b, pix, my_k, depth_ok, bounds, any_d, any_b);
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int k = k0 + c;
        if (k >= p.views) break;
        const int srcl = (lane & ~3) | c;
        SampleSetup ss;
        ss.off = __shfl_sync(0xffffffffu, mine.off, srcl);
        ss.mask = __shfl_sync(0xffffffffu, mine.mask, srcl);
        ss.w[0] = __shfl_sync(0xffffffffu, mine.w[0], srcl);
        ss.w[1] = __shfl_sync(0xffffffffu, mine.w[1], srcl);
        ss.w[2] = __shfl_sync(0xffffffffu, mine.w[2], srcl);
        ss.w[3] = __shfl_sync(0xffffffffu, mine.w[3], srcl);
        const float m = __shfl_sync(0xffffffffu, my_m, srcl);
        const float* sv = p.src_feats_nhwc + ((long long)b * p.views + k) * HW * kC;
        const float4 wv = sample_apply_nb(sv, q, ss, p.width);
        const float dot = DT_MUL(quad_dot(wv, cur), m);
        total = (k == 0) ? dot : DT_ADD(total, dot);
      }
    }
    {
      int dflag = any_d, bflag = any_b;   // the quad's lanes own different views
      dflag |= __shfl_xor_sync(0xffffffffu, dflag, 1);
      bflag |= __shfl_xor_sync(0xffffffffu, bflag, 1);
      dflag |= __shfl_xor_sync(0xffffffffu, dflag, 2);
      bflag |= __shfl_xor_sync(0xffffffffu, bflag, 2);
      if (live && q == 0) {
        p.volume[((long long)b * p.planes + d) * HW + pix] = total;
        if (last && p.mask_any) p.mask_any[(long long)b * HW + pix] = (dflag && bflag) ? 1 : 0;
      }
    }
    if (besti == 0x7fffffff || better(total, d, best, besti)) {
      best = total;
      besti = d;
    }
  }
  if (q == 0) {
    s_best[warp][lane >> 2] = best;
    s_besti[warp][lane >> 2] = besti;
  }
  __syncthreads();
  if (sub == 0 && q == 0 && live) {
    for (int s = 1; s < S; ++s) {
      float v = s_best[warp + s][lane >> 2];
      int i = s_besti[warp + s][lane >> 2];
      if (i != 0x7fffffff && (besti == 0x7fffffff || better(v, i, best, besti))) {
        best = v;
        besti = i;
      }
    }
    if (p.best_index) p.best_index[(long long)b * HW + pix] = besti;
    if (p.lowest_cost) p.lowest_cost[(long long)b * HW + pix] = plane_depth(p, b, besti, pix);
  }
}

// ============================================================================================================
// DOT, TMA-staged variant (north star: "TMA-staged feature tiles into shared memory"; compared against the __ldg gather
// above in tools/cv_sweep.py).  A block owns a 16 x 4 pixel tile and walks all (plane, view) steps.  For every step a
// producer warp projects the tile's four corners (a projective map with z > 0 keeps the tile's image inside the convex
// hull of the corners), and one cp.async.bulk.tensor box of kBW x kBH texels x 16 channels lands in a 4-deep shared-memory
// ring -- the hardware zero-fills texels outside the image, i.e. grid_sample's zeros padding.  The 8 consumer warps sample
// their bilinear taps from shared memory; a tap that falls outside the staged box (huge parallax, a corner behind the
// camera) is fetched from global memory exactly like the kernel above, so results are bit-identical in every case.
// ============================================================================================================
constexpr int kTmaTW = 16, kTmaTH = 4;          // pixel tile
constexpr int kTmaBW = 24, kTmaBH = 12;         // staged box (texels)
constexpr int kTmaStages = 4;
constexpr int kTmaBoxBytes = kTmaBW * kTmaBH * kC * 4;   // 18 KB
struct DotStageMeta {
  int bx, by, valid;
};

__global__ void __launch_bounds__(9 * 32) cv_dot_tma_kernel(const dtb200_cost_volume_params p, const __grid_constant__ CUtensorMap src_map) {
  extern __shared__ uint8_t dsm_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dsm_raw) + 127) & ~(uintptr_t)127);
  __shared__ ViewConst s_vc[DTB200_MAX_VIEWS];
  __shared__ DotStageMeta s_meta[kTmaStages];
  __shared__ uint64_t s_full[kTmaStages], s_empty[kTmaStages];

  const int b = blockIdx.y;
  const int HW = p.height * p.width;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_x = ceil_div(p.width, kTmaTW);
  const int tx0 = (blockIdx.x % tiles_x) * kTmaTW, ty0 = (blockIdx.x / tiles_x) * kTmaTH;
  if (tid < p.views)
    load_view_const(s_vc[tid], p.src_Ks + ((long long)b * p.views + tid) * 16, p.src_extrinsics + ((long long)b * p.views + tid) * 16,
                    p.src_poses + ((long long)b * p.views + tid) * 16);
  if (tid == 0) {
    for (int s = 0; s < kTmaStages; ++s) tc::mbar_init(&s_full[s], 1), tc::mbar_init(&s_empty[s], 8);
    tc::fence_mbar_init();
  }
  __syncthreads();
  const int steps = p.planes * p.views;
  const float invW = 1.f / (float)p.width, invH = 1.f / (float)p.height;

  if (warp == 8) {
    // ------------------------------------------------------------------ producer: corner projection -> box origin -> TMA
    const int cx = tx0 + ((lane & 1) ? min(kTmaTW, p.width - tx0) - 1 : 0);
    const int cy = ty0 + ((lane & 2) ? min(kTmaTH, p.height - ty0) - 1 : 0);
    float r[3];
    backproject_ray(p.cur_invK + b * 16, cx, cy, r);
    for (int s = 0; s < steps; ++s) {
      const int st = s % kTmaStages;
      tc::mbar_wait(&s_empty[st], (uint32_t)(((s / kTmaStages) & 1) ^ 1), 90, 32);
      const int d = s / p.views, k = s - d * p.views;
      // per-pixel planes: the corner's own plane depth (the bound is a heuristic there; the fallback path keeps it exact)
      const float depth = plane_depth(p, b, d, min(cy, p.height - 1) * p.width + min(cx, p.width - 1));
      const Projected pr = project_point(s_vc[k], DT_MUL(depth, r[0]), DT_MUL(depth, r[1]), DT_MUL(depth, r[2]));
      const float ix = pr.u - 0.5f, iy = pr.v - 0.5f;
      bool ok = (lane < 4) ? (pr.zp > 1e-6f && fabsf(ix) < 1e6f && fabsf(iy) < 1e6f) : true;
      float xmin = lane < 4 ? ix : 3e38f, xmax = lane < 4 ? ix : -3e38f, ymin = lane < 4 ? iy : 3e38f, ymax = lane < 4 ? iy : -3e38f;
#pragma unroll
      for (int o = 1; o < 4; o <<= 1) {
        xmin = fminf(xmin, __shfl_xor_sync(0xffffffffu, xmin, o)), xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymin = fminf(ymin, __shfl_xor_sync(0xffffffffu, ymin, o)), ymax = fmaxf(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
      }
      ok = __all_sync(0xffffffffu, ok);
      if (lane == 0) {
        int bx = 0, by = 0, valid = 0;
        if (ok) {
          bx = (int)floorf(xmin) - 1, by = (int)floorf(ymin) - 1;
          valid = ((int)floorf(xmax) + 2 - bx < kTmaBW) && ((int)floorf(ymax) + 2 - by < kTmaBH) ? 1 : 0;
          // a box entirely outside the image would be all zeros: skip the copy, the consumers' range tests already give 0
          if (bx >= p.width || by >= p.height || bx + kTmaBW <= 0 || by + kTmaBH <= 0) valid = 0;
        }
        s_meta[st].bx = bx, s_meta[st].by = by, s_meta[st].valid = valid;
        if (valid) {
          tc::mbar_arrive_expect_tx(&s_full[st], kTmaBoxBytes);
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                  tc::smem_u32(ring + (size_t)st * kTmaBoxBytes)),
              "l"(&src_map), "r"(0), "r"(bx), "r"(by), "r"(b * p.views + k), "r"(tc::smem_u32(&s_full[st]))
              : "memory");
        } else {
          tc::mbar_arrive(&s_full[st]);   // meta only (the release of the arrive orders the s_meta stores)
        }
      }
      __syncwarp();
    }
    return;
  }

  // -------------------------------------------------------------------- consumers: 64 pixels x 4 channel quads
  const int q = lane & 3;
  const int pi = tid >> 2;
  const int x = tx0 + (pi & 15), y = ty0 + (pi >> 4);
  const bool live = x < p.width && y < p.height;
  const int xc = min(x, p.width - 1), yc = min(y, p.height - 1);
  const int pixc = yc * p.width + xc, pix = pixc;
  float r[3];
  backproject_ray(p.cur_invK + b * 16, xc, yc, r);
  float4 cur;
  {
    const float* c = p.cur_feats + ((long long)b * kC + q * 4) * HW + pixc;
    cur = make_float4(c[0], c[HW], c[2 * HW], c[3 * HW]);
  }
  float best = 0.f;
  int besti = 0x7fffffff;
  int s = 0;
  for (int d = 0; d < p.planes; ++d) {
    const float depth = plane_depth(p, b, d, pixc);
    const float X0 = DT_MUL(depth, r[0]), X1 = DT_MUL(depth, r[1]), X2 = DT_MUL(depth, r[2]);
    float total = 0.f;
    bool any_d = false, any_b = false;
    const bool last = (d == p.planes - 1);
    for (int k0 = 0; k0 < p.views; k0 += 4) {
      SampleSetup mine;
      mine.off = 0, mine.mask = 0, mine.w[0] = mine.w[1] = mine.w[2] = mine.w[3] = 0.f;
      float my_m = 0.f;
      int my_x0 = 0, my_y0 = 0;   // texel of the nw tap (may be -1: the flat offset alone cannot tell x = -1 from x = W - 1)
      const int my_k = k0 + q;
      if (my_k < p.views) {
        const Projected pr = project_point(s_vc[my_k], X0, X1, X2);
        mine = sample_setup(pr.u, pr.v, p.height, p.width, invW, invH);
        if (mine.mask) {  // the same expressions sample_setup evaluates
          const float gx = DT_SUB(DT_MUL(DT_MUL(2.f, pr.u), invW), 1.f), gy = DT_SUB(DT_MUL(DT_MUL(2.f, pr.v), invH), 1.f);
          my_x0 = (int)floorf(DT_MUL(DT_SUB(DT_MUL(DT_ADD(gx, 1.f), (float)p.width), 1.f), 0.5f));
          my_y0 = (int)floorf(DT_MUL(DT_SUB(DT_MUL(DT_ADD(gy, 1.f), (float)p.height), 1.f), 0.5f));
        }
        const bool depth_ok = pr.zp > 0.f;
        my_m = depth_ok ? 1.f : 0.f;
        if (last && live) {
          const bool bounds = (pr.u > 2.f) && (pr.u < (float)(p.width - 2)) && (pr.v > 2.f) && (pr.v < (float)(p.height - 2));
          write_masks(p, b, pix, my_k, depth_ok, bounds, any_d, any_b);
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int k = k0 + c;
        if (k >= p.views) break;
        const int srcl = (lane & ~3) | c;
        SampleSetup ss;
        const int x0t = __shfl_sync(0xffffffffu, my_x0, srcl), y0t = __shfl_sync(0xffffffffu, my_y0, srcl);
        ss.mask = __shfl_sync(0xffffffffu, mine.mask, srcl);
        ss.w[0] = __shfl_sync(0xffffffffu, mine.w[0], srcl);
        ss.w[1] = __shfl_sync(0xffffffffu, mine.w[1], srcl);
        ss.w[2] = __shfl_sync(0xffffffffu, mine.w[2], srcl);
        ss.w[3] = __shfl_sync(0xffffffffu, mine.w[3], srcl);
        const float m = __shfl_sync(0xffffffffu, my_m, srcl);
        const int st = s % kTmaStages;
        tc::mbar_wait(&s_full[st], (uint32_t)((s / kTmaStages) & 1), 91);
        const int bx = s_meta[st].bx, by = s_meta[st].by, valid = s_meta[st].valid;
        const float* box = reinterpret_cast<const float*>(ring + (size_t)st * kTmaBoxBytes);
        const float* sv = p.src_feats_nhwc + ((long long)b * p.views + k) * HW * kC;
        const bool v0 = ss.mask & 1, v1 = ss.mask & 2, v2 = ss.mask & 4, v3 = ss.mask & 8;
        const float w0 = v0 ? ss.w[0] : 0.f, w1 = v1 ? ss.w[1] : 0.f, w2 = v2 ? ss.w[2] : 0.f, w3 = v3 ? ss.w[3] : 0.f;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        auto tap = [&](bool on, int dx, int dy) -> float4 {
          if (!on) return z;
          const int lx = x0t + dx - bx, ly = y0t + dy - by;
          if (valid && lx >= 0 && lx < kTmaBW && ly >= 0 && ly < kTmaBH)
            return *reinterpret_cast<const float4*>(box + ((size_t)(ly * kTmaBW + lx)) * kC + q * 4);
          return __ldg(reinterpret_cast<const float4*>(sv + ((long long)(y0t + dy) * p.width + x0t + dx) * kC + q * 4));
        };
        const float4 t0 = tap(v0, 0, 0), t1 = tap(v1, 1, 0), t2 = tap(v2, 0, 1), t3 = tap(v3, 1, 1);
        float4 wv;
        wv.x = DT_FMA(t3.x, w3, DT_FMA(t2.x, w2, DT_FMA(t1.x, w1, DT_FMA(t0.x, w0, 0.f))));
        wv.y = DT_FMA(t3.y, w3, DT_FMA(t2.y, w2, DT_FMA(t1.y, w1, DT_FMA(t0.y, w0, 0.f))));
        wv.z = DT_FMA(t3.z, w3, DT_FMA(t2.z, w2, DT_FMA(t1.z, w1, DT_FMA(t0.z, w0, 0.f))));
        wv.w = DT_FMA(t3.w, w3, DT_FMA(t2.w, w2, DT_FMA(t1.w, w1, DT_FMA(t0.w, w0, 0.f))));
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&s_empty[st]);
        ++s;
        const float dot = DT_MUL(quad_dot(wv, cur), m);
        total = (k == 0) ? dot : DT_ADD(total, dot);
      }
    }
    int dflag = any_d, bflag = any_b;
    dflag |= __shfl_xor_sync(0xffffffffu, dflag, 1);
    bflag |= __shfl_xor_sync(0xffffffffu, bflag, 1);
    dflag |= __shfl_xor_sync(0xffffffffu, dflag, 2);
    bflag |= __shfl_xor_sync(0xffffffffu, bflag, 2);
    if (live && q == 0) {
      p.volume[((long long)b * p.planes + d) * HW + pix] = total;
      if (last && p.mask_any) p.mask_any[(long long)b * HW + pix] = (dflag && bflag) ? 1 : 0;
    }
    if (besti == 0x7fffffff || better(total, d, best, besti)) best = total, besti = d;
  }
  if (live && q == 0) {
    if (p.best_index) p.best_index[(long long)b * HW + pix] = besti;
    if (p.lowest_cost) p.lowest_cost[(long long)b * HW + pix] = plane_depth(p, b, besti, pix);
  }
}

// ============================================================================================================
// MLP / MLP_HINT, math = EXACT: FeatureVolumeManager / FeatureMeshHintVolumeManager.
// Block = 8 pixels x all planes, processed 8 planes (64 rows) at a time:
//   phase A  gather+metadata: warp w builds the 26K+20 feature vectors of plane d0+w into smem (feature-major)
//   phase B  layer 1 (F->128) and layer 2 (128->128): 4x8 register tiles, weights streamed through smem in
//            32-feature chunks; each output accumulates bias first, then features in ascending order (fp32 FMA)
//   phase C  layer 3 (128->1) as a fixed shuffle tree, hint MLP (3->12->12->1) per row, volume store, running arg-max
// ============================================================================================================
constexpr int kRows = 64;        // rows (pixel, plane) per iteration
constexpr int kRowStride = 68;   // padded row stride of the feature-major tiles (16-B aligned, few bank conflicts)
constexpr int kHidden = 128;
constexpr int kChunk = 32;       // weight rows (features) per staged chunk


// One dense layer on the block's 64-row tile: acc[4 rows][8 cols] per thread.
// xt: smem activations, feature-major [nfeat][kRowStride]; wt_g: global weights (out=128, in=nfeat) row-major.
__device__ __forceinline__ void dense_layer(float (&acc)[4][8], const float* __restrict__ xt, int nfeat,
                                            const float* __restrict__ w_g, const float* __restrict__ b_g,
                                            float* __restrict__ wchunk /* [kChunk][kHidden+4] */) {
  const int tid = threadIdx.x;
  const int tr = tid >> 4, tc = tid & 15;  // 16 row-groups x 16 col-groups
  constexpr int kWS = kHidden + 4;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float bj = b_g[tc * 8 + j];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][j] = bj;
  }
  for (int f0 = 0; f0 < nfeat; f0 += kChunk) {
    const int nf = min(kChunk, nfeat - f0);
    __syncthreads();  // previous chunk fully consumed
    // stage W[:, f0:f0+nf] transposed -> wchunk[f][col]; global read is row-major in `in`, so lanes walk features
    for (int e = tid; e < nf * kHidden; e += kWarps * 32) {
      int col = e / nf, f = e - col * nf;
      wchunk[f * kWS + col] = __ldg(w_g + (long long)col * nfeat + f0 + f);
    }
    __syncthreads();
    for (int f = 0; f < nf; ++f) {
      float4 xv = *reinterpret_cast<const float4*>(xt + (f0 + f) * kRowStride + tr * 4);
      float4 w0 = *reinterpret_cast<const float4*>(wchunk + f * kWS + tc * 8);
      float4 w1 = *reinterpret_cast<const float4*>(wchunk + f * kWS + tc * 8 + 4);
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
      const float ws[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = DT_FMA(xs[i], ws[j], acc[i][j]);
    }
  }
}

template <bool kHint>
__global__ void __launch_bounds__(kWarps * 32) cv_mlp_exact_kernel(const dtb200_cost_volume_params p) {
  extern __shared__ __align__(16) float smem[];
  const int K = p.views;
  const int F = 26 * K + 20;
  float* xt = smem;                                   // [max(F,128)][kRowStride]
  float* wchunk = xt + (size_t)max(F, kHidden) * kRowStride;  // [kChunk][kHidden+4]
  float* s_score = wchunk + kChunk * (kHidden + 4);   // [kRows]
  __shared__ ViewConst s_vc[DTB200_MAX_VIEWS];
  __shared__ float s_best[kPixPerWarp];
  __shared__ int s_besti[kPixPerWarp];
  __shared__ float s_hint[3][kPixPerWarp];  // hint depth, weight, valid per pixel

  const int b = blockIdx.y;
  const int HW = p.height * p.width;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < K) {
    load_view_const(s_vc[tid], p.src_Ks + ((long long)b * K + tid) * 16, p.src_extrinsics + ((long long)b * K + tid) * 16,
                    p.src_poses + ((long long)b * K + tid) * 16);
  }
  const int q = lane & 3, pl = lane >> 2;
  const int pix = blockIdx.x * kPixPerWarp + pl;
  const bool live = pix < HW;
  const int pixc = live ? pix : HW - 1;
  const int y = pixc / p.width, x = pixc - y * p.width;
  const float invW = 1.f / (float)p.width, invH = 1.f / (float)p.height;

  if (tid < kPixPerWarp) {
    s_besti[tid] = 0x7fffffff;
    s_best[tid] = 0.f;
    if (kHint) {
      int ppix = min(blockIdx.x * kPixPerWarp + tid, HW - 1);
      int py = ppix / p.width, px = ppix - py * p.width;
      // F.interpolate(mode="nearest") to (H,W): src = min(floor(dst * in/out), in-1)   (mesh_hint_volume.py:186-204)
      int sy = min((int)floorf((float)py * ((float)p.hint_height / (float)p.height)), p.hint_height - 1);
      int sx = min((int)floorf((float)px * ((float)p.hint_width / (float)p.width)), p.hint_width - 1);
      long long o = ((long long)b * p.hint_height + sy) * p.hint_width + sx;
      bool valid = p.hint_mask[o] != 0.f;
      s_hint[0][tid] = valid ? p.depth_hint[o] : 0.f;
      s_hint[1][tid] = valid ? p.hint_weights[o] : 0.f;
      s_hint[2][tid] = valid ? 1.f : 0.f;
    }
  }
  __syncthreads();

  float r[3];
  backproject_ray(p.cur_invK + b * 16, x, y, r);
  float4 cur;
  {
    const float* c = p.cur_feats + ((long long)b * kC + q * 4) * HW + pixc;
    cur = make_float4(c[0], c[HW], c[2 * HW], c[3 * HW]);
  }
  // feature offsets (channel order of mesh_hint_volume.py:343-367)
  const int oCur = kC * K, oMask = oCur + kC, oDepth = oMask + K, oPlane = oDepth + K, oDot = oPlane + 1;
  const int oAngle = oDot + K, oRayCur = oAngle + K, oRaySrc = oRayCur + 3, oComb = oRaySrc + 3 * K;
  const int oRm = oComb + K, oTm = oRm + K;

  for (int d0 = 0; d0 < p.planes; d0 += kWarps) {
    // ------------------------------------------------------------------ phase A: features of plane d0+warp
    {
      const int d = min(d0 + warp, p.planes - 1);
      const int row = warp * kPixPerWarp + pl;
      float* xr = xt + row;
      float depth = plane_depth(p, b, d, pixc);
      float X0 = DT_MUL(depth, r[0]), X1 = DT_MUL(depth, r[1]), X2 = DT_MUL(depth, r[2]);
      // F.normalize(X, dim=1): X / max(||X||, 1e-12)
      float nn = DT_MUL(X0, X0);
      nn = DT_FMA(X1, X1, nn);
      nn = DT_FMA(X2, X2, nn);
      float nc = fmaxf(sqrtf(nn), 1e-12f);
      float rc0 = DT_DIV(X0, nc), rc1 = DT_DIV(X1, nc), rc2 = DT_DIV(X2, nc);
      // cosine_similarity re-normalises both rays with eps=1e-5 (ATen: (x1/max(|x1|,eps))*(x2/max(|x2|,eps))).sum)
      float n1 = DT_MUL(rc0, rc0);
      n1 = DT_FMA(rc1, rc1, n1);
      n1 = DT_FMA(rc2, rc2, n1);
      n1 = fmaxf(sqrtf(n1), 1e-5f);
      float a0 = DT_DIV(rc0, n1), a1 = DT_DIV(rc1, n1), a2 = DT_DIV(rc2, n1);
      bool any_d = false, any_b = false;
      const bool last = (d0 + warp == p.planes - 1);
      for (int k = 0; k < K; ++k) {
        const ViewConst& vc = s_vc[k];
        Projected pr = project_point(vc, X0, X1, X2);
        const float* sv = p.src_feats_nhwc + ((long long)b * K + k) * HW * kC;
        float4 wv = sample_quad(sv, q, pr.u, pr.v, p.height, p.width, invW, invH);
        float dot = quad_dot(wv, cur);
        bool depth_ok = pr.zp > 0.f;
        float m = depth_ok ? 1.f : 0.f;
        dot = DT_MUL(dot, m);
        float* xk = xr + (size_t)(kC * k + 4 * q) * kRowStride;
        xk[0] = wv.x;
        xk[kRowStride] = wv.y;
        xk[2 * kRowStride] = wv.z;
        xk[3 * kRowStride] = wv.w;
        if (q == 0) {
          xr[(size_t)(oMask + k) * kRowStride] = m;
          xr[(size_t)(oDepth + k) * kRowStride] = pr.zp;
          if (last && live) {
            bool bounds =
                (pr.u > 2.f) && (pr.u < (float)(p.width - 2)) && (pr.v > 2.f) && (pr.v < (float)(p.height - 2));
            write_masks(p, b, pix, k, depth_ok, bounds, any_d, any_b);
          }
        } else if (q == 1) {
          xr[(size_t)(oDot + k) * kRowStride] = dot;
          xr[(size_t)(oComb + k) * kRowStride] = vc.comb;
          xr[(size_t)(oRm + k) * kRowStride] = vc.rm;
          xr[(size_t)(oTm + k) * kRowStride] = vc.tm;
        } else {
          // source ray: normalize(X - t_src)   (geometry_utils.py:178-182)
          float y0 = DT_SUB(X0, vc.t[0]), y1 = DT_SUB(X1, vc.t[1]), y2 = DT_SUB(X2, vc.t[2]);
          float sn = DT_MUL(y0, y0);
          sn = DT_FMA(y1, y1, sn);
          sn = DT_FMA(y2, y2, sn);
          float sc = fmaxf(sqrtf(sn), 1e-12f);
          float rs0 = DT_DIV(y0, sc), rs1 = DT_DIV(y1, sc), rs2 = DT_DIV(y2, sc);
          if (q == 2) {
            xr[(size_t)(oRaySrc + 3 * k + 0) * kRowStride] = rs0;
            xr[(size_t)(oRaySrc + 3 * k + 1) * kRowStride] = rs1;
            xr[(size_t)(oRaySrc + 3 * k + 2) * kRowStride] = rs2;
          } else {
            float n2 = DT_MUL(rs0, rs0);
            n2 = DT_FMA(rs1, rs1, n2);
            n2 = DT_FMA(rs2, rs2, n2);
            n2 = fmaxf(sqrtf(n2), 1e-5f);
            float c0 = DT_MUL(a0, DT_DIV(rs0, n2)), c1 = DT_MUL(a1, DT_DIV(rs1, n2)), c2 = DT_MUL(a2, DT_DIV(rs2, n2));
            xr[(size_t)(oAngle + k) * kRowStride] = DT_ADD(DT_ADD(c0, c1), c2);
          }
        }
      }
      float* xc = xr + (size_t)(oCur + 4 * q) * kRowStride;
      xc[0] = cur.x;
      xc[kRowStride] = cur.y;
      xc[2 * kRowStride] = cur.z;
      xc[3 * kRowStride] = cur.w;
      if (q == 0) {
        xr[(size_t)oPlane * kRowStride] = depth;
        if (last && live && p.mask_any) p.mask_any[(long long)b * HW + pix] = any_d && any_b;
      } else if (q == 1) {
        xr[(size_t)(oRayCur + 0) * kRowStride] = rc0;
        xr[(size_t)(oRayCur + 1) * kRowStride] = rc1;
        xr[(size_t)(oRayCur + 2) * kRowStride] = rc2;
      }
    }
    __syncthreads();
    // ------------------------------------------------------------------ phase B: 202->128->128
    float acc[4][8];
    const int tr = tid >> 4, tc = tid & 15;
    dense_layer(acc, xt, F, p.w1, p.b1, wchunk);
    __syncthreads();  // every thread done reading the feature tile
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 v = make_float4(leaky01(acc[0][j]), leaky01(acc[1][j]), leaky01(acc[2][j]), leaky01(acc[3][j]));
      *reinterpret_cast<float4*>(xt + (size_t)(tc * 8 + j) * kRowStride + tr * 4) = v;
    }
    // dense_layer syncs before touching smem again
    dense_layer(acc, xt, kHidden, p.w2, p.b2, wchunk);
    // ------------------------------------------------------------------ phase C: 128->1 (+ hint MLP), store, arg-max
    {
      float w3[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) w3[j] = __ldg(p.w3 + tc * 8 + j);
      float b3 = __ldg(p.b3);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float s = DT_MUL(leaky01(acc[i][0]), w3[0]);
#pragma unroll
        for (int j = 1; j < 8; ++j) s = DT_FMA(leaky01(acc[i][j]), w3[j], s);
        // fixed tree over the 16 column groups (lanes differing in bits 0..3)
        s = DT_ADD(s, __shfl_xor_sync(0xffffffffu, s, 1));
        s = DT_ADD(s, __shfl_xor_sync(0xffffffffu, s, 2));
        s = DT_ADD(s, __shfl_xor_sync(0xffffffffu, s, 4));
        s = DT_ADD(s, __shfl_xor_sync(0xffffffffu, s, 8));
        if (tc == 0) s_score[tr * 4 + i] = DT_ADD(s, b3);
      }
    }
    __syncthreads();
    if (tid < kRows) {
      const int row = tid, rw = row / kPixPerWarp, rp = row % kPixPerWarp;
      const int d = d0 + rw;
      const int opix = blockIdx.x * kPixPerWarp + rp;
      float score = s_score[row];
      if (kHint) {
        float depth = plane_depth(p, b, min(d, p.planes - 1), min(opix, HW - 1));
        bool valid = s_hint[2][rp] != 0.f;
        float in[3] = {score, valid ? fabsf(DT_SUB(s_hint[0][rp], depth)) : -1.f, s_hint[1][rp]};
        float h1[12], h2[12];
#pragma unroll
        for (int o = 0; o < 12; ++o) {
          float a = __ldg(p.hb1 + o);
#pragma unroll
          for (int i = 0; i < 3; ++i) a = DT_FMA(in[i], __ldg(p.hw1 + o * 3 + i), a);
          h1[o] = leaky01(a);
        }
#pragma unroll
        for (int o = 0; o < 12; ++o) {
          float a = __ldg(p.hb2 + o);
#pragma unroll
          for (int i = 0; i < 12; ++i) a = DT_FMA(h1[i], __ldg(p.hw2 + o * 12 + i), a);
          h2[o] = leaky01(a);
        }
        float a = __ldg(p.hb3);
#pragma unroll
        for (int i = 0; i < 12; ++i) a = DT_FMA(h2[i], __ldg(p.hw3 + i), a);
        score = a;
      }
      if (d < p.planes && opix < HW) p.volume[((long long)b * p.planes + d) * HW + opix] = score;
      s_score[row] = score;
    }
    __syncthreads();
    if (tid < kPixPerWarp) {
      float best = s_best[tid];
      int besti = s_besti[tid];
      for (int w = 0; w < kWarps && d0 + w < p.planes; ++w) {
        float v = s_score[w * kPixPerWarp + tid];
        if (besti == 0x7fffffff || better(v, d0 + w, best, besti)) {
          best = v;
          besti = d0 + w;
        }
      }
      s_best[tid] = best;
      s_besti[tid] = besti;
    }
    // next iteration's phase A writes xt; every reader of xt/s_score is behind the barriers above + the one below
    __syncthreads();
  }
  if (tid < kPixPerWarp) {
    int opix = blockIdx.x * kPixPerWarp + tid;
    if (opix < HW) {
      if (p.best_index) p.best_index[(long long)b * HW + opix] = s_besti[tid];
      if (p.lowest_cost) p.lowest_cost[(long long)b * HW + opix] = plane_depth(p, b, s_besti[tid], opix);
    }
  }
}

static size_t mlp_exact_smem(int K) {
  int F = 26 * K + 20;
  int nf = F > kHidden ? F : kHidden;
  return ((size_t)nf * kRowStride + (size_t)kChunk * (kHidden + 4) + kRows) * sizeof(float);
}

int launch_cost_volume_tc(const dtb200_cost_volume_params& p, cudaStream_t stream);       // cost_volume_tc.cu
int prepare_cost_volume_tc(const dtb200_cost_volume_params& p, cudaStream_t stream);      // cost_volume_tc.cu
uint64_t cost_volume_tc_workspace_bytes(const dtb200_cost_volume_params& p);              // cost_volume_tc.cu
int launch_cost_volume_tch(const dtb200_cost_volume_params& p, cudaStream_t stream);      // cost_volume_tch.cu
int prepare_cost_volume_tch(const dtb200_cost_volume_params& p, cudaStream_t stream);     // cost_volume_tch.cu
uint64_t cost_volume_tch_workspace_bytes(const dtb200_cost_volume_params& p);             // cost_volume_tch.cu

}  // namespace dtb200

namespace dtb200 {

// development / measurement switch (tools/cv_sweep.py): DTB200_CV_DOT_VARIANT=tma selects the TMA-staged dot-product kernel
static int dot_variant() {
  const char* e = getenv("DTB200_CV_DOT_VARIANT");   // read per launch: the sweep tool switches it between points
  return (e && e[0] == 't') ? 1 : 0;
}

static int launch_cv_dot_tma(const dtb200_cost_volume_params& p, cudaStream_t stream) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    cudaDriverEntryPointQueryResult q;
    void* ptr = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess) encode = (EncodeFn)ptr;
  }
  if (!encode) return fail(DTB200_ERR_CUDA, "cost volume (dot, tma): cuTensorMapEncodeTiled entry point not available%s");
  CUtensorMap map;
  cuuint64_t dims[4] = {(cuuint64_t)kC, (cuuint64_t)p.width, (cuuint64_t)p.height, (cuuint64_t)p.batch * p.views};
  cuuint64_t strides[3] = {(cuuint64_t)kC * 4, (cuuint64_t)p.width * kC * 4, (cuuint64_t)p.height * p.width * kC * 4};
  cuuint32_t box[4] = {(cuuint32_t)kC, kTmaBW, kTmaBH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(p.src_feats_nhwc), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(DTB200_ERR_CUDA, "cost volume (dot, tma): cuTensorMapEncodeTiled failed with code %s%lld", "", (long long)r);
  const size_t smem = (size_t)kTmaStages * kTmaBoxBytes + 128;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(cv_dot_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  dim3 grid(ceil_div(p.width, kTmaTW) * ceil_div(p.height, kTmaTH), p.batch);
  cv_dot_tma_kernel<<<grid, 9 * 32, smem, stream>>>(p, map);
  return check_launch("cv_dot_tma_kernel");
}

}  // namespace dtb200

using namespace dtb200;

extern "C" uint64_t dtb200_cost_volume_workspace_bytes(const dtb200_cost_volume_params* p) {
  if (!p || p->kind == DTB200_VOLUME_DOT || p->views < 1 || p->views > DTB200_MAX_VIEWS) return 0;
  if (p->math == DTB200_MATH_TCH) return (p->batch < 1 || p->height < 1 || p->width < 1) ? 0 : cost_volume_tch_workspace_bytes(*p);
  if (p->math != DTB200_MATH_TC3X) return 0;
  return cost_volume_tc_workspace_bytes(*p);
}

extern "C" int dtb200_cost_volume_prepare(const dtb200_cost_volume_params* p, dtb200_stream_t stream) {
  if (!p) return fail(DTB200_ERR_INVALID, "cost_volume_prepare: null params%s");
  if ((p->math != DTB200_MATH_TC3X && p->math != DTB200_MATH_TCH) || p->kind == DTB200_VOLUME_DOT) return DTB200_OK;
  if (!p->w1 || !p->w2 || !p->b1 || p->views < 1 || p->views > DTB200_MAX_VIEWS)
    return fail(DTB200_ERR_INVALID, "cost_volume_prepare: weights / views missing%s");
  if (p->math == DTB200_MATH_TCH) return prepare_cost_volume_tch(*p, (cudaStream_t)stream);
  return prepare_cost_volume_tc(*p, (cudaStream_t)stream);
}

extern "C" int dtb200_cost_volume(const dtb200_cost_volume_params* pp, dtb200_stream_t stream_) {
  if (!pp) return fail(DTB200_ERR_INVALID, "dtb200_cost_volume: null params%s");
  const dtb200_cost_volume_params& p = *pp;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (p.channels != kC)
    return fail(DTB200_ERR_UNSUPPORTED, "cost volume: channels must be 16, got %s%lld", "", p.channels);
  if (p.views < 1 || p.views > DTB200_MAX_VIEWS)
    return fail(DTB200_ERR_INVALID, "cost volume: views must be in [1,16], got %s%lld", "", p.views);
  if (p.batch < 1 || p.height < 1 || p.width < 1 || p.planes < 1)
    return fail(DTB200_ERR_INVALID, "cost volume: empty shape%s");
  if (!p.cur_feats || !p.src_feats_nhwc || !p.src_extrinsics || !p.src_poses || !p.src_Ks || !p.cur_invK ||
      !p.plane_depths || !p.volume)
    return fail(DTB200_ERR_INVALID, "cost volume: null tensor pointer%s");
  const int HW = p.height * p.width;
  if (p.kind == DTB200_VOLUME_DOT && dot_variant() == 1) return launch_cv_dot_tma(p, stream);
  if (p.kind == DTB200_VOLUME_DOT) {
    // plane split: keep >= ~4 blocks of 8 warps per SM in flight when the map is small
    long long pixel_groups = (long long)ceil_div(HW, kPixPerWarp) * p.batch;
    int S = pixel_groups >= 148LL * 8 * 8 ? 1 : (pixel_groups >= 148LL * 8 * 4 ? 2 : (pixel_groups >= 148LL * 8 * 2 ? 4 : 8));
    while (S > p.planes) S >>= 1;
    dim3 grid(ceil_div(HW, kPixPerWarp * (kWarps / S)), p.batch);
    switch (S) {
      case 1: cv_dot_kernel<1><<<grid, kWarps * 32, 0, stream>>>(p); break;
      case 2: cv_dot_kernel<2><<<grid, kWarps * 32, 0, stream>>>(p); break;
      case 4: cv_dot_kernel<4><<<grid, kWarps * 32, 0, stream>>>(p); break;
      default: cv_dot_kernel<8><<<grid, kWarps * 32, 0, stream>>>(p); break;
    }
    return check_launch("cv_dot_kernel");
  }
  if (p.kind != DTB200_VOLUME_MLP && p.kind != DTB200_VOLUME_MLP_HINT)
    return fail(DTB200_ERR_INVALID, "cost volume: unknown kind %s%lld", "", p.kind);
  if (!p.w1 || !p.b1 || !p.w2 || !p.b2 || !p.w3 || !p.b3)
    return fail(DTB200_ERR_INVALID, "cost volume: null MLP weight pointer%s");
  const bool hint = p.kind == DTB200_VOLUME_MLP_HINT;
  if (hint && (!p.depth_hint || !p.hint_weights || !p.hint_mask || !p.hw1 || !p.hb1 || !p.hw2 || !p.hb2 || !p.hw3 ||
               !p.hb3 || p.hint_height < 1 || p.hint_width < 1))
    return fail(DTB200_ERR_INVALID, "cost volume: hint inputs / hint MLP weights missing%s");
  if (p.math == DTB200_MATH_TC3X) return launch_cost_volume_tc(p, stream);
  if (p.math == DTB200_MATH_TCH) return launch_cost_volume_tch(p, stream);
  if (p.math != DTB200_MATH_EXACT) return fail(DTB200_ERR_INVALID, "cost volume: unknown math mode %s%lld", "", p.math);
  size_t smem = mlp_exact_smem(p.views);
  dim3 grid(ceil_div(HW, kPixPerWarp), p.batch);
  cudaError_t e;
  if (hint) {
    e = cudaFuncSetAttribute(cv_mlp_exact_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(DTB200_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    cv_mlp_exact_kernel<true><<<grid, kWarps * 32, smem, stream>>>(p);
  } else {
    e = cudaFuncSetAttribute(cv_mlp_exact_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(DTB200_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    cv_mlp_exact_kernel<false><<<grid, kWarps * 32, smem, stream>>>(p);
  }
  return check_launch("cv_mlp_exact_kernel");
}
