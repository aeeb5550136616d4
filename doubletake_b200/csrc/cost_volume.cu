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
#include "common.cuh"
#include "cv_common.cuh"

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
    for (int k = 0; k < p.views; ++k) {
      Projected pr = project_point(s_vc[k], X0, X1, X2);
      const float* sv = p.src_feats_nhwc + ((long long)b * p.views + k) * HW * kC;
      float4 wv = sample_quad(sv, q, pr.u, pr.v, p.height, p.width, invW, invH);
      float dot = quad_dot(wv, cur);
      bool depth_ok = pr.zp > 0.f;
      dot = DT_MUL(dot, depth_ok ? 1.f : 0.f);
      total = (k == 0) ? dot : DT_ADD(total, dot);
      if (last && live && q == 0) {
        bool bounds = (pr.u > 2.f) && (pr.u < (float)(p.width - 2)) && (pr.v > 2.f) && (pr.v < (float)(p.height - 2));
        write_masks(p, b, pix, k, depth_ok, bounds, any_d, any_b);
      }
    }
    if (live && q == 0) {
      p.volume[((long long)b * p.planes + d) * HW + pix] = total;
      if (last && p.mask_any) p.mask_any[(long long)b * HW + pix] = any_d && any_b;
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
