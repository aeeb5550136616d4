// Fused plane-sweep feature volume on 5th-gen tensor cores (tcgen05 + TMEM), math = TC3X, sm_100a.
//
// FeatureVolumeManager / FeatureMeshHintVolumeManager (reference modules/feature_volume.py:186-352,
// modules/mesh_hint_volume.py:209-393) for 16 pixels x all planes per CTA, 8 planes (= 128 rows) per iteration:
//
//   producers (8 warps)   warp the K source maps (same fp32 geometry / bilinear code as the exact kernel), build the
//                         26K+20 metadata vector of every (pixel, plane) row and write it, split 3xTF32, straight into
//                         SWIZZLE_128B K-major operand tiles -- the 202-channel tensor of the reference never exists
//   GEMM1  (tcgen05)      D1[128 x 128] = rows x W1^T      (K = 26K+20 padded to 32; W1 tiles by cp.async.bulk)
//   epilogue-1            D1 (TMEM) -> +b1, LeakyReLU(0.01) -> split -> operand tiles of GEMM2 (never leaves the SM)
//   GEMM2  (tcgen05)      D2[128 x 128] = h1 x W2^T
//   epilogue-2            D2 -> +b2, LeakyReLU -> dot w3 + b3 -> hint MLP (3-12-12-1, fp32) -> volume store, running
//                         arg-max over planes -> lowest_cost, source-view masks of the last plane
//
// GEMM1 and GEMM2 share one 2-stage operand ring (K blocks 0..nkb1-1 feed D1, the next four feed D2); the MMA warp also
// issues the weight-tile bulk copies.  3xTF32: small*big + big*small + big*big per K step, fp32 accumulation in TMEM.
#include "common.cuh"
#include "cv_common.cuh"
#include "tc_common.cuh"

namespace dtb200 {

using namespace tc;

constexpr int kTRows = 128;      // rows per iteration = 8 planes x 16 pixels
constexpr int kTPix = 16;
constexpr int kTPlanes = 8;
constexpr int kTHidden = 128;
constexpr int kTProducerWarps = 8;
constexpr int kTThreads = (kTProducerWarps + 1) * 32;
constexpr int kTATile = kTRows * 128;            // 16 KB (one of big / small)
constexpr int kTBTile = kTHidden * 128;          // 16 KB
constexpr int kTAStage = 2 * kTATile;            // A_big | A_small, 32 KB
constexpr int kTBStage = 2 * kTBTile;            // B_big | B_small, 32 KB
constexpr int kTStages = 2;                      // A stages (producers <-> MMA)

__host__ __device__ inline int cvtc_nkb1(int K) { return (26 * K + 20 + 31) / 32; }
__host__ __device__ inline int cvtc_meta_stride(int K) { return ((10 * K + 4 + 3) / 4) * 4 + 4; }

// weight-tile ring depth: as deep as shared memory allows (the tiles are prefetched ahead of the MMAs)
__host__ __device__ inline int cvtc_bstages(int K) {
  const int fixed = 1024 + kTStages * kTAStage + kTRows * cvtc_meta_stride(K) * 4 + 4096;
  int n = (227 * 1024 - fixed) / kTBStage;
  return n > 4 ? 4 : n;
}
size_t cvtc_smem_bytes(int K) {
  return 1024 + (size_t)kTStages * kTAStage + (size_t)cvtc_bstages(K) * kTBStage + (size_t)kTRows * cvtc_meta_stride(K) * 4 + 4096;
}

__device__ __forceinline__ void producer_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ void store_split(uint8_t* a_big, uint32_t off, float4 v) {
  float4 big = make_float4(tf32_big(v.x), tf32_big(v.y), tf32_big(v.z), tf32_big(v.w));
  float4 small = make_float4(v.x - big.x, v.y - big.y, v.z - big.z, v.w - big.w);
  *reinterpret_cast<float4*>(a_big + off) = big;
  *reinterpret_cast<float4*>(a_big + kTATile + off) = small;
}

template <bool kHint>
__global__ void __launch_bounds__(kTThreads, 1) cv_mlp_tc_kernel(const dtb200_cost_volume_params p, const float* __restrict__ w1p,
                                                                 const float* __restrict__ w2p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int K = p.views;
  const int nmeta = 10 * K + 4;
  const int nkb1 = cvtc_nkb1(K);
  const int nkb = nkb1 + 4;
  const int MS = cvtc_meta_stride(K);
  const int SB = cvtc_bstages(K);
  uint8_t* stages = smem;                                   // A stages
  uint8_t* b_ring = smem + kTStages * kTAStage;             // weight-tile ring
  float* meta = reinterpret_cast<float*>(b_ring + (size_t)SB * kTBStage);  // [128][MS]
  uint8_t* tail = reinterpret_cast<uint8_t*>(meta + (size_t)kTRows * MS);
  ViewConst* s_vc = reinterpret_cast<ViewConst*>(tail);                       // [16] x 72 B = 1152
  float* s_partial = reinterpret_cast<float*>(tail + 1280);                   // [2][128]
  float* s_score = s_partial + 2 * kTRows;                                    // [128]
  float* s_best = s_score + kTRows;                                           // [16]
  int* s_besti = reinterpret_cast<int*>(s_best + kTPix);                      // [16]
  float* s_hint = reinterpret_cast<float*>(s_besti + kTPix);                  // [3][16]
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail + 1280 + 4 * (2 * kTRows + kTRows + kTPix + kTPix + 3 * kTPix) + 16);
  bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(bars) + 7) & ~(uintptr_t)7);
  uint64_t* full = bars;            // [2]  A stage filled by the producers
  uint64_t* empty = bars + 2;       // [2]  A stage consumed (tcgen05.commit)
  uint64_t* d1_full = bars + 4;
  uint64_t* d2_full = bars + 5;
  uint64_t* b_full = bars + 6;      // [4]  weight tile landed (complete_tx)
  uint64_t* b_empty = bars + 10;    // [4]  weight tile consumed (tcgen05.commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int HW = p.height * p.width;
  const int pix0 = blockIdx.x * kTPix;

  if (tid < K) {
    load_view_const(s_vc[tid], p.src_Ks + ((long long)b * K + tid) * 16, p.src_extrinsics + ((long long)b * K + tid) * 16,
                    p.src_poses + ((long long)b * K + tid) * 16);
  }
  if (tid < kTPix) {
    s_besti[tid] = 0x7fffffff;
    s_best[tid] = 0.f;
    if (kHint) {
      int ppix = min(pix0 + tid, HW - 1);
      int py = ppix / p.width, px = ppix - py * p.width;
      int sy = min((int)floorf((float)py * ((float)p.hint_height / (float)p.height)), p.hint_height - 1);
      int sx = min((int)floorf((float)px * ((float)p.hint_width / (float)p.width)), p.hint_width - 1);
      long long o = ((long long)b * p.hint_height + sy) * p.hint_width + sx;
      bool valid = p.hint_mask[o] != 0.f;
      s_hint[0 * kTPix + tid] = valid ? p.depth_hint[o] : 0.f;
      s_hint[1 * kTPix + tid] = valid ? p.hint_weights[o] : 0.f;
      s_hint[2 * kTPix + tid] = valid ? 1.f : 0.f;
    }
  }
  if (tid == 0) {
    for (int s = 0; s < kTStages; ++s) {
      mbar_init(&full[s], kTProducerWarps);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    mbar_init(d1_full, 1);
    mbar_init(d2_full, 1);
    fence_mbar_init();
  }
  if (warp == kTProducerWarps) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_d1 = tmem_base, tmem_d2 = tmem_base + kTHidden;
  const int num_iters = (p.planes + kTPlanes - 1) / kTPlanes;

  if (warp < kTProducerWarps) {
    // ================================================================================ producers / epilogues
    const int q = tid & 3;
    const int rh = tid >> 2;                 // 0..63 -> rows rh and rh + 64 (same pixel, planes dp and dp + 4)
    const int pi = rh & 15, dp0 = rh >> 4;
    const int pix = pix0 + pi;
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
    // meta offsets (channel order of mesh_hint_volume.py:343-367 after the 16(K+1) visual channels)
    const int oMask = 0, oDepth = K, oPlane = 2 * K, oDot = 2 * K + 1, oAngle = 3 * K + 1, oRayCur = 4 * K + 1;
    const int oRaySrc = 4 * K + 4, oComb = 7 * K + 4, oRm = 8 * K + 4, oTm = 9 * K + 4;
    const int b_meta = (K + 1) / 2;  // first K block that reads metadata channels
    // shared-memory byte offsets of this thread's two chunks in its two rows
    uint32_t soff[2][2];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int row = rh + 64 * rr;
        soff[rr][hh] = (uint32_t)row * 128u + (uint32_t)(((4 * hh + q) ^ (row & 7)) << 4);
      }
    // epilogue coordinates: TMEM lane == row
    const int erow = (warp & 3) * 32 + lane;
    const int ehalf = warp >> 2;

    int stage = 0, phase = 0;
    for (int iter = 0; iter < num_iters; ++iter) {
      const int d0 = iter * kTPlanes;
      // ---- per-row plane state
      float X[2][3], rc[2][3], an[2][3], depth[2];
      bool lastp[2];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int dreal = d0 + dp0 + 4 * rr;
        const int d = min(dreal, p.planes - 1);
        lastp[rr] = (dreal == p.planes - 1);
        depth[rr] = plane_depth(p, b, d, pixc);
        X[rr][0] = DT_MUL(depth[rr], r[0]), X[rr][1] = DT_MUL(depth[rr], r[1]), X[rr][2] = DT_MUL(depth[rr], r[2]);
        float nn = DT_MUL(X[rr][0], X[rr][0]);
        nn = DT_FMA(X[rr][1], X[rr][1], nn);
        nn = DT_FMA(X[rr][2], X[rr][2], nn);
        float nc = fmaxf(sqrtf(nn), 1e-12f);
        rc[rr][0] = DT_DIV(X[rr][0], nc), rc[rr][1] = DT_DIV(X[rr][1], nc), rc[rr][2] = DT_DIV(X[rr][2], nc);
        float n1 = DT_MUL(rc[rr][0], rc[rr][0]);
        n1 = DT_FMA(rc[rr][1], rc[rr][1], n1);
        n1 = DT_FMA(rc[rr][2], rc[rr][2], n1);
        n1 = fmaxf(sqrtf(n1), 1e-5f);
        an[rr][0] = DT_DIV(rc[rr][0], n1), an[rr][1] = DT_DIV(rc[rr][1], n1), an[rr][2] = DT_DIV(rc[rr][2], n1);
        float* mrow = meta + (size_t)(rh + 64 * rr) * MS;
        if (q == 0) mrow[oPlane] = depth[rr];
        if (q == 1) {
          mrow[oRayCur + 0] = rc[rr][0];
          mrow[oRayCur + 1] = rc[rr][1];
          mrow[oRayCur + 2] = rc[rr][2];
        }
      }
      bool any_d[2] = {false, false}, any_b[2] = {false, false};

      // ---- GEMM1 operand K blocks
      for (int kb = 0; kb < nkb1; ++kb) {
        if (kb == b_meta) producer_bar();  // every view's metadata has been written
        // The quad (4 lanes of one pixel) has 4 (row, view) combos in this K block: c = 2*rr + hh -> row rr, slot 2kb+hh.
        // Lane q does the per-combo scalar work of combo q exactly once (projection, sampling setup, masks, depths, source
        // ray, ray angle, pose constants) and the quad shares the sampling setup by shuffles; then every lane gathers its
        // own 4 channels for all 4 combos.
        float4 val[2][2];
        SampleSetup mine;
        mine.off = 0, mine.mask = 0, mine.w[0] = mine.w[1] = mine.w[2] = mine.w[3] = 0.f;
        float my_m = 0.f;
        {
          const int rr = q >> 1, slot = 2 * kb + (q & 1);
          if (slot < K) {
            const ViewConst& vc = s_vc[slot];
            const float X0 = rr ? X[1][0] : X[0][0], X1 = rr ? X[1][1] : X[0][1], X2 = rr ? X[1][2] : X[0][2];
            const Projected pr = project_point(vc, X0, X1, X2);
            mine = sample_setup(pr.u, pr.v, p.height, p.width, invW, invH);
            const bool depth_ok = pr.zp > 0.f;
            my_m = depth_ok ? 1.f : 0.f;
            float* mrow = meta + (size_t)(rh + 64 * rr) * MS;
            mrow[oMask + slot] = my_m;
            mrow[oDepth + slot] = pr.zp;
            mrow[oComb + slot] = vc.comb;
            mrow[oRm + slot] = vc.rm;
            mrow[oTm + slot] = vc.tm;
            if ((rr ? lastp[1] : lastp[0]) && live) {
              bool bounds = (pr.u > 2.f) && (pr.u < (float)(p.width - 2)) && (pr.v > 2.f) && (pr.v < (float)(p.height - 2));
              write_masks(p, b, pix, slot, depth_ok, bounds, any_d[0], any_b[0]);  // per-lane accumulators, merged below
            }
            // source ray normalize(X - t_src) and cos(cur ray, src ray); reciprocal square roots instead of the exact
            // kernel's IEEE divisions (the operands feed a 3xTF32 GEMM; ~1 ulp differences are irrelevant here)
            const float y0 = X0 - vc.t[0], y1 = X1 - vc.t[1], y2 = X2 - vc.t[2];
            const float inv = rsqrtf(fmaxf(y0 * y0 + y1 * y1 + y2 * y2, 1e-24f));
            const float rs0 = y0 * inv, rs1 = y1 * inv, rs2 = y2 * inv;
            mrow[oRaySrc + 3 * slot + 0] = rs0;
            mrow[oRaySrc + 3 * slot + 1] = rs1;
            mrow[oRaySrc + 3 * slot + 2] = rs2;
            const float inv2 = rsqrtf(fmaxf(rs0 * rs0 + rs1 * rs1 + rs2 * rs2, 1e-10f));
            const float a0 = rr ? an[1][0] : an[0][0], a1 = rr ? an[1][1] : an[0][1], a2 = rr ? an[1][2] : an[0][2];
            mrow[oAngle + slot] = (a0 * rs0 + a1 * rs1 + a2 * rs2) * inv2;
          }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int rr = c >> 1, hh = c & 1;
          const int slot = 2 * kb + hh;
          if (slot < K) {
            const int srcl = (lane & ~3) | c;
            SampleSetup ss;
            ss.off = __shfl_sync(0xffffffffu, mine.off, srcl);
            ss.mask = __shfl_sync(0xffffffffu, mine.mask, srcl);
            ss.w[0] = __shfl_sync(0xffffffffu, mine.w[0], srcl);
            ss.w[1] = __shfl_sync(0xffffffffu, mine.w[1], srcl);
            ss.w[2] = __shfl_sync(0xffffffffu, mine.w[2], srcl);
            ss.w[3] = __shfl_sync(0xffffffffu, mine.w[3], srcl);
            const float m = __shfl_sync(0xffffffffu, my_m, srcl);
            const float* sv = p.src_feats_nhwc + ((long long)b * K + slot) * HW * kC;
            const float4 wv = sample_apply(sv, q, ss, p.width);
            const float dot = DT_MUL(quad_dot(wv, cur), m);
            val[rr][hh] = wv;
            if (q == c) meta[(size_t)(rh + 64 * rr) * MS + oDot + slot] = dot;
          } else if (slot == K) {
            val[rr][hh] = cur;
          } else {
            const int m0 = kC * (slot - K - 1) + 4 * q;  // metadata channels m0..m0+3
            const float* mrow = meta + (size_t)(rh + 64 * rr) * MS;
            float4 v;
            v.x = (m0 + 0 < nmeta) ? mrow[m0 + 0] : 0.f;
            v.y = (m0 + 1 < nmeta) ? mrow[m0 + 1] : 0.f;
            v.z = (m0 + 2 < nmeta) ? mrow[m0 + 2] : 0.f;
            v.w = (m0 + 3 < nmeta) ? mrow[m0 + 3] : 0.f;
            val[rr][hh] = v;
          }
        }
        mbar_wait(&empty[stage], phase ^ 1, 10 + kb);
        uint8_t* a_big = stages + stage * kTAStage;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) store_split(a_big, soff[rr][hh], val[rr][hh]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[stage]);
        if (++stage == kTStages) stage = 0, phase ^= 1;
      }
      {
        // lanes q and q^1 handled the even / odd view slots of row q>>1: merge their any-view flags
        const int d_other = __shfl_xor_sync(0xffffffffu, (int)any_d[0], 1);  // unconditional: every lane must shuffle
        const int b_other = __shfl_xor_sync(0xffffffffu, (int)any_b[0], 1);
        const bool d_all = any_d[0] || (d_other != 0);
        const bool b_all = any_b[0] || (b_other != 0);
        const int rr = q >> 1;
        if ((q & 1) == 0 && p.mask_any && (rr ? lastp[1] : lastp[0]) && live) p.mask_any[(long long)b * HW + pix] = d_all && b_all;
      }

      // ---- epilogue-1: D1 -> LeakyReLU(D1 + b1) -> GEMM2 operand K blocks (4 x 32 columns)
      mbar_wait(d1_full, iter & 1, 20);
      tc_fence_after();
      for (int j = 0; j < 4; ++j) {
        float v[32];
        const bool mine = (j & 1) == ehalf;
        if (mine) {
          tmem_ld32(tmem_d1 + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(32 * j), v);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = leaky01(v[i] + __ldg(p.b1 + 32 * j + i));
        }
        mbar_wait(&empty[stage], phase ^ 1, 30 + j);
        if (mine) {
          uint8_t* a_big = stages + stage * kTAStage;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t off = (uint32_t)erow * 128u + (uint32_t)((c ^ (erow & 7)) << 4);
            store_split(a_big, off, make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]));
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[stage]);
        if (++stage == kTStages) stage = 0, phase ^= 1;
      }

      // ---- epilogue-2: D2 -> LeakyReLU(D2 + b2) . w3 + b3 -> hint MLP -> volume, running arg-max
      mbar_wait(d2_full, iter & 1, 21);
      tc_fence_after();
      {
        float s = 0.f;
#pragma unroll
        for (int cc = 0; cc < 64; cc += 32) {
          float v[32];
          const int n0 = 64 * ehalf + cc;
          tmem_ld32(tmem_d2 + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)n0, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) s = DT_FMA(leaky01(v[i] + __ldg(p.b2 + n0 + i)), __ldg(p.w3 + n0 + i), s);
        }
        s_partial[ehalf * kTRows + erow] = s;
      }
      tc_fence_before();
      producer_bar();
      if (tid < kTRows) {
        const int row = tid, dp = row >> 4, rpi = row & 15;
        const int d = d0 + dp;
        const int opix = pix0 + rpi;
        float score = DT_ADD(DT_ADD(s_partial[row], s_partial[kTRows + row]), __ldg(p.b3));
        if (kHint) {
          float dd = plane_depth(p, b, min(d, p.planes - 1), min(opix, HW - 1));
          bool valid = s_hint[2 * kTPix + rpi] != 0.f;
          float in[3] = {score, valid ? fabsf(DT_SUB(s_hint[rpi], dd)) : -1.f, s_hint[kTPix + rpi]};
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
      producer_bar();
      if (tid < kTPix) {
        float best = s_best[tid];
        int besti = s_besti[tid];
        for (int w = 0; w < kTPlanes && d0 + w < p.planes; ++w) {
          float v = s_score[w * kTPix + tid];
          if (besti == 0x7fffffff || better(v, d0 + w, best, besti)) best = v, besti = d0 + w;
        }
        s_best[tid] = best;
        s_besti[tid] = besti;
      }
      // the next iteration's meta / s_partial writes are ordered behind the barriers above by program order + the
      // producer_bar at its b_meta block and before its s_partial store
    }
    producer_bar();
    if (tid < kTPix) {
      int opix = pix0 + tid;
      if (opix < HW) {
        if (p.best_index) p.best_index[(long long)b * HW + opix] = s_besti[tid];
        if (p.lowest_cost) p.lowest_cost[(long long)b * HW + opix] = plane_depth(p, b, s_besti[tid], opix);
      }
    }
  } else {
    // ================================================================================ weight copies + MMA issue
    // whole warp convergent; single-thread instructions are issued under elect.sync (see tc_common.cuh).
    // Weight tiles run SB-1 K blocks ahead of the MMAs through their own ring, so their L2 latency is never exposed.
    constexpr uint32_t idesc = umma_idesc_tf32(kTRows, kTHidden);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t stages_u = smem_u32(stages), b_ring_u = smem_u32(b_ring);
    const long long total_u = (long long)num_iters * nkb;
    auto wtile = [&](long long g) -> const uint8_t* {  // global K-block sequence number -> packed weight tile
      const int u = (int)(g % nkb);
      return (u < nkb1) ? reinterpret_cast<const uint8_t*>(w1p) + (size_t)u * kTBStage
                        : reinterpret_cast<const uint8_t*>(w2p) + (size_t)(u - nkb1) * kTBStage;
    };
    auto issue_b = [&](long long g) {  // all lanes; one lane issues
      const int sb = (int)(g % SB), pb = (int)((g / SB) & 1);
      mbar_wait(&b_empty[sb], pb ^ 1, 80);
      if (elect_one()) {
        mbar_arrive_expect_tx(&b_full[sb], kTBStage);
        bulk_g2s(b_ring + (size_t)sb * kTBStage, wtile(g), kTBStage, &b_full[sb]);
      }
      __syncwarp();
    };
    for (long long g = 0; g < SB - 1 && g < total_u; ++g) issue_b(g);
    int stage = 0, phase = 0;
    long long g = 0;
    for (int iter = 0; iter < num_iters; ++iter) {
      for (int u = 0; u < nkb; ++u, ++g) {
        if (g + SB - 1 < total_u) issue_b(g + SB - 1);
        const int sb = (int)(g % SB), pb = (int)((g / SB) & 1);
        const uint32_t a_big_u = stages_u + stage * kTAStage, a_small_u = a_big_u + kTATile;
        const uint32_t b_big_u = b_ring_u + sb * kTBStage, b_small_u = b_big_u + kTBTile;
        mbar_wait(&full[stage], phase, 40 + u);
        mbar_wait(&b_full[sb], pb, 60 + u);
        tc_fence_after();
        const uint32_t dst = (u < nkb1) ? tmem_u : tmem_u + kTHidden;
        const bool first = (u == 0) || (u == nkb1);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t ko = ks * 32;
            const uint64_t da_b = umma_desc_k128(a_big_u + ko), da_s = umma_desc_k128(a_small_u + ko);
            const uint64_t db_b = umma_desc_k128(b_big_u + ko), db_s = umma_desc_k128(b_small_u + ko);
            umma_tf32(dst, da_s, db_b, idesc, !(first && ks == 0));
            umma_tf32(dst, da_b, db_s, idesc, true);
            umma_tf32(dst, da_b, db_b, idesc, true);
          }
          umma_commit(&empty[stage]);
          umma_commit(&b_empty[sb]);
          if (u == nkb1 - 1) umma_commit(d1_full);
          if (u == nkb - 1) umma_commit(d2_full);
        }
        __syncwarp();
        if (++stage == kTStages) stage = 0, phase ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kTProducerWarps) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// (128, in) row-major Linear weight -> nkb tiles of [big | small], each [128 rows][32 fp32] SWIZZLE_128B K-major
__global__ void pack_linear_tc_kernel(const float* __restrict__ w, float* __restrict__ packed, int in_features, int nkb) {
  const long long tile_floats = (long long)kTHidden * 32;
  const long long total = (long long)nkb * tile_floats;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int kb = (int)(i / tile_floats);
    int e = (int)(i - (long long)kb * tile_floats);
    int row = e / 32, kk = e % 32;
    int f = kb * 32 + kk;
    float x = f < in_features ? w[(long long)row * in_features + f] : 0.f;
    float big = tf32_big(x);
    float* tile = packed + (long long)kb * 2 * tile_floats;
    uint32_t off = sw128_offset(row, kk) / 4;
    tile[off] = big;
    tile[tile_floats + off] = x - big;
  }
}

uint64_t cost_volume_tc_workspace_bytes(const dtb200_cost_volume_params& p) {
  return (uint64_t)(cvtc_nkb1(p.views) + 4) * 2 * kTBTile;
}

// packs w1 / w2 into the workspace (idempotent; the host layer calls it once per weight version)
int prepare_cost_volume_tc(const dtb200_cost_volume_params& p, cudaStream_t stream) {
  if (!p.workspace || p.workspace_bytes < cost_volume_tc_workspace_bytes(p))
    return fail(DTB200_ERR_INVALID, "cost volume (tc3x): workspace too small (dtb200_cost_volume_workspace_bytes)%s");
  const int nkb1 = cvtc_nkb1(p.views);
  float* w1p = reinterpret_cast<float*>(p.workspace);
  float* w2p = w1p + (size_t)nkb1 * 2 * kTBTile / 4;
  pack_linear_tc_kernel<<<64, 256, 0, stream>>>(p.w1, w1p, 26 * p.views + 20, nkb1);
  int rc = check_launch("pack_linear_tc_kernel");
  if (rc != DTB200_OK) return rc;
  pack_linear_tc_kernel<<<64, 256, 0, stream>>>(p.w2, w2p, kTHidden, 4);
  return check_launch("pack_linear_tc_kernel");
}

int launch_cost_volume_tc(const dtb200_cost_volume_params& p, cudaStream_t stream) {
  if (!p.workspace || p.workspace_bytes < cost_volume_tc_workspace_bytes(p))
    return fail(DTB200_ERR_INVALID, "cost volume (tc3x): workspace missing/too small (dtb200_cost_volume_workspace_bytes)%s");
  if (!p.workspace_prepared) {
    int rc = prepare_cost_volume_tc(p, stream);
    if (rc != DTB200_OK) return rc;
  }
  const int nkb1 = cvtc_nkb1(p.views);
  const float* w1p = reinterpret_cast<const float*>(p.workspace);
  const float* w2p = w1p + (size_t)nkb1 * 2 * kTBTile / 4;
  const int HW = p.height * p.width;
  const size_t smem = cvtc_smem_bytes(p.views);
  if (smem > 227 * 1024) return fail(DTB200_ERR_UNSUPPORTED, "cost volume (tc3x): too many views for shared memory%s");
  dim3 grid(ceil_div(HW, kTPix), p.batch);
  cudaError_t e;
  if (p.kind == DTB200_VOLUME_MLP_HINT) {
    e = cudaFuncSetAttribute(cv_mlp_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(DTB200_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    cv_mlp_tc_kernel<true><<<grid, kTThreads, smem, stream>>>(p, w1p, w2p);
  } else {
    e = cudaFuncSetAttribute(cv_mlp_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(DTB200_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    cv_mlp_tc_kernel<false><<<grid, kTThreads, smem, stream>>>(p, w1p, w2p);
  }
  return check_launch("cv_mlp_tc_kernel");
}

}  // namespace dtb200
