// Fused channels-last convolution on 5th-gen tensor cores, math = TCH: tcgen05 kind::f16 with a 2-term fp16 split of
// activations and weights, sm_100a.
//
//   dst = act( conv_{k x k, stride}( concat_c[ src_i ] ) + bias (+ residual) )
//
// Activations live in HBM ALREADY SPLIT ("split16" layout): a (B, H, W, C) map is stored as (B, H, W, 2, C) fp16 -- per
// pixel C "big" values, big = fp16(x), followed by C "small" values, small = fp16((x - big) * 2048) -- the same 4C bytes per
// pixel as fp32, 22 significant bits.  The producing layer's epilogue writes this form, so a consumer needs NO operand
// conversion at all: per K block the loader issues two TMA boxes (one per plane: 4-D maps over the same buffer, the small
// plane's base shifted by 2C bytes) and the tiles land in shared memory ready for tcgen05.mma.  Compared with the 3xTF32
// kernels (conv_tc.cu) there are no splitter warps, half the shared-memory bytes per channel, half the weight bytes and
// kind::f16 runs at twice the TF32 rate.
//
// Per 16-wide K step:  D_main += A_big * W_big^T,  D_corr += A_big * W_small^T + A_small * W_big^T  (three products, issued
// as TWO instructions: the weight tile image is [W_big | W_small] = 2*BN contiguous K-major rows, so one N = 2*BN MMA
// writes [main | corr] and one N = BN MMA adds A_small * W_big^T to the corr columns).  Epilogue: main + corr / 2048
// (+ bias, residual, activation), split again, stored.  Dropped term small*small ~ 2^-22 |x||w|.
#include <cuda.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "conv_common.cuh"
#include "tc_common.cuh"

namespace dtb200 {

using namespace tc;

constexpr int kCM = 128;              // pixels per tile (UMMA M)
constexpr int kCK = 64;               // fp16 channels per K block = one 128-byte swizzled row
constexpr int kCTile = kCM * 128;     // 16 KB: [128 pixels][64 fp16]
constexpr int kCEpiWarps = 8;         // warps 0-7: TMEM lane quadrant = warp & 3, the two warps of a quadrant split the columns
constexpr int kCMmaWarp = 8, kCLoadA = 9, kCLoadB = 10;
constexpr int kCThreads = 11 * 32;

// K layout shared by the weight packer and the kernels: tap-major; inside a tap the sources in order, each cut into
// 64-channel chunks (the last chunk of a source may overhang: TMA zero-fills, the packer writes zero weights).
struct KLayoutH {
  int num_src, taps, kb_per_tap, num_kb;
  int src_c[DTB200_CONV_MAX_SRC], chunk_end[DTB200_CONV_MAX_SRC], c_begin[DTB200_CONV_MAX_SRC];
};
__host__ __device__ inline KLayoutH make_klayout_h(int num_src, const int32_t* src_c, int ksize) {
  KLayoutH k;
  k.num_src = num_src;
  k.taps = ksize * ksize;
  int chunks = 0, cb = 0;
  for (int s = 0; s < DTB200_CONV_MAX_SRC; ++s) {
    k.src_c[s] = s < num_src ? src_c[s] : 0;
    k.c_begin[s] = cb;
    cb += k.src_c[s];
    chunks += (k.src_c[s] + kCK - 1) / kCK;
    k.chunk_end[s] = chunks;
  }
  k.kb_per_tap = chunks;
  k.num_kb = chunks * k.taps;
  return k;
}
__host__ __device__ inline void klayout_h_decode(const KLayoutH& k, int kbi, int& tap, int& src, int& c0) {
  tap = kbi / k.kb_per_tap;
  const int r = kbi - tap * k.kb_per_tap;
  src = r < k.chunk_end[0] ? 0 : (r < k.chunk_end[1] ? 1 : 2);
  const int base = src == 0 ? 0 : (src == 1 ? k.chunk_end[0] : k.chunk_end[1]);
  c0 = (r - base) * kCK;
}
__host__ __device__ inline int klayout_h_src_c(const KLayoutH& k, int src) {
  return src == 0 ? k.src_c[0] : (src == 1 ? k.src_c[1] : k.src_c[2]);
}

struct HWork {  // persistent tile scheduler: item -> (m tile = (batch, tile row, tile col), n tile, K split)
  int tw, th, tiles_x, tiles_y;
  int m_tiles, n_tiles, splits, kb_per_split, num_kb_total;
  long long total;
};

struct HMaps {
  CUtensorMap big[DTB200_CONV_MAX_SRC], small[DTB200_CONV_MAX_SRC];
};

__device__ __forceinline__ void tma_load_4d_h(uint32_t smem_dst, const CUtensorMap* tm, int c, int x, int y, int b, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
          smem_dst),
      "l"(tm), "r"(c), "r"(x), "r"(y), "r"(b), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ uint4 ldg128u(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void stg128u(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  *reinterpret_cast<uint4*>(p) = make_uint4(a, b, c, d);
}
__device__ __forceinline__ float2 join_pair(uint32_t big, uint32_t small) {
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&big));
  const float2 s = __half22float2(*reinterpret_cast<const __half2*>(&small));
  return make_float2(fmaf(s.x, kHalfSplitInv, b.x), fmaf(s.y, kHalfSplitInv, b.y));
}

// Epilogue of one tile row (one output pixel) for 32 consecutive output channels starting at channel c0:
// v[] = main + corr / 2048 already combined.  Adds bias / residual, activates, splits, stores both planes.
__device__ __forceinline__ void tch_finish_chunk(const dtb200_conv_params& p, float (&v)[32], long long m, int c0) {
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 bv = ld4(p.bias + c0 + j);
      v[j] += bv.x, v[j + 1] += bv.y, v[j + 2] += bv.z, v[j + 3] += bv.w;
    }
  }
  const size_t row_bytes = (size_t)p.out_c * 4;   // per pixel: C big halves then C small halves
  if (p.residual) {
    const uint8_t* r = reinterpret_cast<const uint8_t*>(p.residual) + (size_t)m * row_bytes + (size_t)c0 * 2;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 rb = ldg128u(r + 16 * j), rs = ldg128u(r + (size_t)p.out_c * 2 + 16 * j);
      const float2 a0 = join_pair(rb.x, rs.x), a1 = join_pair(rb.y, rs.y), a2 = join_pair(rb.z, rs.z), a3 = join_pair(rb.w, rs.w);
      v[8 * j + 0] += a0.x, v[8 * j + 1] += a0.y, v[8 * j + 2] += a1.x, v[8 * j + 3] += a1.y;
      v[8 * j + 4] += a2.x, v[8 * j + 5] += a2.y, v[8 * j + 6] += a3.x, v[8 * j + 7] += a3.y;
    }
  }
  uint8_t* d = reinterpret_cast<uint8_t*>(p.dst) + (size_t)m * row_bytes + (size_t)c0 * 2;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t bg[4], sl[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      split_half2(activate(v[8 * j + 2 * i], p.act, p.act_slope), activate(v[8 * j + 2 * i + 1], p.act, p.act_slope), bg[i], sl[i]);
    stg128u(d + 16 * j, bg[0], bg[1], bg[2], bg[3]);
    stg128u(d + (size_t)p.out_c * 2 + 16 * j, sl[0], sl[1], sl[2], sl[3]);
  }
}

template <int BN>
struct TchCfg {
  static constexpr int kBBytes = 2 * BN * 128;                      // W_big | W_small: 2*BN K-major rows
  static constexpr int kStageBytes = 2 * kCTile + kBBytes;          // A_big | A_small | W_big | W_small
  static constexpr int kStages = BN == 64 ? 4 : 3;
  static constexpr int kAccCols = 2 * BN;                           // main | corr
  static constexpr int kTmemCols = 2 * kAccCols;                    // two accumulators
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 512;
};

// ======================================================================================================================
// Tap-major kernel: any 1x1 / 3x3, stride 1 / 2, 64- or 128-channel N tiles, optional split-K.  Persistent, one CTA per SM.
//   warps 0-3  epilogue (double-buffered TMEM accumulator)
//   warp 4     MMA issuer (elect.sync): per K block up to 4 K steps x 2 tcgen05.mma, tcgen05.commit frees the stage
//   warp 5     A loader: two TMA boxes per K block (big / small plane) = 64 channels x (TW x TH = 128) output pixels of one
//              source at one tap, SWIZZLE_128B, hardware zero fill outside the image and beyond the source's channels
//   warp 6     B loader: one cp.async.bulk of the pre-packed weight tile
// ======================================================================================================================
template <int BN>
__global__ void __launch_bounds__(kCThreads, 1) conv_tch_kernel(const dtb200_conv_params p, const __grid_constant__ HMaps maps,
                                                                KLayoutH kl, long long m_total, HWork wk, float* __restrict__ partial) {
  using Cfg = TchCfg<BN>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + S * Cfg::kStageBytes);
  uint64_t* full = bars;                // [S] A boxes + weight tile landed (2 arrivals + tx bytes)
  uint64_t* empty = full + S;           // [S] tcgen05.commit
  uint64_t* acc_full = empty + S;       // [2]
  uint64_t* acc_empty = acc_full + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 2), mbar_init(&empty[s], 1);
    for (int s = 0; s < 2; ++s) mbar_init(&acc_full[s], 1), mbar_init(&acc_empty[s], kCEpiWarps);
    fence_mbar_init();
  }
  if (warp == kCMmaWarp) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](long long item, int& bb, int& y0, int& x0, int& n_tile, int& kb_begin, int& num_kb, int& split) {
    split = (int)(item % wk.splits);
    const long long r = item / wk.splits;
    n_tile = (int)(r % wk.n_tiles);
    const int m_tile = (int)(r / wk.n_tiles);
    const int per_img = wk.tiles_x * wk.tiles_y;
    bb = m_tile / per_img;
    const int t = m_tile - bb * per_img;
    y0 = (t / wk.tiles_x) * wk.th;
    x0 = (t % wk.tiles_x) * wk.tw;
    kb_begin = split * wk.kb_per_split;
    num_kb = min(wk.kb_per_split, wk.num_kb_total - kb_begin);
  };

  if (warp < kCEpiWarps) {
    // ============================================================ epilogue
    const int qd = warp & 3, chalf = warp >> 2;
    const int row = qd * 32 + lane;
    const int ty = row / wk.tw, tx = row - ty * wk.tw;
    int use = 0;
    for (long long item = blockIdx.x; item < wk.total; item += gridDim.x, ++use) {
      int bb, y0, x0, n_tile, kb_begin, num_kb, split;
      decode(item, bb, y0, x0, n_tile, kb_begin, num_kb, split);
      const int buf = use & 1;
      const int oy = y0 + ty, ox = x0 + tx;
      const bool live = oy < p.out_h && ox < p.out_w;
      const long long m = ((long long)bb * p.out_h + oy) * p.out_w + ox;
      const int n_base = n_tile * BN;
      mbar_wait(&acc_full[buf], (use >> 1) & 1, 1, 128);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(buf * Cfg::kAccCols) + ((uint32_t)(qd * 32) << 16);
#pragma unroll 1
      for (int cc = chalf * (BN / 2); cc < (chalf + 1) * (BN / 2); cc += 32) {
        float v[32], c[32];
        tmem_ld32(taddr + (uint32_t)cc, v);
        tmem_ld32(taddr + (uint32_t)(BN + cc), c);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaf(c[j], kHalfSplitInv, v[j]);
        if (!live) continue;
        if (partial) {
          float* dst = partial + ((long long)split * m_total + m) * p.out_c + n_base + cc;
#pragma unroll
          for (int j = 0; j < 32; j += 8) stg256(dst + j, v + j);
        } else {
          tch_finish_chunk(p, v, m, n_base + cc);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  } else if (warp == kCMmaWarp) {
    // ============================================================ MMA issuer
    constexpr uint32_t idesc = umma_idesc_f16(kCM, BN), idesc2 = umma_idesc_f16(kCM, 2 * BN);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    constexpr uint32_t kDescHi = 64u | (1u << 14) | (2u << 29);   // SBO = 1024 B, version 1, SWIZZLE_128B
    const uint32_t lo_ring = ((smem_u32(ring) & 0x3FFFFu) >> 4) | (1u << 16);
    auto desc = [](uint32_t lo) { return ((uint64_t)kDescHi << 32) | lo; };
    int stage = 0, phase = 0, use = 0;
    for (long long item = blockIdx.x; item < wk.total; item += gridDim.x, ++use) {
      int bb, y0, x0, n_tile, kb_begin, num_kb, split;
      decode(item, bb, y0, x0, n_tile, kb_begin, num_kb, split);
      int tap, src, c0;
      klayout_h_decode(kl, kb_begin, tap, src, c0);
      const int buf = use & 1;
      mbar_wait(&acc_empty[buf], ((use >> 1) & 1) ^ 1, 3);
      tc_fence_after();
      const uint32_t tmem_d = tmem_u + (uint32_t)(buf * Cfg::kAccCols);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int cvalid = klayout_h_src_c(kl, src) - c0;
        const int ksteps = cvalid >= kCK ? 4 : (cvalid + 15) >> 4;
        mbar_wait(&full[stage], phase, 4);
        tc_fence_after();
        const uint32_t lo_a_big = lo_ring + (uint32_t)stage * (Cfg::kStageBytes >> 4);
        const uint32_t lo_a_small = lo_a_big + (kCTile >> 4), lo_b = lo_a_big + (2 * kCTile >> 4);
        if (elect_one()) {
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint32_t ko = ks * 2;   // 16 fp16 = 32 bytes along K inside the swizzled row, in 16-byte units
            umma_f16(tmem_d, desc(lo_a_big + ko), desc(lo_b + ko), idesc2, (kb | ks) != 0);        // [big x big | big x small]
            umma_f16(tmem_d + BN, desc(lo_a_small + ko), desc(lo_b + ko), idesc, true);             // small x big -> corr
          }
          umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == S) stage = 0, phase ^= 1;
        c0 += kCK;
        if (c0 >= klayout_h_src_c(kl, src)) {
          c0 = 0;
          if (++src == kl.num_src) src = 0;
        }
      }
      if (elect_one()) umma_commit(&acc_full[buf]);
      __syncwarp();
    }
  } else if (warp == kCLoadA) {
    // ============================================================ A loader
    const int pad = p.ksize / 2;
    const uint32_t ring_u = smem_u32(ring);
    int stage = 0, phase = 0;
    for (long long item = blockIdx.x; item < wk.total; item += gridDim.x) {
      int bb, y0, x0, n_tile, kb_begin, num_kb, split;
      decode(item, bb, y0, x0, n_tile, kb_begin, num_kb, split);
      int tap, src, c0;
      klayout_h_decode(kl, kb_begin, tap, src, c0);
      int ky = tap / p.ksize, kx = tap - ky * p.ksize;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1, 6, 32);
        if (elect_one()) {
          const uint32_t a_big = ring_u + (uint32_t)stage * Cfg::kStageBytes;
          mbar_arrive_expect_tx(&full[stage], 2 * kCTile);
          tma_load_4d_h(a_big, &maps.big[src], c0, x0 * p.stride + kx - pad, y0 * p.stride + ky - pad, bb, &full[stage]);
          tma_load_4d_h(a_big + kCTile, &maps.small[src], c0, x0 * p.stride + kx - pad, y0 * p.stride + ky - pad, bb, &full[stage]);
        }
        __syncwarp();
        if (++stage == S) stage = 0, phase ^= 1;
        c0 += kCK;
        if (c0 >= klayout_h_src_c(kl, src)) {
          c0 = 0;
          if (++src == kl.num_src) {
            src = 0;
            if (++kx == p.ksize) kx = 0, ++ky;
          }
        }
      }
    }
  } else if (warp == kCLoadB) {
    // ============================================================ B loader
    int stage = 0, phase = 0;
    for (long long item = blockIdx.x; item < wk.total; item += gridDim.x) {
      int bb, y0, x0, n_tile, kb_begin, num_kb, split;
      decode(item, bb, y0, x0, n_tile, kb_begin, num_kb, split);
      const uint8_t* wbase = reinterpret_cast<const uint8_t*>(p.weight) + ((size_t)n_tile * wk.num_kb_total + kb_begin) * Cfg::kBBytes;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1, 7, 32);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full[stage], Cfg::kBBytes);
          bulk_g2s(ring + (size_t)stage * Cfg::kStageBytes + 2 * kCTile, wbase + (size_t)kb * Cfg::kBBytes, Cfg::kBBytes, &full[stage]);
        }
        __syncwarp();
        if (++stage == S) stage = 0, phase ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kCMmaWarp) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ======================================================================================================================
// Halo-tile kernel for 3x3 / stride-1 layers with 64-channel N tiles and at least one full round of 8 x 16 tiles.
// The A side of a stage is ONE patch of (16+2) x (8+2) pixels x 64 channels per plane (2 x 23 KB), loaded by two TMA boxes; the
// 9 taps read it through shared-memory descriptors that are only SHIFTED by whole pixels (start + (ky*10 + kx) * 128 B, 8-row
// core-matrix stride = one patch row = 1280 B; SWIZZLE_128B works on absolute address bits, tools/halo_probe.cu).  Weight
// tiles stream through their own ring, three taps per stage.  Per 64-channel chunk: 2 TMA boxes + 72 MMAs.
// ======================================================================================================================
constexpr int kHTW = 8, kHTH = 16;
constexpr int kHPW = kHTW + 2, kHPH = kHTH + 2;
constexpr int kHPatchBytes = kHPW * kHPH * 128;                     // 23040
constexpr int kHSlotBytes = (kHPatchBytes + 1023) / 1024 * 1024;    // 23552
constexpr int kHAStageBytes = 2 * kHSlotBytes;                      // big | small
constexpr int kHAStagesN = 2;
constexpr int kHTaps = 3;                                           // taps per weight-ring stage
constexpr int kHBBytes = 2 * 64 * 128;                              // W_big | W_small of one (tap, chunk), BN = 64
constexpr int kHBStageBytes = kHTaps * kHBBytes;
constexpr int kHBStagesN = 2;
constexpr int kHaloSmemBytes = kHAStagesN * kHAStageBytes + kHBStagesN * kHBStageBytes + 1024 + 512;

__global__ void __launch_bounds__(kCThreads, 1) conv_tch_halo_kernel(const dtb200_conv_params p, const __grid_constant__ HMaps maps,
                                                                     KLayoutH kl, HWork wk) {
  constexpr int BN = 64, SA = kHAStagesN, SB = kHBStagesN;
  constexpr int kAccCols = 2 * BN, kTmemCols = 2 * kAccCols;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ring_a = smem;
  uint8_t* ring_b = ring_a + SA * kHAStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring_b + SB * kHBStageBytes);
  uint64_t* a_full = bars;              // [SA] both patch planes landed (1 arrival + tx)
  uint64_t* a_empty = a_full + SA;      // [SA] tcgen05.commit after the 9th tap
  uint64_t* b_full = a_empty + SA;      // [SB]
  uint64_t* b_empty = b_full + SB;      // [SB]
  uint64_t* acc_full = b_empty + SB;    // [2]
  uint64_t* acc_empty = acc_full + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < SA; ++s) mbar_init(&a_full[s], 1), mbar_init(&a_empty[s], 1);
    for (int s = 0; s < SB; ++s) mbar_init(&b_full[s], 1), mbar_init(&b_empty[s], 1);
    for (int s = 0; s < 2; ++s) mbar_init(&acc_full[s], 1), mbar_init(&acc_empty[s], kCEpiWarps);
    fence_mbar_init();
  }
  if (warp == kCMmaWarp) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int chunks = kl.kb_per_tap;     // 64-channel chunks of the concatenated sources = A stages per item

  auto decode = [&](long long item, int& bb, int& y0, int& x0, int& n_tile) {
    n_tile = (int)(item % wk.n_tiles);
    const int m_tile = (int)(item / wk.n_tiles);
    const int per_img = wk.tiles_x * wk.tiles_y;
    bb = m_tile / per_img;
    const int t = m_tile - bb * per_img;
    y0 = (t / wk.tiles_x) * kHTH;
    x0 = (t % wk.tiles_x) * kHTW;
  };

  if (warp < kCEpiWarps) {
    // ============================================================ epilogue
    const int qd = warp & 3, chalf = warp >> 2;
    const int row = qd * 32 + lane;
    const int ty = row / kHTW, tx = row - ty * kHTW;
    int use = 0;
    for (long long item = blockIdx.x; item < wk.total; item += gridDim.x, ++use) {
      int bb, y0, x0, n_tile;
      decode(item, bb, y0, x0, n_tile);
      const int buf = use & 1;
      const int oy = y0 + ty, ox = x0 + tx;
      const bool live = oy < p.out_h && ox < p.out_w;
      const long long m = ((long long)bb * p.out_h + oy) * p.out_w + ox;
      const int n_base = n_tile * BN;
      mbar_wait(&acc_full[buf], (use >> 1) & 1, 1, 128);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(buf * kAccCols) + ((uint32_t)(qd * 32) << 16);
      {
        const int cc = chalf * 32;   // BN = 64: one 32-column chunk per warp
        float v[32], c[32];
        tmem_ld32(taddr + (uint32_t)cc, v);
        tmem_ld32(taddr + (uint32_t)(BN + cc), c);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaf(c[j], kHalfSplitInv, v[j]);
        if (live) tch_finish_chunk(p, v, m, n_base + cc);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  } else if (warp == kCMmaWarp) {
    // ============================================================ MMA issuer
    constexpr uint32_t idesc = umma_idesc_f16(kCM, BN), idesc2 = umma_idesc_f16(kCM, 2 * BN);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    // A: SBO = one patch row (10 pixels = 1280 B); B: SBO = 1024 B.  High words are constant, low words move by plain adds.
    constexpr uint32_t kDescHiA = (uint32_t)(kHPW * 128 / 16) | (1u << 14) | (2u << 29);
    constexpr uint32_t kDescHiB = 64u | (1u << 14) | (2u << 29);
    const uint32_t lo_ring_a = ((smem_u32(ring_a) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t lo_ring_b = ((smem_u32(ring_b) & 0x3FFFFu) >> 4) | (1u << 16);
    auto desc_a = [](uint32_t lo) { return ((uint64_t)kDescHiA << 32) | lo; };
    auto desc_b = [](uint32_t lo) { return ((uint64_t)kDescHiB << 32) | lo; };
    int sa = 0, pa = 0, sb = 0, pb = 0, use = 0;
    for (long long item = blockIdx.x; item < wk.total; item += gridDim.x, ++use) {
      const int buf = use & 1;
      mbar_wait(&acc_empty[buf], ((use >> 1) & 1) ^ 1, 3);
      tc_fence_after();
      const uint32_t tmem_d = tmem_u + (uint32_t)(buf * kAccCols);
      int src = 0, c0 = 0;
      for (int ch = 0; ch < chunks; ++ch) {
        const int cvalid = klayout_h_src_c(kl, src) - c0;
        const int ksteps = cvalid >= kCK ? 4 : (cvalid + 15) >> 4;
        mbar_wait(&a_full[sa], pa, 4);
        tc_fence_after();
        const uint32_t lo_big = lo_ring_a + (uint32_t)sa * (kHAStageBytes >> 4);
        const uint32_t lo_small = lo_big + (kHSlotBytes >> 4);
#pragma unroll 1
        for (int g = 0; g < 9 / kHTaps; ++g) {
          mbar_wait(&b_full[sb], pb, 5);
          tc_fence_after();
          const uint32_t lo_b_stage = lo_ring_b + (uint32_t)sb * (kHBStageBytes >> 4);
          if (elect_one()) {
#pragma unroll
            for (int t = 0; t < kHTaps; ++t) {
              const int tap = g * kHTaps + t;
              const int ky = tap / 3, kx = tap - ky * 3;
              const uint32_t shift = (uint32_t)(ky * kHPW + kx) * (128u >> 4);   // whole pixels, in 16-byte units
              const uint32_t lo_b = lo_b_stage + (uint32_t)t * (kHBBytes >> 4);
              for (int ks = 0; ks < ksteps; ++ks) {
                const uint32_t ko = ks * 2;
                umma_f16(tmem_d, desc_a(lo_big + shift + ko), desc_b(lo_b + ko), idesc2, (ch | tap | ks) != 0);
                umma_f16(tmem_d + BN, desc_a(lo_small + shift + ko), desc_b(lo_b + ko), idesc, true);
              }
            }
            umma_commit(&b_empty[sb]);
            if (g == 9 / kHTaps - 1) umma_commit(&a_empty[sa]);
          }
          __syncwarp();
          if (++sb == SB) sb = 0, pb ^= 1;
        }
        if (++sa == SA) sa = 0, pa ^= 1;
        c0 += kCK;
        if (c0 >= klayout_h_src_c(kl, src)) c0 = 0, ++src;
      }
      if (elect_one()) umma_commit(&acc_full[buf]);
      __syncwarp();
    }
  } else if (warp == kCLoadA) {
    // ============================================================ A loader: two TMA boxes (patch planes) per 64-channel chunk
    const uint32_t ring_u = smem_u32(ring_a);
    int stage = 0, phase = 0;
    for (long long item = blockIdx.x; item < wk.total; item += gridDim.x) {
      int bb, y0, x0, n_tile;
      decode(item, bb, y0, x0, n_tile);
      int src = 0, c0 = 0;
      for (int ch = 0; ch < chunks; ++ch) {
        mbar_wait(&a_empty[stage], phase ^ 1, 6, 64);
        if (elect_one()) {
          const uint32_t dst = ring_u + (uint32_t)stage * kHAStageBytes;
          mbar_arrive_expect_tx(&a_full[stage], 2 * kHPatchBytes);
          tma_load_4d_h(dst, &maps.big[src], c0, x0 - 1, y0 - 1, bb, &a_full[stage]);
          tma_load_4d_h(dst + kHSlotBytes, &maps.small[src], c0, x0 - 1, y0 - 1, bb, &a_full[stage]);
        }
        __syncwarp();
        if (++stage == SA) stage = 0, phase ^= 1;
        c0 += kCK;
        if (c0 >= klayout_h_src_c(kl, src)) c0 = 0, ++src;
      }
    }
  } else if (warp == kCLoadB) {
    // ============================================================ B loader: weight tiles of (tap, chunk), packed tap-major
    int stage = 0, phase = 0;
    for (long long item = blockIdx.x; item < wk.total; item += gridDim.x) {
      int bb, y0, x0, n_tile;
      decode(item, bb, y0, x0, n_tile);
      const uint8_t* wbase = reinterpret_cast<const uint8_t*>(p.weight) + (size_t)n_tile * wk.num_kb_total * kHBBytes;
      for (int ch = 0; ch < chunks; ++ch) {
        for (int g = 0; g < 9 / kHTaps; ++g) {
          mbar_wait(&b_empty[stage], phase ^ 1, 7, 32);
          if (elect_one()) {
            mbar_arrive_expect_tx(&b_full[stage], kHBStageBytes);
#pragma unroll
            for (int t = 0; t < kHTaps; ++t)
              bulk_g2s(ring_b + (size_t)stage * kHBStageBytes + (size_t)t * kHBBytes,
                       wbase + (size_t)((g * kHTaps + t) * chunks + ch) * kHBBytes, kHBBytes, &b_full[stage]);
          }
          __syncwarp();
          if (++stage == SB) stage = 0, phase ^= 1;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kCMmaWarp) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ======================================================================================================================
// small kernels of the split16 world
// ======================================================================================================================
// OIHW (out_c, in_c, k, k) -> per (N tile, K block): [W_big | W_small], each [BN rows][64 fp16] in the SWIZZLE_128B K-major
// shared-memory image; K blocks follow KLayoutH.
__global__ void pack_weight_tch_kernel(const float* __restrict__ oihw, uint8_t* __restrict__ packed, int out_c, int in_c, KLayoutH kl,
                                       int bn) {
  const long long tile_elems = (long long)bn * kCK;
  const long long total = (long long)(out_c / bn) * kl.num_kb * tile_elems;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long t = i / tile_elems;
    const int e = (int)(i - t * tile_elems);
    const int n_tile = (int)(t / kl.num_kb), kbi = (int)(t - (long long)n_tile * kl.num_kb);
    const int row = e / kCK, kk = e % kCK;
    int tap, src, c0;
    klayout_h_decode(kl, kbi, tap, src, c0);
    const int c = c0 + kk;
    float x = 0.f;
    if (c < kl.src_c[src]) x = oihw[((long long)(n_tile * bn + row) * in_c + kl.c_begin[src] + c) * kl.taps + tap];
    x = fminf(fmaxf(x, -kHalfMax), kHalfMax);
    const __half big = __float2half_rn(x);
    const __half small = __float2half_rn((x - __half2float(big)) * kHalfSplitScale);
    uint8_t* tile = packed + (size_t)t * (size_t)(2 * bn * 128);
    *reinterpret_cast<__half*>(tile + sw128_offset_h(row, kk)) = big;
    *reinterpret_cast<__half*>(tile + (size_t)bn * 128 + sw128_offset_h(row, kk)) = small;
  }
}

// (N, C, H, W) fp32 -> (N, H, W, 2, C) split16: 32 x 32 shared-memory tile transpose of each sample's (C, H*W) matrix
__global__ void nchw_to_split16_kernel(const float* __restrict__ src, __half* __restrict__ dst, int c, int hw) {
  __shared__ float tile[32][33];
  const float* s = src + (size_t)blockIdx.z * c * hw;
  __half* d = dst + (size_t)blockIdx.z * hw * 2 * c;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int cc = c0 + i, pp = p0 + threadIdx.x;
    if (cc < c && pp < hw) tile[i][threadIdx.x] = s[(size_t)cc * hw + pp];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int pp = p0 + i, cc = c0 + threadIdx.x;
    if (pp < hw && cc < c) {
      const float x = fminf(fmaxf(tile[threadIdx.x][i], -kHalfMax), kHalfMax);
      const __half big = __float2half_rn(x);
      d[(size_t)pp * 2 * c + cc] = big;
      d[(size_t)pp * 2 * c + c + cc] = __float2half_rn((x - __half2float(big)) * kHalfSplitScale);
    }
  }
}

__global__ void split16_to_nchw_kernel(const __half* __restrict__ src, float* __restrict__ dst, int c, int hw) {
  __shared__ float tile[32][33];
  const __half* s = src + (size_t)blockIdx.z * hw * 2 * c;
  float* d = dst + (size_t)blockIdx.z * c * hw;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int pp = p0 + i, cc = c0 + threadIdx.x;
    if (pp < hw && cc < c) tile[i][threadIdx.x] = join_half(s[(size_t)pp * 2 * c + cc], s[(size_t)pp * 2 * c + c + cc]);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int cc = c0 + i, pp = p0 + threadIdx.x;
    if (cc < c && pp < hw) d[(size_t)cc * hw + pp] = tile[threadIdx.x][i];
  }
}

// 8 consecutive channels of a split16 map at pixel (y, x) as fp32
__device__ __forceinline__ void load8_split16(const uint8_t* base, int c_total, long long pix, int c, float (&v)[8]) {
  const uint8_t* r = base + (size_t)pix * c_total * 4 + (size_t)c * 2;
  const uint4 b = ldg128u(r), s = ldg128u(r + (size_t)c_total * 2);
  const float2 a0 = join_pair(b.x, s.x), a1 = join_pair(b.y, s.y), a2 = join_pair(b.z, s.z), a3 = join_pair(b.w, s.w);
  v[0] = a0.x, v[1] = a0.y, v[2] = a1.x, v[3] = a1.y, v[4] = a2.x, v[5] = a2.y, v[6] = a3.x, v[7] = a3.y;
}

// ksize == 0 descriptor: dst = x2 resample (bilinear align_corners=False / nearest) of a split16 map, written once
__global__ void resample_copy_h_kernel(const dtb200_conv_params p) {
  const int C = p.src_c[0], c8 = C / 8;
  const int sh = p.in_h / 2, sw = p.in_w / 2;
  const uint8_t* src = reinterpret_cast<const uint8_t*>(p.src[0]);
  uint8_t* dst = reinterpret_cast<uint8_t*>(p.dst);
  const long long total = (long long)p.batch * p.in_h * p.in_w * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c8) * 8;
    const long long pix = i / c8;
    const int x = (int)(pix % p.in_w);
    const long long r = pix / p.in_w;
    const int y = (int)(r % p.in_h), b = (int)(r / p.in_h);
    const long long sbase = (long long)b * sh * sw;
    float v[8];
    if (p.src_resample[0] == DTB200_RESAMPLE_NEAREST_UP2) {
      load8_split16(src, C, sbase + (long long)(y >> 1) * sw + (x >> 1), c, v);
    } else {
      int y0, y1, x0, x1;
      float hy0, hy1, wx0, wx1;
      up2_coord(y, sh, y0, y1, hy0, hy1);
      up2_coord(x, sw, x0, x1, wx0, wx1);
      float v00[8], v01[8], v10[8], v11[8];
      load8_split16(src, C, sbase + (long long)y0 * sw + x0, c, v00);
      load8_split16(src, C, sbase + (long long)y0 * sw + x1, c, v01);
      load8_split16(src, C, sbase + (long long)y1 * sw + x0, c, v10);
      load8_split16(src, C, sbase + (long long)y1 * sw + x1, c, v11);
#pragma unroll
      for (int j = 0; j < 8; ++j) {  // ATen upsample_bilinear2d: h0*(w0*v00 + w1*v01) + h1*(w0*v10 + w1*v11)
        const float top = DT_FMA(wx1, v01[j], DT_MUL(wx0, v00[j])), bot = DT_FMA(wx1, v11[j], DT_MUL(wx0, v10[j]));
        v[j] = DT_FMA(hy1, bot, DT_MUL(hy0, top));
      }
    }
    uint32_t bg[4], sl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_half2(v[2 * j], v[2 * j + 1], bg[j], sl[j]);
    uint8_t* d = dst + (size_t)pix * C * 4 + (size_t)c * 2;
    stg128u(d, bg[0], bg[1], bg[2], bg[3]);
    stg128u(d + (size_t)C * 2, sl[0], sl[1], sl[2], sl[3]);
  }
}

// 1x1 conv to few output channels from split16 sources, fp32 NHWC output (the log-depth heads).  One warp per pixel.
__global__ void __launch_bounds__(256) conv_head_h_kernel(const dtb200_conv_params p, long long pixels) {
  const int lane = threadIdx.x & 31;
  const long long pix = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (pix >= pixels) return;
  for (int n = 0; n < p.out_c; ++n) {
    float s = 0.f;
    int cg = 0;
    for (int si = 0; si < p.num_src; ++si) {
      const int C = p.src_c[si];
      const __half* x = reinterpret_cast<const __half*>(p.src[si]) + (size_t)pix * 2 * C;
      for (int c = lane; c < C; c += 32) s = DT_FMA(join_half(x[c], x[C + c]), __ldg(p.weight + (long long)(cg + c) * p.out_c + n), s);
      cg += C;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = DT_ADD(s, __shfl_xor_sync(0xffffffffu, s, o));
    if (lane == 0) {
      float v = p.bias ? DT_ADD(s, p.bias[n]) : s;
      if (p.residual) v = DT_ADD(v, p.residual[pix * p.out_c + n]);
      p.dst[pix * p.out_c + n] = activate(v, p.act, p.act_slope);
    }
  }
}

// Deterministic split-K reduction (splits summed in order) fused with bias / residual / activation, split16 output.
__global__ void splitk_epilogue_h_kernel(const dtb200_conv_params p, const float* __restrict__ partial, int splits, long long m_total) {
  const int c32 = p.out_c / 32;
  const long long total = m_total * c32;
  const long long mn = m_total * p.out_c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / c32;
    const int c0 = (int)(i - m * c32) * 32;
    float v[32];
    const float* src = partial + m * p.out_c + c0;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 a = ld4(src + j);
      v[j] = a.x, v[j + 1] = a.y, v[j + 2] = a.z, v[j + 3] = a.w;
    }
    for (int s = 1; s < splits; ++s) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 a = ld4(src + (long long)s * mn + j);
        v[j] += a.x, v[j + 1] += a.y, v[j + 2] += a.z, v[j + 3] += a.w;
      }
    }
    tch_finish_chunk(p, v, m, c0);
  }
}

// ======================================================================================================================
// host side
// ======================================================================================================================
int launch_conv_simt(const dtb200_conv_params& p, int in_c_total, cudaStream_t stream);
int launch_pack_simt(const float* oihw, float* packed, int out_c, int in_c, int ksize, cudaStream_t s);

static inline int tch_bn(int out_c) { return (out_c % 128 == 0) ? 128 : 64; }

uint64_t packed_floats_tch(int out_c, int num_src, const int32_t* src_c, int ksize) {
  int in_c = 0;
  for (int s = 0; s < num_src; ++s) in_c += src_c[s];
  if (out_c % 64 != 0) return (uint64_t)out_c * in_c * ksize * ksize;   // heads: [tap][in_c][out_c] fp32
  return (uint64_t)out_c * make_klayout_h(num_src, src_c, ksize).num_kb * 64;   // 2 planes x 64 fp16 per (row, K block)
}

int launch_pack_tch(const float* oihw, float* packed, int out_c, int num_src, const int32_t* src_c, int ksize, cudaStream_t stream) {
  int in_c = 0;
  for (int s = 0; s < num_src; ++s) in_c += src_c[s];
  if (out_c % 64 != 0) return launch_pack_simt(oihw, packed, out_c, in_c, ksize, stream);
  const KLayoutH kl = make_klayout_h(num_src, src_c, ksize);
  const long long total = (long long)out_c * kl.num_kb * kCK;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_weight_tch_kernel<<<blocks, 256, 0, stream>>>(oihw, reinterpret_cast<uint8_t*>(packed), out_c, in_c, kl, tch_bn(out_c));
  return check_launch("pack_weight_tch_kernel");
}

static void tch_tile_shape(int out_h, int out_w, int& tw, int& th) {
  long long best = -1;
  for (int w = 128; w >= 8; w >>= 1) {
    const int h = 128 / w;
    const long long tiles = (long long)((out_w + w - 1) / w) * ((out_h + h - 1) / h);
    if (best < 0 || tiles < best || (tiles == best && w == 16)) best = tiles, tw = w, th = h;
  }
}
static long long tch_m_tiles(const dtb200_conv_params& p) {
  int tw, th;
  tch_tile_shape(p.out_h, p.out_w, tw, th);
  return (long long)p.batch * ((p.out_w + tw - 1) / tw) * ((p.out_h + th - 1) / th);
}

// split-K plan for maps with fewer (M, N) tiles than SMs (persistent kernel: time ~ rounds x (K blocks per item + fixed
// per-item cost) + reduction cost).  148 on purpose: the plan (and the fp32 summation order it implies) must not depend
// on the device the workspace was sized on.
static int tch_splits(long long m_tiles, int out_c, int num_kb) {
  const long long ctas = m_tiles * (out_c / tch_bn(out_c));
  if (ctas >= 148 || num_kb < 4) return 1;
  int best = 1;
  double best_cost = 1e30;
  for (int sp = 1; sp <= 32 && sp * 2 <= num_kb; ++sp) {
    const int kb_per = (num_kb + sp - 1) / sp;
    if ((num_kb + kb_per - 1) / kb_per != sp) continue;
    const long long rounds = (ctas * sp + 147) / 148;
    const double cost = (double)rounds * (kb_per + 2.0) + (sp > 1 ? 4.0 + 0.5 * sp : 0.0);
    if (cost < best_cost - 1e-9) best_cost = cost, best = sp;
  }
  return best;
}

uint64_t conv_tch_workspace_bytes(const dtb200_conv_params& p) {
  if (p.ksize == 0 || p.out_c % 64 != 0) return 0;
  const KLayoutH kl = make_klayout_h(p.num_src, p.src_c, p.ksize);
  const int splits = tch_splits(tch_m_tiles(p), p.out_c, kl.num_kb);
  if (splits > 1) return (uint64_t)splits * p.batch * p.out_h * p.out_w * p.out_c * sizeof(float);
  return 0;
}

typedef CUresult (*TensorMapEncodeFnH)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFnH g_encode_h = nullptr;
static int g_tch_sms[64] = {0};
static std::once_flag g_tch_once[64];

int conv_tch_init() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  std::call_once(g_tch_once[dev], [dev] {
    if (!g_encode_h) {
      cudaDriverEntryPointQueryResult q;
      void* ptr = nullptr;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess) g_encode_h = (TensorMapEncodeFnH)ptr;
    }
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    g_tch_sms[dev] = sms > 0 ? sms : 148;
    cudaFuncSetAttribute(conv_tch_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TchCfg<64>::kSmemBytes);
    cudaFuncSetAttribute(conv_tch_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TchCfg<128>::kSmemBytes);
    cudaFuncSetAttribute(conv_tch_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHaloSmemBytes);
    cudaGetLastError();
  });
  return g_tch_sms[dev];
}

static int encode_split16_maps(const dtb200_conv_params& p, HMaps& maps, cuuint32_t box_w, cuuint32_t box_h, cuuint32_t estride) {
  memset(&maps, 0, sizeof(maps));
  for (int s = 0; s < p.num_src; ++s) {
    const cuuint64_t C = (cuuint64_t)p.src_c[s];
    cuuint64_t dims[4] = {C, (cuuint64_t)p.in_w, (cuuint64_t)p.in_h, (cuuint64_t)p.batch};
    cuuint64_t strides[3] = {C * 4, (cuuint64_t)p.in_w * C * 4, (cuuint64_t)p.in_h * p.in_w * C * 4};
    cuuint32_t box[4] = {(cuuint32_t)kCK, box_w, box_h, 1};
    cuuint32_t estr[4] = {1, estride, estride, 1};
    for (int plane = 0; plane < 2; ++plane) {
      void* base = const_cast<uint8_t*>(reinterpret_cast<const uint8_t*>(p.src[s]) + (size_t)plane * C * 2);
      CUresult r = g_encode_h(plane ? &maps.small[s] : &maps.big[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(DTB200_ERR_CUDA, "conv (tch): cuTensorMapEncodeTiled failed with code %s%lld", "", (long long)r);
    }
  }
  return DTB200_OK;
}

int launch_resample_copy_h(const dtb200_conv_params& p, cudaStream_t stream) {
  if (p.num_src != 1 || p.src_resample[0] == DTB200_RESAMPLE_NONE || p.src_c[0] % 8 != 0 || p.out_c != p.src_c[0] ||
      p.out_h != p.in_h || p.out_w != p.in_w || (p.in_h & 1) || (p.in_w & 1) || !p.src[0] || !p.dst)
    return fail(DTB200_ERR_INVALID, "resample copy (tch, ksize=0): needs one x2-resampled split16 source (C % 8 == 0) and a matching dst%s");
  const long long total = (long long)p.batch * p.in_h * p.in_w * (p.src_c[0] / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  resample_copy_h_kernel<<<blocks, 256, 0, stream>>>(p);
  return check_launch("resample_copy_h_kernel");
}

int launch_split16_transpose(const void* src, void* dst, int n, int c, int hw, bool to_split, cudaStream_t stream) {
  if (!src || !dst || n < 1 || c < 1 || hw < 1) return fail(DTB200_ERR_INVALID, "split16 transpose: bad arguments%s");
  dim3 grid(ceil_div(hw, 32), ceil_div(c, 32), n);
  if (to_split) nchw_to_split16_kernel<<<grid, dim3(32, 8), 0, stream>>>(reinterpret_cast<const float*>(src), reinterpret_cast<__half*>(dst), c, hw);
  else split16_to_nchw_kernel<<<grid, dim3(32, 8), 0, stream>>>(reinterpret_cast<const __half*>(src), reinterpret_cast<float*>(dst), c, hw);
  return check_launch("split16 transpose");
}

int launch_conv_tch(const dtb200_conv_params& p, int in_c_total, cudaStream_t stream) {
  (void)in_c_total;
  for (int s = 0; s < p.num_src; ++s) {
    if (p.src_c[s] % 8 != 0)
      return fail(DTB200_ERR_UNSUPPORTED, "conv (tch): every source needs a multiple of 8 channels, got %s%lld", "", p.src_c[s]);
    if (p.src_resample[s] != DTB200_RESAMPLE_NONE)
      return fail(DTB200_ERR_UNSUPPORTED,
                  "conv (tch): x2-resampled sources must be materialised first (ksize = 0 descriptor); ConvPlan does this%s");
    if (reinterpret_cast<uintptr_t>(p.src[s]) % 16 != 0) return fail(DTB200_ERR_INVALID, "conv (tch): source pointers must be 16-byte aligned%s");
  }
  if (p.out_c % 64 != 0) {  // heads: fp32 output, CUDA-core dot product over split16 sources
    if (!(p.ksize == 1 && p.stride == 1 && p.out_c < 64))
      return fail(DTB200_ERR_UNSUPPORTED, "conv (tch): out_c must be a multiple of 64 (or a <64-channel 1x1 head), got %s%lld", "", p.out_c);
    const long long pixels = (long long)p.batch * p.out_h * p.out_w;
    conv_head_h_kernel<<<(unsigned)((pixels + 7) / 8), 256, 0, stream>>>(p, pixels);
    return check_launch("conv_head_h_kernel");
  }
  if (reinterpret_cast<uintptr_t>(p.dst) % 16 != 0 || reinterpret_cast<uintptr_t>(p.residual) % 16 != 0 ||
      reinterpret_cast<uintptr_t>(p.workspace) % 32 != 0)
    return fail(DTB200_ERR_INVALID, "conv (tch): dst / residual must be 16-byte and workspace 32-byte aligned%s");
  const int num_sms = conv_tch_init();
  if (!g_encode_h) return fail(DTB200_ERR_CUDA, "conv (tch): cuTensorMapEncodeTiled entry point not available%s");
  const long long m_total = (long long)p.batch * p.out_h * p.out_w;
  const KLayoutH kl = make_klayout_h(p.num_src, p.src_c, p.ksize);
  const int bn = tch_bn(p.out_c);
  HWork wk;
  tch_tile_shape(p.out_h, p.out_w, wk.tw, wk.th);
  wk.tiles_x = (p.out_w + wk.tw - 1) / wk.tw;
  wk.tiles_y = (p.out_h + wk.th - 1) / wk.th;
  wk.m_tiles = p.batch * wk.tiles_x * wk.tiles_y;
  const int splits = tch_splits(wk.m_tiles, p.out_c, kl.num_kb);
  float* partial = nullptr;
  if (splits > 1) {
    const uint64_t need = (uint64_t)splits * m_total * p.out_c * sizeof(float);
    if (!p.workspace || p.workspace_bytes < need)
      return fail(DTB200_ERR_INVALID, "conv (tch): split-K needs %s%lld workspace bytes (dtb200_conv_workspace_bytes)", "", (long long)need);
    partial = reinterpret_cast<float*>(p.workspace);
  }
  wk.num_kb_total = kl.num_kb;
  wk.kb_per_split = (kl.num_kb + splits - 1) / splits;
  wk.splits = (kl.num_kb + wk.kb_per_split - 1) / wk.kb_per_split;
  wk.n_tiles = p.out_c / bn;
  wk.total = (long long)wk.m_tiles * wk.n_tiles * wk.splits;

  HMaps maps;
  // 3x3 / stride-1 layers with 64-channel N tiles and at least one full round of 8 x 16 tiles: halo-tile kernel
  if (p.ksize == 3 && p.stride == 1 && bn == 64 && splits == 1 && p.in_h == p.out_h && p.in_w == p.out_w) {
    HWork hw = wk;
    hw.tw = kHTW, hw.th = kHTH;
    hw.tiles_x = (p.out_w + kHTW - 1) / kHTW;
    hw.tiles_y = (p.out_h + kHTH - 1) / kHTH;
    hw.m_tiles = p.batch * hw.tiles_x * hw.tiles_y;
    hw.total = (long long)hw.m_tiles * hw.n_tiles;
    if (hw.total >= 148) {
      int rc = encode_split16_maps(p, maps, kHPW, kHPH, 1);
      if (rc != DTB200_OK) return rc;
      const unsigned grid = (unsigned)(hw.total < num_sms ? hw.total : num_sms);
      conv_tch_halo_kernel<<<grid, kCThreads, kHaloSmemBytes, stream>>>(p, maps, kl, hw);
      return check_launch("conv_tch_halo_kernel");
    }
  }
  int rc = encode_split16_maps(p, maps, (cuuint32_t)(wk.tw * p.stride), (cuuint32_t)(wk.th * p.stride), (cuuint32_t)p.stride);
  if (rc != DTB200_OK) return rc;
  const unsigned grid = (unsigned)(wk.total < num_sms ? wk.total : num_sms);
  if (bn == 128) conv_tch_kernel<128><<<grid, kCThreads, TchCfg<128>::kSmemBytes, stream>>>(p, maps, kl, m_total, wk, partial);
  else conv_tch_kernel<64><<<grid, kCThreads, TchCfg<64>::kSmemBytes, stream>>>(p, maps, kl, m_total, wk, partial);
  rc = check_launch("conv_tch_kernel");
  if (rc != DTB200_OK || !partial) return rc;
  const long long work = m_total * (p.out_c / 32);
  int blocks = (int)((work + 127) / 128);
  if (blocks > 148 * 8) blocks = 148 * 8;
  splitk_epilogue_h_kernel<<<blocks, 128, 0, stream>>>(p, partial, wk.splits, m_total);
  return check_launch("splitk_epilogue_h_kernel");
}

}  // namespace dtb200
