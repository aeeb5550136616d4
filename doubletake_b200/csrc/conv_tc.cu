// Fused channels-last convolution on 5th-gen tensor cores (tcgen05, TMEM accumulators), math = TC3X, sm_100a.
//
//   dst = act( conv_{k x k, stride}( concat_c[ resample_i(src_i) ] ) + bias (+ residual) )
//
// Implicit GEMM per CTA: D[128 pixels x BN channels] += A[128 x K] * B[BN x K]^T with K = taps x concatenated input
// channels, walked in blocks of 32 (4 groups of 8 channels; a group never straddles a source, so torch.cat is free).
//
// Precision: 3xTF32.  Every fp32 operand x is split into big = tf32-truncated x and small = x - big (both exact);
// per K step three kind::tf32 MMAs accumulate small*big + big*small + big*big in fp32 TMEM.  Per-product error ~2^-21:
// fp32-class, which the 1e-4 relative depth bar needs through ~20 stacked convs (plain TF32/BF16 does not meet it).
//
// Warp roles (288 threads):
//   warps 0-7  producers: im2col gather (zero padding, concat, x2 up-sampling on load) -> split -> st.shared into the
//              SWIZZLE_128B K-major A_big / A_small tiles; fence.proxy.async; one mbarrier arrival per warp.
//              Afterwards the same warps run the epilogue: tcgen05.ld -> bias/residual/activation -> NHWC store.
//   warp 8     lane 0: cp.async.bulk of the pre-swizzled, pre-split weight tile (UBLKCP, complete_tx on the same
//              mbarrier), then tcgen05.mma issue (12 per K block) and tcgen05.commit to release the stage.
// Weights are packed once (dtb200_pack_conv_weight) into the exact shared-memory image of each (N tile, K block).
#include "common.cuh"
#include "conv_common.cuh"
#include "tc_common.cuh"

namespace dtb200 {

using namespace tc;

constexpr int kBM = 128;                 // pixels per tile (UMMA M)
constexpr int kBK = 32;                  // fp32 per K block = one 128-byte swizzled row
constexpr int kLoadWarps = 4;            // cp.async im2col gather (never fence: they keep many loads in flight)
constexpr int kSplitWarps = 4;           // smem-only 3xTF32 split + proxy fence
constexpr int kProducerWarps = kLoadWarps + kSplitWarps;
constexpr int kEpilogueWarps = 4;
constexpr int kMmaWarp = kProducerWarps + kEpilogueWarps;      // 12
constexpr int kAuxWarp = kMmaWarp + 1;                         // 13: weight tiles + row info
constexpr int kThreads = (kAuxWarp + 1) * 32;                  // 448
constexpr int kATileBytes = kBM * 128;   // 16 KB (one of big / small)

template <int BN>
struct TcCfg {
  static constexpr int kBTileBytes = BN * 128;
  static constexpr int kBBytes = 2 * kBTileBytes;                    // B_big | B_small
  static constexpr int kStageBytes = 2 * kATileBytes + kBBytes;      // A_big | A_small | B_big | B_small
  static constexpr int kStages = BN <= 64 ? 4 : 3;
  static constexpr int kTmemCols = 2 * BN;                           // two accumulators
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 2048 /*row info*/ + 512 /*barriers*/;
};

__host__ __device__ inline int tc_num_kblocks(int in_c, int ksize) { return (in_c * ksize * ksize + kBK - 1) / kBK; }
__host__ __device__ inline int tc_bn(int out_c) { return (out_c % 128 == 0) ? 128 : 64; }

struct TcWork {  // persistent tile scheduler: item -> (m tile, n tile, K split)
  int m_tiles, n_tiles, splits, kb_per_split, num_kb_total;
  long long total;
};

// Persistent warp-specialised implicit-GEMM conv.  One CTA per SM loops over work items (static stride).
//   warps 0-7   A producers: im2col gather -> 3xTF32 split -> SWIZZLE_128B tiles (4-deep ring)
//   warps 8-11  epilogue: TMEM -> bias/residual/activation -> NHWC store (or split-K partials); double-buffered accumulator
//   warp 12     MMA issuer (one lane): 12 tcgen05.mma per K block, tcgen05.commit frees the A and B stages
//   warp 13     aux: weight tiles by cp.async.bulk into the same stage/barrier as A (one MMA-side wait per K block), and
//               the NEXT tile's row info (pixel index + zero-padding tap mask per row) so loaders start tiles without a prologue
template <int BN>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const dtb200_conv_params p, int in_c_total, long long m_total,
                                                              TcWork wk, float* __restrict__ partial) {
  using Cfg = TcCfg<BN>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = smem;                                                       // S x [A_big|A_small|B_big|B_small]
  int2* rowinfo = reinterpret_cast<int2*>(ring + S * Cfg::kStageBytes);       // [2][128] {centre pixel index, tap mask}
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(rowinfo) + 2048);
  uint64_t* raw_full = bars;            // [S] loaders -> splitters (cp.async completion, 128 arrivals)
  uint64_t* full = raw_full + S;        // [S] splitters (4) + weight copy (1 + tx) -> MMA: ONE wait per K block
  uint64_t* empty = full + S;           // [S] MMA (tcgen05.commit) -> loaders, weight loader
  uint64_t* acc_full = empty + S;       // [2]
  uint64_t* acc_empty = acc_full + 2;   // [2]
  uint64_t* ri_full = acc_empty + 2;    // [2] row info of the next tile written
  uint64_t* ri_empty = ri_full + 2;     // [2] loaders done with it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ri_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < S; ++s)
      mbar_init(&raw_full[s], kLoadWarps * 32), mbar_init(&full[s], kSplitWarps + 1), mbar_init(&empty[s], 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1), mbar_init(&acc_empty[s], kEpilogueWarps);
      mbar_init(&ri_full[s], 1), mbar_init(&ri_empty[s], kLoadWarps);
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int hw = p.out_h * p.out_w;

  auto decode = [&](long long item, int& m_tile, int& n_tile, int& kb_begin, int& num_kb, int& split) {
    split = (int)(item % wk.splits);
    long long r = item / wk.splits;
    n_tile = (int)(r % wk.n_tiles);
    m_tile = (int)(r / wk.n_tiles);
    kb_begin = split * wk.kb_per_split;
    num_kb = min(wk.kb_per_split, wk.num_kb_total - kb_begin);
  };

  if (warp < kLoadWarps) {
    // ============================================================ A loaders: raw fp32 im2col rows by cp.async
    SrcView sv[DTB200_CONV_MAX_SRC];
    int grp_end[DTB200_CONV_MAX_SRC];
    int acc_g = 0;
#pragma unroll
    for (int s = 0; s < DTB200_CONV_MAX_SRC; ++s) {
      sv[s].ptr = p.src[s];
      sv[s].c = s < p.num_src ? p.src_c[s] : 0;
      sv[s].resample = p.src_resample[s];
      const bool up = sv[s].resample != DTB200_RESAMPLE_NONE;
      sv[s].h = up ? p.in_h / 2 : p.in_h;
      sv[s].w = up ? p.in_w / 2 : p.in_w;
      acc_g += sv[s].c / 8;
      grp_end[s] = acc_g;
    }
    const int groups_per_tap = in_c_total / 8;
    const int taps = p.ksize * p.ksize;
    const int pad = p.ksize / 2;
    const int q = tid & 7;          // 16-byte chunk of the 128-byte row this thread fills
    const int prow = tid >> 3;      // rows prow + 16*it, it = 0..7
    uint32_t soff[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int row = prow + it * 16;
      soff[it] = (uint32_t)row * 128u + (uint32_t)((q ^ (row & 7)) << 4);
    }
    int stage = 0, phase = 0, use = 0;
    for (long long item = blockIdx.x; item < wk.total; item += gridDim.x, ++use) {
      int m_tile, n_tile, kb_begin, num_kb, split;
      decode(item, m_tile, n_tile, kb_begin, num_kb, split);
      const int rb = use & 1;
      mbar_wait(&ri_full[rb], (use >> 1) & 1);   // row info of this tile (written one tile ahead by the aux warp)
      int2 info[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) info[it] = rowinfo[rb * kBM + prow + it * 16];
      __syncwarp();
      if (lane == 0) mbar_arrive(&ri_empty[rb]);  // copied to registers: the buffer can be rewritten
      // group cursor of this thread's chunk: (tap, group-in-tap), advanced by 4 groups per K block
      int g_tap, g_r;
      {
        const int g = kb_begin * 4 + (q >> 1);
        g_tap = g / groups_per_tap;
        g_r = g - g_tap * groups_per_tap;
      }
#pragma unroll 1
      for (int kb = 0; kb < num_kb; ++kb) {
        const bool g_ok = g_tap < taps;
        const int tap = g_ok ? g_tap : 0;
        const int src_i = g_r < grp_end[0] ? 0 : (g_r < grp_end[1] ? 1 : 2);
        const int base = src_i == 0 ? 0 : (src_i == 1 ? grp_end[0] : grp_end[1]);
        const int c0 = (g_r - base) * 8 + (q & 1) * 4;
        const int ky = tap / p.ksize, kx = tap - ky * p.ksize;
        SrcView my = sv[0];
        if (src_i == 1) my = sv[1];
        if (src_i == 2) my = sv[2];
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* a_big = ring + stage * Cfg::kStageBytes;
        if (my.resample == DTB200_RESAMPLE_NONE) {
          const int dpix = (ky - pad) * p.in_w + (kx - pad);
          const float* bp = my.ptr + c0;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const bool ok = g_ok && ((info[it].y >> tap) & 1);
            const long long off = ok ? (long long)(info[it].x + dpix) * my.c : 0;
            cp_async16(a_big + soff[it], bp + off, ok ? 16u : 0u);
          }
        } else {
          // generic path: x2 up-sampling on load (the TC plans normally materialise up-sampled maps instead)
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g_ok && info[it].y != 0) {
              const long long m = (long long)m_tile * kBM + prow + it * 16;
              const int bb = (int)(m / hw);
              const int r = (int)(m - (long long)bb * hw);
              const int oy = r / p.out_w, ox = r - oy * p.out_w;
              v = load_input4(my, bb, oy * p.stride + ky - pad, ox * p.stride + kx - pad, p.in_h, p.in_w, c0);
            }
            *reinterpret_cast<float4*>(a_big + soff[it]) = v;
          }
          __threadfence_block();
        }
        cp_async_mbar_arrive(&raw_full[stage]);  // fires when this thread's copies have landed; the thread moves on
        if (++stage == S) stage = 0, phase ^= 1;
        g_r += 4;
        while (g_r >= groups_per_tap) g_r -= groups_per_tap, ++g_tap;
      }
    }
  } else if (warp < kProducerWarps) {
    // ============================================================ A splitters: small = x - tf32(x), shared memory only
    // The raw fp32 tile stays in place as the "big" operand: kind::tf32 ignores the low 13 mantissa bits of its
    // operands (verified on B200: every parity test passes bit-for-bit as with an explicit mask), so big = trunc(x)
    // needs no store.  Define DTB200_MASK_BIG to write the masked value anyway.
    const int st = tid - kLoadWarps * 32;
    const int q = st & 7, prow = st >> 3;
    uint32_t soff[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int row = prow + it * 16;
      soff[it] = (uint32_t)row * 128u + (uint32_t)((q ^ (row & 7)) << 4);
    }
    int stage = 0, phase = 0;
    for (long long item = blockIdx.x; item < wk.total; item += gridDim.x) {
      int m_tile, n_tile, kb_begin, num_kb, split;
      decode(item, m_tile, n_tile, kb_begin, num_kb, split);
#pragma unroll 1
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&raw_full[stage], phase);
        uint8_t* a_big = ring + stage * Cfg::kStageBytes;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const float4 v = *reinterpret_cast<const float4*>(a_big + soff[it]);
          float4 big = make_float4(tf32_big(v.x), tf32_big(v.y), tf32_big(v.z), tf32_big(v.w));
          float4 small = make_float4(v.x - big.x, v.y - big.y, v.z - big.z, v.w - big.w);
#ifdef DTB200_MASK_BIG
          *reinterpret_cast<float4*>(a_big + soff[it]) = big;
#endif
          *reinterpret_cast<float4*>(a_big + kATileBytes + soff[it]) = small;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[stage]);
        if (++stage == S) stage = 0, phase ^= 1;
      }
    }
  } else if (warp < kMmaWarp) {
    // ============================================================ epilogue warps
    const int ew = warp - kProducerWarps;  // == warp % 4 (kProducerWarps is a multiple of 4): TMEM lane quadrant
    const int row = ew * 32 + lane;
    int use = 0;
    for (long long item = blockIdx.x; item < wk.total; item += gridDim.x, ++use) {
      int m_tile, n_tile, kb_begin, num_kb, split;
      decode(item, m_tile, n_tile, kb_begin, num_kb, split);
      const int buf = use & 1;
      const long long m = (long long)m_tile * kBM + row;
      const bool live = m < m_total;
      const int n_base = n_tile * BN;
      float* dst = nullptr;
      const float* res = nullptr;
      if (partial) {
        dst = partial + ((long long)split * m_total + m) * p.out_c + n_base;
      } else {
        dst = p.dst + m * p.out_c + n_base;                 // NHWC: pixel index m is the row index
        res = p.residual ? p.residual + m * p.out_c + n_base : nullptr;
      }
      mbar_wait(&acc_full[buf], (use >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(buf * BN) + ((uint32_t)(ew * 32) << 16);
#pragma unroll 1
      for (int cc = 0; cc < BN; cc += 32) {
        float v[32];
        tmem_ld32(taddr + (uint32_t)cc, v);
        if (!live) continue;
        if (partial) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + cc + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const int n = cc + j;
            float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            if (p.bias) {
              float4 bb = ld4(p.bias + n_base + n);
              o.x += bb.x, o.y += bb.y, o.z += bb.z, o.w += bb.w;
            }
            if (res) {
              float4 rr = ld4(res + n);
              o.x += rr.x, o.y += rr.y, o.z += rr.z, o.w += rr.w;
            }
            o.x = activate(o.x, p.act, p.act_slope);
            o.y = activate(o.y, p.act, p.act_slope);
            o.z = activate(o.z, p.act, p.act_slope);
            o.w = activate(o.w, p.act, p.act_slope);
            *reinterpret_cast<float4*>(dst + n) = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  } else if (warp == kMmaWarp) {
    // ============================================================ MMA issuer (whole warp convergent; elect.sync issues)
    constexpr uint32_t idesc = umma_idesc_tf32(kBM, BN);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t ring_u = smem_u32(ring);
    int stage = 0, phase = 0, use = 0;
    for (long long item = blockIdx.x; item < wk.total; item += gridDim.x, ++use) {
      int m_tile, n_tile, kb_begin, num_kb, split;
      decode(item, m_tile, n_tile, kb_begin, num_kb, split);
      const int buf = use & 1;
      mbar_wait(&acc_empty[buf], ((use >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_u + (uint32_t)(buf * BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[stage], phase);                  // A (split) and B (weight tile) of this K block are in place
        tc_fence_after();
        const uint32_t a_big_u = ring_u + stage * Cfg::kStageBytes, a_small_u = a_big_u + kATileBytes;
        const uint32_t b_big_u = a_big_u + 2 * kATileBytes, b_small_u = b_big_u + Cfg::kBTileBytes;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t ko = ks * 32;  // 8 tf32 = 32 bytes along K inside the swizzled row
            const uint64_t da_b = umma_desc_k128(a_big_u + ko), da_s = umma_desc_k128(a_small_u + ko);
            const uint64_t db_b = umma_desc_k128(b_big_u + ko), db_s = umma_desc_k128(b_small_u + ko);
            umma_tf32(tmem_d, da_s, db_b, idesc, (kb | ks) != 0);
            umma_tf32(tmem_d, da_b, db_s, idesc, true);
            umma_tf32(tmem_d, da_b, db_b, idesc, true);
          }
          umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == S) stage = 0, phase ^= 1;
      }
      if (elect_one()) umma_commit(&acc_full[buf]);
      __syncwarp();
    }
  } else {
    // ============================================================ aux warp: next tile's row info + weight tiles
    const int pad = p.ksize / 2;
    auto write_rowinfo = [&](long long item, int use) {
      int m_tile, n_tile, kb_begin, num_kb, split;
      decode(item, m_tile, n_tile, kb_begin, num_kb, split);
      const int rb = use & 1;
      mbar_wait(&ri_empty[rb], ((use >> 1) & 1) ^ 1);
      // lane handles rows 4*lane .. 4*lane+3 (consecutive output pixels): one division, then carry
      long long m = (long long)m_tile * kBM + lane * 4;
      int bb = 0, oy = 0, ox = 0;
      if (m < m_total) {
        bb = (int)(m / hw);
        const int r = (int)(m - (long long)bb * hw);
        oy = r / p.out_w;
        ox = r - oy * p.out_w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j, ++m) {
        int2 info = make_int2(0, 0);
        if (m < m_total) {
          const int cy = oy * p.stride, cx = ox * p.stride;
          int mask = 0, t = 0;
          for (int ky = 0; ky < p.ksize; ++ky)
            for (int kx = 0; kx < p.ksize; ++kx, ++t) {
              const int iy = cy + ky - pad, ix = cx + kx - pad;
              if (iy >= 0 && iy < p.in_h && ix >= 0 && ix < p.in_w) mask |= 1 << t;
            }
          info = make_int2((bb * p.in_h + cy) * p.in_w + cx, mask);
          if (++ox == p.out_w) {
            ox = 0;
            if (++oy == p.out_h) oy = 0, ++bb;
          }
        }
        rowinfo[rb * kBM + lane * 4 + j] = info;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&ri_full[rb]);
    };
    int stage = 0, phase = 0, use = 0;
    if ((long long)blockIdx.x < wk.total) write_rowinfo(blockIdx.x, 0);
    for (long long item = blockIdx.x; item < wk.total; item += gridDim.x, ++use) {
      if (item + gridDim.x < wk.total) write_rowinfo(item + gridDim.x, use + 1);
      int m_tile, n_tile, kb_begin, num_kb, split;
      decode(item, m_tile, n_tile, kb_begin, num_kb, split);
      const uint8_t* wbase = reinterpret_cast<const uint8_t*>(p.weight) +
                             ((size_t)n_tile * wk.num_kb_total + kb_begin) * Cfg::kBBytes;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full[stage], Cfg::kBBytes);
          bulk_g2s(ring + stage * Cfg::kStageBytes + 2 * kATileBytes, wbase + (size_t)kb * Cfg::kBBytes, Cfg::kBBytes,
                   &full[stage]);
        }
        __syncwarp();
        if (++stage == S) stage = 0, phase ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// OIHW (out_c, in_c, k, k) -> per (N tile, K block): [B_big tile | B_small tile], each [BN rows][32 fp32] in the
// SWIZZLE_128B K-major shared-memory image; K flattened tap-major / channel-minor, zero padded to a multiple of 32.
__global__ void pack_weight_tc_kernel(const float* __restrict__ oihw, float* __restrict__ packed, int out_c, int in_c,
                                      int taps, int bn, int num_kb) {
  const long long tile_floats = (long long)bn * kBK;
  const long long total = (long long)(out_c / bn) * num_kb * tile_floats;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i / tile_floats;
    int e = (int)(i - t * tile_floats);
    int n_tile = (int)(t / num_kb), kb = (int)(t - (long long)n_tile * num_kb);
    int row = e / kBK, kk = e % kBK;
    int kflat = kb * kBK + kk;
    float x = 0.f;
    if (kflat < taps * in_c) {
      int tap = kflat / in_c, c = kflat - tap * in_c;
      x = oihw[((long long)(n_tile * bn + row) * in_c + c) * taps + tap];
    }
    float big = tf32_big(x);
    float* tile = packed + t * 2 * tile_floats;
    uint32_t off = sw128_offset(row, kk) / 4;
    tile[off] = big;
    tile[tile_floats + off] = x - big;
  }
}

int launch_conv_simt(const dtb200_conv_params& p, int in_c_total, cudaStream_t stream);

int conv_tc_debug_set(int) { return DTB200_OK; }  // development hook (knock-out switches live on the experiment branch)

uint64_t packed_floats_tc(int out_c, int in_c, int ksize) {
  if (out_c % 64 != 0) return (uint64_t)out_c * in_c * ksize * ksize;  // heads stay on the SIMT layout
  return (uint64_t)out_c * tc_num_kblocks(in_c, ksize) * kBK * 2;
}

int launch_pack_simt(const float* oihw, float* packed, int out_c, int in_c, int ksize, cudaStream_t s);

int launch_pack_tc(const float* oihw, float* packed, int out_c, int in_c, int ksize, cudaStream_t stream) {
  if (out_c % 64 != 0) return launch_pack_simt(oihw, packed, out_c, in_c, ksize, stream);
  int bn = tc_bn(out_c), num_kb = tc_num_kblocks(in_c, ksize);
  long long total = (long long)out_c * num_kb * kBK;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_weight_tc_kernel<<<blocks, 256, 0, stream>>>(oihw, packed, out_c, in_c, ksize * ksize, bn, num_kb);
  return check_launch("pack_weight_tc_kernel");
}

// split-K plan: enough CTAs to fill the machine when the (M, N) grid alone cannot, >= 4 K blocks per split
static int tc_splits(long long m_total, int out_c, int num_kb) {
  const int bn = tc_bn(out_c);
  const long long ctas = ((m_total + kBM - 1) / kBM) * (out_c / bn);
  if (ctas >= 148) return 1;
  int splits = (int)((2 * 148 + ctas - 1) / ctas);
  int max_splits = num_kb / 4;
  if (splits > max_splits) splits = max_splits;
  if (splits > 32) splits = 32;
  return splits < 1 ? 1 : splits;
}

uint64_t conv_tc_workspace_bytes(const dtb200_conv_params& p, int in_c_total) {
  if (p.ksize == 0 || p.out_c % 64 != 0) return 0;
  const long long m_total = (long long)p.batch * p.out_h * p.out_w;
  const int splits = tc_splits(m_total, p.out_c, tc_num_kblocks(in_c_total, p.ksize));
  return splits > 1 ? (uint64_t)splits * m_total * p.out_c * sizeof(float) : 0;
}

// Deterministic split-K reduction (splits summed in order) fused with bias / residual / activation.
__global__ void splitk_epilogue_kernel(const float* __restrict__ partial, int splits, long long mn, int out_c,
                                       const float* __restrict__ bias, const float* __restrict__ residual, int act,
                                       float slope, float* __restrict__ dst) {
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4; i < mn; i += (long long)gridDim.x * blockDim.x * 4) {
    float4 a = ld4(partial + i);
    for (int s = 1; s < splits; ++s) {
      float4 b = ld4(partial + (long long)s * mn + i);
      a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
    }
    const int n = (int)(i % out_c);
    if (bias) {
      float4 b = ld4(bias + n);
      a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
    }
    if (residual) {
      float4 r = ld4(residual + i);
      a.x += r.x, a.y += r.y, a.z += r.z, a.w += r.w;
    }
    a.x = activate(a.x, act, slope), a.y = activate(a.y, act, slope);
    a.z = activate(a.z, act, slope), a.w = activate(a.w, act, slope);
    *reinterpret_cast<float4*>(dst + i) = a;
  }
}

// ksize == 0 descriptor: dst = resample(src[0]) (x2 bilinear / nearest), written once so the tensor-core convs that
// consume it (3x3 conv1 and the 1x1 skip projection) read plain channels-last data instead of interpolating 10 times.
__global__ void resample_copy_kernel(const dtb200_conv_params p) {
  SrcView sv;
  sv.ptr = p.src[0];
  sv.c = p.src_c[0];
  sv.resample = p.src_resample[0];
  sv.h = p.in_h / 2;
  sv.w = p.in_w / 2;
  const int c4 = sv.c / 4;
  const long long total = (long long)p.batch * p.in_h * p.in_w * c4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % c4) * 4;
    long long pix = i / c4;
    int x = (int)(pix % p.in_w);
    long long r = pix / p.in_w;
    int y = (int)(r % p.in_h), b = (int)(r / p.in_h);
    float4 v = load_input4(sv, b, y, x, p.in_h, p.in_w, c);
    *reinterpret_cast<float4*>(p.dst + pix * sv.c + c) = v;
  }
}

int launch_resample_copy(const dtb200_conv_params& p, cudaStream_t stream) {
  if (p.num_src != 1 || p.src_resample[0] == DTB200_RESAMPLE_NONE || p.src_c[0] % 4 != 0 || p.out_c != p.src_c[0] ||
      p.out_h != p.in_h || p.out_w != p.in_w || (p.in_h & 1) || (p.in_w & 1) || !p.src[0] || !p.dst)
    return fail(DTB200_ERR_INVALID, "resample copy (ksize=0): needs one x2-resampled source and a matching dst%s");
  long long total = (long long)p.batch * p.in_h * p.in_w * (p.src_c[0] / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  resample_copy_kernel<<<blocks, 256, 0, stream>>>(p);
  return check_launch("resample_copy_kernel");
}

int launch_conv_tc(const dtb200_conv_params& p, int in_c_total, cudaStream_t stream) {
  if (p.out_c % 64 != 0) return launch_conv_simt(p, in_c_total, stream);  // 1-channel heads: CUDA-core dot product
  for (int s = 0; s < p.num_src; ++s)
    if (p.src_c[s] % 8 != 0)
      return fail(DTB200_ERR_UNSUPPORTED, "conv (tc3x): every source needs a multiple of 8 channels, got %s%lld", "", p.src_c[s]);
  const long long m_total = (long long)p.batch * p.out_h * p.out_w;
  if (m_total * (long long)in_c_total >= (1LL << 31) || (long long)p.batch * p.in_h * p.in_w >= (1LL << 31) / 1024)
    return fail(DTB200_ERR_UNSUPPORTED, "conv (tc3x): tensor too large for 32-bit pixel indexing%s");
  const int num_kb = tc_num_kblocks(in_c_total, p.ksize);
  const int bn = tc_bn(p.out_c);
  const int splits = tc_splits(m_total, p.out_c, num_kb);
  float* partial = nullptr;
  if (splits > 1) {
    uint64_t need = (uint64_t)splits * m_total * p.out_c * sizeof(float);
    if (!p.workspace || p.workspace_bytes < need)
      return fail(DTB200_ERR_INVALID, "conv (tc3x): split-K needs %s%lld workspace bytes (dtb200_conv_workspace_bytes)", "",
                  (long long)need);
    partial = reinterpret_cast<float*>(p.workspace);
  }
  TcWork wk;
  wk.num_kb_total = num_kb;
  wk.kb_per_split = (num_kb + splits - 1) / splits;
  wk.splits = (num_kb + wk.kb_per_split - 1) / wk.kb_per_split;
  wk.m_tiles = (int)((m_total + kBM - 1) / kBM);
  wk.n_tiles = p.out_c / bn;
  wk.total = (long long)wk.m_tiles * wk.n_tiles * wk.splits;
  const int zsplits = wk.splits;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  const unsigned grid = (unsigned)(wk.total < num_sms ? wk.total : num_sms);
  cudaError_t e;
  if (bn == 128) {
    e = cudaFuncSetAttribute(conv_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<128>::kSmemBytes);
    if (e != cudaSuccess) return fail(DTB200_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    conv_tc_kernel<128><<<grid, kThreads, TcCfg<128>::kSmemBytes, stream>>>(p, in_c_total, m_total, wk, partial);
  } else {
    e = cudaFuncSetAttribute(conv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<64>::kSmemBytes);
    if (e != cudaSuccess) return fail(DTB200_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    conv_tc_kernel<64><<<grid, kThreads, TcCfg<64>::kSmemBytes, stream>>>(p, in_c_total, m_total, wk, partial);
  }
  int rc = check_launch("conv_tc_kernel");
  if (rc != DTB200_OK || !partial) return rc;
  const long long mn = m_total * p.out_c;
  int blocks = (int)((mn / 4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  splitk_epilogue_kernel<<<blocks, 256, 0, stream>>>(partial, zsplits, mn, p.out_c, p.bias, p.residual, p.act, p.act_slope,
                                                     p.dst);
  return check_launch("splitk_epilogue_kernel");
}

}  // namespace dtb200
