// Fused channels-last convolution on 5th-gen tensor cores (tcgen05, TMEM accumulators), math = TC3X, sm_100a.
//
//   dst = act( conv_{k x k, stride}( concat_c[ src_i ] ) + bias (+ residual) )
//
// Implicit GEMM per tile: D[128 pixels x BN channels] += A[128 x K] * B[BN x K]^T, K = taps x 32-channel chunks of the
// concatenated sources (torch.cat is free: every K block comes from one source).  The 128 pixels are a TW x TH patch of the
// output map; the A tile of a K block is ONE TMA box (cp.async.bulk.tensor.4d over the NHWC tensor) shifted by the tap
// offset -- hardware address generation, SWIZZLE_128B placement and zero fill (= the conv's zero padding).
//
// Precision: 3xTF32.  x = big + small with big = tf32-truncated x (the raw fp32 tile itself: kind::tf32 ignores the low
// 13 mantissa bits) and small = x - big; per K step small*big + big*small + big*big accumulate in fp32 TMEM.
// Weights are packed once (dtb200_pack_conv_weight_srcs) into the exact shared-memory image of each (N tile, K block).
#include <cuda.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "conv_common.cuh"
#include "tc_common.cuh"

namespace dtb200 {

using namespace tc;

constexpr int kBM = 128;                 // pixels per tile (UMMA M) = TW x TH output pixels
constexpr int kBK = 32;                  // fp32 per K block = one 128-byte swizzled row = one TMA box row
constexpr int kEpilogueWarps = 4;        // warps 0-3: TMEM lane quadrant == warp index
constexpr int kSplitWarps = 4;           // warps 4-7
constexpr int kMmaWarp = kEpilogueWarps + kSplitWarps;   // 8
constexpr int kLoadWarp = kMmaWarp + 1;                  // 9: TMA boxes (A); 10: bulk copies (weight tiles)
constexpr int kThreads = (kLoadWarp + 2) * 32;           // 352
constexpr int kATileBytes = kBM * 128;   // 16 KB (one of big / small)

// kMerged (BN = 64): the weight tile image [B_big | B_small] is 128 contiguous K-major rows, so ONE N=128 MMA computes
// A_big x [B_big | B_small]^T into two 64-column halves of a 128-column accumulator and a second N=64 MMA adds
// A_small x B_big^T to the first half; the epilogue sums the halves.  2 MMAs (64 + 48 clk, 14 KB of operand reads) per
// K step instead of 3 (3 x 48 clk, 18 KB): the N=64 tiles are bound by shared-memory operand bandwidth.
template <int BN, bool kMerged, int kS = (BN <= 64 ? 4 : 3)>
struct TcCfg {
  static constexpr int kBTileBytes = BN * 128;
  static constexpr int kBBytes = 2 * kBTileBytes;                    // B_big | B_small
  static constexpr int kStageBytes = 2 * kATileBytes + kBBytes;      // A_big(raw) | A_small | B_big | B_small
  static constexpr int kStages = kS;
  static constexpr int kAccCols = kMerged ? 2 * BN : BN;             // TMEM columns of one accumulator
  static constexpr int kTmemCols = 2 * kAccCols;                     // two accumulators
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 512 /*barriers*/;
};

// development switches (dtb200_debug_set, or DTB200_CONV_FLAGS in the environment at first use):
//   bit 0: issue the 64-channel tiles as three N=64 MMAs per K step instead of the merged N=128 + N=64 pair
//   bit 1: same for the 128-channel tiles (their merged form, N=256 + N=128, takes all 512 TMEM columns)
//   bit 2: previous split-K rule (many short items) instead of the round-count cost model
//   bit 7: no halo-tile kernel (3x3 / stride-1 layers use the tap-major kernel too); bit 3: halo kernel with 1 tap per
//          weight-ring stage instead of 3
//   bit 4: (tch) resident-weights halo kernel for 3x3 / stride-1 layers with <= 64 input and 64 output channels (conv_tch.cu)
//   bits 8..13: timing knock-outs of the tensor-core conv pipeline (results are WRONG; tools/conv_bench.py --debug)
static int g_flags = -1;
// the timing knock-outs (bits 8 and up) make the kernels skip work, i.e. produce WRONG results: they are honoured only
// when DTB200_DEVELOPMENT=1 is set in the environment (tools/conv_bench.py sets it), never by a stray debug_set call
static int sanitize_flags(int flags) {
  static int dev = -1;
  if (dev < 0) {
    const char* e = getenv("DTB200_DEVELOPMENT");
    dev = (e && atoi(e) != 0) ? 1 : 0;
  }
  if (flags < 0) flags = 0;
  return dev ? flags : (flags & 0xff);
}
static int conv_flags() {
  if (g_flags < 0) {
    const char* e = getenv("DTB200_CONV_FLAGS");
    g_flags = sanitize_flags(e ? atoi(e) : 0);
  }
  return g_flags;
}
int conv_debug_flags() { return conv_flags(); }   // read by the tch kernels' knock-outs too
int conv_tc_debug_set(int flags) {
  g_flags = sanitize_flags(flags);
  return DTB200_OK;
}

__host__ __device__ inline int tc_bn(int out_c) { return (out_c % 128 == 0) ? 128 : 64; }

// K layout shared by the weight packer and the kernel: tap-major; inside a tap the sources in order, each cut into
// 32-channel chunks (the last chunk of a source may overhang: TMA zero-fills, the packer writes zero weights).
struct KLayout {
  int num_src, taps, kb_per_tap, num_kb;
  int src_c[DTB200_CONV_MAX_SRC], chunk_end[DTB200_CONV_MAX_SRC], c_begin[DTB200_CONV_MAX_SRC];
};
__host__ __device__ inline KLayout make_klayout(int num_src, const int32_t* src_c, int ksize) {
  KLayout k;
  k.num_src = num_src;
  k.taps = ksize * ksize;
  int chunks = 0, cb = 0;
  for (int s = 0; s < DTB200_CONV_MAX_SRC; ++s) {
    k.src_c[s] = s < num_src ? src_c[s] : 0;
    k.c_begin[s] = cb;
    cb += k.src_c[s];
    chunks += (k.src_c[s] + kBK - 1) / kBK;
    k.chunk_end[s] = chunks;
  }
  k.kb_per_tap = chunks;
  k.num_kb = chunks * k.taps;
  return k;
}
__host__ __device__ inline void klayout_decode(const KLayout& k, int kbi, int& tap, int& src, int& c0) {
  tap = kbi / k.kb_per_tap;
  const int r = kbi - tap * k.kb_per_tap;
  src = r < k.chunk_end[0] ? 0 : (r < k.chunk_end[1] ? 1 : 2);
  const int base = src == 0 ? 0 : (src == 1 ? k.chunk_end[0] : k.chunk_end[1]);
  c0 = (r - base) * kBK;
}

struct TcWork {  // persistent tile scheduler: item -> (m tile = (batch, tile row, tile col), n tile, K split)
  int tw, th, tiles_x, tiles_y;     // pixel tile shape (tw * th == 128) and tile grid per image
  int m_tiles, n_tiles, splits, kb_per_split, num_kb_total;
  long long total;
};

struct TcMaps {
  CUtensorMap m[DTB200_CONV_MAX_SRC];
};

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tm, int c, int x, int y, int b, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tm), "r"(c), "r"(x), "r"(y), "r"(b), "r"(smem_u32(bar))
      : "memory");
}

// dbg bit 5 (flag value 32 << 8): per-role cycle accounting, printed by CTA 0 (development aid, see tools/conv_bench.py)
struct RoleProf {
  long long t0 = 0, waited = 0;
  int n = 0;
};
__device__ __forceinline__ void prof_wait(uint64_t* bar, uint32_t parity, bool on, RoleProf& pr, uint32_t sleep_ns = 0) {
  if (!on) {
    mbar_wait(bar, parity, 0, sleep_ns);
    return;
  }
  const long long a = clock64();
  mbar_wait(bar, parity);
  pr.waited += clock64() - a;
  ++pr.n;
}
__device__ __forceinline__ void prof_report(const char* role, bool on, const RoleProf& pr) {
  if (on && blockIdx.x == 0 && (threadIdx.x & 31) == 0)
    printf("role %-8s warp %2d: %8lld clk in loop, %8lld clk waiting, %5d waits\n", role, (int)(threadIdx.x >> 5),
           clock64() - pr.t0, pr.waited, pr.n);
}

// Persistent warp-specialised implicit-GEMM conv, TMA-fed.  One CTA per SM loops over work items (static stride).
//   warps 0-3  epilogue: TMEM -> bias/residual/activation -> NHWC store (or split-K partials); double-buffered accumulator
//   warps 4-7  splitters: the raw fp32 tile stays in place as the tf32 "big" operand (the tensor core ignores the low 13
//              mantissa bits); they add small = x - tf32(x) next to it, fence.proxy.async, arrive
//   warp 8     MMA issuer (convergent warp, elect.sync): 12 tcgen05.mma per K block, tcgen05.commit frees the stage
//   warp 9     loader: per K block ONE cp.async.bulk.tensor (TMA) box = 32 channels x (TW x TH) pixels of one source at one
//              tap, written by hardware in the SWIZZLE_128B layout with zero fill outside the image (= conv padding, also
//              the channel overhang of sources that are not multiples of 32), plus one cp.async.bulk of the weight tile
template <int BN, bool kMerged, int kS = (BN <= 64 ? 4 : 3)>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const dtb200_conv_params p, const __grid_constant__ TcMaps maps,
                                                              KLayout kl, long long m_total, TcWork wk,
                                                              float* __restrict__ partial, int dbg) {
  using Cfg = TcCfg<BN, kMerged, kS>;
  constexpr int S = Cfg::kStages;
  static_assert(!kMerged || Cfg::kTmemCols <= 512, "merged accumulators exceed TMEM");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = smem;                                        // S x [A_big|A_small|B_big|B_small]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + S * Cfg::kStageBytes);
  uint64_t* raw_full = bars;            // [S] TMA box landed (1 arrival + tx)            loader   -> splitters
  uint64_t* full = raw_full + S;        // [S] small written (4) + weight tile (1 + tx)     -> MMA (one wait per K block)
  uint64_t* empty = full + S;           // [S] tcgen05.commit                               MMA      -> loader
  uint64_t* acc_full = empty + S;       // [2]
  uint64_t* acc_empty = acc_full + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  long long* stamp = reinterpret_cast<long long*>(tmem_slot + 2);  // [2][S] profiling time stamps (dbg bit 5 only)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&raw_full[s], 1), mbar_init(&full[s], ((dbg & 64) ? kSplitWarps : kSplitWarps / 2) + 1), mbar_init(&empty[s], 1);
    for (int s = 0; s < 2; ++s) mbar_init(&acc_full[s], 1), mbar_init(&acc_empty[s], kEpilogueWarps);
    for (int s = 0; s < 2 * S; ++s) stamp[s] = 0;
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool prof = (dbg & 32) != 0;
  RoleProf pr, pr2, pr3;
  if (prof) pr.t0 = pr2.t0 = clock64();

  auto decode = [&](long long item, int& bb, int& y0, int& x0, int& n_tile, int& kb_begin, int& num_kb, int& split) {
    split = (int)(item % wk.splits);
    long long r = item / wk.splits;
    n_tile = (int)(r % wk.n_tiles);
    int m_tile = (int)(r / wk.n_tiles);
    const int per_img = wk.tiles_x * wk.tiles_y;
    bb = m_tile / per_img;
    const int t = m_tile - bb * per_img;
    y0 = (t / wk.tiles_x) * wk.th;
    x0 = (t % wk.tiles_x) * wk.tw;
    kb_begin = split * wk.kb_per_split;
    num_kb = min(wk.kb_per_split, wk.num_kb_total - kb_begin);
  };

  if (warp < kEpilogueWarps) {
    // ============================================================ epilogue warps
    const int row = warp * 32 + lane;          // TMEM lane == tile row == (ty, tx) = (row / tw, row % tw)
    const int ty = row / wk.tw, tx = row - ty * wk.tw;
    int use = 0;
    for (long long item = blockIdx.x; item < wk.total; item += gridDim.x, ++use) {
      int bb, y0, x0, n_tile, kb_begin, num_kb, split;
      decode(item, bb, y0, x0, n_tile, kb_begin, num_kb, split);
      const int buf = use & 1;
      const int oy = y0 + ty, ox = x0 + tx;
      const bool live = oy < p.out_h && ox < p.out_w;
      const long long m = ((long long)bb * p.out_h + oy) * p.out_w + ox;   // NHWC pixel index
      const int n_base = n_tile * BN;
      float* dst = nullptr;
      const float* res = nullptr;
      if (partial) {
        dst = partial + ((long long)split * m_total + m) * p.out_c + n_base;
      } else {
        dst = p.dst + m * p.out_c + n_base;
        res = p.residual ? p.residual + m * p.out_c + n_base : nullptr;
      }
      prof_wait(&acc_full[buf], (use >> 1) & 1, prof, pr, 128);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(buf * Cfg::kAccCols) + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
      for (int cc = 0; cc < BN; cc += 32) {
        float r[32];
        const bool use_res = res != nullptr && live && !(dbg & 16);
        if (use_res) {  // residual row segment first: its latency overlaps the TMEM loads
#pragma unroll
          for (int j = 0; j < 32; j += 8) ldg256(res + cc + j, r + j);
        }
        float v[32];
        tmem_ld32(taddr + (uint32_t)cc, v);
        if (kMerged) {  // + A_big x B_small^T, accumulated in the upper half
          float v2[32];
          tmem_ld32(taddr + (uint32_t)(BN + cc), v2);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += v2[j];
        }
        if (!live || (dbg & 16)) continue;
        if (!partial) {
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bv = ld4(p.bias + n_base + cc + j);
              v[j] += bv.x, v[j + 1] += bv.y, v[j + 2] += bv.z, v[j + 3] += bv.w;
            }
          }
          if (use_res) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += r[j];
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = activate(v[j], p.act, p.act_slope);
        }
#pragma unroll
        for (int j = 0; j < 32; j += 8) stg256(dst + cc + j, v + j);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  } else if (warp < kMmaWarp) {
    // ============================================================ splitters: small = x - tf32(x), shared memory only
    // Two groups of two warps take alternate K blocks: the per-K-block cost of a group (barrier wake-up, shared-memory
    // round trip, proxy fence) is latency, not work, so two blocks are in flight at once.
    // A stage is always split by the same warps (a parity wait must observe every phase of its barrier).
    const bool split_all = (dbg & 64) != 0;  // experiment: all four warps on every K block instead of two groups
    const int group = split_all ? 0 : (warp - kEpilogueWarps) >> 1;
    const int st = tid - (kEpilogueWarps + 2 * group) * 32;  // 0..63 inside a group (0..127 with split_all)
    const int q = st & 7, prow = st >> 3;
    const uint32_t ring_u = smem_u32(ring);
    auto soff = [&](int row) { return (uint32_t)row * 128u + (uint32_t)((q ^ (row & 7)) << 4); };
    auto split4 = [&](uint32_t a_big_u, const float4& x, int row) {
      const float4 big = make_float4(tf32_big(x.x), tf32_big(x.y), tf32_big(x.z), tf32_big(x.w));
      sts128(a_big_u + kATileBytes + soff(row), make_float4(x.x - big.x, x.y - big.y, x.z - big.z, x.w - big.w));
    };
    int stage = 0, phase = 0;
    for (long long item = blockIdx.x; item < wk.total; item += gridDim.x) {
      int bb, y0, x0, n_tile, kb_begin, num_kb, split;
      decode(item, bb, y0, x0, n_tile, kb_begin, num_kb, split);
#pragma unroll 1
      for (int kb = 0; kb < num_kb; ++kb) {
        if (split_all || (stage & 1) == group) {
          prof_wait(&raw_full[stage], phase, prof, pr, 32);
          if (prof && st == 0) pr2.waited += clock64() - stamp[stage], ++pr2.n;  // TMA issue -> box landed and seen
          if (!(dbg & 2)) {
            // all loads first (explicit ld.shared: the generic-pointer form serialises load -> store -> load on possible
            // aliasing and pays the generic-address translation), then split and store
            const uint32_t a_big_u = ring_u + (uint32_t)(stage * Cfg::kStageBytes);
            if (split_all) {
              float4 v[8];
#pragma unroll
              for (int it = 0; it < 8; ++it) v[it] = lds128(a_big_u + soff(prow + it * 16));
#pragma unroll
              for (int it = 0; it < 8; ++it) split4(a_big_u, v[it], prow + it * 16);
            } else {
              float4 v[16];
#pragma unroll
              for (int it = 0; it < 16; ++it) v[it] = lds128(a_big_u + soff(prow + it * 8));
#pragma unroll
              for (int it = 0; it < 16; ++it) split4(a_big_u, v[it], prow + it * 8);
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&full[stage]);
        }
        if (++stage == S) stage = 0, phase ^= 1;
      }
    }
    if (prof && st == 0 && blockIdx.x == 0)
      printf("latency  TMA issue -> splitter sees the box: %lld clk avg over %d\n", pr2.waited / max(pr2.n, 1), pr2.n);
  } else if (warp == kMmaWarp) {
    // ============================================================ MMA issuer (whole warp convergent; elect.sync issues)
    constexpr uint32_t idesc = umma_idesc_tf32(kBM, BN);
    constexpr uint32_t idesc2 = umma_idesc_tf32(kBM, 2 * BN);  // merged: N covers [B_big | B_small]
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    // shared-memory descriptors (tc_common.cuh: umma_desc_k128): the high word is constant, the low word is
    // (address >> 4) | LBO; per K block only the low word moves (stage base + tile + 32-byte K step), by plain adds
    constexpr uint32_t kDescHi = 64u | (1u << 14) | (2u << 29);
    const uint32_t lo_ring = ((smem_u32(ring) & 0x3FFFFu) >> 4) | (1u << 16);
    auto desc = [](uint32_t lo) { return ((uint64_t)kDescHi << 32) | lo; };
    int stage = 0, phase = 0, use = 0;
    for (long long item = blockIdx.x; item < wk.total; item += gridDim.x, ++use) {
      int bb, y0, x0, n_tile, kb_begin, num_kb, split;
      decode(item, bb, y0, x0, n_tile, kb_begin, num_kb, split);
      const int buf = use & 1;
      prof_wait(&acc_empty[buf], ((use >> 1) & 1) ^ 1, prof, pr2);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_u + (uint32_t)(buf * Cfg::kAccCols);
      for (int kb = 0; kb < num_kb; ++kb) {
        prof_wait(&full[stage], phase, prof, pr);        // A (raw + small) and B (weight tile) of this K block are in place
        tc_fence_after();
        if (prof && lane == 0) pr3.waited += clock64() - stamp[stage], ++pr3.n;  // TMA issue -> MMA warp sees `full`
        const uint32_t lo_a_big = lo_ring + (uint32_t)stage * (Cfg::kStageBytes >> 4);
        const uint32_t lo_a_small = lo_a_big + (kATileBytes >> 4), lo_b_big = lo_a_big + (2 * kATileBytes >> 4);
        const uint32_t lo_b_small = lo_b_big + (Cfg::kBTileBytes >> 4);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (dbg & 1) break;
            const uint32_t ko = ks * 2;  // 8 tf32 = 32 bytes along K inside the swizzled row, in 16-byte units
            const uint64_t da_b = desc(lo_a_big + ko), da_s = desc(lo_a_small + ko);
            const uint64_t db_b = desc(lo_b_big + ko), db_s = desc(lo_b_small + ko);
            if (kMerged) {
              umma_tf32(tmem_d, da_b, db_b, idesc2, (kb | ks) != 0);  // [big x big | big x small]
              umma_tf32(tmem_d, da_s, db_b, idesc, true);             // small x big into the first half
            } else {
              umma_tf32(tmem_d, da_s, db_b, idesc, (kb | ks) != 0);
              umma_tf32(tmem_d, da_b, db_s, idesc, true);
              umma_tf32(tmem_d, da_b, db_b, idesc, true);
            }
          }
          umma_commit(&empty[stage]);
        }
        if (prof && lane == 0) stamp[S + stage] = clock64();
        __syncwarp();
        if (++stage == S) stage = 0, phase ^= 1;
      }
      if (elect_one()) umma_commit(&acc_full[buf]);
      __syncwarp();
    }
  } else {
    // ============================================================ loader: TMA boxes (A) + weight tiles (B)
    const int pad = p.ksize / 2;
    int stage = 0, phase = 0;
    for (long long item = blockIdx.x; item < wk.total; item += gridDim.x) {
      int bb, y0, x0, n_tile, kb_begin, num_kb, split;
      decode(item, bb, y0, x0, n_tile, kb_begin, num_kb, split);
      const uint8_t* wbase = reinterpret_cast<const uint8_t*>(p.weight) +
                             ((size_t)n_tile * wk.num_kb_total + kb_begin) * Cfg::kBBytes;
      // (tap, source, channel chunk) of the item's first K block by division, then advanced incrementally
      int tap, src, c0;
      klayout_decode(kl, kb_begin, tap, src, c0);
      int ky = tap / p.ksize, kx = tap - ky * p.ksize;
      for (int kb = 0; kb < num_kb; ++kb) {
        prof_wait(&empty[stage], phase ^ 1, prof, pr, 32);
        if (prof && warp == kLoadWarp && lane == 0) {
          const long long now = clock64();
          if (stamp[S + stage] != 0) pr2.waited += now - stamp[S + stage], ++pr2.n;  // commit issued -> loader sees `empty`
          stamp[stage] = now;
        }
        if (elect_one()) {
          uint8_t* a_big = ring + stage * Cfg::kStageBytes;
          if (warp == kLoadWarp) {
            if (dbg & 4) {
              mbar_arrive(&raw_full[stage]);
            } else {
              mbar_arrive_expect_tx(&raw_full[stage], kATileBytes);
              tma_load_4d(a_big, &maps.m[src], c0, x0 * p.stride + kx - pad, y0 * p.stride + ky - pad, bb, &raw_full[stage]);
            }
          } else {
            if (dbg & 8) {
              mbar_arrive(&full[stage]);
            } else {
              mbar_arrive_expect_tx(&full[stage], Cfg::kBBytes);
              bulk_g2s(a_big + 2 * kATileBytes, wbase + (size_t)kb * Cfg::kBBytes, Cfg::kBBytes, &full[stage]);
            }
          }
        }
        __syncwarp();
        if (++stage == S) stage = 0, phase ^= 1;
        c0 += kBK;
        if (c0 >= (src == 0 ? kl.src_c[0] : (src == 1 ? kl.src_c[1] : kl.src_c[2]))) {
          c0 = 0;
          if (++src == kl.num_src) {
            src = 0;
            if (++kx == p.ksize) kx = 0, ++ky;
          }
        }
      }
    }
  }
  if (prof && blockIdx.x == 0 && lane == 0) {
    if (warp == kMmaWarp) printf("latency  TMA issue -> MMA warp sees full:    %lld clk avg over %d\n", pr3.waited / max(pr3.n, 1), pr3.n);
    if (warp == kLoadWarp) printf("latency  MMA commit issued -> loader sees empty: %lld clk avg over %d\n", pr2.waited / max(pr2.n, 1), pr2.n);
  }
  if (prof) {
    prof_report(warp < kEpilogueWarps ? "epilogue" : warp < kMmaWarp ? "split" : warp == kMmaWarp ? "mma" : "load", true, pr);
    if (warp == kMmaWarp) prof_report("mma/acc", true, pr2);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ======================================================================================================================
// Halo-tile variant for 3x3 / stride-1 layers (the bulk of the decoder's time at 240x320 and 120x160).
//
// The tap-major kernel above re-reads every activation 9 times from L2 (one TMA box + one split per tap) and its 48 KB
// stages leave room for only 4 K blocks in flight.  Here the A side of a stage is ONE patch of
// (TH+2) x (TW+2) = 18 x 10 pixels x 32 channels (22.5 KB raw + 22.5 KB small), loaded by one TMA box and split once; the 9
// taps read it through shared-memory descriptors that are only SHIFTED by whole pixels: tile row r = (ty, tx) of tap
// (ky, kx) is patch pixel (ty + ky) * 10 + tx + kx, i.e. start address + (ky * 10 + kx) * 128 B with the 8-row core-matrix
// stride (SBO) = one patch row = 1280 B.  tcgen05 applies SWIZZLE_128B on absolute shared-memory address bits, so a
// shifted start needs no base_offset (measured: tools/halo_probe.cu, profiles/r01d_halo_probe.txt).  Weight tiles (one per
// tap and chunk, the packed layout is unchanged) stream through their own ring.  Per 32-channel chunk: 1 TMA box + 1 split
// + 9 weight tiles + 72 MMAs, instead of 9 boxes + 9 splits.
constexpr int kHaloTW = 8, kHaloTH = 16;                       // 128 output pixels, 8 wide: one core-matrix group per tile row
constexpr int kHaloPW = kHaloTW + 2, kHaloPH = kHaloTH + 2;    // patch
constexpr int kHaloPatchBytes = kHaloPW * kHaloPH * 128;       // 23040 = the TMA box
constexpr int kHaloSlotBytes = (kHaloPatchBytes + 1023) / 1024 * 1024;  // 23552: slots stay 1024-byte aligned
constexpr int kHaloAStageBytes = 2 * kHaloSlotBytes;           // raw | small
constexpr int kHaloAStages = 2;
// kTB = taps per weight-ring stage: the issuing thread pays one mbarrier wait (~190 clk, never hidden: MMA issue is
// synchronous with the tensor pipe, tools/umma_bench.cu) and one commit per stage, so kTB = 3 amortises them over 24 MMAs.
template <int BN, int kTB = 1>
struct HaloCfg {
  static constexpr int kBBytes = 2 * BN * 128;                 // B_big | B_small of one (tap, chunk)
  static constexpr int kBStageBytes = kTB * kBBytes;
  static constexpr int kBStages = kTB == 1 ? (BN <= 64 ? 6 : 4) : 2;
  static constexpr int kAccCols = 2 * BN;                      // merged accumulator (see kMerged above)
  static constexpr int kTmemCols = 2 * kAccCols;
  static constexpr int kSmemBytes = kHaloAStages * kHaloAStageBytes + kBStages * kBStageBytes + 1024 /*align*/ + 512 /*barriers*/;
};

template <int BN, int kTB>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_halo_kernel(const dtb200_conv_params p, const __grid_constant__ TcMaps maps,
                                                                   KLayout kl, TcWork wk) {
  using Cfg = HaloCfg<BN, kTB>;
  constexpr int SA = kHaloAStages, SB = Cfg::kBStages;
  static_assert(Cfg::kTmemCols <= 512, "accumulators exceed TMEM");
  static_assert(9 % kTB == 0, "a weight-ring stage holds a whole number of tap groups");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ring_a = smem;                                   // SA x [raw patch | small patch]
  uint8_t* ring_b = ring_a + SA * kHaloAStageBytes;         // SB x kTB x [B_big | B_small]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring_b + SB * Cfg::kBStageBytes);
  uint64_t* raw_full = bars;            // [SA] TMA box landed (1 arrival + tx)          A loader  -> splitters
  uint64_t* a_full = raw_full + SA;     // [SA] small patch written (4 warps)             splitters -> MMA
  uint64_t* a_empty = a_full + SA;      // [SA] tcgen05.commit after the 9th tap          MMA       -> A loader
  uint64_t* b_full = a_empty + SA;      // [SB] weight tile landed (1 arrival + tx)       B loader  -> MMA
  uint64_t* b_empty = b_full + SB;      // [SB] tcgen05.commit after the tap              MMA       -> B loader
  uint64_t* acc_full = b_empty + SB;    // [2]
  uint64_t* acc_empty = acc_full + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < SA; ++s) mbar_init(&raw_full[s], 1), mbar_init(&a_full[s], kSplitWarps), mbar_init(&a_empty[s], 1);
    for (int s = 0; s < SB; ++s) mbar_init(&b_full[s], 1), mbar_init(&b_empty[s], 1);
    for (int s = 0; s < 2; ++s) mbar_init(&acc_full[s], 1), mbar_init(&acc_empty[s], kEpilogueWarps);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int chunks = kl.kb_per_tap;     // 32-channel chunks of the concatenated sources = A stages per item

  auto decode = [&](long long item, int& bb, int& y0, int& x0, int& n_tile) {
    n_tile = (int)(item % wk.n_tiles);
    const int m_tile = (int)(item / wk.n_tiles);
    const int per_img = wk.tiles_x * wk.tiles_y;
    bb = m_tile / per_img;
    const int t = m_tile - bb * per_img;
    y0 = (t / wk.tiles_x) * kHaloTH;
    x0 = (t % wk.tiles_x) * kHaloTW;
  };

  if (warp < kEpilogueWarps) {
    // ============================================================ epilogue (as in conv_tc_kernel, merged accumulator)
    const int row = warp * 32 + lane;
    const int ty = row / kHaloTW, tx = row - ty * kHaloTW;
    int use = 0;
    for (long long item = blockIdx.x; item < wk.total; item += gridDim.x, ++use) {
      int bb, y0, x0, n_tile;
      decode(item, bb, y0, x0, n_tile);
      const int buf = use & 1;
      const int oy = y0 + ty, ox = x0 + tx;
      const bool live = oy < p.out_h && ox < p.out_w;
      const long long m = ((long long)bb * p.out_h + oy) * p.out_w + ox;
      const int n_base = n_tile * BN;
      float* dst = p.dst + m * p.out_c + n_base;
      const float* res = p.residual ? p.residual + m * p.out_c + n_base : nullptr;
      mbar_wait(&acc_full[buf], (use >> 1) & 1, 1, 128);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(buf * Cfg::kAccCols) + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
      for (int cc = 0; cc < BN; cc += 32) {
        float r[32];
        const bool use_res = res != nullptr && live;
        if (use_res) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) ldg256(res + cc + j, r + j);
        }
        float v[32], v2[32];
        tmem_ld32(taddr + (uint32_t)cc, v);
        tmem_ld32(taddr + (uint32_t)(BN + cc), v2);  // + A_big x B_small^T, accumulated in the upper half
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += v2[j];
        if (!live) continue;
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bv = ld4(p.bias + n_base + cc + j);
            v[j] += bv.x, v[j + 1] += bv.y, v[j + 2] += bv.z, v[j + 3] += bv.w;
          }
        }
        if (use_res) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += r[j];
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = activate(v[j], p.act, p.act_slope);
#pragma unroll
        for (int j = 0; j < 32; j += 8) stg256(dst + cc + j, v + j);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  } else if (warp < kMmaWarp) {
    // ============================================================ splitters: small patch = x - tf32(x), same layout as raw
    // (elementwise on 16-byte units, so the swizzle never has to be computed); all four warps share every patch
    const int st = tid - kEpilogueWarps * 32;  // 0..127
    constexpr int kUnits = kHaloPatchBytes / 16;  // 1440
    const uint32_t ring_u = smem_u32(ring_a);
    int stage = 0, phase = 0;
    for (long long item = blockIdx.x; item < wk.total; item += gridDim.x) {
#pragma unroll 1
      for (int ch = 0; ch < chunks; ++ch) {
        mbar_wait(&raw_full[stage], phase, 2, 32);
        const uint32_t raw_u = ring_u + (uint32_t)(stage * kHaloAStageBytes);
        float4 v[12];
#pragma unroll
        for (int it = 0; it < 12; ++it) {
          const int u = st + it * 128;
          if (u < kUnits) v[it] = lds128(raw_u + (uint32_t)u * 16u);
        }
#pragma unroll
        for (int it = 0; it < 12; ++it) {
          const int u = st + it * 128;
          if (u < kUnits) {
            const float4 x = v[it];
            const float4 big = make_float4(tf32_big(x.x), tf32_big(x.y), tf32_big(x.z), tf32_big(x.w));
            sts128(raw_u + kHaloSlotBytes + (uint32_t)u * 16u, make_float4(x.x - big.x, x.y - big.y, x.z - big.z, x.w - big.w));
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[stage]);
        if (++stage == SA) stage = 0, phase ^= 1;
      }
    }
  } else if (warp == kMmaWarp) {
    // ============================================================ MMA issuer
    constexpr uint32_t idesc = umma_idesc_tf32(kBM, BN);
    constexpr uint32_t idesc2 = umma_idesc_tf32(kBM, 2 * BN);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    // A: SBO = one patch row (10 pixels = 1280 B); B: SBO = 1024 B.  High words are constant, low words move by plain adds.
    constexpr uint32_t kDescHiA = (uint32_t)(kHaloPW * 128 / 16) | (1u << 14) | (2u << 29);
    constexpr uint32_t kDescHiB = 64u | (1u << 14) | (2u << 29);
    const uint32_t lo_ring_a = ((smem_u32(ring_a) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t lo_ring_b = ((smem_u32(ring_b) & 0x3FFFFu) >> 4) | (1u << 16);
    auto desc_a = [](uint32_t lo) { return ((uint64_t)kDescHiA << 32) | lo; };
    auto desc_b = [](uint32_t lo) { return ((uint64_t)kDescHiB << 32) | lo; };
    int sa = 0, pa = 0, sb = 0, pb = 0, use = 0;
    for (long long item = blockIdx.x; item < wk.total; item += gridDim.x, ++use) {
      const int buf = use & 1;
      mbar_wait(&acc_empty[buf], ((use >> 1) & 1) ^ 1, 3);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_u + (uint32_t)(buf * Cfg::kAccCols);
      for (int ch = 0; ch < chunks; ++ch) {
        mbar_wait(&a_full[sa], pa, 4);                      // raw patch landed and small patch written
        tc_fence_after();
        const uint32_t lo_raw = lo_ring_a + (uint32_t)sa * (kHaloAStageBytes >> 4);
        const uint32_t lo_small = lo_raw + (kHaloSlotBytes >> 4);
#pragma unroll 1
        for (int g = 0; g < 9 / kTB; ++g) {
          mbar_wait(&b_full[sb], pb, 5);
          tc_fence_after();
          const uint32_t lo_b_stage = lo_ring_b + (uint32_t)sb * (Cfg::kBStageBytes >> 4);
          if (elect_one()) {
#pragma unroll
            for (int t = 0; t < kTB; ++t) {
              const int tap = g * kTB + t;
              const int ky = tap / 3, kx = tap - ky * 3;
              const uint32_t shift = (uint32_t)(ky * kHaloPW + kx) * (128u >> 4);  // whole pixels, in 16-byte units
              const uint32_t lo_b_big = lo_b_stage + (uint32_t)t * (Cfg::kBBytes >> 4);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint32_t ko = ks * 2;  // 8 tf32 = 32 bytes along K inside the swizzled row
                umma_tf32(tmem_d, desc_a(lo_raw + shift + ko), desc_b(lo_b_big + ko), idesc2, (ch | tap | ks) != 0);  // [big x big | big x small]
                umma_tf32(tmem_d, desc_a(lo_small + shift + ko), desc_b(lo_b_big + ko), idesc, true);                 // small x big
              }
            }
            umma_commit(&b_empty[sb]);
            if (g == 9 / kTB - 1) umma_commit(&a_empty[sa]);
          }
          __syncwarp();
          if (++sb == SB) sb = 0, pb ^= 1;
        }
        if (++sa == SA) sa = 0, pa ^= 1;
      }
      if (elect_one()) umma_commit(&acc_full[buf]);
      __syncwarp();
    }
  } else if (warp == kLoadWarp) {
    // ============================================================ A loader: one TMA box (patch) per 32-channel chunk
    int stage = 0, phase = 0;
    for (long long item = blockIdx.x; item < wk.total; item += gridDim.x) {
      int bb, y0, x0, n_tile;
      decode(item, bb, y0, x0, n_tile);
      int src = 0, c0 = 0;
      for (int ch = 0; ch < chunks; ++ch) {
        mbar_wait(&a_empty[stage], phase ^ 1, 6, 64);
        if (elect_one()) {
          mbar_arrive_expect_tx(&raw_full[stage], kHaloPatchBytes);
          tma_load_4d(ring_a + stage * kHaloAStageBytes, &maps.m[src], c0, x0 - 1, y0 - 1, bb, &raw_full[stage]);
        }
        __syncwarp();
        if (++stage == SA) stage = 0, phase ^= 1;
        c0 += kBK;
        if (c0 >= (src == 0 ? kl.src_c[0] : (src == 1 ? kl.src_c[1] : kl.src_c[2]))) c0 = 0, ++src;
      }
    }
  } else {
    // ============================================================ B loader: weight tile of (tap, chunk), packed tap-major
    int stage = 0, phase = 0;
    for (long long item = blockIdx.x; item < wk.total; item += gridDim.x) {
      int bb, y0, x0, n_tile;
      decode(item, bb, y0, x0, n_tile);
      const uint8_t* wbase = reinterpret_cast<const uint8_t*>(p.weight) + (size_t)n_tile * wk.num_kb_total * Cfg::kBBytes;
      for (int ch = 0; ch < chunks; ++ch) {
        for (int g = 0; g < 9 / kTB; ++g) {
          mbar_wait(&b_empty[stage], phase ^ 1, 7, 32);
          if (elect_one()) {
            mbar_arrive_expect_tx(&b_full[stage], Cfg::kBStageBytes);
#pragma unroll
            for (int t = 0; t < kTB; ++t)
              bulk_g2s(ring_b + stage * Cfg::kBStageBytes + t * Cfg::kBBytes,
                       wbase + (size_t)((g * kTB + t) * chunks + ch) * Cfg::kBBytes, Cfg::kBBytes, &b_full[stage]);
          }
          __syncwarp();
          if (++stage == SB) stage = 0, phase ^= 1;
        }
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
// SWIZZLE_128B K-major shared-memory image; K blocks follow KLayout (tap-major, per source 32-channel chunks, zero padded).
__global__ void pack_weight_tc_kernel(const float* __restrict__ oihw, float* __restrict__ packed, int out_c, int in_c,
                                      KLayout kl, int bn) {
  const long long tile_floats = (long long)bn * kBK;
  const long long total = (long long)(out_c / bn) * kl.num_kb * tile_floats;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i / tile_floats;
    int e = (int)(i - t * tile_floats);
    int n_tile = (int)(t / kl.num_kb), kbi = (int)(t - (long long)n_tile * kl.num_kb);
    int row = e / kBK, kk = e % kBK;
    int tap, src, c0;
    klayout_decode(kl, kbi, tap, src, c0);
    const int c = c0 + kk;
    float x = 0.f;
    if (c < kl.src_c[src]) x = oihw[((long long)(n_tile * bn + row) * in_c + kl.c_begin[src] + c) * kl.taps + tap];
    float big = tf32_big(x);
    float* tile = packed + t * 2 * tile_floats;
    uint32_t off = sw128_offset(row, kk) / 4;
    tile[off] = big;
    tile[tile_floats + off] = x - big;
  }
}

int launch_conv_simt(const dtb200_conv_params& p, int in_c_total, cudaStream_t stream);
int launch_pack_simt(const float* oihw, float* packed, int out_c, int in_c, int ksize, cudaStream_t s);

uint64_t packed_floats_tc(int out_c, int num_src, const int32_t* src_c, int ksize) {
  int in_c = 0;
  for (int s = 0; s < num_src; ++s) in_c += src_c[s];
  if (out_c % 64 != 0) return (uint64_t)out_c * in_c * ksize * ksize;  // heads stay on the SIMT layout
  return (uint64_t)out_c * make_klayout(num_src, src_c, ksize).num_kb * kBK * 2;
}

int launch_pack_tc(const float* oihw, float* packed, int out_c, int num_src, const int32_t* src_c, int ksize,
                   cudaStream_t stream) {
  int in_c = 0;
  for (int s = 0; s < num_src; ++s) in_c += src_c[s];
  if (out_c % 64 != 0) return launch_pack_simt(oihw, packed, out_c, in_c, ksize, stream);
  KLayout kl = make_klayout(num_src, src_c, ksize);
  int bn = tc_bn(out_c);
  long long total = (long long)out_c * kl.num_kb * kBK;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_weight_tc_kernel<<<blocks, 256, 0, stream>>>(oihw, packed, out_c, in_c, kl, bn);
  return check_launch("pack_weight_tc_kernel");
}

// pixel tile shape (tw x th = 128) that covers an out_h x out_w map with the fewest tiles
static void tc_tile_shape(int out_h, int out_w, int& tw, int& th) {
  long long best = -1;
  for (int w = 128; w >= 8; w >>= 1) {
    const int h = 128 / w;
    const long long tiles = (long long)((out_w + w - 1) / w) * ((out_h + h - 1) / h);
    if (best < 0 || tiles < best || (tiles == best && w == 16)) best = tiles, tw = w, th = h;
  }
}
static long long tc_m_tiles(const dtb200_conv_params& p) {
  int tw, th;
  tc_tile_shape(p.out_h, p.out_w, tw, th);
  return (long long)p.batch * ((p.out_w + tw - 1) / tw) * ((p.out_h + th - 1) / th);
}

typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFn tensor_map_encoder() {
  static TensorMapEncodeFn fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    void* ptr = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess) fn = (TensorMapEncodeFn)ptr;
  }
  return fn;
}

// Per-device host state (SM count, opt-in shared memory), resolved on first use of each device outside any stream capture.
// cudaFuncSetAttribute is per device, so a process that drives several GPUs initialises each of them; the once-flags make
// concurrent first calls from several host threads safe.
static int g_num_sms_dev[64] = {0};
static std::once_flag g_init_once[64];
int conv_tc_init() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  std::call_once(g_init_once[dev], [dev] {
    tensor_map_encoder();
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    g_num_sms_dev[dev] = sms > 0 ? sms : 148;
    cudaFuncSetAttribute(conv_tc_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<128, false>::kSmemBytes);
    cudaFuncSetAttribute(conv_tc_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<64, false>::kSmemBytes);
    cudaFuncSetAttribute(conv_tc_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<128, true>::kSmemBytes);
    cudaFuncSetAttribute(conv_tc_kernel<64, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<64, true, 2>::kSmemBytes);
    cudaFuncSetAttribute(conv_tc_kernel<64, true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<64, true, 3>::kSmemBytes);
    cudaFuncSetAttribute(conv_tc_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<64, true>::kSmemBytes);
    cudaFuncSetAttribute(conv_tc_halo_kernel<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, HaloCfg<64, 1>::kSmemBytes);
    cudaFuncSetAttribute(conv_tc_halo_kernel<64, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, HaloCfg<64, 3>::kSmemBytes);
    cudaGetLastError();
  });
  return g_num_sms_dev[dev];
}

// split-K plan for maps with fewer (M, N) tiles than SMs.  The kernel is persistent, so time ~ rounds x (K blocks per
// item + fixed per-item cost) + reduction cost: pick the split count that minimises it (one full round of longer items
// beats three rounds of short ones).  148 is used as the machine width on purpose: the plan (and the fp32 summation
// order it implies) must not depend on the device the workspace was sized on.
static int tc_splits(long long m_tiles, int out_c, int num_kb) {
  const int bn = tc_bn(out_c);
  const long long ctas = m_tiles * (out_c / bn);
  if (ctas >= 148 || num_kb < 4) return 1;
  if (conv_flags() & 4) {  // round-1 rule: >= 2 x 148 items of >= 4 K blocks
    int splits = (int)((2 * 148 + ctas - 1) / ctas);
    if (splits > num_kb / 4) splits = num_kb / 4;
    if (splits > 32) splits = 32;
    return splits < 1 ? 1 : splits;
  }
  int best = 1;
  double best_cost = 1e30;
  for (int sp = 1; sp <= 32 && sp * 2 <= num_kb; ++sp) {
    const int kb_per = (num_kb + sp - 1) / sp;
    if ((num_kb + kb_per - 1) / kb_per != sp) continue;  // not a distinct partition
    const long long rounds = (ctas * sp + 147) / 148;
    const double cost = (double)rounds * (kb_per + 2.0) + (sp > 1 ? 4.0 + 0.5 * sp : 0.0);
    if (cost < best_cost - 1e-9) best_cost = cost, best = sp;
  }
  return best;
}

// 3x3 / stride-1 layers with 64 output channels, no split-K and at least one full round of 8 x 16 tiles run on the
// halo-tile kernel; fills the tile grid of that kernel.
static bool halo_tiles(const dtb200_conv_params& p, int bn, int splits, TcWork* hw) {
  if ((conv_flags() & 128) || p.ksize != 3 || p.stride != 1 || bn != 64 || splits != 1 || p.in_h != p.out_h || p.in_w != p.out_w)
    return false;
  hw->tw = kHaloTW, hw->th = kHaloTH;
  hw->tiles_x = (p.out_w + kHaloTW - 1) / kHaloTW;
  hw->tiles_y = (p.out_h + kHaloTH - 1) / kHaloTH;
  hw->m_tiles = p.batch * hw->tiles_x * hw->tiles_y;
  hw->n_tiles = p.out_c / bn;
  hw->total = (long long)hw->m_tiles * hw->n_tiles;
  return hw->total >= 148;
}

uint64_t conv_tc_workspace_bytes(const dtb200_conv_params& p, int in_c_total) {
  if (p.ksize == 0 || p.out_c % 64 != 0) return 0;
  const long long m_total = (long long)p.batch * p.out_h * p.out_w;
  (void)in_c_total;
  const KLayout kl = make_klayout(p.num_src, p.src_c, p.ksize);
  const int splits = tc_splits(tc_m_tiles(p), p.out_c, kl.num_kb);
  if (splits > 1) return (uint64_t)splits * m_total * p.out_c * sizeof(float);
  return 0;
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
  for (int s = 0; s < p.num_src; ++s) {
    if (p.src_c[s] % 4 != 0)
      return fail(DTB200_ERR_UNSUPPORTED, "conv (tc3x): every source needs a multiple of 4 channels, got %s%lld", "", p.src_c[s]);
    if (p.src_resample[s] != DTB200_RESAMPLE_NONE)
      return fail(DTB200_ERR_UNSUPPORTED,
                  "conv (tc3x): x2-resampled sources must be materialised first (ksize = 0 descriptor); ConvPlan does this%s");
    if (reinterpret_cast<uintptr_t>(p.src[s]) % 16 != 0)
      return fail(DTB200_ERR_INVALID, "conv (tc3x): source pointers must be 16-byte aligned%s");
  }
  if (reinterpret_cast<uintptr_t>(p.dst) % 32 != 0 || reinterpret_cast<uintptr_t>(p.residual) % 32 != 0 ||
      reinterpret_cast<uintptr_t>(p.workspace) % 32 != 0)
    return fail(DTB200_ERR_INVALID, "conv (tc3x): dst / residual / workspace must be 32-byte aligned (256-bit stores)%s");
  TensorMapEncodeFn encode = tensor_map_encoder();
  if (!encode) return fail(DTB200_ERR_CUDA, "conv (tc3x): cuTensorMapEncodeTiled entry point not available%s");
  const long long m_total = (long long)p.batch * p.out_h * p.out_w;
  const KLayout kl = make_klayout(p.num_src, p.src_c, p.ksize);
  const int num_kb = kl.num_kb;
  const int bn = tc_bn(p.out_c);
  TcWork wk;
  tc_tile_shape(p.out_h, p.out_w, wk.tw, wk.th);
  wk.tiles_x = (p.out_w + wk.tw - 1) / wk.tw;
  wk.tiles_y = (p.out_h + wk.th - 1) / wk.th;
  wk.m_tiles = p.batch * wk.tiles_x * wk.tiles_y;
  const int splits = tc_splits(wk.m_tiles, p.out_c, num_kb);
  float* partial = nullptr;
  if (splits > 1) {
    uint64_t need = (uint64_t)splits * m_total * p.out_c * sizeof(float);
    if (!p.workspace || p.workspace_bytes < need)
      return fail(DTB200_ERR_INVALID, "conv (tc3x): split-K needs %s%lld workspace bytes (dtb200_conv_workspace_bytes)", "",
                  (long long)need);
    partial = reinterpret_cast<float*>(p.workspace);
  }
  wk.num_kb_total = num_kb;
  wk.kb_per_split = (num_kb + splits - 1) / splits;
  wk.splits = (num_kb + wk.kb_per_split - 1) / wk.kb_per_split;
  wk.n_tiles = p.out_c / bn;
  wk.total = (long long)wk.m_tiles * wk.n_tiles * wk.splits;
  const int zsplits = wk.splits;

  // 3x3 / stride-1 layers with at least one full round of 8 x 16 tiles: halo-tile kernel (one patch load + one split per
  // 9 taps).  Development switches: bit 7 of the flags turns it OFF (tap-major kernel everywhere), bit 3 selects the
  // one-tap-per-weight-stage variant (default: 3 taps per stage).
  {
    TcWork hw = wk;
    if (halo_tiles(p, bn, splits, &hw)) {
      TcMaps hmaps;
      memset(&hmaps, 0, sizeof(hmaps));
      for (int s = 0; s < p.num_src; ++s) {
        const cuuint64_t C = (cuuint64_t)p.src_c[s];
        cuuint64_t dims[4] = {C, (cuuint64_t)p.in_w, (cuuint64_t)p.in_h, (cuuint64_t)p.batch};
        cuuint64_t strides[3] = {C * 4, (cuuint64_t)p.in_w * C * 4, (cuuint64_t)p.in_h * p.in_w * C * 4};
        cuuint32_t box[4] = {(cuuint32_t)kBK, (cuuint32_t)kHaloPW, (cuuint32_t)kHaloPH, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&hmaps.m[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(p.src[s]), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(DTB200_ERR_CUDA, "conv (tc3x halo): cuTensorMapEncodeTiled failed with code %s%lld", "", (long long)r);
      }
      const int g_num_sms = conv_tc_init();
      const unsigned hgrid = (unsigned)(hw.total < g_num_sms ? hw.total : g_num_sms);
      if (conv_flags() & 8)  // one tap per weight-ring stage (6 stages): 51.2 us on the 240x320 64->64 layer vs 45.1 us
        conv_tc_halo_kernel<64, 1><<<hgrid, kThreads, HaloCfg<64, 1>::kSmemBytes, stream>>>(p, hmaps, kl, hw);
      else
        conv_tc_halo_kernel<64, 3><<<hgrid, kThreads, HaloCfg<64, 3>::kSmemBytes, stream>>>(p, hmaps, kl, hw);
      return check_launch("conv_tc_halo_kernel");
    }
  }

  // one 4-D tensor map (C, W, H, B) per source: box = 32 channels x (tw x th) output pixels; for stride 2 the box spans
  // 2*tw x 2*th input pixels of which every second one is written (elementStrides)
  TcMaps maps;
  memset(&maps, 0, sizeof(maps));
  for (int s = 0; s < p.num_src; ++s) {
    const cuuint64_t C = (cuuint64_t)p.src_c[s];
    cuuint64_t dims[4] = {C, (cuuint64_t)p.in_w, (cuuint64_t)p.in_h, (cuuint64_t)p.batch};
    cuuint64_t strides[3] = {C * 4, (cuuint64_t)p.in_w * C * 4, (cuuint64_t)p.in_h * p.in_w * C * 4};
    cuuint32_t box[4] = {(cuuint32_t)kBK, (cuuint32_t)(wk.tw * p.stride), (cuuint32_t)(wk.th * p.stride), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)p.stride, (cuuint32_t)p.stride, 1};
    CUresult r = encode(&maps.m[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(p.src[s]), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(DTB200_ERR_CUDA, "conv (tc3x): cuTensorMapEncodeTiled failed with code %s%lld", "", (long long)r);
  }
  const int num_sms = conv_tc_init();
  const unsigned grid = (unsigned)(wk.total < num_sms ? wk.total : num_sms);
  const int dbg = conv_flags() >> 8;  // timing knock-outs (results are wrong): 1 no MMA, 2 no split, 4 no A box, 8 no B tile, 16 no store
  if (bn == 128 && !(conv_flags() & 2)) {
    conv_tc_kernel<128, true><<<grid, kThreads, TcCfg<128, true>::kSmemBytes, stream>>>(p, maps, kl, m_total, wk, partial, dbg);
  } else if (bn == 128) {
    conv_tc_kernel<128, false><<<grid, kThreads, TcCfg<128, false>::kSmemBytes, stream>>>(p, maps, kl, m_total, wk, partial, dbg);
  } else if (conv_flags() & 1) {
    conv_tc_kernel<64, false><<<grid, kThreads, TcCfg<64, false>::kSmemBytes, stream>>>(p, maps, kl, m_total, wk, partial, dbg);
  } else if (conv_flags() & 32) {  // experiment: ring depth 2 / 3 instead of 4
    conv_tc_kernel<64, true, 2><<<grid, kThreads, TcCfg<64, true, 2>::kSmemBytes, stream>>>(p, maps, kl, m_total, wk, partial, dbg);
  } else if (conv_flags() & 64) {
    conv_tc_kernel<64, true, 3><<<grid, kThreads, TcCfg<64, true, 3>::kSmemBytes, stream>>>(p, maps, kl, m_total, wk, partial, dbg);
  } else {
    conv_tc_kernel<64, true><<<grid, kThreads, TcCfg<64, true>::kSmemBytes, stream>>>(p, maps, kl, m_total, wk, partial, dbg);
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
