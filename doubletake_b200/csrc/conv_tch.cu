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
  int debug;   // timing knock-outs (development only, DTB200_DEVELOPMENT=1): 0x100 no weight copies, 0x200 no patch loads,
               // 0x400 no MMAs, 0x800 no epilogue stores -- after the first ring fill; results are wrong by design
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

// development timeline (dtb200_debug_trace): every tch conv kernel stamps the earliest start and the latest end of its CTAs
__device__ __forceinline__ unsigned long long tch_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void tch_trace_begin(unsigned long long* trace) {
  if (trace && threadIdx.x == 0) atomicMin(trace, tch_globaltimer());
}
__device__ __forceinline__ void tch_trace_end(unsigned long long* trace) {
  if (trace && threadIdx.x == 0) atomicMax(trace + 1, tch_globaltimer());
}

__device__ __forceinline__ uint4 ldg128u(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void stg128u(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  *reinterpret_cast<uint4*>(p) = make_uint4(a, b, c, d);
}
__device__ __forceinline__ void ldg256u(const void* p, uint32_t (&v)[8]) {   // 32-byte aligned, read-only path
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256u(void* p, const uint32_t (&v)[8]) {   // 32-byte aligned: one full sector
  asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]),
               "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ float2 join_pair(uint32_t big, uint32_t small) {
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&big));
  const float2 s = __half22float2(*reinterpret_cast<const __half2*>(&small));
  return make_float2(fmaf(s.x, kHalfSplitInv, b.x), fmaf(s.y, kHalfSplitInv, b.y));
}

// Epilogue of one tile row (one output pixel) for 32 consecutive output channels starting at channel c0:
// v[] = main + corr / 2048 already combined.  Adds bias / residual, activates, splits, stores both planes.
// The activation is a template parameter: with a run-time switch per element the epilogue of a tile (770 instructions per
// thread, two warps per scheduler) took longer than the tile's MMAs (tools/conv_bench.py --debug 4096 timeline, profiles/r02d_*).
template <int ACT>
__device__ __forceinline__ float tch_act(float v, float slope) {
  if (ACT == DTB200_ACT_LEAKY) return v > 0.f ? v : v * slope;
  if (ACT == DTB200_ACT_ELU) return v > 0.f ? v : expm1f(v);
  return v;
}
// The residual of one (pixel, 32-channel chunk): 64 bytes of each plane, fetched BEFORE the epilogue warp waits for the tile's
// accumulator -- the loads do not depend on it, and issued after the wait their L2 latency made every "+res" layer's epilogue
// longer than the tile's MMAs (25 vs 20 us per 240x320 layer in the graph timeline, profiles/r02d_*).
struct TchResid {
  uint32_t big[16], small[16];
};
__device__ __forceinline__ void tch_resid_load(const dtb200_conv_params& p, long long m, int c0, TchResid& r) {
  const uint8_t* src = reinterpret_cast<const uint8_t*>(p.residual) + (size_t)m * ((size_t)p.out_c * 4) + (size_t)c0 * 2;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    ldg256u(src + 32 * j, *reinterpret_cast<uint32_t(*)[8]>(&r.big[8 * j]));
    ldg256u(src + (size_t)p.out_c * 2 + 32 * j, *reinterpret_cast<uint32_t(*)[8]>(&r.small[8 * j]));
  }
}
template <int ACT>
__device__ __forceinline__ void tch_finish_chunk_t(const dtb200_conv_params& p, float (&v)[32], long long m, int c0, const TchResid* res,
                                                   bool knock_stores) {
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 bv = ld4(p.bias + c0 + j);
      v[j] += bv.x, v[j + 1] += bv.y, v[j + 2] += bv.z, v[j + 3] += bv.w;
    }
  }
  if (res) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float2 a = join_pair(res->big[i], res->small[i]);
      v[2 * i] += a.x, v[2 * i + 1] += a.y;
    }
  }
  // 256-bit stores: a thread owns 64 contiguous bytes of each plane of its pixel = two full 32-byte sectors per plane
  uint8_t* d = reinterpret_cast<uint8_t*>(p.dst) + (size_t)m * ((size_t)p.out_c * 4) + (size_t)c0 * 2;
  const float slope = p.act_slope;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    uint32_t bg[8], sl[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) split_half2_sat(tch_act<ACT>(v[16 * j + 2 * i], slope), tch_act<ACT>(v[16 * j + 2 * i + 1], slope), bg[i], sl[i]);
    if (knock_stores && (bg[0] ^ sl[7]) != 0x7e57c0deu) continue;   // development timing: all the arithmetic, (practically) no store
    stg256u(d + 32 * j, bg);
    stg256u(d + (size_t)p.out_c * 2 + 32 * j, sl);
  }
}
// `res`: the prefetched residual, or nullptr when the layer has none (p.residual == nullptr)
__device__ __forceinline__ void tch_finish_chunk(const dtb200_conv_params& p, float (&v)[32], long long m, int c0, const TchResid* res,
                                                 bool knock_stores = false) {
  if (p.act == DTB200_ACT_LEAKY) tch_finish_chunk_t<DTB200_ACT_LEAKY>(p, v, m, c0, res, knock_stores);
  else if (p.act == DTB200_ACT_ELU) tch_finish_chunk_t<DTB200_ACT_ELU>(p, v, m, c0, res, knock_stores);
  else tch_finish_chunk_t<DTB200_ACT_NONE>(p, v, m, c0, res, knock_stores);
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

// ---- MMA issue.  tcgen05.mma issue is synchronous with the tensor pipe (tools/umma_bench.cu), so every instruction the issuing
// thread spends between two MMAs is tensor idle time.  A K-step loop with a run-time trip count costs two R2UR, a compare and
// two branches per MMA -- measured ~88 clk per issue where the pipe needs 64 (N = 128) or 48 (N = 64), profiles/r02d_*.  The
// issue sequences below are therefore fully unrolled over a COMPILE-TIME number of K steps (descriptor low words = one uniform
// base + immediates) and the callers dispatch on the step count once per stage.
constexpr uint32_t kDescHiK128 = 64u | (1u << 14) | (2u << 29);   // SBO = 1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint64_t tch_desc(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; }

// one K block of the tap-major kernel: KS x { [main | corr] += A_big x [W_big | W_small]^T ; corr += A_small x W_big^T }
template <int BN, int KS>
__device__ __forceinline__ void tch_issue_kblock(uint32_t tmem_d, uint32_t lo_a_big, uint32_t lo_a_small, uint32_t lo_b, bool first_kb) {
  constexpr uint32_t idesc = umma_idesc_f16(kCM, BN), idesc2 = umma_idesc_f16(kCM, 2 * BN);
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const uint32_t ko = ks * 2;   // 16 fp16 = 32 bytes along K inside the swizzled row, in 16-byte units
    umma_f16(tmem_d, tch_desc(kDescHiK128, lo_a_big + ko), tch_desc(kDescHiK128, lo_b + ko), idesc2, ks != 0 || !first_kb);
    umma_f16(tmem_d + BN, tch_desc(kDescHiK128, lo_a_small + ko), tch_desc(kDescHiK128, lo_b + ko), idesc, true);
  }
}
template <int BN>
__device__ __forceinline__ void tch_issue_kblock_n(int ksteps, uint32_t tmem_d, uint32_t lo_a_big, uint32_t lo_a_small, uint32_t lo_b,
                                                   bool first_kb) {
  switch (ksteps) {
    case 4: tch_issue_kblock<BN, 4>(tmem_d, lo_a_big, lo_a_small, lo_b, first_kb); break;
    case 3: tch_issue_kblock<BN, 3>(tmem_d, lo_a_big, lo_a_small, lo_b, first_kb); break;
    case 2: tch_issue_kblock<BN, 2>(tmem_d, lo_a_big, lo_a_small, lo_b, first_kb); break;
    default: tch_issue_kblock<BN, 1>(tmem_d, lo_a_big, lo_a_small, lo_b, first_kb); break;
  }
}

// 4 consecutive output channels of one pixel (cluster split-K reducer): the operation order of tch_finish_chunk.  Consecutive
// lanes own consecutive 16 bytes of the fp32 staging row and consecutive 8 bytes of each output plane.
template <int ACT>
__device__ __forceinline__ void tch_finish4_t(const dtb200_conv_params& p, float4 v, long long m, int c0) {
  if (p.bias) {
    const float4 b = ld4(p.bias + c0);
    v.x += b.x, v.y += b.y, v.z += b.z, v.w += b.w;
  }
  const size_t row_bytes = (size_t)p.out_c * 4;
  if (p.residual) {
    const uint8_t* r = reinterpret_cast<const uint8_t*>(p.residual) + (size_t)m * row_bytes + (size_t)c0 * 2;
    const uint2 rb = __ldg(reinterpret_cast<const uint2*>(r)), rs = __ldg(reinterpret_cast<const uint2*>(r + (size_t)p.out_c * 2));
    const float2 a0 = join_pair(rb.x, rs.x), a1 = join_pair(rb.y, rs.y);
    v.x += a0.x, v.y += a0.y, v.z += a1.x, v.w += a1.y;
  }
  const float slope = p.act_slope;
  uint2 bg, sl;
  split_half2_sat(tch_act<ACT>(v.x, slope), tch_act<ACT>(v.y, slope), bg.x, sl.x);
  split_half2_sat(tch_act<ACT>(v.z, slope), tch_act<ACT>(v.w, slope), bg.y, sl.y);
  uint8_t* d = reinterpret_cast<uint8_t*>(p.dst) + (size_t)m * row_bytes + (size_t)c0 * 2;
  *reinterpret_cast<uint2*>(d) = bg;
  *reinterpret_cast<uint2*>(d + (size_t)p.out_c * 2) = sl;
}
__device__ __forceinline__ void tch_finish4(const dtb200_conv_params& p, float4 v, long long m, int c0) {
  if (p.act == DTB200_ACT_LEAKY) tch_finish4_t<DTB200_ACT_LEAKY>(p, v, m, c0);
  else if (p.act == DTB200_ACT_ELU) tch_finish4_t<DTB200_ACT_ELU>(p, v, m, c0);
  else tch_finish4_t<DTB200_ACT_NONE>(p, v, m, c0);
}

// ======================================================================================================================
// Tap-major kernel: any 1x1 / 3x3, stride 1 / 2, 64- or 128-channel N tiles.  Persistent over (M, N) tiles.
//   warps 0-7  epilogue (double-buffered TMEM accumulator)
//   warp 8     MMA issuer (elect.sync): per K block up to 4 K steps x 2 tcgen05.mma, tcgen05.commit frees the stage
//   warp 9     A loader: two TMA boxes per K block (big / small plane) = 64 channels x (TW x TH = 128) output pixels of one
//              source at one tap, SWIZZLE_128B, hardware zero fill outside the image and beyond the source's channels
//   warp 10    B loader: one cp.async.bulk of the pre-packed weight tile
// Split-K for maps with fewer tiles than SMs runs ACROSS A THREAD-BLOCK CLUSTER: the `splits` CTAs of a cluster own the same
// (M, N) tile and consecutive K ranges; every CTA parks its fp32 partial tile in its own shared memory (over the idle operand
// ring), and after a cluster-scope mbarrier round CTA r sums rows [128 r / splits, 128 (r+1) / splits) of all partial tiles
// through distributed shared memory -- IN SPLIT ORDER, so the result does not depend on timing -- and runs the epilogue for
// them.  No HBM round trip of partial sums and no second kernel (the first tch version wrote `splits` fp32 copies of the map to a
// workspace and launched a reduce kernel: 9-13 us per layer on 62 of the 169 layers of the cfg-2 plan).
// ======================================================================================================================
template <int BN>
__global__ void __launch_bounds__(kCThreads, 1) conv_tch_kernel(const dtb200_conv_params p, const __grid_constant__ HMaps maps,
                                                                KLayoutH kl, HWork wk, unsigned long long* trace) {
  tch_trace_begin(trace);
  using Cfg = TchCfg<BN>;
  constexpr int S = Cfg::kStages;
  constexpr int kRowF = BN + 4;         // staging row pitch in floats: 16-byte lane skew -> conflict-free v4 stores
  static_assert(kCM * kRowF * 4 <= S * Cfg::kStageBytes, "the split-K staging tile must fit into the operand ring");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + S * Cfg::kStageBytes);
  uint64_t* full = bars;                // [S] A boxes + weight tile landed (2 arrivals + tx bytes)
  uint64_t* empty = full + S;           // [S] tcgen05.commit
  uint64_t* acc_full = empty + S;       // [2]
  uint64_t* acc_empty = acc_full + 2;   // [2]
  uint64_t* red_full = acc_empty + 2;   // [1] split-K: every CTA of the cluster has staged its partial tile (8 warps x splits)
  uint64_t* red_done = red_full + 1;    // [1] split-K: every CTA of the cluster has finished reading this CTA's staging
  uint64_t* ring_free = red_done + 1;   // [1] split-K: epilogue warps hand the ring back to the loaders
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ring_free + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == kCLoadA && lane < 2 * p.num_src)   // descriptor fetch off the first load's critical path
    prefetch_tensormap(lane & 1 ? &maps.small[lane >> 1] : &maps.big[lane >> 1]);
  // development timeline (debug bit 0x1000): CTA 0 records clock64 at the hand-over points of its first item and prints them
  __shared__ long long tr[24];
  __shared__ long long tr_t0;
  const bool tracing = (wk.debug & 0x1000) && blockIdx.x == 0;
#define TM_TRACE(idx) do { if (tracing && lane == 0) tr[idx] = clock64() - tr_t0; } while (0)
  if (tid == 0) tr_t0 = clock64();
  const int splits = wk.splits;
  const int rank = splits > 1 ? (int)cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / splits, num_clusters = gridDim.x / splits;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 2), mbar_init(&empty[s], 1);
    for (int s = 0; s < 2; ++s) mbar_init(&acc_full[s], 1), mbar_init(&acc_empty[s], kCEpiWarps);
    mbar_init(red_full, kCEpiWarps * splits), mbar_init(red_done, kCEpiWarps * splits), mbar_init(ring_free, kCEpiWarps);
    fence_mbar_init();
  }
  if (warp == kCMmaWarp) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) TM_TRACE(0);
  if (splits > 1) cluster_sync_all();   // a peer may arrive on red_full / red_done only after this CTA has initialised them
  if (warp == 0) TM_TRACE(1);
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: the next layer may be scheduled from here on (its prologue and weight prefetch overlap this
  // grid's tail); everything of THIS grid that touches activations waits for its own prerequisites first.  The weight loader
  // does not wait: weights are never written inside a plan.
  if (tid == 0) griddep_launch_dependents();
  if (warp != kCLoadB) griddep_wait();
  if (warp == 0) TM_TRACE(2);

  auto decode = [&](long long tile, int& bb, int& y0, int& x0, int& n_tile, int& kb_begin, int& num_kb) {
    n_tile = (int)(tile % wk.n_tiles);
    const int m_tile = (int)(tile / wk.n_tiles);
    const int per_img = wk.tiles_x * wk.tiles_y;
    bb = m_tile / per_img;
    const int t = m_tile - bb * per_img;
    y0 = (t / wk.tiles_x) * wk.th;
    x0 = (t % wk.tiles_x) * wk.tw;
    kb_begin = rank * wk.kb_per_split;
    num_kb = min(wk.kb_per_split, wk.num_kb_total - kb_begin);
  };

  if (warp < kCEpiWarps) {
    // ============================================================ epilogue
    const int qd = warp & 3, chalf = warp >> 2;
    const int row = qd * 32 + lane;
    const int ty = row / wk.tw, tx = row - ty * wk.tw;
    const uint32_t stage_u = smem_u32(ring);
    int use = 0;
    for (long long tile = cluster_id; tile < wk.total; tile += num_clusters, ++use) {
      int bb, y0, x0, n_tile, kb_begin, num_kb;
      decode(tile, bb, y0, x0, n_tile, kb_begin, num_kb);
      const int buf = use & 1;
      const int n_base = n_tile * BN;
      mbar_wait(&acc_full[buf], (use >> 1) & 1, 1, 128);
      tc_fence_after();
      if (warp == 0 && use == 0) TM_TRACE(3);
      const uint32_t taddr = tmem_base + (uint32_t)(buf * Cfg::kAccCols) + ((uint32_t)(qd * 32) << 16);
      if (splits == 1) {
        const int oy = y0 + ty, ox = x0 + tx;
        const bool live = oy < p.out_h && ox < p.out_w;
        const long long m = ((long long)bb * p.out_h + oy) * p.out_w + ox;
#pragma unroll 1
        for (int cc = chalf * (BN / 2); cc < (chalf + 1) * (BN / 2); cc += 32) {
          float v[32], c[32];
          tmem_ld32(taddr + (uint32_t)cc, v);
          tmem_ld32(taddr + (uint32_t)(BN + cc), c);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaf(c[j], kHalfSplitInv, v[j]);
          if (live) {
            TchResid res;
            if (p.residual) tch_resid_load(p, m, n_base + cc, res);
            tch_finish_chunk(p, v, m, n_base + cc, p.residual ? &res : nullptr);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
        continue;
      }
      // ---- cluster split-K.  1: park this CTA's partial tile ([128 rows][BN] fp32) in the ring: every MMA of the item has
      // completed (acc_full), and the loaders do not touch the ring again before ring_free.
#pragma unroll 1
      for (int cc = chalf * (BN / 2); cc < (chalf + 1) * (BN / 2); cc += 32) {
        float v[32], c[32];
        tmem_ld32(taddr + (uint32_t)cc, v);
        tmem_ld32(taddr + (uint32_t)(BN + cc), c);
        const uint32_t a = stage_u + (uint32_t)(row * kRowF + cc) * 4u;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          sts128(a + 4u * j, make_float4(fmaf(c[j], kHalfSplitInv, v[j]), fmaf(c[j + 1], kHalfSplitInv, v[j + 1]),
                                         fmaf(c[j + 2], kHalfSplitInv, v[j + 2]), fmaf(c[j + 3], kHalfSplitInv, v[j + 3])));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      // 2: tell every CTA of the cluster (release), wait until all of them have staged (acquire)
      if (warp == 0 && use == 0) TM_TRACE(4);
      if (lane < splits) mbar_arrive_cluster(mapa_shared(smem_u32(red_full), (uint32_t)lane));
      mbar_wait_cluster(red_full, use & 1, 8);
      if (warp == 0 && use == 0) TM_TRACE(5);
      // 3: rows [r0, r1) of the tile are reduced here, partial tiles summed in split order
      const int r0 = rank * kCM / splits, r1 = (rank + 1) * kCM / splits;
      // Lane l of a warp owns floats [4 l', 4 l' + 4) of a staging row: one ld.shared::cluster.v4 of the warp reads 512 contiguous
      // bytes of the peer's shared memory (DSMEM moves ~20 B/clk per SM, so the 64 KB of a 128 x 128 tile bound this phase).
      // Every partial tile's 16 bytes are requested before the first one is used (8 = the largest cluster).
      constexpr int kPerRow = BN / 4;
      for (int it = tid; it < (r1 - r0) * kPerRow; it += kCEpiWarps * 32) {
        const int rr = r0 + it / kPerRow, c4 = (it % kPerRow) * 4;
        const uint32_t a = stage_u + (uint32_t)(rr * kRowF + c4) * 4u;
        float4 px[8];
#pragma unroll
        for (int s = 0; s < 8; ++s)
          if (s < splits) px[s] = s == rank ? lds128(a) : ld_dsmem128(mapa_shared(a, (uint32_t)s));   // own partial: local port
        float4 acc = px[0];
#pragma unroll
        for (int s = 1; s < 8; ++s)
          if (s < splits) acc.x += px[s].x, acc.y += px[s].y, acc.z += px[s].z, acc.w += px[s].w;   // split order
        const int ry = rr / wk.tw, rx = rr - ry * wk.tw;
        const int oy = y0 + ry, ox = x0 + rx;
        if (oy < p.out_h && ox < p.out_w) tch_finish4(p, acc, ((long long)bb * p.out_h + oy) * p.out_w + ox, n_base + c4);
      }
      // 4: nobody may overwrite (next item's loads) or retire (exit) a staging tile a peer still reads
      __syncwarp();
      if (warp == 0 && use == 0) TM_TRACE(6);
      if (lane < splits) mbar_arrive_cluster(mapa_shared(smem_u32(red_done), (uint32_t)lane));
      mbar_wait_cluster(red_done, use & 1, 9);
      if (warp == 0 && use == 0) TM_TRACE(7);
      fence_proxy_async_smem();
      if (lane == 0) mbar_arrive(ring_free);
    }
  } else if (warp == kCMmaWarp) {
    // ============================================================ MMA issuer
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t lo_ring = ((smem_u32(ring) & 0x3FFFFu) >> 4) | (1u << 16);
    int stage = 0, phase = 0, use = 0;
    for (long long tile = cluster_id; tile < wk.total; tile += num_clusters, ++use) {
      int bb, y0, x0, n_tile, kb_begin, num_kb;
      decode(tile, bb, y0, x0, n_tile, kb_begin, num_kb);
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
        if (use == 0 && kb < 12) TM_TRACE(8 + kb);
        const uint32_t lo_a_big = lo_ring + (uint32_t)stage * (Cfg::kStageBytes >> 4);
        if (elect_one()) {
          tch_issue_kblock_n<BN>(ksteps, tmem_d, lo_a_big, lo_a_big + (kCTile >> 4), lo_a_big + (2 * kCTile >> 4), kb == 0);
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
      if (use == 0) TM_TRACE(20);
    }
  } else if (warp == kCLoadA) {
    // ============================================================ A loader
    const int pad = p.ksize / 2;
    const uint32_t ring_u = smem_u32(ring);
    int stage = 0, phase = 0, use = 0;
    for (long long tile = cluster_id; tile < wk.total; tile += num_clusters, ++use) {
      int bb, y0, x0, n_tile, kb_begin, num_kb;
      decode(tile, bb, y0, x0, n_tile, kb_begin, num_kb);
      int tap, src, c0;
      klayout_h_decode(kl, kb_begin, tap, src, c0);
      int ky = tap / p.ksize, kx = tap - ky * p.ksize;
      if (splits > 1 && use > 0) mbar_wait(ring_free, (use - 1) & 1, 10, 32);
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
    int stage = 0, phase = 0, use = 0;
    for (long long tile = cluster_id; tile < wk.total; tile += num_clusters, ++use) {
      int bb, y0, x0, n_tile, kb_begin, num_kb;
      decode(tile, bb, y0, x0, n_tile, kb_begin, num_kb);
      const uint8_t* wbase = reinterpret_cast<const uint8_t*>(p.weight) + ((size_t)n_tile * wk.num_kb_total + kb_begin) * Cfg::kBBytes;
      if (splits > 1 && use > 0) mbar_wait(ring_free, (use - 1) & 1, 11, 32);
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
  tch_trace_end(trace);
  if (warp == kCMmaWarp) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
  if (tracing && tid == 0) {
    const int nkb = min(wk.kb_per_split, 12);
    printf("conv_tch<%d> trace (clk since kernel start, CTA 0, grid %d, splits %d, %d K blocks): init+tmem %lld, cluster sync %lld, dependency wait %lld | "
           "K block full seen:", BN, (int)gridDim.x, splits, wk.kb_per_split, tr[0], tr[1], tr[2]);
    for (int i = 0; i < nkb; ++i) printf(" %lld", tr[8 + i]);
    printf(" | all issued %lld | epi: acc_full seen %lld, staged %lld, all staged %lld, reduced+stored %lld, peers done %lld | end %lld\n", tr[20], tr[3],
           tr[4], tr[5], tr[6], tr[7], clock64() - tr_t0);
  }
#undef TM_TRACE
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
constexpr uint32_t kDescHiPatch = (uint32_t)(kHPW * 128 / 16) | (1u << 14) | (2u << 29);   // A: SBO = one patch row (1280 B)
constexpr int kHBBytes = 2 * 64 * 128;                              // W_big | W_small of one (tap, chunk), BN = 64

// one kernel row (3 taps) x KS K steps against one 64-channel patch stage (both planes); `lo_big` / `lo_small` already point at
// patch row ky, `lo_b` at the row's first weight tile.  Fully unrolled, see tch_issue_kblock.
template <int KS>
__device__ __forceinline__ void halo_issue_row(uint32_t tmem_d, uint32_t lo_big, uint32_t lo_small, uint32_t lo_b, bool first) {
  constexpr uint32_t idesc = umma_idesc_f16(kCM, 64), idesc2 = umma_idesc_f16(kCM, 128);
#pragma unroll
  for (int t = 0; t < 3; ++t) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const uint32_t a_off = (uint32_t)t * (128u >> 4) + ks * 2, b_off = (uint32_t)t * (kHBBytes >> 4) + ks * 2;
      umma_f16(tmem_d, tch_desc(kDescHiPatch, lo_big + a_off), tch_desc(kDescHiK128, lo_b + b_off), idesc2, (t | ks) != 0 || !first);
      umma_f16(tmem_d + 64, tch_desc(kDescHiPatch, lo_small + a_off), tch_desc(kDescHiK128, lo_b + b_off), idesc, true);
    }
  }
}
__device__ __forceinline__ void halo_issue_row_n(int ksteps, uint32_t tmem_d, uint32_t lo_big, uint32_t lo_small, uint32_t lo_b, bool first) {
  switch (ksteps) {
    case 4: halo_issue_row<4>(tmem_d, lo_big, lo_small, lo_b, first); break;
    case 3: halo_issue_row<3>(tmem_d, lo_big, lo_small, lo_b, first); break;
    case 2: halo_issue_row<2>(tmem_d, lo_big, lo_small, lo_b, first); break;
    case 1: halo_issue_row<1>(tmem_d, lo_big, lo_small, lo_b, first); break;
    default: break;   // timing knock-out
  }
}
// resident-weights kernel: all 9 taps x KS K steps of ONE patch plane.  kBig: [main | corr] += A_big x [W_big | W_small]^T (the very
// first MMA of a tile overwrites); else corr += A_small x W_big^T.
template <int KS, bool kBig>
__device__ __forceinline__ void halo_issue_plane(uint32_t tmem_d, uint32_t lo_a, uint32_t lo_w) {
  constexpr uint32_t idesc = umma_idesc_f16(kCM, 64), idesc2 = umma_idesc_f16(kCM, 128);
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const uint32_t a_off = (uint32_t)((tap / 3) * kHPW + tap % 3) * (128u >> 4) + ks * 2;   // whole pixels, in 16-byte units
      const uint32_t b_off = (uint32_t)tap * (kHBBytes >> 4) + ks * 2;
      if (kBig) umma_f16(tmem_d, tch_desc(kDescHiPatch, lo_a + a_off), tch_desc(kDescHiK128, lo_w + b_off), idesc2, (tap | ks) != 0);
      else umma_f16(tmem_d + 64, tch_desc(kDescHiPatch, lo_a + a_off), tch_desc(kDescHiK128, lo_w + b_off), idesc, true);
    }
  }
}
// first tile of a CTA: one kernel row of the big plane at a time, as the resident weights arrive (row ky's barrier has completed)
template <int KS>
__device__ __forceinline__ void halo_issue_big_row(uint32_t tmem_d, uint32_t lo_a_row, uint32_t lo_w_row, bool first) {
  constexpr uint32_t idesc2 = umma_idesc_f16(kCM, 128);
#pragma unroll
  for (int t = 0; t < 3; ++t) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
      umma_f16(tmem_d, tch_desc(kDescHiPatch, lo_a_row + (uint32_t)t * (128u >> 4) + ks * 2),
               tch_desc(kDescHiK128, lo_w_row + (uint32_t)t * (kHBBytes >> 4) + ks * 2), idesc2, (t | ks) != 0 || !first);
  }
}
__device__ __forceinline__ void halo_issue_big_row_n(int ksteps, uint32_t tmem_d, uint32_t lo_a_row, uint32_t lo_w_row, bool first) {
  switch (ksteps) {
    case 4: halo_issue_big_row<4>(tmem_d, lo_a_row, lo_w_row, first); break;
    case 3: halo_issue_big_row<3>(tmem_d, lo_a_row, lo_w_row, first); break;
    case 2: halo_issue_big_row<2>(tmem_d, lo_a_row, lo_w_row, first); break;
    case 1: halo_issue_big_row<1>(tmem_d, lo_a_row, lo_w_row, first); break;
    default: break;
  }
}
template <bool kBig>
__device__ __forceinline__ void halo_issue_plane_n(int ksteps, uint32_t tmem_d, uint32_t lo_a, uint32_t lo_w) {
  switch (ksteps) {
    case 4: halo_issue_plane<4, kBig>(tmem_d, lo_a, lo_w); break;
    case 3: halo_issue_plane<3, kBig>(tmem_d, lo_a, lo_w); break;
    case 2: halo_issue_plane<2, kBig>(tmem_d, lo_a, lo_w); break;
    case 1: halo_issue_plane<1, kBig>(tmem_d, lo_a, lo_w); break;
    default: break;   // timing knock-out
  }
}

constexpr int kHPatchBytes = kHPW * kHPH * 128;                     // 23040
constexpr int kHSlotBytes = (kHPatchBytes + 1023) / 1024 * 1024;    // 23552
constexpr int kHAStageBytes = 2 * kHSlotBytes;                      // big | small
constexpr int kHAStagesN = 2;
constexpr int kHTaps = 3;                                           // taps per weight-ring stage
constexpr int kHBStageBytes = kHTaps * kHBBytes;
constexpr int kHBStagesN = 2;
constexpr int kHaloSmemBytes = kHAStagesN * kHAStageBytes + kHBStagesN * kHBStageBytes + 1024 + 512;

__global__ void __launch_bounds__(kCThreads, 1) conv_tch_halo_kernel(const dtb200_conv_params p, const __grid_constant__ HMaps maps,
                                                                     KLayoutH kl, HWork wk, unsigned long long* trace) {
  tch_trace_begin(trace);
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
  if (warp == kCLoadA && lane < 2 * p.num_src)   // descriptor fetch off the first load's critical path
    prefetch_tensormap(lane & 1 ? &maps.small[lane >> 1] : &maps.big[lane >> 1]);
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
  if (tid == 0) griddep_launch_dependents();   // programmatic dependent launch, see conv_tch_kernel
  if (warp != kCLoadB) griddep_wait();

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
      TchResid res;
      const bool has_res = p.residual != nullptr && live;
      if (has_res) tch_resid_load(p, m, n_base + chalf * 32, res);   // in flight while this warp waits for the accumulator
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
        if (live && !(wk.debug & 0x800)) tch_finish_chunk(p, v, m, n_base + cc, has_res ? &res : nullptr);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  } else if (warp == kCMmaWarp) {
    // ============================================================ MMA issuer
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t lo_ring_a = ((smem_u32(ring_a) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t lo_ring_b = ((smem_u32(ring_b) & 0x3FFFFu) >> 4) | (1u << 16);
    static_assert(kHTaps == 3, "a weight-ring stage is one kernel row");
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
            const uint32_t row = (uint32_t)g * (kHPW * 128u >> 4);   // kernel row ky = g: whole patch rows, in 16-byte units
            halo_issue_row_n((wk.debug & 0x400) ? 0 : ksteps, tmem_d, lo_big + row, lo_small + row, lo_b_stage, (ch | g) == 0);
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
        if ((wk.debug & 0x200) && (phase || item != (long long)blockIdx.x)) {
          if (elect_one()) mbar_arrive(&a_full[stage]);
        } else if (elect_one()) {
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
          if ((wk.debug & 0x100) && (phase || item != (long long)blockIdx.x)) {
            if (elect_one()) mbar_arrive(&b_full[stage]);
          } else if (elect_one()) {
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
  tch_trace_end(trace);
  if (warp == kCMmaWarp) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ======================================================================================================================
// Halo-tile kernel with RESIDENT WEIGHTS for 3x3 / stride-1 layers with at most 64 input channels and 64 output channels (the
// 64->64 and 24->64 blocks: 18 of the 24 layers at 240x320 and most of the 120x160 ones at cfg 2).
// The streaming halo kernel above re-reads the layer's whole weight set (9 taps x 16 KB = 144 KB) for EVERY tile through a
// 2-stage ring: measured (tools/conv_bench.py knock-outs, profiles/r02d_*), the kernel is bound by that ring's round trips
// (tcgen05.commit -> loader wake-up -> bulk copy -> MMA wait), not by bandwidth or MMA time -- tensor pipe 31 % active, and
// skipping every copy, MMA and store still leaves 12 of 25 us.  Here the 144 KB stay in shared memory for the CTA's whole
// life (one bulk copy in the prologue, before the programmatic-dependency wait: weights are static), and the patch ring
// becomes three single-PLANE slots: a tile issues all A_big MMAs (into [main | corr]) and then all A_small MMAs (into corr), so
// the big plane's slot is free for the next tile while the small plane is still being read.  Per tile: 2 TMA boxes, 72
// MMAs issued back to back with two mbarrier waits, no weight traffic.
// ======================================================================================================================
constexpr int kHRWBytes = 9 * kHBBytes;                             // 147456: resident [tap][W_big | W_small]
constexpr int kHRSlots = 3;                                         // patch-plane ring
constexpr int kHaloResSmemBytes = kHRWBytes + kHRSlots * kHSlotBytes + 1024 + 512;
constexpr int kHRMmaWarp2 = 11;                                     // second MMA issuer
constexpr int kHRThreads = 12 * 32;

__global__ void __launch_bounds__(kHRThreads, 1) conv_tch_halo_res_kernel(const dtb200_conv_params p, const __grid_constant__ HMaps maps,
                                                                         KLayoutH kl, HWork wk, unsigned long long* trace) {
  tch_trace_begin(trace);
  constexpr int BN = 64;
  constexpr int kAccCols = 2 * BN, kTmemCols = 2 * kAccCols;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* w_res = smem;
  uint8_t* ring_a = w_res + kHRWBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring_a + kHRSlots * kHSlotBytes);
  uint64_t* a_full = bars;                  // [3] one patch plane landed (1 arrival + tx)
  uint64_t* a_empty = a_full + kHRSlots;    // [3] tcgen05.commit after the plane's last MMA
  uint64_t* w_full = a_empty + kHRSlots;    // [3] resident weights of kernel row ky landed
  uint64_t* acc_full = w_full + 3;          // [2] small-plane issuer -> epilogue
  uint64_t* acc_empty = acc_full + 2;       // [2] epilogue -> big-plane issuer
  uint64_t* big_done = acc_empty + 2;       // [2] big-plane issuer -> small-plane issuer (tcgen05.commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(big_done + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == kCLoadA && lane < 2 * p.num_src)   // descriptor fetch off the first load's critical path
    prefetch_tensormap(lane & 1 ? &maps.small[lane >> 1] : &maps.big[lane >> 1]);
  // development timeline (debug bit 0x1000): CTA 0 records clock64 at the hand-over points of every role and prints them
  __shared__ long long tr[6][24];
  __shared__ long long tr_t0;
  const bool tracing = (wk.debug & 0x1000) && blockIdx.x == 0;
#define HR_TRACE(rowi, idx) do { if (tracing && lane == 0 && (idx) < 24) tr[rowi][idx] = clock64() - tr_t0; } while (0)
  if (tid == 0) {
    tr_t0 = clock64();
    for (int s = 0; s < kHRSlots; ++s) mbar_init(&a_full[s], 1), mbar_init(&a_empty[s], 1);
    for (int s = 0; s < 3; ++s) mbar_init(&w_full[s], 1);
    for (int s = 0; s < 2; ++s) mbar_init(&acc_full[s], 1), mbar_init(&acc_empty[s], kCEpiWarps), mbar_init(&big_done[s], 1);
    fence_mbar_init();
  }
  if (warp == kCMmaWarp) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) griddep_launch_dependents();   // programmatic dependent launch, see conv_tch_kernel
  if (warp != kCLoadB) griddep_wait();

  auto decode = [&](long long item, int& bb, int& y0, int& x0) {
    const int per_img = wk.tiles_x * wk.tiles_y;
    bb = (int)(item / per_img);
    const int t = (int)(item - (long long)bb * per_img);
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
      int bb, y0, x0;
      decode(item, bb, y0, x0);
      const int buf = use & 1;
      const int oy = y0 + ty, ox = x0 + tx;
      const bool live = oy < p.out_h && ox < p.out_w;
      const long long m = ((long long)bb * p.out_h + oy) * p.out_w + ox;
      const int cc = chalf * 32;   // BN = 64: one 32-column chunk per warp
      TchResid res;
      const bool has_res = p.residual != nullptr && live;
      if (has_res) tch_resid_load(p, m, cc, res);   // in flight while this warp waits for the accumulator
      mbar_wait(&acc_full[buf], (use >> 1) & 1, 1, 128);
      tc_fence_after();
      if (warp == 0) HR_TRACE(3, use * 3);
      const uint32_t taddr = tmem_base + (uint32_t)(buf * kAccCols) + ((uint32_t)(qd * 32) << 16);
      float v[32], c[32];
      tmem_ld32(taddr + (uint32_t)cc, v);
      tmem_ld32(taddr + (uint32_t)(BN + cc), c);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);   // the accumulator is in registers: the next-but-one tile may start
      if (warp == 0) HR_TRACE(3, use * 3 + 1);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaf(c[j], kHalfSplitInv, v[j]);
      if (live && !(wk.debug & 0x800)) tch_finish_chunk(p, v, m, cc, has_res ? &res : nullptr, (wk.debug & 0x2000) != 0);
      if (warp == 0) HR_TRACE(3, use * 3 + 2);
    }
  } else if (warp == kCMmaWarp) {
    // ============================================================ MMA issuer 1: the big-plane pass of every tile
    // Two issuing warps on purpose.  tcgen05.mma issue is synchronous with the tensor pipe, so each mbarrier wait of a single
    // issuer (~190 clk, three per tile) and each descriptor preamble is tensor idle time: measured 6000 clk per tile where the
    // MMAs need 4032.  With the passes on two warps the pipe takes whatever is issuable: while this warp waits for tile t+1's
    // accumulator or patch, the other one is issuing tile t's small-plane pass.
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t lo_ring_a = ((smem_u32(ring_a) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t lo_w = ((smem_u32(w_res) & 0x3FFFFu) >> 4) | (1u << 16);
    const int cvalid = kl.src_c[0];
    const int ksteps = (wk.debug & 0x400) ? 0 : (cvalid >= kCK ? 4 : (cvalid + 15) >> 4);
    int use = 0;
    for (long long item = blockIdx.x; item < wk.total; item += gridDim.x, ++use) {
      const int buf = use & 1, q = 2 * use, slot = q % kHRSlots, ph = (q / kHRSlots) & 1;
      mbar_wait(&acc_empty[buf], ((use >> 1) & 1) ^ 1, 3);
      HR_TRACE(0, 1 + use);
      mbar_wait(&a_full[slot], ph, 4);
      tc_fence_after();
      HR_TRACE(1, use * 4);
      const uint32_t tmem_d = tmem_u + (uint32_t)(buf * kAccCols);
      const uint32_t lo_a = lo_ring_a + (uint32_t)slot * (kHSlotBytes >> 4);
      if (use == 0) {   // the CTA's first pass starts as soon as the first kernel row of weights is in
        for (int ky = 0; ky < 3; ++ky) {
          mbar_wait(&w_full[ky], 0, 5);
          tc_fence_after();
          if (ky == 2) HR_TRACE(0, 0);
          if (elect_one())
            halo_issue_big_row_n(ksteps, tmem_d, lo_a + (uint32_t)ky * (kHPW * 128u >> 4), lo_w + (uint32_t)ky * 3u * (kHBBytes >> 4), ky == 0);
          __syncwarp();
        }
      }
      if (elect_one()) {
        if (use != 0) halo_issue_plane_n<true>(ksteps, tmem_d, lo_a, lo_w);
        umma_commit(&a_empty[slot]);
        umma_commit(&big_done[buf]);
      }
      __syncwarp();
      HR_TRACE(1, use * 4 + 1);
    }
  } else if (warp == kHRMmaWarp2) {
    // ============================================================ MMA issuer 2: the small-plane pass (corr += A_small x W_big^T)
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t lo_ring_a = ((smem_u32(ring_a) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t lo_w = ((smem_u32(w_res) & 0x3FFFFu) >> 4) | (1u << 16);
    const int cvalid = kl.src_c[0];
    const int ksteps = (wk.debug & 0x400) ? 0 : (cvalid >= kCK ? 4 : (cvalid + 15) >> 4);
    for (int ky = 0; ky < 3; ++ky) mbar_wait(&w_full[ky], 0, 5);   // this thread's MMAs read the resident weights too
    int use = 0;
    for (long long item = blockIdx.x; item < wk.total; item += gridDim.x, ++use) {
      const int buf = use & 1, q = 2 * use + 1, slot = q % kHRSlots, ph = (q / kHRSlots) & 1;
      mbar_wait(&a_full[slot], ph, 4);
      mbar_wait(&big_done[buf], (use >> 1) & 1, 8);   // the first big-plane MMA overwrites [main | corr]: order behind that pass
      tc_fence_after();
      HR_TRACE(1, use * 4 + 2);
      const uint32_t tmem_d = tmem_u + (uint32_t)(buf * kAccCols);
      const uint32_t lo_a = lo_ring_a + (uint32_t)slot * (kHSlotBytes >> 4);
      if (elect_one()) {
        halo_issue_plane_n<false>(ksteps, tmem_d, lo_a, lo_w);
        umma_commit(&a_empty[slot]);
        umma_commit(&acc_full[buf]);
      }
      __syncwarp();
      HR_TRACE(1, use * 4 + 3);
    }
  } else if (warp == kCLoadA) {
    // ============================================================ A loader: one TMA box per patch plane
    const uint32_t ring_u = smem_u32(ring_a);
    int slot = 0, ph = 0, nq = 0;
    bool first = true;   // planes are loaded in consumption order big(t), small(t), big(t+1), ...: plane q lives in slot q % 3
    for (long long item = blockIdx.x; item < wk.total; item += gridDim.x) {
      int bb, y0, x0;
      decode(item, bb, y0, x0);
      for (int plane = 0; plane < 2; ++plane, ++nq) {
        mbar_wait(&a_empty[slot], ph ^ 1, 6, 64);
        HR_TRACE(2, nq);
        if ((wk.debug & 0x200) && !first) {
          if (elect_one()) mbar_arrive(&a_full[slot]);
        } else if (elect_one()) {
          mbar_arrive_expect_tx(&a_full[slot], kHPatchBytes);
          tma_load_4d_h(ring_u + (uint32_t)slot * kHSlotBytes, plane ? &maps.small[0] : &maps.big[0], 0, x0 - 1, y0 - 1, bb, &a_full[slot]);
        }
        __syncwarp();
        if (++slot == kHRSlots) slot = 0, ph ^= 1, first = false;
      }
    }
  } else if (warp == kCLoadB) {
    // ============================================================ weights: once per CTA (not a dependency of the previous layer)
    if (elect_one()) {
#pragma unroll 1
      for (int ky = 0; ky < 3; ++ky) {
        mbar_arrive_expect_tx(&w_full[ky], 3 * kHBBytes);
        for (int t = 0; t < 3; ++t)
          bulk_g2s(w_res + (size_t)(ky * 3 + t) * kHBBytes, reinterpret_cast<const uint8_t*>(p.weight) + (size_t)(ky * 3 + t) * kHBBytes, kHBBytes,
                   &w_full[ky]);
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  tch_trace_end(trace);
  if (warp == kCMmaWarp) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
  if (tracing && tid == 0) {
    const int tiles = (int)((wk.total - 1 - blockIdx.x) / gridDim.x) + 1;
    printf("halo_res trace (clk since prologue start, CTA 0, %d tiles): end %lld, weights landed %lld\n", tiles, clock64() - tr_t0, tr[0][0]);
    for (int t = 0; t < tiles && t < 6; ++t)
      printf("  tile %d: acc_empty seen %lld | loadA issue big %lld small %lld | big: full seen %lld issued %lld | small: full seen %lld issued %lld | "
             "epi: acc_full seen %lld tmem read %lld stored %lld\n",
             t, tr[0][1 + t], tr[2][2 * t], tr[2][2 * t + 1], tr[1][4 * t], tr[1][4 * t + 1], tr[1][4 * t + 2], tr[1][4 * t + 3], tr[3][3 * t],
             tr[3][3 * t + 1], tr[3][3 * t + 2]);
  }
#undef HR_TRACE
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
__global__ void resample_copy_h_kernel(const dtb200_conv_params p, unsigned long long* trace) {
  tch_trace_begin(trace);
  griddep_launch_dependents();
  griddep_wait();
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
  tch_trace_end(trace);
}

// 1x1 conv to few output channels from split16 sources, fp32 NHWC output (the log-depth heads).  Eight lanes per pixel, each
// with 16-byte loads of 8 channels per plane (the first version read single halves, 64 bytes per warp instruction: 21-36 us for
// the 240x320 head, which is the LAST op of the plan's critical path), fixed 3-step shuffle tree.
__global__ void __launch_bounds__(256) conv_head_h_kernel(const dtb200_conv_params p, long long pixels, unsigned long long* trace) {
  tch_trace_begin(trace);
  griddep_launch_dependents();
  griddep_wait();
  const int t = threadIdx.x & 7;
  const long long pix = (long long)blockIdx.x * 32 + (threadIdx.x >> 3);
  const bool live = pix < pixels;
  const long long pc = live ? pix : pixels - 1;   // whole warps stay converged for the shuffles
  for (int n = 0; n < p.out_c; ++n) {
    float s = 0.f;
    int cg = 0;
    for (int si = 0; si < p.num_src; ++si) {
      const int C = p.src_c[si];
      const uint8_t* x = reinterpret_cast<const uint8_t*>(p.src[si]) + (size_t)pc * 4 * C;
      for (int c = 8 * t; c < C; c += 64) {
        float v[8];
        const uint4 bq = ldg128u(x + 2 * c), sq = ldg128u(x + 2 * C + 2 * c);
        const float2 a0 = join_pair(bq.x, sq.x), a1 = join_pair(bq.y, sq.y), a2 = join_pair(bq.z, sq.z), a3 = join_pair(bq.w, sq.w);
        v[0] = a0.x, v[1] = a0.y, v[2] = a1.x, v[3] = a1.y, v[4] = a2.x, v[5] = a2.y, v[6] = a3.x, v[7] = a3.y;
        const float* w = p.weight + (long long)(cg + c) * p.out_c + n;
#pragma unroll
        for (int j = 0; j < 8; ++j) s = DT_FMA(v[j], __ldg(w + (long long)j * p.out_c), s);
      }
      cg += C;
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s = DT_ADD(s, __shfl_xor_sync(0xffffffffu, s, o));
    if (t == 0 && live) {
      float v = p.bias ? DT_ADD(s, p.bias[n]) : s;
      if (p.residual) v = DT_ADD(v, p.residual[pix * p.out_c + n]);
      p.dst[pix * p.out_c + n] = activate(v, p.act, p.act_slope);
    }
  }
  tch_trace_end(trace);
}

// ======================================================================================================================
// host side
// ======================================================================================================================
int launch_conv_simt(const dtb200_conv_params& p, int in_c_total, cudaStream_t stream);
int conv_debug_flags();   // conv_tc.cu
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

// Split-K plan for maps with fewer (M, N) tiles than SMs: `splits` CTAs of one thread-block cluster share a tile (portable
// cluster size <= 8).  Cost ~ rounds x (K blocks per CTA + fixed per-item cost) + the DSMEM reduction.  Clusters are placed
// inside a GPC, so fewer than 148 / splits of them are co-resident; the table is a fixed, conservative model of a 148-SM B200
// (8 GPCs) on purpose: the plan -- and the fp32 summation order it implies -- must not depend on the chip's floorsweeping.
static int tch_clusters_per_round(int sp) {
  static const int t[9] = {0, 148, 72, 48, 32, 24, 24, 16, 16};
  return t[sp];
}
static int tch_splits(long long m_tiles, int out_c, int num_kb) {
  const long long tiles = m_tiles * (out_c / tch_bn(out_c));
  if (tiles >= 148 || num_kb < 4) return 1;
  int best = 1;
  double best_cost = 1e30;
  for (int sp = 1; sp <= 8 && sp * 2 <= num_kb; ++sp) {
    const int kb_per = (num_kb + sp - 1) / sp;
    if ((num_kb + kb_per - 1) / kb_per != sp) continue;
    const long long rounds = (tiles + tch_clusters_per_round(sp) - 1) / tch_clusters_per_round(sp);
    const double cost = (double)rounds * (kb_per + 2.0) + (sp > 1 ? 1.0 + 0.125 * sp : 0.0);
    if (cost < best_cost - 1e-9) best_cost = cost, best = sp;
  }
  return best;
}

uint64_t conv_tch_workspace_bytes(const dtb200_conv_params& p) {
  (void)p;
  return 0;   // split-K partial tiles live in the cluster's shared memory
}

// Launch with optional thread-block cluster and programmatic dependent launch.  PDL: a kernel launched with the attribute may
// begin before its same-stream predecessor has finished; every kernel launched this way executes griddepcontrol.wait before it
// touches activations (so completion stays transitive along a lane).  Under stream capture the attribute becomes a programmatic
// graph edge.  DTB200_CONV_PDL=0 turns it off (A/B measurements).
// development timeline: dtb200_debug_trace(buf, capacity) hands out one (start, end) pair of the caller's device buffer per
// launch, in launch order; under stream capture the pair's address is baked into the graph node
static unsigned long long* g_trace_buf = nullptr;
static int g_trace_cap = 0, g_trace_next = 0;
int conv_tch_trace_set(unsigned long long* buf, int capacity) {
  g_trace_buf = buf;
  g_trace_cap = buf ? capacity : 0;
  g_trace_next = 0;
  return DTB200_OK;
}
static unsigned long long* tch_trace_slot() {
  if (!g_trace_buf || g_trace_next >= g_trace_cap) return nullptr;
  return g_trace_buf + 2 * (size_t)(g_trace_next++);
}

static bool tch_pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("DTB200_CONV_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}
template <typename... KArgs, typename... Args>
static cudaError_t tch_launch(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t stream, int cluster,
                              Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(block, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  unsigned n = 0;
  if (cluster > 1) {
    attrs[n].id = cudaLaunchAttributeClusterDimension;
    attrs[n].val.clusterDim.x = (unsigned)cluster;
    attrs[n].val.clusterDim.y = 1;
    attrs[n].val.clusterDim.z = 1;
    ++n;
  }
  if (tch_pdl_enabled()) {
    attrs[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attrs;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
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
    cudaFuncSetAttribute(conv_tch_halo_res_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHaloResSmemBytes);
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
  cudaError_t e = tch_launch(resample_copy_h_kernel, (unsigned)blocks, 256, 0, stream, 1, p, tch_trace_slot());
  if (e != cudaSuccess) return fail(DTB200_ERR_CUDA, "resample_copy_h_kernel: %s", cudaGetErrorString(e));
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
    cudaError_t e = tch_launch(conv_head_h_kernel, (unsigned)((pixels + 31) / 32), 256, 0, stream, 1, p, pixels, tch_trace_slot());
    if (e != cudaSuccess) return fail(DTB200_ERR_CUDA, "conv_head_h_kernel: %s", cudaGetErrorString(e));
    return check_launch("conv_head_h_kernel");
  }
  if (reinterpret_cast<uintptr_t>(p.dst) % 32 != 0 || reinterpret_cast<uintptr_t>(p.residual) % 32 != 0)
    return fail(DTB200_ERR_INVALID, "conv (tch): dst / residual must be 32-byte aligned%s");
  const int num_sms = conv_tch_init();
  if (!g_encode_h) return fail(DTB200_ERR_CUDA, "conv (tch): cuTensorMapEncodeTiled entry point not available%s");
  const KLayoutH kl = make_klayout_h(p.num_src, p.src_c, p.ksize);
  const int bn = tch_bn(p.out_c);
  HWork wk;
  tch_tile_shape(p.out_h, p.out_w, wk.tw, wk.th);
  wk.tiles_x = (p.out_w + wk.tw - 1) / wk.tw;
  wk.tiles_y = (p.out_h + wk.th - 1) / wk.th;
  wk.m_tiles = p.batch * wk.tiles_x * wk.tiles_y;
  const int splits = tch_splits(wk.m_tiles, p.out_c, kl.num_kb);
  wk.num_kb_total = kl.num_kb;
  wk.kb_per_split = (kl.num_kb + splits - 1) / splits;
  wk.splits = (kl.num_kb + wk.kb_per_split - 1) / wk.kb_per_split;
  wk.n_tiles = p.out_c / bn;
  wk.total = (long long)wk.m_tiles * wk.n_tiles;   // (M, N) tiles; a cluster of `splits` CTAs works on each
  wk.debug = conv_debug_flags() & ~0xff;

  HMaps maps;
  cudaError_t e;
  // 3x3 / stride-1 layers with 64-channel N tiles and at least one full round of 8 x 16 tiles: halo-tile kernel
  if (p.ksize == 3 && p.stride == 1 && bn == 64 && wk.splits == 1 && p.in_h == p.out_h && p.in_w == p.out_w) {
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
      // Resident-weights variant (development switch bit 4; <= 64 input channels, 64 output channels, at least three tiles per
      // CTA).  Measured in round 2 (profiles/r02d_*): alone it is the faster kernel (two issuing warps: 4700 clk per tile against
      // 6000; 19.9 vs 21.2 us per 240x320 64->64 layer), but inside the DAG graph the whole plan is 43 us SLOWER with it
      // (1890 vs 1846 us, same box, repeated): where its layers alternate with another lane's 148-CTA kernels the scheduler
      // leaves 5.6 us between them instead of 2.  The streaming kernel stays the default; tests/test_gpu_networks.py runs both.
      if (kl.kb_per_tap == 1 && p.out_c == 64 && hw.total >= 3 * 148 && (conv_debug_flags() & 16)) {
        e = tch_launch(conv_tch_halo_res_kernel, grid, kHRThreads, kHaloResSmemBytes, stream, 1, p, maps, kl, hw, tch_trace_slot());
        if (e != cudaSuccess) return fail(DTB200_ERR_CUDA, "conv_tch_halo_res_kernel: %s", cudaGetErrorString(e));
        return check_launch("conv_tch_halo_res_kernel");
      }
      e = tch_launch(conv_tch_halo_kernel, grid, kCThreads, kHaloSmemBytes, stream, 1, p, maps, kl, hw, tch_trace_slot());
      if (e != cudaSuccess) return fail(DTB200_ERR_CUDA, "conv_tch_halo_kernel: %s", cudaGetErrorString(e));
      return check_launch("conv_tch_halo_kernel");
    }
  }
  int rc = encode_split16_maps(p, maps, (cuuint32_t)(wk.tw * p.stride), (cuuint32_t)(wk.th * p.stride), (cuuint32_t)p.stride);
  if (rc != DTB200_OK) return rc;
  // one cluster per tile, as many clusters as a round holds (a cluster of `splits` CTAs lives inside one GPC)
  const long long max_clusters = wk.splits > 1 ? tch_clusters_per_round(wk.splits) : num_sms;
  const unsigned grid = (unsigned)(wk.total < max_clusters ? wk.total : max_clusters) * (unsigned)wk.splits;
  if (bn == 128) e = tch_launch(conv_tch_kernel<128>, grid, kCThreads, TchCfg<128>::kSmemBytes, stream, wk.splits, p, maps, kl, wk, tch_trace_slot());
  else e = tch_launch(conv_tch_kernel<64>, grid, kCThreads, TchCfg<64>::kSmemBytes, stream, wk.splits, p, maps, kl, wk, tch_trace_slot());
  if (e != cudaSuccess) return fail(DTB200_ERR_CUDA, "conv_tch_kernel: %s", cudaGetErrorString(e));
  return check_launch("conv_tch_kernel");
}

}  // namespace dtb200
