// sm_100a tensor-core plumbing shared by the tcgen05 kernels: mbarrier, bulk async copy, TMEM alloc/ld, UMMA descriptors,
// the 3xTF32 operand split.  Inline PTX only (no CUTLASS); the bit layouts follow cute/arch/mma_sm100_desc.hpp.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace dtb200 {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a CONVERGED warp.  Single-thread instructions on the uniform datapath (tcgen05.mma / commit, bulk copies)
// must be issued from warp-uniform control flow and predicated with elect.sync: under `if (lane == 0)` the compiler wraps
// each of them in an ELECT + R2UR waterfall loop (hundreds of cycles per instruction).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xFFFFFFFF;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug traps (CUDA error after ~2 s) instead of hanging the GPU box.
// `sleep_ns` > 0 backs off between polls (exponentially, capped at 8 x sleep_ns).  A polling warp issues ~8 instructions
// and one shared-memory access every ~30 clk: measured in round 2, the waiting roles of the persistent kernels (epilogue,
// loaders, the idle MMA warp) took ~30 % of all issue slots away from the producer warps.  Roles whose wake-up latency is
// not on the critical path therefore sleep; the default (0) is the tight loop.
__device__ __forceinline__ bool mbar_try(uint32_t addr, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(addr), "r"(parity), "r"(0x989680u)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag = 0, uint32_t sleep_ns = 0) {
  const uint32_t addr = smem_u32(bar);
  if (mbar_try(addr, parity)) return;
  const long long t0 = clock64();
  uint32_t ns = sleep_ns;
  for (;;) {
    if (ns) {
      __nanosleep(ns);
      if (ns < 8 * sleep_ns) ns <<= 1;
    }
    if (mbar_try(addr, parity)) return;
    const long long waited = clock64() - t0;
    if (waited > 2000000000LL) {
      if (blockIdx.x == 0 && (threadIdx.x & 31) == 0)
        printf("mbar timeout: warp %d tag %d bar@%u parity %u\n", (int)(threadIdx.x >> 5), tag, addr, parity);
      if (waited > 2400000000LL) __trap();
    }
  }
}

// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// explicit shared-memory 128-bit accesses by 32-bit shared address
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// 256-bit global accesses (sm_100: LDG/STG.256), 32-byte aligned: one full sector per lane
__device__ __forceinline__ void ldg256(const float* p, float* v) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

// Ampere-style 16-byte async copy global -> shared (LDGSTS); src_bytes = 0 zero-fills (used for conv zero padding)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one (pre-counted) arrival once all prior cp.async of this thread have landed
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier (UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------------------------- clusters / DSMEM / PDL
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// all threads of all CTAs of the cluster (prologue: makes every CTA's mbarrier init visible before a peer arrives on it)
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of THIS CTA -> shared::cluster address of the same variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {   // release at cluster scope
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_acq_cluster(uint32_t addr, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(addr), "r"(parity), "r"(0x989680u)
      : "memory");
  return done != 0;
}
// bounded like mbar_wait; acquire at cluster scope (pairs with mbar_arrive_cluster of peer CTAs)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, int tag) {
  const uint32_t addr = smem_u32(bar);
  if (mbar_try_acq_cluster(addr, parity)) return;
  const long long t0 = clock64();
  for (;;) {
    if (mbar_try_acq_cluster(addr, parity)) return;
    const long long waited = clock64() - t0;
    if (waited > 2000000000LL) {
      if ((threadIdx.x & 31) == 0) printf("cluster mbar timeout: block %d warp %d tag %d parity %u\n", (int)blockIdx.x, (int)(threadIdx.x >> 5), tag, parity);
      if (waited > 2400000000LL) __trap();
    }
  }
}
__device__ __forceinline__ float4 ld_dsmem128(uint32_t cluster_addr) {
  float4 v;
  // volatile keeps the loads behind the cluster-scope acquire that precedes them; no "memory" clobber, so independent global
  // loads of the epilogue (bias, residual) may be scheduled around them
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(cluster_addr));
  return v;
}
__device__ __forceinline__ void prefetch_tensormap(const void* tm) { asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory"); }
// Programmatic dependent launch.  wait: returns once every prerequisite grid has completed and its writes are visible (at
// once when the launch carries no programmatic edge).  launch_dependents: lets the runtime schedule the dependent grid as
// soon as every CTA of this grid has issued it or exited -- the dependent's prologue then overlaps this grid's tail.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------- TMEM
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(kCols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// tcgen05.commit: the mbarrier receives one arrival when every previously issued MMA of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets TMEM lane (lane_base + i), columns [col, col+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, "
      "[%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------------------------------------- UMMA descriptors
// K-major operand tile, rows of 128 bytes (32 fp32), SWIZZLE_128B, 8-row groups 1024 B apart; tile base 1024-B aligned.
//   bits [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (=64) | [46,48) version=1
//   | [61,64) layout type (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor, kind::tf32, fp32 accumulate, A and B K-major:
//   [4,6) c_format=1 (F32) | [7,10) a_format=2 (TF32) | [10,13) b_format=2 | [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T ; one thread issues
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}

// byte offset of element (row, k) inside a [rows][32 fp32] SWIZZLE_128B K-major tile
__host__ __device__ __forceinline__ uint32_t sw128_offset(int row, int k) {
  return (uint32_t)row * 128u + (uint32_t)((((k >> 2) ^ (row & 7)) << 4) + ((k & 3) << 2));
}

// 3xTF32 split: x = big + small with big = x truncated to TF32 (exact), small = x - big (exact in fp32).
// a*b ~= big_a*big_b + big_a*small_b + small_a*big_b, error ~2^-21 |a||b|.
__host__ __device__ __forceinline__ float tf32_big(float x) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
#else
  union {
    float f;
    uint32_t u;
  } v;
  v.f = x;
  v.u &= 0xFFFFE000u;
  return v.f;
#endif
}


// ---------------------------------------------------------------------------------------------- 2-term fp16 split (TCH)
// x = big + small' / 2048 with big = fp16(x) (round to nearest) and small' = fp16((x - big) * 2048): 22 significant bits,
// |x - (big + small'/2048)| <= 2^-22 |x| for |x| in the fp16 range (denormals of both terms keep the ABSOLUTE error below
// 2^-36).  a*b ~= big_a*big_b + (big_a*small'_b + small'_a*big_b) / 2048: three kind::f16 MMAs (twice the TF32 rate, half
// the operand bytes of 3xTF32), the two correction products accumulate in their own TMEM columns and are scaled in the
// epilogue.  Values are clamped to the finite fp16 range first (inf * 0 inside an MMA would poison whole rows).
constexpr float kHalfSplitScale = 2048.f;
constexpr float kHalfSplitInv = 1.f / 2048.f;
constexpr float kHalfMax = 65504.f;

__device__ __forceinline__ void split_half2(float x0, float x1, uint32_t& big, uint32_t& small) {
  x0 = fminf(fmaxf(x0, -kHalfMax), kHalfMax);  // NaN stays NaN (fminf/fmaxf return the non-NaN operand: clamp maps NaN to
  x1 = fminf(fmaxf(x1, -kHalfMax), kHalfMax);  // -65504; callers that must keep NaN handle it before the split)
  const __half2 b = __floats2half2_rn(x0, x1);
  const float2 bf = __half22float2(b);
  const __half2 s = __floats2half2_rn((x0 - bf.x) * kHalfSplitScale, (x1 - bf.y) * kHalfSplitScale);
  big = *reinterpret_cast<const uint32_t*>(&b);
  small = *reinterpret_cast<const uint32_t*>(&s);
}
// The same split with one saturating pack instead of four clamps (F2FP.SATFINITE): finite values beyond +-65504 saturate, NaN
// stays NaN in both terms (the reference's convolutions propagate NaN too; an MMA row that holds a NaN is exactly the set of
// outputs whose window contains it).  Used by the conv epilogues, where the split is the bulk of the per-element work.
__device__ __forceinline__ void split_half2_sat(float x0, float x1, uint32_t& big, uint32_t& small) {
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(big) : "f"(x1), "f"(x0));
  const float2 bf = __half22float2(*reinterpret_cast<const __half2*>(&big));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(small) : "f"((x1 - bf.y) * kHalfSplitScale), "f"((x0 - bf.x) * kHalfSplitScale));
}
__device__ __forceinline__ float join_half(__half b, __half s) { return fmaf(__half2float(s), kHalfSplitInv, __half2float(b)); }

// byte offset of element (row, k) inside a [rows][64 fp16] SWIZZLE_128B K-major tile
__host__ __device__ __forceinline__ uint32_t sw128_offset_h(int row, int k) {
  return (uint32_t)row * 128u + (uint32_t)((((k >> 3) ^ (row & 7)) << 4) + ((k & 7) << 1));
}
// Instruction descriptor, kind::f16 with fp16 A/B (format 0), fp32 accumulate, A and B K-major
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}

}  // namespace tc
}  // namespace dtb200
