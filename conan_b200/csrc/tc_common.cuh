// tcgen05 / TMEM / TMA / mbarrier building blocks shared by the tensor-core kernels (sm_100a inline PTX), and the
// host helpers around them (tensor-map cache, co-residency computation).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace conan {

constexpr int TILE_M = 128;

// ------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
// Polling costs shared-memory bandwidth that the tensor core also needs for its A / B operand reads (SS-mode MMAs
// stream 64 B/cycle from shared memory): waits that are expected to be long are done by ONE lane per warp with a
// sleep between polls; the rest of the warp parks on __syncwarp, which also orders memory among the lanes.
__device__ __forceinline__ void mbar_wait_lane0(uint64_t* bar, uint32_t parity, unsigned ns) {
  if ((threadIdx.x & 31) == 0) {
    while (!mbar_try_wait(bar, parity)) { if (ns) __nanosleep(ns); }
  }
  __syncwarp();
}
// wait executed by a whole converged warp with a warp-uniform exit condition (vote): the code after it stays provably
// convergent, which lets ptxas keep loop-carried descriptors / counters of the issuing loops in uniform registers
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
  while (!__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) {}
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// Programmatic dependent launch: the next tensor-core kernel of the step may start its prologue (barrier
// init, TMEM allocation, bias staging, resident-weight TMA) while this one drains; it touches activations
// only after pdl_wait(), which returns when every prerequisite grid has completed and flushed.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// one deterministic leader lane of a converged warp.  tcgen05.mma / tcgen05.commit / TMA issues are wrapped in
// `if (elect_one_sync())` inside WARP-UNIFORM loops: descriptors and coordinates then live in uniform registers.
// (Issuing them from an `if (lane == 0)` region instead makes every operand a per-thread value and the compiler
// wraps each UTCHMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop -- ~100 cycles of single-thread issue per MMA.)
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// same load without the wait: several loads can be in flight before one tcgen05.wait::ld
__device__ __forceinline__ void tc_ld_32x32b_x16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

// NK consecutive K = 16 steps of one tap (A and B advance 32 bytes inside the swizzle atom per step) issued by one
// elected lane from a single PTX block: one elect per tap instead of one per MMA, and a straight-line sequence for ptxas.
// `first` != 0 makes the very first MMA overwrite the accumulator.
template <int NK>
__device__ __forceinline__ void tc_mma_f16_tap(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t first) {
  static_assert(NK == 2 || NK == 4, "C = 32 or 64");
  if (NK == 4) {
    asm volatile(
        "{\n"
        ".reg .pred p, q, t;\n"
        ".reg .b64 a1, b1, a2, b2, a3, b3;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "setp.eq.b32 q, %4, 0;\n"
        "setp.eq.b32 t, 0, 0;\n"
        "add.u64 a1, %1, 2;\n add.u64 b1, %2, 2;\n"
        "add.u64 a2, %1, 4;\n add.u64 b2, %2, 4;\n"
        "add.u64 a3, %1, 6;\n add.u64 b3, %2, 6;\n"
        "@p tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n"
        "@p tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, t;\n"
        "@p tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, t;\n"
        "@p tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, t;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(first) : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p, q, t;\n"
        ".reg .b64 a1, b1;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "setp.eq.b32 q, %4, 0;\n"
        "setp.eq.b32 t, 0, 0;\n"
        "add.u64 a1, %1, 2;\n add.u64 b1, %2, 2;\n"
        "@p tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n"
        "@p tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, t;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(first) : "memory");
  }
}

// ------------------------------------------------------------------------------ CTA pair (cta_group::2) building blocks
// Two CTAs of a cluster (the two SMs of a TPC) execute one M = 256 MMA: each holds its own 128 rows of A, half of the B rows and
// the accumulator of its own rows in its own TMEM; the LEADER (cluster rank 0) issues the instruction.  Barrier addresses with
// the peer bit cleared name the leader's copy of a barrier from either CTA (same convention as CUTLASS's Sm100MmaPeerBitMask).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA loads whose completion bytes are counted on the LEADER's mbarrier
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// commit of the pair's MMAs, arriving on the barrier at this shared-memory offset in BOTH CTAs
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// arrive on the leader's copy of a barrier (from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}

// distributed shared memory: address of `p` (a pointer into this CTA's shared memory) in the shared memory of CTA `rank` of the cluster
__device__ __forceinline__ uint32_t dsmem_addr(const void* p, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(p)), "r"(rank));
  return ra;
}
__device__ __forceinline__ float4 dsmem_ld_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// K-major, swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4, [16,30) LBO >> 4 (= 1 for swizzled K-major), [32,46) SBO >> 4
//   (8 rows * swizzle span), [46,48) version = 1, [61,64) layout (2 = SWIZZLE_128B, 4 = SWIZZLE_64B)
template <int SWIZZLE_BYTES>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  constexpr uint64_t layout = SWIZZLE_BYTES == 128 ? 2 : 4;
  constexpr uint64_t sbo = (8 * SWIZZLE_BYTES) >> 4;
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | (sbo << 32) | ((uint64_t)1 << 46) | (layout << 61);
}

// instruction descriptor for kind::f16: fp16 A/B (K-major), fp32 accumulate, M = 128 (256 for a CTA pair), N = BN
template <int BN, int M = TILE_M>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
  return (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// generic-proxy writes to shared memory -> visible to the async proxy (tcgen05.mma / TMA) after the next barrier
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- host side (defined in conv_gemm_tc.cu)
// cached cuTensorMapEncodeTiled of an fp16 tensor: rank 2 or 3, dims d0 (innermost) .. d2, strides in bytes, box b0..b2
int get_tensor_map(CUtensorMap* out, const void* ptr, int rank, unsigned long long d0, unsigned long long d1, unsigned long long d2,
                   unsigned long long s1_bytes, unsigned long long s2_bytes, unsigned b0, unsigned b1, unsigned b2, int swz_bytes);
// CTAs of a kernel that fit on one SM with the carve-out preference "max shared"
int resident_ctas(const void* func, int threads, size_t dyn_smem, int tmem_cols);
int num_sms();

}  // namespace conan
